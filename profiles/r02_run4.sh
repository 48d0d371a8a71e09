#!/bin/bash
# round 2, GPU run 4 (2 GPUs): rsag exchange with one fence per CTA and phase; the round-1 F=100 line at N=1/2 with finish traces
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_multigpu.py -q -x > gpurun_out/r02/multigpu_tests_n2_b.log 2>&1; echo "rc=$?" >> gpurun_out/r02/multigpu_tests_n2_b.log
tail -3 gpurun_out/r02/multigpu_tests_n2_b.log
for N in 1 2; do
MSG_TRACE_FINISH=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --records 20000000 --steps 10 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02/bench_n${N}_run4.json 2> gpurun_out/r02/bench_n${N}_run4.err; echo "bench N=$N rc=$?"
grep '^{' gpurun_out/r02/bench_n${N}_run4.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg5', d['n_gpus'], d['value'], d['ms_per_step'])"; grep "msg finish" gpurun_out/r02/bench_n${N}_run4.err | tail -2
MSG_TRACE_FINISH=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --config 12 --gpus $N --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02/bench12_n${N}_run4.json 2> gpurun_out/r02/bench12_n${N}_run4.err; echo "bench12 N=$N rc=$?"
grep '^{' gpurun_out/r02/bench12_n${N}_run4.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg12', d['n_gpus'], d['value'], d['ms_per_step'])"; grep "msg finish" gpurun_out/r02/bench12_n${N}_run4.err | tail -4
done
