#!/bin/bash
# round 2, GPU run 9 (2 GPUs): rsag exchange phase breakdown (in-kernel globaltimer stamps of CTA 0), fence.acq_rel.sys
mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_multigpu.py -q -x -k "bigF or genes1m or late" > gpurun_out/r02/multigpu_tests_n2_c.log 2>&1; tail -2 gpurun_out/r02/multigpu_tests_n2_c.log
MSG_TRACE_FINISH=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --records 20000000 --steps 10 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r02/bench_n2_run9.json 2> gpurun_out/r02/bench_n2_run9.err; echo "bench N=2 rc=$?"
grep '^{' gpurun_out/r02/bench_n2_run9.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg5', d['n_gpus'], d['value'], d['ms_per_step'])"; grep "msg finish" gpurun_out/r02/bench_n2_run9.err | tail -4
