#!/bin/bash
# round 2, GPU run 3 (2 GPUs): multi-GPU parity tests, bench at N=2 (config 5, 2 chunks), N=2 latency probe on the round-1 shape
mkdir -p gpurun_out/r02
nvidia-smi topo -m > gpurun_out/r02/topo_2gpu.txt 2>&1; nvidia-smi -L >> gpurun_out/r02/topo_2gpu.txt
timeout 900 python -m pytest tests/test_multigpu.py -q -x > gpurun_out/r02/multigpu_tests_n2.log 2>&1; echo "rc=$?" >> gpurun_out/r02/multigpu_tests_n2.log
tail -25 gpurun_out/r02/multigpu_tests_n2.log
for N in 1 2; do
MSG_TRACE_FINISH=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --records 20000000 --steps 5 --warmup 3 --e2e-chunks 2 --no-cpu-baseline > gpurun_out/r02/bench_n${N}_run3.json 2> gpurun_out/r02/bench_n${N}_run3.err; echo "bench N=$N rc=$?"
tail -c 2500 gpurun_out/r02/bench_n${N}_run3.json; grep "msg finish" gpurun_out/r02/bench_n${N}_run3.err | tail -4
done
