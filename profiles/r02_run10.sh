#!/bin/bash
# round 2, GPU run 10 (8 GPUs): multi-GPU parity at 8 and 4 ranks, the F=100 line at N=2 on an 8-GPU node (round-1 anomaly), config-5 scaling lines
mkdir -p gpurun_out/r02
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/r02/topo_8gpu.txt; nproc >> gpurun_out/r02/topo_8gpu.txt; free -g | head -2 >> gpurun_out/r02/topo_8gpu.txt
MSG_TEST_WORLD=8 timeout 600 python -m pytest tests/test_multigpu.py -q -k "genes1m or bigF or mid or fused or late or equal" > gpurun_out/r02/multigpu_tests_n8.log 2>&1; echo "rc=$?" >> gpurun_out/r02/multigpu_tests_n8.log; tail -3 gpurun_out/r02/multigpu_tests_n8.log
MSG_TEST_WORLD=4 timeout 600 python -m pytest tests/test_multigpu.py -q -k "genes1m or fused or cov" > gpurun_out/r02/multigpu_tests_n4.log 2>&1; echo "rc=$?" >> gpurun_out/r02/multigpu_tests_n4.log; tail -3 gpurun_out/r02/multigpu_tests_n4.log
run() { # N config extra-args tag
  MSG_TRACE_FINISH=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $1 --config $2 $3 --no-cpu-baseline > gpurun_out/r02/scale_$4.json 2> gpurun_out/r02/scale_$4.err; echo "$4 rc=$?"
  grep '^{' gpurun_out/r02/scale_$4.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$4', d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'parity', d['parity_checked'])"
  grep "msg finish" gpurun_out/r02/scale_$4.err | tail -2 | cut -c1-330
}
run 1 12 "--steps 20 --warmup 5 --e2e-steps 0" cfg12_n1
run 2 12 "--steps 20 --warmup 5 --e2e-steps 0" cfg12_n2
CUDA_VISIBLE_DEVICES=0,4 run 2 12 "--steps 20 --warmup 5 --e2e-steps 0" cfg12_n2_dev04
run 4 12 "--steps 20 --warmup 5 --e2e-steps 0" cfg12_n4
run 8 5 "--steps 10 --warmup 3" cfg5_n8
run 4 5 "--records 40000000 --steps 10 --warmup 3 --e2e-steps 0" cfg5_40M_n4
run 2 5 "--records 40000000 --steps 10 --warmup 3 --e2e-steps 0" cfg5_40M_n2
run 1 5 "--records 40000000 --steps 10 --warmup 3 --e2e-steps 0" cfg5_40M_n1
