#!/bin/bash
# round 2, GPU run 1: large-F parity tests + launch list / ncu captures of the config-5 shape (state at round start)
mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_largeF.py -x -q > gpurun_out/r02/largeF_tests.log 2>&1; echo "largeF rc=$?" >> gpurun_out/r02/largeF_tests.log
python profiles/configs_probe.py 20000000 1,3,4,5 10 20 > gpurun_out/r02/configs_probe_start.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/launches_cfg5_start.csv \
    python profiles/configs_probe.py 10000000 5 2 2 > gpurun_out/r02/ncu_cfg5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'em_loop_kernel|fused_warp_kernel|fused_walk_kernel|decode_kernel' -s 8 -c 4 \
    -o gpurun_out/r02/prof_cfg5_start python profiles/configs_probe.py 10000000 5 2 2 > gpurun_out/r02/ncu_cfg5_full.log 2>&1
tail -3 gpurun_out/r02/largeF_tests.log; cat gpurun_out/r02/configs_probe_start.txt
