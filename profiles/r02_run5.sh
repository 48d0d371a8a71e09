#!/bin/bash
# round 2, GPU run 5: decode v3 (LDGSTS staging, straight-line fast path), bitmap coverage, word-wise gather: tests + probes + ncu
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02/gpu_tests_run5.log 2>&1; echo "rc=$?" >> gpurun_out/r02/gpu_tests_run5.log
tail -12 gpurun_out/r02/gpu_tests_run5.log
python profiles/configs_probe.py 20000000 1,3,4,5 10 20 > gpurun_out/r02/configs_probe_run5.txt 2>&1; cat gpurun_out/r02/configs_probe_run5.txt
ncu --set full --clock-control none --import-source on -k regex:'decode_kernel' -s 2 -c 1 \
    -o gpurun_out/r02/prof_decode_v3 python profiles/configs_probe.py 10000000 5 2 2 > gpurun_out/r02/ncu_decode_v3.log 2>&1
tail -2 gpurun_out/r02/ncu_decode_v3.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/launches_cfg14_run5.csv \
    python profiles/configs_probe.py 10000000 1,4 2 2 > gpurun_out/r02/ncu_cfg14.log 2>&1
