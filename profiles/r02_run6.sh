#!/bin/bash
# round 2, GPU run 6: full test suite, config 1/4 probes after the gather / coverage changes, the DEFAULT bench (config 5, 100 M records)
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02/gpu_tests_run6.log 2>&1; echo "rc=$?" >> gpurun_out/r02/gpu_tests_run6.log
tail -4 gpurun_out/r02/gpu_tests_run6.log
python profiles/configs_probe.py 20000000 1,4 10 20 > gpurun_out/r02/configs_probe_run6.txt 2>&1; cat gpurun_out/r02/configs_probe_run6.txt | cut -c1-330
/usr/bin/time -v python bench.py > gpurun_out/r02/bench_default_run6.json 2> gpurun_out/r02/bench_default_run6.err; echo "bench rc=$?"
grep -E "Elapsed|Maximum resident" gpurun_out/r02/bench_default_run6.err; tail -c 4500 gpurun_out/r02/bench_default_run6.json
nvidia-smi --query-gpu=memory.total,memory.used --format=csv; free -g | head -2; nproc
