#!/bin/bash
# round 2, GPU run 2: whole GPU test suite with the async push / reader-thread CLI, then a small config-5 bench
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02/gpu_tests_run2.log 2>&1; echo "rc=$?" >> gpurun_out/r02/gpu_tests_run2.log
tail -15 gpurun_out/r02/gpu_tests_run2.log
echo skip bench

