#!/bin/bash
# round 2, GPU run 7: the DEFAULT bench (config 5, 100 M records per GPU) + reference arm, wall-clocked
mkdir -p gpurun_out/r02
S=$(date +%s); python bench.py > gpurun_out/r02/bench_default_run7.json 2> gpurun_out/r02/bench_default_run7.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))s"
tail -c 5000 gpurun_out/r02/bench_default_run7.json; tail -3 gpurun_out/r02/bench_default_run7.err
S=$(date +%s); python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/bench_reference_run7.json 2> gpurun_out/r02/bench_reference_run7.err; echo "ref rc=$? wall=$(( $(date +%s) - S ))s"
tail -c 1500 gpurun_out/r02/bench_reference_run7.json; tail -3 gpurun_out/r02/bench_reference_run7.err
