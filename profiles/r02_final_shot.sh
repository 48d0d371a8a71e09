#!/bin/bash
# last GPU call of round 2 (3 GPU-minutes left): the CLI's new host pipeline against the real library, then the ingest leg
mkdir -p gpurun_out/r02
( timeout 100 python -m pytest tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -15; echo "rc=${PIPESTATUS[0]}" ) > gpurun_out/r02/final_cli_tests.log
( timeout 60 python profiles/ingest_probe_r02.py 10000000 16 ) > gpurun_out/r02/final_ingest_16.json 2> gpurun_out/r02/final_ingest_16.err
( timeout 120 python -m pytest tests -m gpu -x -q --deselect tests/test_cli_gpu.py 2>&1 | tail -8; echo "rc=${PIPESTATUS[0]}" ) > gpurun_out/r02/final_rest_tests.log
cat gpurun_out/r02/final_cli_tests.log; head -c 3000 gpurun_out/r02/final_ingest_16.json; tail -3 gpurun_out/r02/final_rest_tests.log
