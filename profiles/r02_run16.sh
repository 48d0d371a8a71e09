#!/bin/bash
# round 2, GPU run 16 (8-GPU node): is the 0.42 / 0.69 ms bimodality of decode_kernel a property of the GPU, or of 8 ranks running at once?
mkdir -p gpurun_out/r02
echo "== one GPU at a time"
for d in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$d python profiles/configs_probe.py 5000000 5 5 20 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        x=json.loads(l); print('seq gpu $d decode_ms', round(x['decode_ms'],4), 'step', round(x['ms_per_step'],3))"
done
echo "== all eight at once"
for d in 0 1 2 3 4 5 6 7; do
  ( CUDA_VISIBLE_DEVICES=$d python profiles/configs_probe.py 5000000 5 5 200 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        x=json.loads(l); print('par gpu $d decode_ms', round(x['decode_ms'],4), 'step', round(x['ms_per_step'],3))" ) &
done
wait
