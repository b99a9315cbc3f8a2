#!/bin/bash
# A/B: LayerNorm CTA shape for d = 2048 (three 4-warp CTAs per SM vs one 8-warp CTA), stand-alone and in a 1.6B step.
set -u
mkdir -p gpurun_out
for w in 4 8 4 8; do echo "LN_WPB=$w"; LADCAST_B200_LN_WPB=$w timeout 200 python tools/bench_ln.py 2>&1 | grep layernorm | tail -2; done
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "layernorm" 2>&1 | tail -1
for w in 4 8; do echo "bsweep LN_WPB=$w"; LADCAST_B200_LN_WPB=$w timeout 400 python tools/bsweep.py 1.6B 3,13 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['members'], round(d['ms_per_step'],1), 'ln', d['class_ms']['layernorm'], 'gemm', d['tflops']['gemm_tc'])
"; done
