#!/bin/bash
# A/B of the attention kernel variants (polynomial-exp2 offload period x split P hand-over): correctness first, then
# stand-alone timing of the production shapes, then the in-step class number of the best candidates.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/attn_ab.log; : > $OUT
for split in 0 1; do for poly in 0 8 4 3 2; do
  export LADCAST_B200_ATTN_POLY=$poly LADCAST_B200_ATTN_SPLIT=$split
  echo "=== poly=$poly split=$split" | tee -a $OUT
  case "$poly$split" in 01|40|41|21|30)
    timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "test_attention and bf16" 2>&1 | tail -1 | tee -a $OUT
    grep attention gpurun_out/measured.jsonl 2>/dev/null | tail -7 | tr '\n' ' ' >> $OUT; echo >> $OUT;;
  esac
  timeout 300 python tools/bench_attn.py 2>&1 | grep -v Warning | tail -4 | tee -a $OUT
done; done
