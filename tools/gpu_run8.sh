#!/bin/bash
# attention timeline with the split hand-over; single-pass fused qkv epilogue: parity + in-step A/B
set -u
mkdir -p gpurun_out
echo "=== attention trace"; timeout 120 python tools/attn_trace.py > gpurun_out/attn_trace_split.log 2>&1; sed -n 6,14p gpurun_out/attn_trace_split.log; tail -5 gpurun_out/attn_trace_split.log | cut -c1-400
echo "=== parity with FUSE_QK=1 (single pass)"; LADCAST_B200_FUSE_QK=1 timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
grep -E "denoiser_375M|denoiser_1p6B_T4|denoiser_tiny_golden/bf16" gpurun_out/measured.jsonl | tail -3
echo "=== bench A/B"
for i in 1 2; do
  for fq in 1 0; do
    LADCAST_B200_FUSE_QK=$fq timeout 600 python bench.py --no-cpu-baseline --no-strong --no-e2e --no-metrics > gpurun_out/r02j_bench_fq${fq}_$i.json 2> gpurun_out/r02j_bench.err; echo "fq=$fq rc=$?"
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02j_bench_fq*_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, round(d['value'],2), round(d['ms_per_step'],2), 'gemm',r['achieved'],r['ms'], {s['class']:(s['ms'],s['achieved']) for s in r['secondary'] if s['class'] in ('attention_tc','qk_norm_rope','layernorm')})
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r02j_bench.err
