#!/bin/bash
# A/B: qk-norm + RoPE fused into the (now staged) qkv GEMM epilogue vs the stand-alone kernel; parity of the fused path.
set -u
mkdir -p gpurun_out
echo "=== parity with FUSE_QK=1"; LADCAST_B200_FUSE_QK=1 timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
grep -E "denoiser_375M|denoiser_1p6B_T4|rollout_metrics/375M" gpurun_out/measured.jsonl | tail -6
echo "=== bench A/B"
for i in 1 2; do
  for fq in 1 0; do
    LADCAST_B200_FUSE_QK=$fq timeout 600 python bench.py --no-cpu-baseline --no-strong --no-e2e --no-metrics > gpurun_out/r02i_bench_fq${fq}_$i.json 2> gpurun_out/r02i_bench.err; echo "fq=$fq rc=$?"
  done
done
for fq in 0 1; do
  LADCAST_B200_PDL=1 LADCAST_B200_FUSE_QK=$fq timeout 600 python bench.py --no-cpu-baseline --no-strong --no-e2e --no-metrics > gpurun_out/r02i_bench_fq${fq}_pdl.json 2> gpurun_out/r02i_bench.err; echo "pdl fq=$fq rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02i_bench_fq*_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, round(d['value'],2), round(d['ms_per_step'],2), 'gemm',r['achieved'],r['ms'], {s['class']:(s['ms'],s['achieved']) for s in r['secondary'] if s['class'] in ('attention_tc','qk_norm_rope','layernorm')})
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r02i_bench.err
