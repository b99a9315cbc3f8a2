#!/bin/bash
# A/B: GEMM epilogue L2 prefetch of residual tiles; GEMM CTA-pair timeline for the K=1536 shapes; new parity tests.
set -u
mkdir -p gpurun_out
echo "=== new tests"; timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "=== gemm trace"
timeout 120 python tools/gemm_trace.py 36000 6144 1536 1 > gpurun_out/trace_mlp_up.log 2>&1; sed -n 1,3p gpurun_out/trace_mlp_up.log; sed -n 12,24p gpurun_out/trace_mlp_up.log
timeout 120 python tools/gemm_trace.py 36000 4608 1536 0 > gpurun_out/trace_qkv.log 2>&1; sed -n 12,20p gpurun_out/trace_qkv.log
echo "=== bench A/B"
for i in 1 2; do
  for pf in 1 0; do
    LADCAST_B200_EPI_PREFETCH=$pf timeout 600 python bench.py --no-cpu-baseline --no-strong --no-e2e --no-metrics > gpurun_out/r02g_bench_pf${pf}_$i.json 2> gpurun_out/r02g_bench.err; echo "pf=$pf rc=$?"
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02g_bench_pf*_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, round(d['value'],2), round(d['ms_per_step'],2), 'gemm',r['achieved'],r['ms'], {s['class']:(s['ms'],s['achieved']) for s in r['secondary'] if s['class'] in ('attention_tc','layernorm','qk_norm_rope','sphere_conv_tc')})
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r02g_bench.err
