#!/bin/bash
# A/B: staged (coalesced) bf16 GEMM epilogue; attention speculative-max variant; full parity suite on the new build.
set -u
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "=== attention SPEC A/B"
for spec in 0 1 0 1; do
  echo "spec=$spec"; LADCAST_B200_ATTN_SPEC=$spec timeout 200 python tools/bench_attn.py 2>&1 | grep "attn B" 
done
LADCAST_B200_ATTN_SPEC=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "test_attention" 2>&1 | tail -1
echo "=== gemm trace (staged / direct)"
for st in 1 0; do
  LADCAST_B200_EPI_STAGE=$st timeout 120 python tools/gemm_trace.py 36000 6144 1536 1 > gpurun_out/trace_mlp_up_st$st.log 2>&1; echo "stage=$st"; sed -n 14,22p gpurun_out/trace_mlp_up_st$st.log
done
echo "=== bench A/B"
for i in 1 2; do
  for st in 1 0; do
    LADCAST_B200_EPI_STAGE=$st timeout 600 python bench.py --no-cpu-baseline --no-strong --no-e2e --no-metrics > gpurun_out/r02h_bench_st${st}_$i.json 2> gpurun_out/r02h_bench.err; echo "st=$st rc=$?"
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02h_bench_st*_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, round(d['value'],2), round(d['ms_per_step'],2), 'gemm',r['achieved'],r['ms'], {s['class']:(s['ms'],s['achieved']) for s in r['secondary'] if s['class'] in ('attention_tc','sphere_conv_tc')})
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r02h_bench.err
