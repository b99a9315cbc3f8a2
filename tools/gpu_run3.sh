#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== full GPU suite"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "=== attention standalone v7 / v6"; python tools/bench_attn.py 2>&1 | tail -2; LADCAST_B200_ATTN=6 python tools/bench_attn.py 2>&1 | tail -2
echo "=== bench (short) v7 / v6"
timeout 600 python bench.py --no-cpu-baseline --no-strong --no-e2e --no-metrics > gpurun_out/r02c_bench_v7.json 2> gpurun_out/r02c_bench.err; echo rc=$?
LADCAST_B200_ATTN=6 timeout 600 python bench.py --no-cpu-baseline --no-strong --no-e2e --no-metrics > gpurun_out/r02c_bench_v6.json 2>> gpurun_out/r02c_bench.err; echo rc=$?
python - <<'PY'
import json
for f in ('gpurun_out/r02c_bench_v7.json','gpurun_out/r02c_bench_v6.json'):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, round(d['value'],2), round(d['ms_per_step'],2), 'gemm',r['achieved'],r['ms'], {s['class']:(s['ms'],s['achieved']) for s in r['secondary'] if s['class'] in ('attention_tc','layernorm','qk_norm_rope','dec_rmsnorm','sphere_conv_tc','dec_multiscale','dec_linear_attn','dec_dwconv_glu')})
    except Exception as e: print(f,'ERR',e)
PY
tail -5 gpurun_out/r02c_bench.err
