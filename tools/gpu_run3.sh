#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== full GPU suite"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "=== bench (short)"
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-strong --no-e2e --no-metrics > gpurun_out/r02d_bench_$i.json 2> gpurun_out/r02d_bench.err; echo rc=$?
done
python - <<'PY'
import json
for f in ('gpurun_out/r02d_bench_1.json','gpurun_out/r02d_bench_2.json'):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, round(d['value'],2), round(d['ms_per_step'],2), 'gemm',r['achieved'],r['ms'], {s['class']:(s['ms'],s['achieved']) for s in r['secondary'] if s['class'] in ('attention_tc','layernorm','qk_norm_rope','sphere_conv_tc')})
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r02d_bench.err
echo "=== gemm shapes"; python tools/bench_gemm_shapes.py 2>&1 | tail -12
