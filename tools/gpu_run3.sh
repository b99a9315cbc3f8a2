#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== GPU suite (kernels, denoiser, parity)"; timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_denoiser_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -4
echo "=== attention standalone v7 / v6"; python tools/bench_attn.py 2>&1 | tail -2; LADCAST_B200_ATTN=6 python tools/bench_attn.py 2>&1 | tail -2
echo "=== trace"; python tools/attn_trace.py | tail -5 | cut -c1-420
echo "=== bench (short)"
timeout 600 python bench.py --no-cpu-baseline --no-strong --no-e2e --no-metrics > gpurun_out/r02c_bench_v7.json 2> gpurun_out/r02c_bench.err; echo rc=$?
python - <<'PY'
import json
for f in ('gpurun_out/r02c_bench_v7.json',):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, round(d['value'],2), round(d['ms_per_step'],2), 'gemm',r['achieved'],r['ms'], {s['class']:(s['ms'],s['achieved']) for s in r['secondary'] if s['class'] in ('attention_tc','layernorm','qk_norm_rope','sphere_conv_tc')})
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r02c_bench.err
