#!/bin/bash
# Final confirmation + full bench line of the last build of round 2 (MMA linear attention / multiscale in the decoder).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/measured.jsonl
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
cp gpurun_out/measured.jsonl gpurun_out/r02q_measured.jsonl 2>/dev/null
echo "=== bench"; timeout 1200 python bench.py > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02q_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02q_bench.json')); r=d['roofline']
print(round(d['value'],2), round(d['e2e']['value'],2), round(d['ms_per_step'],2), 'gemm',r['achieved'],r['frac'],r['ms'], {s['class']:(s['ms'],s['frac']) for s in r['secondary']})
for s in d['strong']: print(s['ensemble_total'], round(s['value'],2), s['ms_per_step'])
PY
