#!/bin/bash
# Final confirmation of the committed build: smoke + full GPU parity suite (incl. the world-size-1 peer-memory test).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/measured.jsonl
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
cp gpurun_out/measured.jsonl gpurun_out/r02l_measured.jsonl 2>/dev/null
echo "=== bench (short, final build)"; timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-metrics --strong-ens 20 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench.json')); r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],2), 'gemm',r['achieved'],r['frac'])
for s in d['strong']: print(s['ensemble_total'], round(s['value'],2), s['ms_per_step'])
PY
