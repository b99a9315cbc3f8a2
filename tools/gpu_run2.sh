#!/bin/bash
# A/B run: parity suite (default build), the same with programmatic dependent launch, short bench lines for both.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/measured.jsonl
echo "=== pytest (default)"; date
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02b_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -8 gpurun_out/r02b_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
echo "=== pytest (PDL=1)"; date
LADCAST_B200_PDL=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_pdl.log 2>&1; echo "pytest pdl rc=$?"; tail -8 gpurun_out/r02b_pytest_pdl.log
echo "=== bench A/B"; date
for i in 1 2; do
  timeout 600 python bench.py --no-cpu-baseline --no-strong --no-e2e --no-metrics > gpurun_out/r02b_bench_pdl0_$i.json 2> gpurun_out/r02b_bench_pdl0.err; echo "bench pdl0 rc=$?"
  LADCAST_B200_PDL=1 timeout 600 python bench.py --no-cpu-baseline --no-strong --no-e2e --no-metrics > gpurun_out/r02b_bench_pdl1_$i.json 2> gpurun_out/r02b_bench_pdl1.err; echo "bench pdl1 rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_bench_pdl*_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, round(d['value'],2), round(d['ms_per_step'],2), 'gemm',r['achieved'], {s['class']:(s['ms'],s['achieved']) for s in r['secondary'] if s['class'] in ('attention_tc','layernorm','qk_norm_rope','dec_rmsnorm','sphere_conv_tc')})
    except Exception as e: print(f, 'ERR', e)
PY
echo "=== small batch A/B (1.6B, B=3)"; date
timeout 600 python tools/bsweep.py 1.6B 3 > gpurun_out/r02b_bsweep_pdl0.log 2>&1; tail -1 gpurun_out/r02b_bsweep_pdl0.log | cut -c1-200
LADCAST_B200_PDL=1 timeout 600 python tools/bsweep.py 1.6B 3 > gpurun_out/r02b_bsweep_pdl1.log 2>&1; tail -1 gpurun_out/r02b_bsweep_pdl1.log | cut -c1-200
date
