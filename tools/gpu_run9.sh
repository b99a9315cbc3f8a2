#!/bin/bash
# Final build of round 2: full parity suite, smoke, full bench line, GEMM per-launch metrics (light ncu pass), 1.6B small-batch sweep.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/measured.jsonl
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
cp gpurun_out/measured.jsonl gpurun_out/r02k_measured.jsonl 2>/dev/null
echo "=== bench"; timeout 1200 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k_bench.json')); r=d['roofline']
print(round(d['value'],2), round(d['e2e']['value'],2), round(d['ms_per_step'],2), 'gemm',r['achieved'],r['frac'],r['ms'], {s['class']:(s['ms'],s['achieved'],s['frac']) for s in r['secondary']})
for s in d['strong']: print(s['ensemble_total'], round(s['value'],2), s['ms_per_step'])
print(d['metrics'])
print(d['cpu_baseline']['value'], d['cpu_baseline'].get('config1'))
PY
echo "=== ncu gemm (light)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
  --clock-control none --profile-from-start off -k regex:gemm_tc -c 40 --csv --page raw --log-file gpurun_out/r02k_ncu_gemm_light.csv python tools/prof_all.py den > gpurun_out/r02k_ncu_gemm_light.log 2>&1; echo "ncu rc=$?"
echo "=== bsweep 1.6B small batches"; timeout 600 python tools/bsweep.py 1.6B 2,3 > gpurun_out/r02k_bsweep.log 2>&1; cut -c1-330 gpurun_out/r02k_bsweep.log | tail -2
echo "=== bsweep 1.6B B=3 with PDL"; LADCAST_B200_PDL=1 timeout 300 python tools/bsweep.py 1.6B 3 > gpurun_out/r02k_bsweep_pdl.log 2>&1; cut -c1-330 gpurun_out/r02k_bsweep_pdl.log | tail -1
