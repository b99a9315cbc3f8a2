#!/bin/bash
# ncu launch list of the final build (same bench command as the headline, 1 timed + 3 warm-up AR steps).
set -u
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02m_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-strong --no-cpu-baseline --no-e2e --no-metrics --no-roofline \
  > gpurun_out/r02m_launches_bench.log 2>&1; echo "launch list rc=$?"
gzip -f gpurun_out/r02m_launches.csv; ls -la gpurun_out/r02m_launches.csv.gz
