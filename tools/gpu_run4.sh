#!/bin/bash
# Round-2 evidence run of the final build: parity suite, full bench line, ncu launch list of the same bench command,
# `ncu --set full` captures of every kernel class (raw-page CSVs travel back; .ncu-rep stay on the box), batch sweep.
set -u
mkdir -p gpurun_out
T=${TAG:-r02f}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_gpu.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  echo "=== pytest"; date
  timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
  tail -12 gpurun_out/${T}_pytest.log
  cat gpurun_out/measured.jsonl 2>/dev/null | tail -40
fi
echo "=== bench"; date
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 800 gpurun_out/${T}_bench.err; head -c 9000 gpurun_out/${T}_bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
  echo "=== ncu launch list"; date
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-strong --no-cpu-baseline --no-e2e --no-metrics --no-roofline \
    > gpurun_out/${T}_launches_bench.log 2>&1; echo "launch list rc=$?"
  gzip -f gpurun_out/${T}_launches.csv
  echo "=== ncu full"; date
  NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
  timeout 900 $NCU -k regex:'attention_tc|layernorm|qk_norm_rope|dpmpp2m|scale_kernel|latent_feedback|patchify|heun|metrics_sorted' \
    -o /tmp/${T}_den -f python tools/prof_all.py den met heun attn16 > gpurun_out/${T}_ncu_den.log 2>&1; echo "ncu den rc=$?"
  timeout 900 $NCU -k regex:'rmsnorm_rows|multiscale|linear_attn|dwconv3|pixel_shuffle|halo_fill|pad_from|in_shortcut' \
    -o /tmp/${T}_dec -f python tools/prof_all.py dec > gpurun_out/${T}_ncu_dec.log 2>&1; echo "ncu dec rc=$?"
  timeout 900 $NCU -k regex:'gemm_tc' -c 24 \
    -o /tmp/${T}_conv -f python tools/prof_all.py dec > gpurun_out/${T}_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
  timeout 900 $NCU -k regex:'gemm_tc' -c 40 \
    -o /tmp/${T}_gemm -f python tools/prof_all.py den > gpurun_out/${T}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
  for f in den dec conv gemm; do
    ncu -i /tmp/${T}_$f.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${T}_ncu_$f.csv.gz
  done
  cp gpurun_out/prof_all_classes.json gpurun_out/${T}_prof_all_classes.json 2>/dev/null
fi
if [ "${SKIP_SWEEP:-0}" != "1" ]; then
  echo "=== bsweep"; date
  timeout 900 python tools/bsweep.py > gpurun_out/${T}_bsweep.log 2>&1; echo "bsweep rc=$?"; cut -c1-400 gpurun_out/${T}_bsweep.log | tail -12
fi
du -sh gpurun_out; date
