#!/bin/bash
# Round-2 GPU run: parity suite first (stop there when it is red), then bench line, batch sweep and ncu --set full
# captures of every non-GEMM kernel.  .ncu-rep files stay in /tmp on the box (gpurun_out is capped at 64 MiB): only
# their raw-page CSVs and logs travel back.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/measured.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_gpu.txt 2>&1
echo "=== smoke"; date
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
echo "=== pytest" ; date
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -25 gpurun_out/r02_pytest.log
cat gpurun_out/measured.jsonl 2>/dev/null
echo "=== bench" ; date
timeout 900 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_a.err; head -c 6000 gpurun_out/r02_bench_a.json
if [ $rc -ne 0 ] && [ "${FORCE:-0}" != "1" ]; then
  echo "parity suite red: decoder tests again without the fused norm epilogue, then stop"
  LADCAST_B200_FUSE_NORM=0 timeout 600 python -m pytest tests/test_dcae_gpu.py -m gpu -q 2>&1 | tail -8
  exit 0
fi
echo "=== bsweep" ; date
timeout 600 python tools/bsweep.py > gpurun_out/r02_bsweep.log 2>&1; echo "bsweep rc=$?"; tail -3 gpurun_out/r02_bsweep.log
echo "=== ncu" ; date
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 900 $NCU -k regex:'attention_tc|layernorm|qk_norm_rope|dpmpp2m|scale_kernel|latent_feedback|patchify|heun|metrics_sorted' \
  -o /tmp/r02_ncu_den -f python tools/prof_all.py den met heun > gpurun_out/r02_ncu_den.log 2>&1; echo "ncu den rc=$?"
timeout 900 $NCU -k regex:'rmsnorm_rows|multiscale|linear_attn|dwconv3|pixel_shuffle|halo_fill|pad_from|in_shortcut' \
  -o /tmp/r02_ncu_dec -f python tools/prof_all.py dec > gpurun_out/r02_ncu_dec.log 2>&1; echo "ncu dec rc=$?"
timeout 900 $NCU -k regex:'gemm_tc2' -c 16 \
  -o /tmp/r02_ncu_conv -f python tools/prof_all.py dec > gpurun_out/r02_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
for f in den dec conv; do
  ncu -i /tmp/r02_ncu_$f.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$f.csv 2>/dev/null
done
du -sh gpurun_out; ls -la gpurun_out | tail -20
date
