#!/bin/bash
# Round-2 GPU run 1: parity suite, bench line, batch sweep, ncu --set full captures of every non-GEMM kernel.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/measured.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_gpu.txt 2>&1
echo "=== pytest" ; date
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest.log
tail -5 gpurun_out/r02_pytest.log
echo "=== bench" ; date
timeout 900 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_a.err
echo "=== bsweep" ; date
timeout 600 python tools/bsweep.py > gpurun_out/r02_bsweep.log 2>&1; echo "bsweep rc=$?"
echo "=== ncu" ; date
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 900 $NCU -k regex:'attention_tc|layernorm|qk_norm_rope|dpmpp2m|scale_kernel|latent_feedback|patchify|heun|metrics_sorted' \
  -o gpurun_out/r02_ncu_den -f python tools/prof_all.py den met heun > gpurun_out/r02_ncu_den.log 2>&1; echo "ncu den rc=$?"
timeout 900 $NCU -k regex:'rmsnorm_rows|multiscale|linear_attn|dwconv3|pixel_shuffle|halo_fill|pad_from|in_shortcut' \
  -o gpurun_out/r02_ncu_dec -f python tools/prof_all.py dec > gpurun_out/r02_ncu_dec.log 2>&1; echo "ncu dec rc=$?"
timeout 900 $NCU -k regex:'gemm_tc2' -c 16 \
  -o gpurun_out/r02_ncu_conv -f python tools/prof_all.py dec > gpurun_out/r02_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
for f in den dec conv; do
  ncu -i gpurun_out/r02_ncu_$f.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$f.csv 2>/dev/null
done
ls -la gpurun_out | tail -20
date
