#!/bin/bash
# 2-GPU check of the only exchange of the path (member-sharded metrics, NCCL re-shard vs fused peer-memory kernel) and
# the driver's multi-GPU bench command (weak headline + strong 1.6B legs with the metrics exchange timed and checked).
set -u
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/topo_${N}gpu.txt
echo "=== pointer-form kernel on one GPU"; timeout 300 python -m pytest tests/test_metrics_gpu.py -m gpu -q -k pointer 2>&1 | tail -2
echo "=== dist metrics check ($N GPUs)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tools/dist_metrics_check.py > gpurun_out/dist_metrics_${N}gpu.log 2>&1; echo "dist check rc=$?"
grep -v "Warning\|warn" gpurun_out/dist_metrics_${N}gpu.log | tail -16
echo "=== bench --gpus $N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/bench_${N}gpu.json') if l.startswith('{')][-1]
    print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])
    for s in d['strong']: print(json.dumps({k:s[k] for k in ('ensemble_total','members_per_rank','value','ms_per_step','metrics','limiter')})[:1500])
except Exception as e: print('ERR',e)
PY
