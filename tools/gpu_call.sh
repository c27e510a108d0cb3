#!/bin/bash
# Round-2 GPU call 5 (2 GPUs): the sharded engine on real NCCL.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "== slab parity, 2 ranks, eager (no graph)"
SPH_SHARD_NO_GRAPH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 tools/check_slab_parity.py --counts 64 24 24 --steps 60 2>&1 | grep -v "^\*\|OMP_NUM" | tail -6
echo "== slab parity, 2 ranks, graph"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29534 tools/check_slab_parity.py --counts 128 48 48 --steps 120 --rebalance-every 4 2>&1 | grep -v "^\*\|OMP_NUM" | tail -6
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench --gpus 2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; tail -c 1500 gpurun_out/bench_r02_n2.err; python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_n2.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','parity_check','strong_scaling','halo','stage_ms_slowest_rank','e2e','sharding'): print(k, json.dumps(d.get(k))[:400])
except Exception as e: print("bench parse failed", e, open('gpurun_out/bench_r02_n2.json').read()[-800:])
P
} > gpurun_out/call05.log 2>&1
tail -50 gpurun_out/call05.log
