#!/bin/bash
# Round-2 GPU call 18 (4 GPUs): split density pass (exchange overlaps interior density + force): parity, timing, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=4
{
echo "== parity $N ranks (launches)"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29534 tools/check_slab_parity.py --counts 160 48 48 --steps 120 --rebalance-every 4 2>&1 | grep "^{" | cut -c1-330
echo "== parity $N ranks (graph)"; SPH_SHARD_GRAPH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29533 tools/check_slab_parity.py --counts 160 48 48 --steps 60 --rebalance-every 4 2>&1 | grep "^{" | cut -c1-200
echo "== timing"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 tools/shard_timing.py --tag n$N --steps 50 2>&1 | grep "^{"
echo "== bench --gpus $N (20 / 5)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; tail -c 300 gpurun_out/bench_r02_n$N.err; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_n$N.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','parity_check','strong_scaling','halo','stage_ms_slowest_rank','e2e','clocks'): print(k, json.dumps(d.get(k))[:300])
except Exception as e: print("bench parse failed", e, open('gpurun_out/bench_r02_n$N.json').read()[-800:])
P
} > gpurun_out/call18.log 2>&1
tail -30 gpurun_out/call18.log
