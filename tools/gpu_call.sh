#!/bin/bash
# Round-2 GPU call 19 (8 GPUs): box_16m, split density pass on (default at 8 ranks) vs off; bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=8
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 tools/shard_timing.py --scene box_16m --steps 40 --warm 25 "${@:2}" 2>&1 | grep "^{"; }
{
echo "== split density ON"; SPH_SHARD_SPLIT_DENSITY=1 run 29541 --tag split
echo "== split density OFF"; SPH_SHARD_SPLIT_DENSITY=0 run 29542 --tag nosplit
echo "== bench --gpus $N (20 / 5), default"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; tail -c 300 gpurun_out/bench_r02_n$N.err; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_n$N.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','parity_check','strong_scaling','halo','stage_ms_slowest_rank','e2e','clocks'): print(k, json.dumps(d.get(k))[:300])
except Exception as e: print("bench parse failed", e, open('gpurun_out/bench_r02_n$N.json').read()[-800:])
P
} > gpurun_out/call19.log 2>&1
tail -30 gpurun_out/call19.log
