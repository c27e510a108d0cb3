#!/bin/bash
# Round-2 GPU call 17 (8 GPUs): box_16m sharded over 8 ranks: bench lines (driver setting and 200/50) + stage timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=8
{
echo "== bench --gpus $N (20 / 5)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; tail -c 300 gpurun_out/bench_r02_n$N.err; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_n$N.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','parity_check','strong_scaling','halo','stage_ms_slowest_rank','e2e','clocks','sharding'): print(k, json.dumps(d.get(k))[:400])
except Exception as e: print("bench parse failed", e, open('gpurun_out/bench_r02_n$N.json').read()[-800:])
P
echo "== timing tool"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 tools/shard_timing.py --scene box_16m --tag n8 --steps 50 2>&1 | grep "^{"
echo "== bench --gpus $N (200 / 50)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $N --steps 200 --warmup 50 > gpurun_out/bench_r02_n${N}_200.json 2>/dev/null; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_n${N}_200.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','strong_scaling','stage_ms_slowest_rank','clocks'): print(k, json.dumps(d.get(k))[:300])
except Exception as e: print("bench parse failed", e)
P
} > gpurun_out/call17.log 2>&1
tail -30 gpurun_out/call17.log
