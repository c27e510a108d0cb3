#!/bin/bash
# Round-2 GPU call 16 (2 GPUs): cost of NVML polling during a short timed region; bench.py --gpus 2 after the OMP-binding fix.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=2
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 tools/shard_timing.py "${@:2}" 2>&1 | grep "^{\|^clocks" | cut -c1-220; }
{
echo "== no sampler, 40 steps"; run 29541 --tag none --steps 40
echo "== both, 10 ms"; run 29542 --tag both10 --steps 40 --sampler
echo "== clock only, 10 ms"; SPH_BENCH_SAMPLER=clock run 29543 --tag clock10 --steps 40 --sampler
echo "== reasons only, 10 ms"; SPH_BENCH_SAMPLER=reasons run 29544 --tag reasons10 --steps 40 --sampler
echo "== both, 2 ms"; SPH_BENCH_SAMPLER_PERIOD_S=0.002 run 29545 --tag both2 --steps 40 --sampler
echo "== bench --gpus $N (20 / 5)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; tail -c 300 gpurun_out/bench_r02_n$N.err; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_n$N.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','parity_check','strong_scaling','stage_ms_slowest_rank','e2e','clocks'): print(k, json.dumps(d.get(k))[:300])
except Exception as e: print("bench parse failed", e, open('gpurun_out/bench_r02_n$N.json').read()[-800:])
P
} > gpurun_out/call16.log 2>&1
tail -30 gpurun_out/call16.log
