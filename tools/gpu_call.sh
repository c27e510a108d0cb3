#!/bin/bash
# Round-2 GPU call 15 (4 GPUs): bench.py --gpus 4 with eager sharded steps and the early-initialised NVML sampler.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=4
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 tools/shard_timing.py "${@:2}" 2>&1 | grep "^{\|^parity\|^clocks"; }
{
echo "== eager"; run 29541 --tag eager --steps 50
echo "== eager + sampler 20 steps"; run 29542 --tag eager-sampler --steps 20 --sampler
echo "== graph"; SPH_SHARD_GRAPH=1 run 29543 --tag graph --steps 50
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; tail -c 400 gpurun_out/bench_r02_n$N.err; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_n$N.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','parity_check','strong_scaling','halo','stage_ms_slowest_rank','e2e','clocks'): print(k, json.dumps(d.get(k))[:400])
except Exception as e: print("bench parse failed", e, open('gpurun_out/bench_r02_n$N.json').read()[-800:])
P
echo "== bench --gpus $N --steps 200 --warmup 50"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $N --steps 200 --warmup 50 > gpurun_out/bench_r02_n${N}_200.json 2>/dev/null; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_n${N}_200.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','strong_scaling','clocks'): print(k, json.dumps(d.get(k))[:300])
except Exception as e: print("bench parse failed", e)
P
} > gpurun_out/call15.log 2>&1
tail -30 gpurun_out/call15.log
