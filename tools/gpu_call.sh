#!/bin/bash
# Round-2 GPU call 14 (2 GPUs): graph replay vs eager launches of the sharded step; PDL; in-process NVML sampler.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=2
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 tools/shard_timing.py "${@:2}" 2>&1 | grep "^{\|^parity\|^clocks"; }
{
echo "== graph"; run 29541 --tag graph --steps 50
echo "== eager"; SPH_SHARD_NO_GRAPH=1 run 29542 --tag eager --steps 50
echo "== graph, no PDL"; SPH_PDL=0 run 29543 --tag graph-nopdl --steps 50
echo "== eager, no PDL"; SPH_PDL=0 SPH_SHARD_NO_GRAPH=1 run 29544 --tag eager-nopdl --steps 50
echo "== graph + NVML sampler, 20 steps"; run 29545 --tag graph-sampler --steps 20 --sampler
echo "== eager + NVML sampler, 20 steps"; SPH_SHARD_NO_GRAPH=1 run 29546 --tag eager-sampler --steps 20 --sampler
} > gpurun_out/call14.log 2>&1
tail -30 gpurun_out/call14.log
