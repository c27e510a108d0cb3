#!/bin/bash
# Round-2 GPU call 11 (2 GPUs): persistent (contiguous tile runs) vs one-block-per-capacity-tile pair kernels in sharded steps.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port $1 tools/shard_timing.py "${@:2}" 2>&1 | grep "^{"; }
{
echo "== persistent (contiguous runs), cap 1.35"; run 29541 --tag pers
echo "== persistent, cap 1.15"; run 29542 --tag pers-cap1.15 --capacity 1.15
echo "== one block per tile, cap 1.35"; SPH_SHARD_PERSISTENT=0 run 29543 --tag plain
echo "== one block per tile, cap 1.15"; SPH_SHARD_PERSISTENT=0 run 29544 --tag plain-cap1.15 --capacity 1.15
echo "== parity (persistent)"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29534 tools/check_slab_parity.py --counts 128 48 48 --steps 120 --rebalance-every 4 2>&1 | grep "^{" | cut -c1-230
} > gpurun_out/call11.log 2>&1
tail -30 gpurun_out/call11.log
