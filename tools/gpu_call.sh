#!/bin/bash
# Round-2 GPU call 6 (2 GPUs): where the sharded step spends its time; NCCL channel settings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port $1 tools/shard_timing.py "${@:2}" 2>&1 | grep "^{"; }
{
echo "== default"; run 29541 --tag default
echo "== capacity 1.1"; run 29542 --tag cap1.1 --capacity 1.1
echo "== p2p nchannels 16..32"; NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32 run 29543 --tag ch16-32
echo "== p2p nchannels 32"; NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 run 29544 --tag ch32
echo "== p2p nchannels 4"; NCCL_MIN_P2P_NCHANNELS=4 NCCL_MAX_P2P_NCHANNELS=4 run 29545 --tag ch4
echo "== NCCL_P2P_LEVEL NVL + chunk"; NCCL_P2P_NET_CHUNKSIZE=524288 NCCL_BUFFSIZE=16777216 run 29546 --tag buff16m
echo "== pytest new tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "emitter or inverted or slab" 2>&1 | tail -3
} > gpurun_out/call06.log 2>&1
tail -30 gpurun_out/call06.log
