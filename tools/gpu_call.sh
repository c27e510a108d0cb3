#!/bin/bash
# per-call scratch script (GPU box): 2-GPU long run of the sharded engine + the final N = 2 bench line
mkdir -p gpurun_out
L=gpurun_out/call22.log
: > $L
echo "== sharded soak, 2 GPUs, 2 M particles, 3000 steps" >> $L
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
  tools/check_slab_parity.py --counts 200 100 100 --steps 3000 --soak 250 >> $L 2>&1
echo "rc=$?" >> $L
echo "== bench N=2" >> $L
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_final.json 2>> $L
echo "rc=$?" >> $L
tail -c 1500 gpurun_out/bench_n2_final.json >> $L
