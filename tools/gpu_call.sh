#!/bin/bash
# per-call scratch script (GPU box): launch list of DFSPH steps (which kernels carry the 2 ms)
mkdir -p gpurun_out
L=gpurun_out/call24.log
: > $L
export SPH_BENCH_CPU_BUDGET_S=1 SPH_BENCH_SKIP_EXTRA=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/dfsph_launches.csv \
  python bench.py --scene dragon_bath_dfsph --steps 4 --warmup 20 > gpurun_out/dfsph_ncu_bench.log 2>&1
echo "rc=$?" >> $L
python tools/launch_shares.py gpurun_out/dfsph_launches.csv >> $L 2>&1
