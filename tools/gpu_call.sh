#!/bin/bash
# Round-2 GPU call 3: density v11 (flat prefetched hit loop) parity + timing, column orders, bench.py shake-down.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest dv=10 (v11)"; SPH_DENSITY_VARIANT=10 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== pytest dv=10 order sym"; SPH_COLUMN_ORDER=026841357 SPH_DENSITY_VARIANT=10 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_golden.py tests/test_gpu_dfsph.py -m gpu -x -q 2>&1 | tail -4
echo "== sweep default lib (v11), orders"
timeout 900 python tools/sweep_variants.py --pairs 1:1,10:1 --orders 012345678,026841357,413570268,135702684,408172635,876543210 --scene dragon_bath 2>&1 | grep -v Warning
for lib in v11packed v11s18 v11mb6 v11t64; do
  echo "== sweep $lib"
  SPH_B200_LIB=$PWD/build_exp/libsph_$lib.so timeout 300 python tools/sweep_variants.py --pairs 10:1 --orders 012345678,026841357 --scene dragon_bath 2>&1 | grep -v Warning
done
echo "== developed flow (warm 400)"
timeout 600 python tools/sweep_variants.py --pairs 1:1,10:1 --orders 012345678,026841357 --scene dragon_bath --warm 400 2>&1 | grep -v Warning
SPH_B200_LIB=$PWD/build_exp/libsph_v11packed.so timeout 300 python tools/sweep_variants.py --pairs 10:1 --orders 026841357 --scene dragon_bath --warm 400 2>&1 | grep -v Warning
echo "== ncu v11"
SPH_DENSITY_VARIANT=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_density_soa' -s 100 -c 1 -f -o gpurun_out/prof_r02_v11 python tools/profile_step.py --warm 100 --steps 2 2>&1 | tail -2
echo "== bench.py"
SPH_BENCH_CPU_BUDGET_S=4 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; tail -c 600 gpurun_out/bench_r02_a.err
} > gpurun_out/call03.log 2>&1
tail -60 gpurun_out/call03.log
