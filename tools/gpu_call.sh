#!/bin/bash
# Round-2 GPU call 4: density v12 (column pipeline + L1 prefetch + SoA hits) and the force prefetch pipeline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest dv=10 (v12 + force prefetch)"; SPH_DENSITY_VARIANT=10 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== sweep default lib (v12, force prefetch mb8)"
timeout 900 python tools/sweep_variants.py --pairs 1:1,10:1 --orders 012345678,413570268 --scene dragon_bath 2>&1 | grep -v Warning
for lib in v12nopf v12fmb7 v12fmb6 v12packed; do
  echo "== sweep $lib"
  SPH_B200_LIB=$PWD/build_exp/libsph_$lib.so timeout 300 python tools/sweep_variants.py --pairs 10:1 --orders 012345678,413570268 --scene dragon_bath 2>&1 | grep -v Warning
done
echo "== developed flow (warm 400)"
timeout 600 python tools/sweep_variants.py --pairs 1:1,10:1 --orders 413570268 --scene dragon_bath --warm 400 2>&1 | grep -v Warning
SPH_B200_LIB=$PWD/build_exp/libsph_v12nopf.so timeout 300 python tools/sweep_variants.py --pairs 10:1 --orders 413570268 --scene dragon_bath --warm 400 2>&1 | grep -v Warning
SPH_B200_LIB=$PWD/build_exp/libsph_v12fmb7.so timeout 300 python tools/sweep_variants.py --pairs 10:1 --orders 413570268 --scene dragon_bath --warm 400 2>&1 | grep -v Warning
echo "== ncu v12"
SPH_COLUMN_ORDER=413570268 SPH_DENSITY_VARIANT=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_density_soa|k_force_packed' -s 200 -c 2 -f -o gpurun_out/prof_r02_v12 python tools/profile_step.py --warm 100 --steps 2 2>&1 | tail -2
} > gpurun_out/call04.log 2>&1
tail -60 gpurun_out/call04.log
