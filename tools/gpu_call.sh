#!/bin/bash
# Round-2 GPU call 2: new parity tests, density v10 parity + timing variants, pipe probes, ncu captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest default"; timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -15
echo "== pytest dv=10"; SPH_DENSITY_VARIANT=10 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke dv=10"; SPH_DENSITY_VARIANT=10 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pipes"; build_exp/pipes
echo "== sweep default lib"
timeout 600 python tools/sweep_variants.py --pairs 1:1,10:1 --scene dragon_bath 2>&1 | grep -v Warning
for lib in v10packed v10sym v10mb9 v10mb6 v10t256; do
  echo "== sweep $lib"
  SPH_B200_LIB=$PWD/build_exp/libsph_$lib.so timeout 300 python tools/sweep_variants.py --pairs 10:1 --scene dragon_bath 2>&1 | grep -v Warning
done
echo "== sweep armadillo"
timeout 600 python tools/sweep_variants.py --pairs 1:1,10:1 --scene armadillo_bath_dynamic --warm 50 2>&1 | grep -v Warning
echo "== ncu v10"
SPH_DENSITY_VARIANT=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_density_soa|k_force_packed' -s 200 -c 2 -f -o gpurun_out/prof_r02_v10 python tools/profile_step.py --warm 100 --steps 2 2>&1 | tail -3
echo "== ncu v10sym"
SPH_DENSITY_VARIANT=10 SPH_B200_LIB=$PWD/build_exp/libsph_v10sym.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_density_soa|k_force_packed' -s 200 -c 2 -f -o gpurun_out/prof_r02_v10sym python tools/profile_step.py --warm 100 --steps 2 2>&1 | tail -3
} > gpurun_out/call02.log 2>&1
tail -70 gpurun_out/call02.log
