#!/bin/bash
# Round-2 GPU call 1: parity of the default build, parity of the parked variants, timing sweep.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest default"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for pair in 3:1 1:2 4:2 3:2; do
  dv=${pair%%:*}; fv=${pair##*:}
  echo "== pytest variants dv=$dv fv=$fv"
  SPH_DENSITY_VARIANT=$dv SPH_FORCE_VARIANT=$fv timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_golden.py -m gpu -x -q 2>&1 | tail -4
done
echo "== sweep default lib"
timeout 600 python tools/sweep_variants.py --pairs 1:1,3:1,1:2,3:2,4:1,4:2 --scene dragon_bath 2>&1 | grep -v Warning
for lib in f7 f6; do
  echo "== sweep $lib"
  SPH_B200_LIB=$PWD/build_exp/libsph_$lib.so timeout 300 python tools/sweep_variants.py --pairs 1:1,1:2,4:2 --scene dragon_bath 2>&1 | grep -v Warning
done
for lib in a448 a320b6 a256b8; do
  echo "== sweep $lib"
  SPH_B200_LIB=$PWD/build_exp/libsph_$lib.so timeout 300 python tools/sweep_variants.py --pairs 4:1,4:2 --scene dragon_bath 2>&1 | grep -v Warning
done
} > gpurun_out/call01.log 2>&1
tail -60 gpurun_out/call01.log
