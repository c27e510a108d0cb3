#!/bin/bash
# Round-2 GPU call 9 (1 GPU): programmatic dependent launch A/B; full suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest gpu (PDL on)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== sweep PDL on"; timeout 600 python tools/sweep_variants.py --pairs 1:1 --scene dragon_bath 2>&1 | grep -v Warning
echo "== sweep PDL off"; SPH_PDL=0 timeout 600 python tools/sweep_variants.py --pairs 1:1 --scene dragon_bath 2>&1 | grep -v Warning
echo "== sweep PDL on (again)"; timeout 600 python tools/sweep_variants.py --pairs 1:1 --scene dragon_bath 2>&1 | grep -v Warning
echo "== armadillo PDL on/off"; timeout 600 python tools/sweep_variants.py --pairs 1:1 --scene armadillo_bath_dynamic --warm 50 2>&1 | grep -v Warning
SPH_PDL=0 timeout 600 python tools/sweep_variants.py --pairs 1:1 --scene armadillo_bath_dynamic --warm 50 2>&1 | grep -v Warning
} > gpurun_out/call09.log 2>&1
tail -30 gpurun_out/call09.log
