#!/bin/bash
# Round-2 GPU call 7 (1 GPU): sort-chain changes (8192-cell scan tiles, warp-aggregated histogram), full test suite, bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== sweep"; timeout 600 python tools/sweep_variants.py --pairs 1:1 --scene dragon_bath 2>&1 | grep -v Warning
timeout 600 python tools/sweep_variants.py --pairs 1:1 --scene dragon_bath --warm 400 2>&1 | grep -v Warning
echo "== stage profile"; timeout 300 python tools/profile_step.py --warm 100 --steps 3 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; SPH_BENCH_CPU_BUDGET_S=6 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; tail -c 400 gpurun_out/bench_r02_b.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r02_b.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','steady','e2e','stage_ms'): print(k, json.dumps(d.get(k))[:300])
print('extra', json.dumps(d['extra_configs'][0].get('ms_per_step')), json.dumps(d['extra_configs'][0].get('stage_ms')))
P
} > gpurun_out/call07.log 2>&1
tail -40 gpurun_out/call07.log
