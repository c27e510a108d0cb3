#!/bin/bash
# per-call scratch script (GPU box): gpu tests + DFSPH step driven from one library call vs op by op with host loops
mkdir -p gpurun_out
L=gpurun_out/call25.log
: > $L
echo "== pytest gpu" >> $L
timeout 900 python -m pytest tests -x -q -m gpu >> $L 2>&1
echo "rc=$?" >> $L
export SPH_BENCH_CPU_BUDGET_S=4 SPH_BENCH_SKIP_EXTRA=1
echo "== DFSPH dragon_bath_dfsph, one library call per step" >> $L
timeout 400 python bench.py --scene dragon_bath_dfsph --steps 60 --warmup 20 > gpurun_out/bench_dfsph_fused_step.json 2>> $L
echo "rc=$?" >> $L
echo "== DFSPH dragon_bath_dfsph, op by op with host loops (reference structure)" >> $L
SPH_DFSPH_HOST_LOOPS=1 timeout 400 python bench.py --scene dragon_bath_dfsph --steps 60 --warmup 20 > gpurun_out/bench_dfsph_host_loops.json 2>> $L
echo "rc=$?" >> $L
python - >> $L 2>&1 <<'P'
import json
for f in ("fused_step", "host_loops"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_dfsph_{f}.json") if l.startswith("{")][-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "steady", round(d["steady"]["ms_per_step"], 4), "launches", d["gpu_launches"], "e2e steps/s", round(d["e2e"]["steps_per_s"], 1))
    except Exception as e:
        print(f, "failed", e)
P
