#!/bin/bash
# per-call scratch script (GPU box): gpu tests with packed and unpacked DFSPH sweeps + timing of both
mkdir -p gpurun_out
L=gpurun_out/call26.log
: > $L
echo "== pytest gpu (default: packed sweeps)" >> $L
timeout 900 python -m pytest tests -x -q -m gpu >> $L 2>&1
echo "rc=$?" >> $L
echo "== pytest gpu, DFSPH tests with SPH_DFSPH_PACKED=0" >> $L
SPH_DFSPH_PACKED=0 timeout 600 python -m pytest tests/test_gpu_dfsph.py tests/test_gpu_reference_golden.py -x -q -m gpu >> $L 2>&1
echo "rc=$?" >> $L
export SPH_BENCH_CPU_BUDGET_S=2 SPH_BENCH_SKIP_EXTRA=1
echo "== DFSPH dragon_bath_dfsph, packed sweeps" >> $L
timeout 400 python bench.py --scene dragon_bath_dfsph --steps 60 --warmup 20 > gpurun_out/bench_dfsph_packed.json 2>> $L
echo "rc=$?" >> $L
echo "== DFSPH dragon_bath_dfsph, unpacked sweeps" >> $L
SPH_DFSPH_PACKED=0 timeout 400 python bench.py --scene dragon_bath_dfsph --steps 60 --warmup 20 > gpurun_out/bench_dfsph_unpacked.json 2>> $L
echo "rc=$?" >> $L
python - >> $L 2>&1 <<'P'
import json
for f in ("packed", "unpacked"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_dfsph_{f}.json") if l.startswith("{")][-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "steady", round(d["steady"]["ms_per_step"], 4), "launches", d["gpu_launches"], "e2e steps/s", round(d["e2e"]["steps_per_s"], 1))
    except Exception as e:
        print(f, "failed", e)
P
