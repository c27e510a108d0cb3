#!/bin/bash
# Round-2 GPU call 21 (1 GPU box, CPU work): OpenMP binding of the CPU reference arm: close vs spread vs unbound.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nproc; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA" | head -8
for bind in close spread false; do
  echo "== OMP_PROC_BIND=$bind"
  OMP_PROC_BIND=$bind OMP_PLACES=cores timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['steps_per_s'], d['cpu_baseline']['cores'], d['cpu_baseline']['threads_tried'])"
done
echo "== unset"; env -u OMP_PROC_BIND -u OMP_PLACES python - <<'P'
import os, time, sys
sys.path.insert(0, '.')
os.environ.pop('OMP_PROC_BIND', None); os.environ.pop('OMP_PLACES', None)
from oracle.sph_oracle import OracleSim, set_threads
from sph_taichi_b200 import scene
o = OracleSim(scene.dragon_bath()); o.initialize()
for n in (128, 64, 32, 16):
    set_threads(n); o.step(); t0 = time.perf_counter()
    for _ in range(3): o.step()
    print(n, 3 / (time.perf_counter() - t0))
P
} > gpurun_out/call21.log 2>&1
tail -30 gpurun_out/call21.log
