"""Long-run robustness: thousands of steps of a scene; status word, bounds, neighbour-list statistics."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sph_taichi_b200 import ParticleSystem, SimConfig, scene

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="dragon_bath")
ap.add_argument("--steps", type=int, default=5000)
ap.add_argument("--every", type=int, default=500)
a = ap.parse_args()
ps = ParticleSystem(SimConfig(scene.NAMED_SCENES[a.scene]()))
s = ps.build_solver(); s.initialize()
done = 0
while done < a.steps:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.step(a.every) if s._fused_step_ok() else [s.step() for _ in range(a.every)]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    done += a.every
    st = ps._engine.read_status()
    ns = ps._engine.neighbor_stats()
    fl = ps.material.to_numpy() == 1
    x, v, rho = ps.x.to_numpy()[fl], ps.v.to_numpy()[fl], ps.density.to_numpy()[fl]
    print(f"step {done:6d} {a.every / dt:8.1f} steps/s status={st} nbr max={ns['max']} mean={ns['mean']:.1f} "
          f"overflow={ns['overflow']} rho max={rho.max():.1f} |v| max={np.abs(v).max():.2f} "
          f"finite={bool(np.isfinite(x).all())} y range=[{x[:,1].min():.3f},{x[:,1].max():.3f}]", flush=True)
