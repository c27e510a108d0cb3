"""Run a few un-graphed steps of a scene so that ncu can attribute kernels (profiling helper)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sph_taichi_b200 import ParticleSystem, SimConfig, scene

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="dragon_bath")
ap.add_argument("--warm", type=int, default=100)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
ps = ParticleSystem(SimConfig(scene.NAMED_SCENES[a.scene]()))
s = ps.build_solver(); s.initialize()
s.step(a.warm)
torch.cuda.synchronize()
for _ in range(a.steps):
    print(ps._engine.profile_step())
torch.cuda.synchronize()
