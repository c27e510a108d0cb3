"""Time the pair-kernel variants (SPH_DENSITY_VARIANT / SPH_FORCE_VARIANT) on one scene."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sph_taichi_b200 import ParticleSystem, SimConfig, scene

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="dragon_bath")
ap.add_argument("--dv", type=int, nargs="*", default=[0, 1])
ap.add_argument("--fv", type=int, nargs="*", default=[0, 1])
ap.add_argument("--pairs", type=str, default="", help="explicit dv:fv pairs, comma separated")
ap.add_argument("--warm", type=int, default=100)
ap.add_argument("--orders", type=str, default="", help="SPH_COLUMN_ORDER values to try for every pair, comma separated")
a = ap.parse_args()
combos = [(d, 0) for d in a.dv] + [(0, f) for f in a.fv if f != 0]
if a.pairs:
    combos = [tuple(int(v) for v in p.split(":")) for p in a.pairs.split(",")]
orders = [o for o in a.orders.split(",") if o] or [None]
for dv, fv, order in [(d, f, o) for d, f in combos for o in orders]:
    if order:
        os.environ["SPH_COLUMN_ORDER"] = order
    os.environ["SPH_DENSITY_VARIANT"] = str(dv)
    os.environ["SPH_FORCE_VARIANT"] = str(fv)
    ps = ParticleSystem(SimConfig(scene.NAMED_SCENES[a.scene]()))
    s = ps.build_solver(); s.initialize(); s.step(a.warm)
    torch.cuda.synchronize()
    acc = {}
    P = 20
    for _ in range(P):
        for k, v in ps._engine.profile_step().items():
            acc[k] = acc.get(k, 0.0) + v / P
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.step(100); e1.record(); torch.cuda.synchronize()
    print(f"dv={dv} fv={fv} order={order or '-'} density={acc['density']*1e3:7.1f}us force={acc['force']*1e3:7.1f}us "
          f"sort={(acc['zero']+acc['hash']+acc['scan']+acc['bucket']+acc['rank_move'])*1e3:6.1f}us "
          f"graph_step={e0.elapsed_time(e1)*10:7.1f}us", flush=True)
    del s, ps
    torch.cuda.empty_cache()
