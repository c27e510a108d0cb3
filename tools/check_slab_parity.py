"""Sharded engine vs single-GPU engine on the same scene (run under torchrun, 1 rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_slab_parity.py --counts 64 24 24 --steps 60

Rank 0 also runs the plain single-GPU engine and compares every particle (matched by x_0).
Prints one JSON line and exits non-zero on mismatch.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EMU = os.environ.get("SPH_EMU_LIB")  # host-emulated library (tests/emu/): CPU tensors, gloo instead of NCCL
if EMU:
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import emu_engine
    emu_engine.install(EMU)
import numpy as np
import torch
import torch.distributed as dist

from sph_taichi_b200 import ParticleSystem, SimConfig, scene, slab


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--counts", type=int, nargs=3, default=[64, 24, 24])
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--vx", type=float, default=1.5)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cpu") if EMU else torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        if EMU:
            dist.init_process_group("gloo", rank=rank, world_size=world)
        else:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    d = 0.02
    c = a.counts
    sc = scene.dam_break_box(c, domain_end=[2.5 * c[0] * d + 0.2, c[1] * d + 0.4, c[2] * d + 0.12], start=[0.06] * 3)
    sc["FluidBlocks"][0]["velocity"] = [a.vx, 0.0, 0.0]
    sim, n_total = slab.build_sharded(sc, rank, world, dev)
    owned_first = sim.owned_count()
    for _ in range(a.steps):
        sim.step()
    x, v, x0 = sim.owned_state()
    owned_last = sim.owned_count()
    blob = {"x": x, "v": v, "x0": x0, "owned": (owned_first, owned_last), "halo": sim.halo_bytes}
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(blob, gathered, dst=0)
    else:
        gathered = [blob]
    ok = True
    if rank == 0:
        X = np.concatenate([g["x"] for g in gathered]); V = np.concatenate([g["v"] for g in gathered])
        X0 = np.concatenate([g["x0"] for g in gathered])
        ps = ParticleSystem(SimConfig(sc), device=dev)
        s = ps.build_solver(); s.initialize(); s.step(a.steps)
        rx, rv, rx0 = ps.x.to_numpy(), ps.v.to_numpy(), ps.x_0.to_numpy()
        ks = np.lexsort((X0[:, 2], X0[:, 1], X0[:, 0])); kr = np.lexsort((rx0[:, 2], rx0[:, 1], rx0[:, 0]))
        same_set = X0.shape == rx0.shape and np.array_equal(X0[ks], rx0[kr])
        dx = float(np.abs(X[ks] - rx[kr]).max() / d) if same_set else float("inf")
        dv = float(np.abs(V[ks] - rv[kr]).max()) if same_set else float("inf")
        migrated = any(g["owned"][0] != g["owned"][1] for g in gathered)
        ok = same_set and dx < 1e-3 and dv < 1e-2
        print(json.dumps({"world": world, "particles": int(n_total), "steps": a.steps, "same_particle_set": bool(same_set),
                          "max_dx_over_d": dx, "max_dv": dv, "migrated": bool(migrated),
                          "owned_first_last": [list(map(int, g["owned"])) for g in gathered],
                          "halo_bytes": [int(g["halo"]) for g in gathered], "ok": bool(ok)}))
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, src=0)
        ok = bool(flag.item())
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
