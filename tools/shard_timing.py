"""Sharded step timing (torchrun, one rank per GPU): ms/step of graph replays + the stage times of un-graphed steps.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 \
        tools/shard_timing.py --scene box_4m --steps 50
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from sph_taichi_b200 import scene, slab

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="box_4m")
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--warm", type=int, default=30)
ap.add_argument("--capacity", type=float, default=1.2)
ap.add_argument("--tag", default="")
ap.add_argument("--pre-parity", action="store_true", help="gather + single-GPU engine on rank 0 first (what bench.py does)")
ap.add_argument("--sampler", action="store_true", help="run the nvidia-smi clock sampler during the timed steps")
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
sim, n_total = slab.build_sharded(scene.NAMED_SCENES[a.scene](), rank, world, dev, capacity_factor=a.capacity)
sim.step(a.warm)
if a.pre_parity:
    from sph_taichi_b200 import ParticleSystem, SimConfig
    g = slab.gather_owned(sim)
    if rank == 0:
        ps = ParticleSystem(SimConfig(scene.NAMED_SCENES[a.scene]()), device=dev)
        sv = ps.build_solver(); sv.initialize(); sv.step(a.warm)
        print("parity", slab.compare_with_single(g, ps, 0.02), flush=True)
        del sv, ps, g
        torch.cuda.empty_cache()
    flag = torch.tensor([1], device=dev); dist.broadcast(flag, src=0)
    sim.step(3)
torch.cuda.synchronize(); dist.barrier()
sampler = None
if a.sampler and rank == 0:
    import bench
    sampler = bench.ClockSampler(int(os.environ["LOCAL_RANK"]))
sim.step(5)
torch.cuda.synchronize(); dist.barrier()
if sampler:
    sampler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sim.step(a.steps); e1.record(); torch.cuda.synchronize()
if sampler:
    print("clocks", sampler.stop(), flush=True)
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
acc = {}
for _ in range(6):
    for k, v in sim.profile_step().items():
        acc[k] = acc.get(k, 0.0) + v / 6
info = sim.info()
t = torch.tensor([acc[k] for k in sorted(acc)] + [float(info["n_live"]), float(info["owned"])], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    out = {"tag": a.tag, "scene": a.scene, "world": world, "ms_per_step": round(float(ms.item()), 4), "halo_cap": sim.halo_cap,
           "n_cap": sim.n_cap}
    out.update({k: round(float(v), 4) for k, v in zip(sorted(acc) + ["n_live_max", "owned_max"], t.tolist())})
    out["split_density"] = os.environ.get("SPH_SHARD_SPLIT_DENSITY", "default")
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
