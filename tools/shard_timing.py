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
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
sim, n_total = slab.build_sharded(scene.NAMED_SCENES[a.scene](), rank, world, dev, capacity_factor=a.capacity)
sim.step(a.warm)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sim.step(a.steps); e1.record(); torch.cuda.synchronize()
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
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
