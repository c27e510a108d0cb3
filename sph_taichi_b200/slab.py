"""x-slab sharding of one WCSPH scene across the GPUs of a node (one process per GPU).

The reference is single-device; this is the multi-GPU extension BASELINE.json asks for.

* The grid is cut into slabs of whole x cell-layers, balanced by particle count (``plan_slabs``).  After the
  counting sort (x-major flattening, reference ``particle_system.py:292-294``) every cell layer is ONE contiguous
  index range of the packed arrays.
* Interaction radius = one cell, so each rank keeps 2 ghost layers per side: densities of the first ghost layer are
  recomputed locally from the second, and a SINGLE exchange per step is enough.
* The whole sharded step lives in ``libsph_b200.so`` (``sph_shard_*``, include/sph_b200.h): live counts, slab
  bounds, send / receive ranges and the record counts (in-band headers) are DEVICE state, the exchange is NCCL
  point-to-point issued by the library on its own communicator, and classify + sort + pair passes + exchange are a
  fixed launch sequence (asynchronous launches by default, one captured CUDA graph per step with
  ``SPH_SHARD_GRAPH=1``).  The host (this module) only builds the scene, hands the library its share and launches
  steps; it never waits on the device inside a step.  Cuts are re-balanced on the device every
  ``rebalance_every`` steps.

``tests/test_slab_gloo.py`` runs the same library code on the CPU: the host-emulated build of the kernels
(tests/emu/) with ``GlooTransport`` -- torch.distributed point-to-point over gloo -- in place of NCCL.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

GHOST_LAYERS = 2
SEND_LAYERS = GHOST_LAYERS + 2  # one more than the ghost band needs: a cut may move by a layer in any step
RECORD_ARRAYS = 4  # posm, veld, x0id, misc (acc is recomputed every step and not exchanged)
INFO_KEYS = ("n_live", "owned", "x_lo", "x_hi", "step", "own_begin", "own_end", "sendL_begin", "sendL_end",
             "sendR_begin", "sendR_end", "recv_left", "recv_right", "n_sorted", "status")
STATUS_OUT_OF_GRID, STATUS_HALO_CAPACITY, STATUS_SHARD_CAPACITY = 1, 4, 8


def plan_slabs(layer_counts, world, min_width=SEND_LAYERS + 1):
    """Cut ``len(layer_counts)`` x cell-layers into ``world`` contiguous slabs with balanced
    particle counts; every slab at least ``min_width`` layers wide.  Returns [(lo, hi)] * world."""
    counts = np.asarray(layer_counts, dtype=np.int64)
    gx = len(counts)
    if gx < world * min_width:
        raise ValueError(f"{gx} cell layers cannot host {world} slabs of >= {min_width} layers")
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        c = int(np.searchsorted(cum, target, side="left"))
        # keep room for the slabs on both sides
        c = max(c, cuts[-1] + min_width)
        c = min(c, gx - (world - r) * min_width)
        cuts.append(c)
    cuts.append(gx)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def layer_of(x, h):
    """Cell layer of positions (float32 true division, truncation; particle_system.py:287-289)."""
    return (np.asarray(x, dtype=np.float32)[:, 0] / np.float32(h)).astype(np.int32)


def select_owned(arrays, h, lo, hi):
    lay = layer_of(arrays["x"], h)
    m = (lay >= lo) & (lay < hi)
    return {k: v[m] for k, v in arrays.items()}, int(m.sum())


class GlooTransport:
    """SphTransport on torch.distributed point-to-point over HOST memory (TEST transport: the host-emulated
    library keeps its buffers in host memory; the product path is NCCL inside the library)."""

    def __init__(self, group=None):
        self.group = group
        self.reqs = []
        self.keep = []
        self.send_tag, self.recv_tag = {}, {}

        def view(ptr, nbytes):
            buf = (C.c_uint8 * int(nbytes)).from_address(ptr)
            self.keep.append(buf)
            return torch.frombuffer(buf, dtype=torch.uint8)

        def group_start(_user):
            return 0

        def group_end(_user, _stream):
            for r in self.reqs:
                r.wait()
            self.reqs.clear()
            self.keep.clear()
            return 0

        def send(_user, ptr, nbytes, peer, _stream):
            tag = self.send_tag.get(peer, 0)
            self.send_tag[peer] = tag + 1
            self.reqs.append(dist.isend(view(ptr, nbytes), peer, group=self.group, tag=tag % 30000))
            return 0

        def recv(_user, ptr, nbytes, peer, _stream):
            tag = self.recv_tag.get(peer, 0)
            self.recv_tag[peer] = tag + 1
            self.reqs.append(dist.irecv(view(ptr, nbytes), peer, group=self.group, tag=tag % 30000))
            return 0

        self._cb = (_lib.TRANSPORT_GROUP_START(group_start), _lib.TRANSPORT_GROUP_END(group_end),
                    _lib.TRANSPORT_SEND(send), _lib.TRANSPORT_RECV(recv))
        self.struct = _lib.SphTransport(None, *self._cb)


class ShardedSimulation:
    """This rank's share of a sharded scene: a CUDA engine in slab mode + the library's halo exchange."""

    def __init__(self, cfg, mine, n_mine, slabs, rank, world, device, n_cap, halo_cap, uniform_hint, group=None,
                 rebalance_every=8, transport=None):
        from . import engine as _engine
        self.cfg, self.rank, self.world, self.group = cfg, rank, world, group
        self.device = torch.device(device)
        self.slabs0 = [tuple(int(v) for v in s_) for s_ in slabs]
        ds = np.array(cfg.get_cfg("domainEnd"), dtype=np.float64) - np.array(cfg.get_cfg("domainStart"))
        radius = cfg.get_cfg("particleRadius")
        self.h = radius * 4.0
        self.grid_num = np.ceil(ds / self.h).astype(int)
        params = _engine.make_params(3, self.grid_num, radius, cfg.get_cfg("density0"), cfg.get_cfg("stiffness"),
                                     cfg.get_cfg("exponent"), cfg.get_cfg("timeStepSize"), cfg.get_cfg("gravitation"), ds)
        self.m_V0 = float(np.float32(0.8 * (2 * radius) ** 3))
        self.n_cap, self.halo_cap = int(n_cap), int(halo_cap)
        self.eng = _engine.Engine(params, n_max=self.n_cap, n_solid=0, n_bodies=0, device=self.device)
        e = self.eng
        self._load(mine, n_mine, uniform_hint)
        # ---- transport: NCCL inside the library (its own communicator), or a caller-supplied one (tests) ----
        if transport is not None:
            self._transport = transport
            e._check(e.lib.sph_comm_set_transport(e.ctx, C.byref(transport.struct), rank, world), "sph_comm_set_transport")
        elif world > 1:
            ident = [None]
            if rank == 0:
                buf = C.create_string_buffer(128)
                rc = e.lib.sph_comm_unique_id(buf)
                if rc:
                    raise RuntimeError(f"sph_comm_unique_id failed ({rc}): {e.lib.sph_last_error(None).decode()}")
                ident = [buf.raw]
            dist.broadcast_object_list(ident, src=0, group=group)
            e._check(e.lib.sph_comm_init_nccl(e.ctx, ident[0], rank, world), "sph_comm_init_nccl")
        else:
            e._check(e.lib.sph_comm_set_transport(e.ctx, None, 0, 1), "sph_comm_set_transport")
        lo, hi = self.slabs0[rank]
        e._check(e.lib.sph_shard_configure(e.ctx, lo, hi, GHOST_LAYERS, self.halo_cap, int(rebalance_every)),
                 "sph_shard_configure")
        e._check(e.lib.sph_shard_begin(e.ctx, e._stream()), "sph_shard_begin")
        self.steps_done = 0

    def _load(self, arrays, n, uniform_hint):
        """Pack this rank's initial particles (numpy arrays in the reference's field layout)."""
        dev = self.device

        def t(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)

        if int((np.asarray(arrays["material"]) != 1).sum()) != 0:
            raise NotImplementedError("x-slab sharding supports fluid-only scenes (SURVEY.md section 8e)")
        dens = t(arrays["density"], np.float32)
        f = {
            "object_id": t(arrays["object_id"], np.int32), "x": t(arrays["x"], np.float32),
            "x_0": t(arrays["x"], np.float32), "v": t(arrays["v"], np.float32),
            "acceleration": torch.zeros((n, 3), dtype=torch.float32, device=dev),
            "m_V": torch.full((n,), self.m_V0, dtype=torch.float32, device=dev), "m": dens * self.m_V0,
            "density": dens, "pressure": t(arrays["pressure"], np.float32), "material": t(arrays["material"], np.int32),
            "is_dynamic": t(arrays["is_dynamic"], np.int32), "color": t(arrays["color"], np.int32),
            "grid_ids": None, "solid_id": None,
        }
        self.eng.pack(f, n, 0, False, uniform_hint=uniform_hint)
        self.synchronize()  # the temporaries above may be freed now

    # ---- stepping --------------------------------------------------------------------------------------
    def step(self, n=1):
        e = self.eng
        e._check(e.lib.sph_shard_step(e.ctx, int(n), e._stream()), "sph_shard_step")
        self.steps_done += int(n)

    def profile_step(self):
        """ONE un-graphed step with CUDA events between the stages (synchronises)."""
        e = self.eng
        buf = (C.c_float * 6)()
        e._check(e.lib.sph_shard_profile_step(e.ctx, buf, e._stream()), "sph_shard_profile_step")
        self.steps_done += 1
        return {"sort_ms": buf[0], "density_boundary_ms": buf[1], "force_boundary_ms": buf[2], "density_interior_ms": buf[3],
                "force_interior_ms": buf[4], "exchange_ms": buf[5]}

    def info(self):
        """The device-resident step state (synchronising read); raises on a capacity / out-of-grid flag."""
        e = self.eng
        out = (C.c_int32 * 16)()
        sent = C.c_uint64(0)
        e._check(e.lib.sph_shard_info(e.ctx, out, C.byref(sent), e._stream()), "sph_shard_info")
        d = {k: int(out[i]) for i, k in enumerate(INFO_KEYS)}
        d["halo_records_sent"] = int(sent.value)
        st = d["status"]
        if st & STATUS_OUT_OF_GRID:
            raise RuntimeError("a particle left the grid (NaN or outside [0, domain))")
        if st & (STATUS_HALO_CAPACITY | STATUS_SHARD_CAPACITY):
            raise RuntimeError(f"rank {self.rank}: slab capacity exceeded (status {st}: halo_capacity {self.halo_cap}, "
                               f"capacity {self.n_cap}, live {d['n_live']}); raise capacity_factor")
        return d

    @property
    def slab(self):
        d = self.info()
        return d["x_lo"], d["x_hi"]

    @property
    def halo_bytes(self):
        return 16 * RECORD_ARRAYS * self.info()["halo_records_sent"]

    def owned_count(self):
        return self.info()["owned"]

    def record_views(self):
        """The 4 exchanged packed arrays as [n_cap, 4] float32 views of the engine workspace (current buffer set)."""
        off = (C.c_uint64 * 5)()
        e = self.eng
        e._check(e.lib.sph_state_offsets(e.ctx, off), "sph_state_offsets")
        base = e._ws_ptr - e.workspace.data_ptr()
        views = []
        for k in range(RECORD_ARRAYS):
            b0 = base + int(off[k])
            views.append(e.workspace[b0:b0 + self.n_cap * 16].view(torch.float32).view(self.n_cap, 4))
        return views

    def owned_tensors(self):
        """(x, v, x_0) of the owned particles as device tensors [owned, 3]."""
        d = self.info()
        v = self.record_views()
        b, e_ = d["own_begin"], d["own_end"]
        ghost = (v[3][b:e_, 2].contiguous().view(torch.int32) & 4) != 0
        keep = ~ghost
        return v[0][b:e_, :3][keep], v[1][b:e_, :3][keep], v[2][b:e_, :3][keep]

    def owned_state(self):
        x, v, x0 = self.owned_tensors()
        return x.cpu().numpy(), v.cpu().numpy(), x0.cpu().numpy()

    def launch_count(self):
        return self.eng.launch_count()

    def synchronize(self):
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)


def build_sharded(scene_dict, rank, world, device, group=None, capacity_factor=1.2, rebalance_every=8, slabs=None,
                  transport=None):
    """Assemble the scene on every rank, plan the slabs from the initial layer histogram and load this rank's
    share into a CUDA engine.  Returns (ShardedSimulation, n_total)."""
    from .config_builder import SimConfig
    from .scene import assemble_particles

    cfg = SimConfig(scene_dict)
    radius = cfg.get_cfg("particleRadius")
    h = radius * 4.0
    arrays, _, _, counts = assemble_particles(cfg, 3, 2 * radius)
    ds = np.array(cfg.get_cfg("domainEnd"), dtype=np.float64) - np.array(cfg.get_cfg("domainStart"))
    gx = int(np.ceil(ds / h).astype(int)[0])
    lay = layer_of(arrays["x"], h)
    hist = np.bincount(lay, minlength=gx)[:gx]
    if slabs is None:
        slabs = plan_slabs(hist, world)
    lo, hi = slabs[rank]
    mine, n_mine = select_owned(arrays, h, lo, hi)
    # one decision for the whole scene (a rank may start empty): all fluid particles share m and m_V?
    fluid = arrays["material"] == 1
    dens = arrays["density"][fluid].astype(np.float32)
    m_V0 = np.float32(0.8 * (2 * radius) ** 3)
    uniform = bool(fluid.all() and dens.size and (dens == dens[0]).all())
    hint = (uniform, float(np.float32(m_V0 * dens[0])) if dens.size else 0.0, float(m_V0))
    # the fullest layer bounds every layer the front may reach later: SEND_LAYERS of them per side and direction
    halo_cap = int(SEND_LAYERS * int(hist.max()) * capacity_factor) + 1024
    ghost_est = 2 * GHOST_LAYERS * int(hist.max())
    n_cap = int((n_mine + ghost_est) * capacity_factor) + 4 * halo_cap + 1024  # live + trash + 2 receive regions
    sim = ShardedSimulation(cfg, mine, n_mine, slabs, rank, world, device, n_cap, halo_cap, hint, group=group,
                            rebalance_every=rebalance_every, transport=transport)
    return sim, counts["total"]


# -----------------------------------------------------------------------------------------------
# sharded state against the single-GPU engine (bench.py --gpus N parity_check, tools/check_slab_parity.py)
# -----------------------------------------------------------------------------------------------
def _lexsort_rows(x0):
    """Permutation sorting the rows of a [n, 3] tensor lexicographically (x, then y, then z) on its device."""
    order = torch.arange(x0.shape[0], device=x0.device)
    for col in (2, 1, 0):
        _, idx = torch.sort(x0[order, col], stable=True)
        order = order[idx]
    return order


def gather_owned(sim, dst=0):
    """All ranks' owned (x, v, x_0) gathered on rank `dst` as device tensors (None elsewhere)."""
    x, v, x0 = sim.owned_tensors()
    rec = torch.cat([x, v, x0], dim=1).contiguous()
    if sim.world == 1:
        return rec[:, 0:3], rec[:, 3:6], rec[:, 6:9]
    n = torch.tensor([rec.shape[0]], device=rec.device, dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(sim.world)]
    dist.all_gather(counts, n, group=sim.group)
    counts = [int(c.item()) for c in counts]
    cap = max(counts)
    pad = torch.zeros((cap, 9), device=rec.device, dtype=rec.dtype)
    pad[:rec.shape[0]] = rec
    bufs = [torch.empty_like(pad) for _ in range(sim.world)] if sim.rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=sim.group)
    if sim.rank != dst:
        return None
    full = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    return full[:, 0:3], full[:, 3:6], full[:, 6:9]


def compare_with_single(gathered, ps, d):
    """Rank 0: the gathered sharded state against a ParticleSystem that ran the same steps on ONE GPU."""
    X, V, X0 = gathered
    ps._pull()
    rx, rv, rx0 = ps._t["x"], ps._t["v"], ps._t["x_0"]
    same = tuple(X0.shape) == tuple(rx0.shape)
    dx = dv = float("inf")
    if same:
        ks, kr = _lexsort_rows(X0), _lexsort_rows(rx0)
        same = bool(torch.equal(X0[ks], rx0[kr]))
        if same:
            dx = float((X[ks] - rx[kr]).abs().max().item() / d)
            dv = float((V[ks] - rv[kr]).abs().max().item())
    return {"same_particle_set": bool(same), "max_dx_over_d": dx, "max_dv": dv, "particles": int(X0.shape[0])}


# -----------------------------------------------------------------------------------------------
# bench.py --gpus N (launched by torchrun, one rank per GPU)
# -----------------------------------------------------------------------------------------------
def bench_main(args):
    import bench as _bench  # the repo-root module: shared helpers
    from . import ParticleSystem, SimConfig

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs WORLD_SIZE={args.gpus} (launch with torch.distributed.run); "
                         f"got WORLD_SIZE={world}")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    name, sc = _bench.scene_for(world, args.scene)
    d = 2.0 * sc["Configuration"]["particleRadius"]
    sim, n_total = build_sharded(sc, rank, world, dev)
    K, W = args.steps, args.warmup

    # ---- parity_check + the same-run single-GPU time: rank 0 steps the SAME scene on one GPU ----
    P = int(os.environ.get("SPH_BENCH_PARITY_STEPS", "20"))
    sim.step(P)
    gathered = gather_owned(sim)
    parity, strong = None, None
    if rank == 0:
        ps = ParticleSystem(SimConfig(sc), device=dev)
        solver = ps.build_solver()
        solver.initialize()
        solver.step(P)
        parity = dict(compare_with_single(gathered, ps, d), steps=P, against="the single-GPU engine on rank 0, same scene, "
                      "particles matched by x_0")
        parity["ok"] = bool(parity["same_particle_set"] and parity["max_dx_over_d"] < 1e-3)
        K1 = max(3, min(K, 20))
        solver.step(3)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        solver.step(K1)
        b.record()
        torch.cuda.synchronize()
        strong = {"t1_ms": a.elapsed_time(b) / K1, "t1_steps": K1,
                  "note": "t1: the whole scene on rank 0's GPU alone (other ranks idle), same run, same scene, "
                          "CUDA-graph steps right after the parity steps"}
        del solver, ps, gathered
        torch.cuda.empty_cache()
    flag = torch.tensor([1 if (parity is None or parity["ok"]) else 0], device=dev)
    dist.broadcast(flag, src=0)
    if not bool(flag.item()):
        if rank == 0:
            print(json.dumps({"error": "parity_check failed", "parity_check": parity}))
        dist.destroy_process_group()
        return 3

    sampler = _bench.ClockSampler(local) if rank == 0 else None  # NVML initialised here, well before the timed region
    sim.step(max(W - P, 3))
    sim.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    info0, launches0 = sim.info(), sim.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    a.record()
    sim.step(K)
    b.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    info1 = sim.info()
    owned = torch.tensor([info1["owned"], 64 * (info1["halo_records_sent"] - info0["halo_records_sent"]),
                          sim.launch_count() - launches0], device=dev, dtype=torch.float64)
    tot = owned.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    mx = owned.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None
    slabs = [None] * world
    dist.all_gather_object(slabs, [info1["x_lo"], info1["x_hi"]])

    # ---- per-kernel times on every rank: CUDA events between the stages of un-graphed steps ----
    acc = {}
    R = 10
    for _ in range(R):
        for k_, v_ in sim.profile_step().items():
            acc[k_] = acc.get(k_, 0.0) + v_ / R
    live_now = sim.info()
    peak, peak_src = _bench.measured_hbm_peak()
    roof = torch.tensor([acc["density_boundary_ms"] + acc["density_interior_ms"], acc["force_boundary_ms"] + acc["force_interior_ms"],
                         float(live_now["n_live"]), float(live_now["owned"]), acc["sort_ms"], acc["exchange_ms"]],
                        device=dev, dtype=torch.float64)
    roof_max = roof.clone()
    dist.all_reduce(roof_max, op=dist.ReduceOp.MAX)

    # ---- e2e: every step, H2D of the live records' {x, m_V} and {v, rho} words from pinned host
    # memory, the sharded step, and D2H of the same words (what a host-side consumer reads) ----
    Ke = max(3, min(K, 50))
    cap = sim.n_cap
    h_pos = torch.empty((cap, 4), dtype=torch.float32).pin_memory()
    h_vel = torch.empty((cap, 4), dtype=torch.float32).pin_memory()

    def pull():
        n_live = sim.info()["n_live"]
        v = sim.record_views()
        h_pos[:n_live].copy_(v[0][:n_live], non_blocking=True)
        h_vel[:n_live].copy_(v[1][:n_live], non_blocking=True)
        return n_live

    def push(n_live):
        v = sim.record_views()
        v[0][:n_live].copy_(h_pos[:n_live], non_blocking=True)
        v[1][:n_live].copy_(h_vel[:n_live], non_blocking=True)

    n_live = pull()
    torch.cuda.synchronize()
    dist.barrier()
    moved = 0
    t0 = time.perf_counter()
    for _ in range(Ke):
        push(n_live)
        sim.step()
        moved += 32 * n_live
        n_live = pull()
        moved += 32 * n_live
        torch.cuda.current_stream().synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    moved_t = torch.tensor([float(moved)], device=dev, dtype=torch.float64)
    dist.all_reduce(moved_t, op=dist.ReduceOp.SUM)
    if rank == 0:
        t_ms = float(ms.item())
        val = K / (t_ms * 1e-3)
        halo_per_step = float(tot[1].item()) / K
        strong.update(tN_ms=t_ms / K, speedup=strong["t1_ms"] / (t_ms / K), n_gpus=world)
        d_ms, f_ms = float(roof_max[0].item()), float(roof_max[1].item())
        kernels = []
        for kind, lms, n_part in (("density", d_ms, float(roof_max[2].item())), ("force", f_ms, float(roof_max[3].item()))):
            ach = _bench.ALGO_BYTES[kind] * n_part / (lms * 1e-3) / 1e9
            kernels.append({"kernel": _bench.KERNEL_NAME[kind] + ", slowest rank", "bound": "hbm", "achieved": ach, "peak": peak,
                            "unit": "GB/s", "frac": ach / peak, "traffic": None, "avg_launch_ms": lms,
                            "particles_this_rank": int(n_part),
                            "algorithmic_bytes_per_particle": _bench.ALGO_BYTES[kind]})
        kernels.sort(key=lambda e_: -e_["avg_launch_ms"])
        roofline = dict(kernels[0])
        roofline.update(peak_source=peak_src, kernels=kernels, sort_ms_slowest_rank=float(roof_max[4].item()),
                        exchange_ms_slowest_rank=float(roof_max[5].item()),
                        note="per GPU, slowest rank; CUDA events between the stages of un-graphed sharded steps (10 steps); "
                             "density = boundary + interior launches over owned + first-ghost-layer particles, forces = boundary "
                             "+ interior launches over owned particles (the interior figure includes waiting for the halo "
                             "exchange if it is not hidden behind the interior passes); the pair kernels are fp32-issue / LSU "
                             "bound, not HBM bound (DESIGN.md section 5)")
        line = {
            "metric": _bench.METRIC, "value": val * n_total / 1e6, "unit": _bench.UNIT, "steps_per_s": val,
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _bench.workload_config(name, sc, n_total, n_total),
            "sharding": {"slabs": slabs, "owned_total": int(tot[0].item()), "owned_max_per_rank": int(mx[0].item()),
                         "ghost_layers": GHOST_LAYERS, "halo_capacity_records": sim.halo_cap,
                         "exchange": "one NCCL send/recv group per step issued by libsph_b200.so on its own communicator, "
                                     "overlapped with the interior force pass; no host synchronisation in a step "
                                     "(asynchronous launches; SPH_SHARD_GRAPH=1 replays one CUDA graph per step); "
                                     "cuts re-balanced on the device every 8 steps"},
            "parity_check": parity,
            "strong_scaling": strong,
            "halo": {"bytes_per_step_all_ranks": halo_per_step,
                     "fraction_of_owned_state": halo_per_step / (64.0 * max(tot[0].item(), 1.0))},
            "clocks": clocks,
            "e2e": {"value": Ke / float(e2e_s.item()) * n_total / 1e6, "unit": _bench.UNIT,
                    "steps_per_s": Ke / float(e2e_s.item()), "steps": Ke,
                    "h2d_bytes_per_step": int(moved_t.item() / (2 * Ke)), "d2h_bytes_per_step": int(moved_t.item() / (2 * Ke)),
                    "note": "per rank and step: pinned-host -> device copy of the live records' position and velocity "
                            "words, sharded step (halo exchange inside), device -> pinned-host copy back; max over ranks"},
            "gpu_launches": int(tot[2].item()),
            "roofline": roofline,
            "stage_ms_slowest_rank": {"sort": float(roof_max[4].item()), "density": d_ms, "force": f_ms,
                                      "exchange_on_comm_stream": float(roof_max[5].item())},
            "cpu_baseline": None,
            "notes": "cpu_baseline is reported by the N = 1 run and by --impl reference (bench.py contract)",
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0
