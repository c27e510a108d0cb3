"""N > 1 on the CPU: the sharded engine's library code (sph_shard_*: device-side bookkeeping kernels, halo
exchange, step sequence and graph replay) in the host-emulated build, world_size 2 and 3 over the gloo backend --
``slab.GlooTransport`` stands in for NCCL.  The sharded run must reproduce the single-domain engine particle by
particle (matched by x_0), with particles migrating between ranks and the slab cuts re-balanced on the way."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from sph_taichi_b200 import slab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


def test_plan_slabs_balanced_and_valid():
    hist = np.array([0, 0, 500, 500, 500, 500, 500, 500, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    for world in (1, 2, 3, 4, 5):
        slabs = slab.plan_slabs(hist, world)
        assert slabs[0][0] == 0 and slabs[-1][1] == len(hist)
        assert all(b[0] == a[1] for a, b in zip(slabs, slabs[1:]))
        assert all(hi - lo >= slab.SEND_LAYERS + 1 for lo, hi in slabs)
    s2 = slab.plan_slabs(hist, 2)
    assert abs(hist[s2[0][0]:s2[0][1]].sum() - hist[s2[1][0]:s2[1][1]].sum()) <= 1000
    with pytest.raises(ValueError):
        slab.plan_slabs(np.ones(7), 2)


def _run(nproc, port, extra, graph=False, split_density=False):
    import build_emu
    lib = build_emu.build()
    env = dict(os.environ, SPH_EMU_LIB=lib, PYTHONPATH=ROOT, SPH_SHARD_GRAPH="1" if graph else "0",
               SPH_SHARD_SPLIT_DENSITY="1" if split_density else "0")
    env.pop("SPH_B200_LIB", None)
    script = os.path.join(ROOT, "tools", "check_slab_parity.py")
    if nproc == 1:
        cmd = [sys.executable, script, *extra]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
               "127.0.0.1", "--master-port", str(port), script, *extra]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env, cwd=ROOT)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
    return json.loads(lines[-1])


def test_sharded_two_ranks_equal_the_single_engine():
    """Two emulated ranks over gloo: halo exchange every step, particles cross the cut (12 m/s for 16 steps); the
    steps replay the captured graph (SPH_SHARD_GRAPH=1: capture of the kernels AND of the transport calls)."""
    out = _run(2, 29547, ["--counts", "24", "8", "8", "--steps", "16", "--vx", "12"], graph=True)
    assert out["ok"] and out["same_particle_set"] and out["max_dx_over_d"] < 1e-4, out
    assert out["migrated"] or out["cuts_moved"], out
    assert all(h > 0 for h in out["halo_bytes"]), out


def test_three_ranks_skewed_cuts_are_rebalanced_on_the_device():
    """Deliberately skewed cuts, re-balancing every 2 steps: the cuts must move (decided on the device by both
    ranks of a cut from the exchanged headers), particles migrate, and the result still equals the single engine
    (default mode: asynchronous un-graphed launches; the 8-rank sequence with the split density pass -- boundary
    densities and forces first, the exchange behind the interior densities and forces)."""
    out = _run(3, 29548, ["--counts", "32", "8", "8", "--steps", "24", "--vx", "8", "--skew", "-2", "--rebalance-every", "2"],
               split_density=True)
    assert out["ok"] and out["same_particle_set"] and out["max_dx_over_d"] < 1e-4, out
    assert out["migrated"] and out["cuts_moved"], out
    first, last = out["owned_first_last"], None
    spread0 = max(f[0] for f in first) - min(f[0] for f in first)
    spread1 = max(f[1] for f in first) - min(f[1] for f in first)
    assert spread1 < spread0, out   # better balanced than at the start
