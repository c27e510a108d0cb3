"""Shared test helpers: small scenes, x_0-keyed matching."""
import numpy as np

from sph_taichi_b200 import scene


def mixed_scene(fluid_counts=(8, 8, 8), with_static=True, with_dynamic=True, domain=(0.6, 0.6, 0.6)):
    """Small fluid block next to a static and a dynamic rigid block (RigidBlocks path)."""
    d = 0.02
    sc = scene.dam_break_box(list(fluid_counts), domain_end=list(domain), start=[0.1, 0.06, 0.1])
    blocks = []
    fx = 0.1 + fluid_counts[0] * d
    if with_static:
        blocks.append({"objectId": 1, "start": [fx, 0.06, 0.1], "end": [fx + 2.5 * d, 0.06 + 5.5 * d, 0.1 + 5.5 * d],
                       "translation": [0.0, 0.0, 0.0], "scale": [1, 1, 1], "velocity": [0.0, 0.0, 0.0],
                       "density": 1000.0, "color": [255, 255, 255], "isDynamic": False})
    if with_dynamic:
        blocks.append({"objectId": 2, "start": [0.1, 0.06 + fluid_counts[1] * d, 0.1],
                       "end": [0.1 + 3.5 * d, 0.06 + (fluid_counts[1] + 2.5) * d, 0.1 + 3.5 * d],
                       "translation": [0.0, 0.0, 0.0], "scale": [1, 1, 1], "velocity": [0.0, -0.5, 0.0],
                       "density": 600.0, "color": [255, 0, 0], "isDynamic": True})
    sc["RigidBlocks"] = blocks
    return sc


def order_by_x0(x0):
    """Permutation that sorts particles by their immutable rest position."""
    x0 = np.asarray(x0, dtype=np.float32)
    return np.lexsort((x0[:, 2], x0[:, 1], x0[:, 0]))


def jitter(sim, amplitude, seed=0):
    """Deterministically perturb positions/velocities of dynamic particles (keeps x_0)."""
    rng = np.random.default_rng(seed)
    dyn = sim.is_dynamic != 0
    sim.x[dyn] += (rng.uniform(-1, 1, size=(int(dyn.sum()), 3)) * amplitude).astype(sim.x.dtype)
    sim.v[dyn] += (rng.uniform(-1, 1, size=(int(dyn.sum()), 3)) * 0.5).astype(sim.v.dtype)
