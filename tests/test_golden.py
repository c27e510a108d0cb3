"""Golden vectors (tests/golden/, produced by the fp64 oracle -- the reference cannot run offline):
the fp32 oracle on CPU and, on a GPU, the CUDA engine must both land on them."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


def _load():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.golden_scene(), np.load(os.path.join(G, "dam_break_10x12x10_f64_40steps.npz"))


def _key(x0):
    x0 = np.asarray(x0, np.float32)
    return np.lexsort((x0[:, 2], x0[:, 1], x0[:, 0]))


def test_known_answers_file():
    ka = json.load(open(os.path.join(G, "known_answers.json")))
    assert abs(ka["k_w"] - 39788.7358) < 1e-3 and abs(ka["m_V0_W0"] - 0.254648) < 1e-6
    from sph_taichi_b200 import scene
    assert scene.cube_particle_count([0.1, 0.1, 0.5], [1.2, 2.9, 1.6], 0.02) == ka["dragon_bath_fluid_particles"]


def test_fp32_oracle_reaches_golden_state():
    from oracle.sph_oracle import OracleSim
    sc, g = _load()
    o = OracleSim(sc)
    o.initialize()
    for _ in range(40):
        o.step()
    ko, kg = _key(o.x_0), _key(g["x0"])
    assert np.array_equal(o.x_0[ko], g["x0"][kg])
    assert np.abs(o.x[ko] - g["x"][kg]).max() / 0.02 < 1e-3
    assert np.abs(o.v[ko] - g["v"][kg]).max() < 2e-2
    assert (g["x"][:, 0].min() <= 0.04 + 1e-9) and (g["x"][:, 1].min() <= 0.04 + 1e-9)  # walls were hit


@pytest.mark.gpu
def test_engine_reaches_golden_state():
    from sph_taichi_b200 import ParticleSystem, SimConfig
    sc, g = _load()
    ps = ParticleSystem(SimConfig(sc))
    s = ps.build_solver()
    s.initialize()
    s.step(40)
    x0 = ps.x_0.to_numpy()
    ke, kg = _key(x0), _key(g["x0"])
    assert np.array_equal(x0[ke], g["x0"][kg])
    assert np.abs(ps.x.to_numpy()[ke] - g["x"][kg]).max() / 0.02 < 1e-3
    assert np.abs(ps.v.to_numpy()[ke] - g["v"][kg]).max() < 2e-2


@pytest.mark.gpu
def test_empty_scene_and_single_particle():
    """Edge cases: no particles at all; one particle resting against two walls."""
    from sph_taichi_b200 import ParticleSystem, SimConfig, scene
    sc = scene.dam_break_box([1, 1, 1], domain_end=[0.4, 0.4, 0.4], start=[0.05, 0.05, 0.05])
    empty = dict(sc)
    empty["FluidBlocks"] = []
    ps = ParticleSystem(SimConfig(empty))
    assert ps.particle_max_num == 0
    s = ps.build_solver()
    s.initialize()
    s.step(3)
    assert ps.dump(0)["position"].shape == (0, 3)
    assert ps.x.to_numpy().shape == (0, 3)
    ps1 = ParticleSystem(SimConfig(sc))
    s1 = ps1.build_solver()
    s1.initialize()
    s1.step(400)  # falls 0.01 m under gravity, then bounces (restitution 0.5) on the floor pad
    x = ps1.x.to_numpy()[0]
    assert 0.04 <= x[1] < 0.0405 and abs(x[0] - 0.05) < 1e-6
    assert ps1.density.to_numpy()[0] == pytest.approx(1000.0)  # clamped: a lone particle has rho = 254.6
