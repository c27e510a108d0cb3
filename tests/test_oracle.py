"""Pins for the CPU oracle besides tests/test_golden_reference.py (the reference ships no tests of its own).

(i) closed-form known answers derived from the reference formulas;
(ii) agreement with the independent all-pairs numpy restatement oracle/np_ref.py;
(iii) integer-exact neighbour-build invariants; (iv) physical invariants of a step.
"""
import numpy as np
import pytest

from oracle import np_ref
from oracle.sph_oracle import OracleSim
from sph_taichi_b200 import scene
from tests.helpers import jitter, mixed_scene


def test_known_answers_kernel_constants():
    h = 0.04
    k = 8 / np.pi / h ** 3
    assert abs(k - 39788.7358) < 1e-3
    assert abs(np_ref.w_cubic(0.0, h) - k) < 1e-9
    assert abs(0.8 * 0.02 ** 3 * k - 0.254648) < 1e-6


def test_known_answers_lattice_density():
    o = OracleSim(scene.cube_8k())
    assert o.n == 8000 and o.C == 15625
    o.initialize()
    o.compute_densities()
    # interior particle: 26 neighbours, rho = 799.978; corner: 7 neighbours, rho = 485.249
    assert abs(float(o.density.max()) - 799.978) < 2e-3
    assert abs(float(o.density.min()) - 485.249) < 2e-3
    o64 = OracleSim(scene.cube_8k(), f64=True)
    o64.initialize()
    o64.compute_densities()
    assert abs(float(o64.density.max()) - 799.978) < 1e-3


def test_scene_counts():
    assert scene.cube_particle_count([0.1, 0.1, 0.5], [1.2, 2.9, 1.6], 0.02) == 423500
    assert scene.cube_particle_count([0.04, 0.04, 0.04], [4.96, 1.50, 1.96], 0.02) == 1723968
    assert scene.cube_particle_count([0.0, 0.0, 0.0], [0.6, 5.4, 0.6], 0.02) == 243000
    o = OracleSim(scene.dragon_bath(with_rigid=False))
    assert tuple(o.grid_num) == (125, 75, 50) and o.C == 468750


def test_neighbor_build_is_stable_counting_sort():
    o = OracleSim(mixed_scene())
    jitter(o, 0.004, seed=3)
    before = {k: getattr(o, k).copy() for k in ("x", "x_0", "v", "object_id", "material")}
    cells = (before["x"] / np.float32(o.support_radius)).astype(np.int32)
    flat = cells[:, 0] * o.grid_num[1] * o.grid_num[2] + cells[:, 1] * o.grid_num[2] + cells[:, 2]
    perm = np.argsort(flat, kind="stable")
    o.initialize_particle_system()
    assert np.array_equal(o.grid_ids, flat[perm])
    assert np.array_equal(o.grid_particles_num, np.cumsum(np.bincount(flat, minlength=o.C)))
    for k, a in before.items():
        assert np.array_equal(getattr(o, k), a[perm]), k


@pytest.mark.parametrize("f64", [True, False])
def test_pair_sums_match_numpy_restatement(f64):
    o = OracleSim(mixed_scene(), f64=f64)
    jitter(o, 0.004, seed=1)
    o.initialize()
    tol = 1e-10 if f64 else 3e-5
    h, d, rho0 = o.support_radius, o.particle_diameter, float(o.P.density0)
    x = o.x.astype(np.float64)
    # boundary volumes
    solid = o.material == 0
    _, r, nb = np_ref._pairs(x, h)
    delta = np_ref.w_cubic(0.0, h) + np.where(nb & solid[None, :], np_ref.w_cubic(r, h), 0.0).sum(axis=1)
    assert np.allclose(o.m_V[solid], (3.0 / delta)[solid], rtol=tol)
    o.compute_densities()
    rho = np_ref.densities(x, o.m_V.astype(np.float64), o.material, h, rho0)
    fl = o.material == 1
    assert np.allclose(o.density[fl], rho[fl], rtol=tol)
    o.compute_non_pressure_forces()
    a_np = np_ref.non_pressure_acc(x, o.v.astype(np.float64), o.m.astype(np.float64), o.density.astype(np.float64),
                                   o.material, o.is_dynamic, h, d, [0.0, -9.81, 0.0])
    scale = np.abs(a_np).max()
    assert np.abs(o.acceleration - a_np).max() <= tol * scale * 10
    a_before = o.acceleration.astype(np.float64).copy()
    body_density = o.density.astype(np.float64).copy()
    o.compute_pressure_forces()
    rho_c, p = np_ref.eos(rho, o.material, rho0, float(o.P.stiffness), float(o.P.exponent))
    rho_c = np.where(fl, rho_c, body_density)
    p = np.where(fl, p, 0.0)
    assert np.allclose(o.pressure[fl], p[fl], rtol=tol * 50, atol=tol * 5e4)
    acc, react = np_ref.pressure_acc(x, o.m_V.astype(np.float64), rho_c, p, o.material, o.is_dynamic, body_density,
                                     h, rho0)
    expect = a_before + np.where(fl[:, None], acc, 0.0) + react
    static = (o.material == 0) & (o.is_dynamic == 0)
    expect[static] = 0.0
    scale = np.abs(expect).max()
    assert scale > 100.0  # the jittered state is genuinely compressed
    assert np.abs(o.acceleration - expect).max() <= tol * scale * 50


def test_step_invariants_and_fp32_noise_floor():
    sc = scene.cube_8k()
    o32, o64 = OracleSim(sc), OracleSim(sc, f64=True)
    for o in (o32, o64):
        o.initialize()
        for _ in range(30):
            o.step()
    pad = np.float32(o32.support_radius)
    assert o32.n == 8000
    assert (o32.x >= pad).all() and (o32.x <= np.float32(1.0 - 0.04) + 1e-7).all()
    k32 = np.lexsort((o32.x_0[:, 2], o32.x_0[:, 1], o32.x_0[:, 0]))
    x0_64 = o64.x_0.astype(np.float32)
    k64 = np.lexsort((x0_64[:, 2], x0_64[:, 1], x0_64[:, 0]))
    assert np.array_equal(o32.x_0[k32], x0_64[k64])
    drift = np.abs(o32.x[k32] - o64.x[k64]).max() / 0.02
    assert drift < 1e-3, drift  # fp32 noise floor after 30 steps, in particle diameters


def test_rigid_solve_recovers_rotation():
    o = OracleSim(mixed_scene(with_static=False), f64=True)
    # RigidBlocks are never registered for shape matching by the reference
    # (particle_system.py:171-193 only adds RigidBodies to object_id_rigid_body); register by hand.
    oid = 2
    o.object_id_rigid_body.add(oid)
    o.dyn_ids = [oid]
    o.initialize()
    sel = (o.object_id == oid)
    cm0 = o.rest_cm[oid]
    th = 0.3
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    shift = np.array([0.01, -0.02, 0.005])
    rng = np.random.default_rng(0)
    o.x[sel] = (o.x_0[sel] - cm0) @ Rz.T + cm0 + shift + rng.normal(0, 1e-4, size=(int(sel.sum()), 3))
    R = o.solve_constraints(oid)
    assert np.allclose(R, Rz, atol=2e-3)
    assert np.allclose(o.x[sel], (o.x_0[sel] - cm0) @ R.T + cm0 + shift, atol=1e-4)


def _dfsph_scene():
    sc = mixed_scene()
    sc["Configuration"]["simulationMethod"] = 4
    sc["Configuration"]["timeStepSize"] = 0.004
    return sc


def test_dfsph_kernels_match_numpy_restatement():
    """DFSPH factor / density change / one Jacobi update against dense all-pairs numpy (fp64)."""
    o = OracleSim(_dfsph_scene(), f64=True)
    jitter(o, 0.002, seed=7)
    fl = o.material == 1
    c = o.x[fl].mean(axis=0)
    o.x[fl] = (o.x[fl] - c) * 0.93 + c
    o.initialize()
    fl = o.material == 1  # the sort permuted the arrays
    o.compute_densities()
    h = o.support_radius
    x, v, mV = o.x.copy(), o.v.copy(), o.m_V.copy()
    rvec, r, nb = np_ref._pairs(x, h)
    gW = np.where(nb[..., None], np_ref.grad_w_cubic(rvec, h), 0.0)
    gp = -mV[None, :, None] * gW                                  # grad_p_j, DFSPH.py:143-152
    sum_k = np.where(fl[None, :], (gp ** 2).sum(-1), 0.0).sum(1)
    grad_i = -gp.sum(1)
    tot = sum_k + (grad_i ** 2).sum(-1)
    factor = np.where(tot > 1e-6, -1.0 / tot, 0.0)
    o.compute_DFSPH_factor()
    assert np.allclose(o.dfsph_factor[fl], factor[fl], rtol=1e-10)
    dvel = v[:, None, :] - v[None, :, :]
    acc = (mV[None, :] * np.einsum("ijk,ijk->ij", dvel, gW)).sum(1)
    nn = nb.sum(1)
    want = np.where(nn < 20, 0.0, np.maximum(acc, 0.0))
    o.compute_density_change()
    assert np.allclose(o.density_adv[fl], want[fl], rtol=1e-9, atol=1e-12)
    assert (want[fl] > 0).any()
    # one divergence Jacobi sweep
    dt = 0.004
    o.multiply_time_step_factor(1 / dt)
    k = o.density_adv * o.dfsph_factor
    ksum = np.where(fl[None, :], k[:, None] + k[None, :], k[:, None])
    active = np.abs(ksum) > 1e-5
    dv = -(dt * np.where(active & nb, ksum, 0.0))[..., None] * gp
    v_want = v + np.where(fl[:, None], dv.sum(1), 0.0)
    o.divergence_solver_iteration_kernel()
    assert np.allclose(o.v, v_want, rtol=1e-9, atol=1e-12)


def test_dfsph_oracle_run_is_sane():
    sc = scene.dam_break_box([12, 16, 12], domain_end=[0.7, 0.7, 0.5], start=[0.06, 0.06, 0.06])
    sc["Configuration"]["simulationMethod"] = 4
    sc["Configuration"]["timeStepSize"] = 0.004
    o = OracleSim(sc)
    o.initialize()
    iters = 0
    for _ in range(50):
        o.step()
        iters += o.last_iterations_v + o.last_iterations
    assert iters > 10 and np.isfinite(o.x).all()
    assert float(o.density[o.material == 1].max()) < 1400.0   # incompressibility is enforced (WCSPH-free run)
    pad = np.float32(0.04)
    assert (o.x >= pad).all()
