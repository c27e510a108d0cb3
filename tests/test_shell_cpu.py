"""Host logic of the reference-surface shells on CPU: a recording stand-in replaces the CUDA engine
(only here, in the test) so that call sequences, field coherence and argument checks of
ParticleSystem / SPHBase / WCSPHSolver / DFSPHSolver can be verified without a GPU."""
import numpy as np
import pytest
import torch

from sph_taichi_b200 import SimConfig, scene
from tests.helpers import mixed_scene


class RecordingEngine:
    calls = None

    def __init__(self, params, n_max, n_solid=0, n_bodies=0, device=None):
        self.params, self.n_max = params, n_max
        self.log = []
        RecordingEngine.last = self
        self._err = 1.0e9

    def __getattr__(self, name):
        def f(*a, **k):
            self.log.append(name)
            if name == "compute_com":
                return torch.zeros(3)
            if name == "solve_constraints":
                return torch.eye(3)
            if name in ("check_status", "read_status", "launch_count"):
                return 0
            return None
        return f

    def dfsph_solve(self, mode, max_iterations, eta, offset, n_fluid, first_batch):
        self.log.append(f"dfsph_solve:{mode}:{first_batch}")
        return 2, 3, 0.5 * eta

    def dfsph_op(self, op, arg=0.0, out=None):
        self.log.append(f"dfsph:{op}")
        if op == 4 and out is not None:   # density error: converge after three sweeps
            self._err = self._err / 1.0e6
            out.fill_(self._err)


@pytest.fixture
def fake_engine(monkeypatch):
    from sph_taichi_b200 import engine
    monkeypatch.setattr(engine, "Engine", RecordingEngine)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    yield


def _ps(sc):
    from sph_taichi_b200.particle_system import ParticleSystem
    return ParticleSystem(SimConfig(sc), device="cpu")


def test_surface_attributes_and_counts(fake_engine):
    ps = _ps(mixed_scene())
    for name in ("cfg", "GGUI", "dim", "domain_start", "domain_size", "particle_radius", "particle_diameter",
                 "support_radius", "m_V0", "grid_size", "grid_num", "padding", "material_solid", "material_fluid",
                 "simulation_method", "particle_num", "particle_max_num", "fluid_particle_num", "solid_particle_num",
                 "num_rigid_bodies", "object_collection", "object_id_rigid_body", "object_id", "x", "x_0", "v",
                 "acceleration", "m_V", "m", "density", "pressure", "material", "color", "is_dynamic", "grid_ids",
                 "grid_particles_num"):
        assert hasattr(ps, name), name
    assert ps.material_solid == 0 and ps.material_fluid == 1
    assert ps.particle_num[None] == ps.particle_max_num == ps.fluid_particle_num + ps.solid_particle_num
    assert ps.support_radius == pytest.approx(0.04) and ps.m_V0 == pytest.approx(0.8 * 0.02 ** 3)
    x = ps.x.to_numpy()
    assert x.dtype == np.float32 and np.array_equal(x, ps.x_0.to_numpy())
    assert np.allclose(ps.m.to_numpy(), np.float32(ps.m_V0) * ps.density.to_numpy())
    assert ps.x[0].shape == (3,) and ps.x.shape == (ps.particle_max_num,)


def test_step_sequences(fake_engine):
    from sph_taichi_b200.WCSPH import WCSPHSolver
    ps = _ps(mixed_scene())
    s = ps.build_solver()
    assert isinstance(s, WCSPHSolver) and s.dt[None] == pytest.approx(4e-4)
    s.initialize()
    eng = RecordingEngine.last
    assert eng.log[:2] == ["set_params", "set_rigid_bodies"] or "pack" in eng.log
    assert eng.log[-3:] == ["neighbor_build", "boundary_volume", "boundary_volume"]
    eng.log.clear()
    s.step()                                   # stock substep -> one fused engine call
    assert eng.log == ["step"]

    class Custom(WCSPHSolver):
        def substep(self):
            super().substep()

    c = Custom(ps)
    eng.log.clear()
    c.step()                                   # overridden substep -> the reference's generic sequence
    assert eng.log == ["neighbor_build", "boundary_volume", "compute_densities", "compute_non_pressure_forces",
                       "compute_pressure_forces", "advect", "enforce_boundary"]
    eng.log.clear()
    c.dt[None] = 1e-4                          # solver.dt[None] = ... re-bakes the parameters
    assert eng.log == ["set_params"] and ps._dt == pytest.approx(1e-4)


def test_field_coherence_protocol(fake_engine):
    ps = _ps(mixed_scene())
    s = ps.build_solver()
    s.initialize()
    eng = RecordingEngine.last
    eng.log.clear()
    ps.x.to_numpy()                            # engine ahead -> unpack once, then cached
    ps.v.to_numpy()
    assert eng.log == ["unpack"]
    eng.log.clear()
    x = ps.x.to_numpy()
    ps.x.from_numpy(x)                         # user write -> re-pack before the next engine call
    s.step()
    assert eng.log[-2:] == ["pack", "step"] and "set_rigid_bodies" in eng.log
    with pytest.raises(ValueError):
        ps.grid_ids.fill(0)                    # derived fields are read-only
    with pytest.raises(ValueError):
        ps.x.from_numpy(x[:-1])


def test_argument_checks(fake_engine):
    sc = mixed_scene()
    sc["Configuration"]["domainStart"] = [0.1, 0.0, 0.0]
    with pytest.raises(ValueError, match="domainStart"):
        _ps(sc)
    sc = mixed_scene()
    sc["Configuration"]["simulationMethod"] = 2
    with pytest.raises(NotImplementedError):
        _ps(sc).build_solver()
    ps = _ps(mixed_scene())
    n = ps.particle_max_num
    with pytest.raises(ValueError, match="exceed particle_max_num"):
        ps.add_particles(0, 1, np.zeros((1, 3)), np.zeros((1, 3)), np.ones(1), np.zeros(1), np.ones(1), np.ones(1),
                         np.zeros((1, 3)))
    assert ps.particle_num[None] == n
    assert ps.compute_cube_particle_num([0.1, 0.1, 0.5], [1.2, 2.9, 1.6]) == 423500


def test_dfsph_host_loops(fake_engine):
    from sph_taichi_b200.DFSPH import DFSPHSolver
    sc = mixed_scene()
    sc["Configuration"]["simulationMethod"] = 4
    sc["Configuration"]["timeStepSize"] = 0.004
    ps = _ps(sc)
    s = ps.build_solver()
    assert isinstance(s, DFSPHSolver) and hasattr(ps, "dfsph_factor") and hasattr(ps, "density_adv")
    s.initialize()
    eng = RecordingEngine.last
    # default: the sweeps run in sph_dfsph_solve (loop condition on the device); what the reference does before
    # and after the loop stays in the shell, and the sweep count of one solve sizes the first batch of the next
    assert s.device_side_loops
    eng.log.clear()
    s.divergence_solve()
    assert eng.log == ["dfsph:2", "dfsph:5", "dfsph_solve:0:2", "dfsph:5"] and s.last_iterations_v == 2
    eng.log.clear()
    s.pressure_solve()
    s.pressure_solve()
    assert [c for c in eng.log if c.startswith("dfsph_solve")] == ["dfsph_solve:1:3", "dfsph_solve:1:3"]
    assert s.last_iterations == 2
    eng.log.clear()
    s.step(3)  # the whole step is one library call in this mode
    assert [c for c in eng.log if "dfsph" in c] == ["dfsph_step"]
    # the reference's own loop structure: one density-error read-back per sweep
    s.device_side_loops = False
    eng.log.clear()
    eng._err = 1.0e9
    s.divergence_solve()
    # compute_density_change, scale by 1/dt, then (iteration, density_change, error) until converged, scale back
    assert eng.log[:2] == ["dfsph:2", "dfsph:5"] and eng.log[-1] == "dfsph:5"
    sweeps = eng.log[2:-1]
    assert sweeps[:3] == ["dfsph:6", "dfsph:2", "dfsph:4"] and len(sweeps) % 3 == 0
    assert s.last_iterations_v == len(sweeps) // 3 - 1
    eng.log.clear()
    eng._err = 1.0e9
    s.substep()
    ops = [c for c in eng.log if c.startswith("dfsph")]
    assert ops[0] == "dfsph:0" and ops[1] == "dfsph:1" and ops[-1] == "dfsph:10"
    assert "dfsph:8" in ops and "dfsph:9" in ops and "dfsph:7" in ops and "dfsph:3" in ops
    assert not s._fused_step_ok()


def test_solver_attributes_drive_the_engine_constants(fake_engine):
    """sph_base.py:13-21 / WCSPH.py:9-16: the reference bakes the SOLVER's attributes into its kernels; assigning
    them (before a step) must reach the engine parameters here too, not be silently ignored."""
    ps = _ps(mixed_scene())
    solver = ps.build_solver()
    eng = RecordingEngine.last
    sent = []
    eng.set_params = lambda p: sent.append(p)
    solver.initialize()
    n0 = len(sent)
    solver.viscosity = 0.05
    solver.surface_tension = 0.02
    solver.stiffness = 40000.0
    solver.step()
    assert len(sent) == n0 + 1
    p = sent[-1]
    assert abs(p.viscosity - 0.05) < 1e-7 and abs(p.surface_tension - 0.02) < 1e-7 and abs(p.stiffness - 40000.0) < 1e-2
    solver.step()
    assert len(sent) == n0 + 1  # unchanged attributes: no re-push
    solver.dt[None] = 2e-4
    assert len(sent) == n0 + 2 and abs(sent[-1].dt - 2e-4) < 1e-9


def test_emitter_reserve_extends_the_capacity(fake_engine):
    sc = mixed_scene()
    n_scene = _ps(mixed_scene()).particle_max_num
    sc["Configuration"]["emitterReserve"] = 100
    ps = _ps(sc)
    assert ps.particle_max_num == n_scene + 100 and ps.particle_num[None] == n_scene
    f0 = ps.fluid_particle_num
    ps.add_particles(0, 100, np.full((100, 3), 0.3), np.zeros((100, 3)), np.full(100, 1000.0), np.zeros(100),
                     np.ones(100), np.ones(100), np.zeros((100, 3)))
    assert ps.particle_num[None] == n_scene + 100 and ps.fluid_particle_num == f0 + 100
    with pytest.raises(ValueError, match="emitterReserve"):
        ps.add_particles(0, 1, np.zeros((1, 3)), np.zeros((1, 3)), np.ones(1), np.zeros(1), np.ones(1), np.ones(1),
                         np.zeros((1, 3)))
