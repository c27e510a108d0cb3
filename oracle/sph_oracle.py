"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY; pinned against the reference's own source run
under a Taichi stand-in, tests/test_golden_reference.py -- see the header of sph_oracle.c).

See ``sph_oracle.c`` for what is restated and from where.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module; the product package never does.

``OracleSim`` mirrors the reference's call surface on plain numpy arrays:
``initialize()`` (sph_base.py:80-85), ``step()`` (sph_base.py:263-271) and the individual
kernels, for fp32 (default) or fp64 (noise-floor calibration).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> None:
    """Compile liboracle_f32.so / liboracle_f64.so with the committed Makefile."""
    targets = [os.path.join(_HERE, f) for f in ("liboracle_f32.so", "liboracle_f64.so")]
    src = os.path.join(_HERE, "sph_oracle.c")
    stale = force or any((not os.path.exists(t)) or os.path.getmtime(t) < os.path.getmtime(src) for t in targets)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)


_LIBS = {}


def set_threads(n: int) -> None:
    """omp_set_num_threads for the oracle's OpenMP runtime."""
    C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))


def _params_struct(real):
    class OracleParams(C.Structure):
        _fields_ = [("n", C.c_int32), ("grid_num", C.c_int32 * 3), ("h", real), ("diameter", real), ("m_V0", real),
                    ("density0", real), ("stiffness", real), ("exponent", real), ("viscosity", real),
                    ("surface_tension", real), ("dt", real), ("g", real * 3), ("domain_size", real * 3),
                    ("k_w", real), ("k_dw", real), ("visc_eps", real)]
    return OracleParams


class _State(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "object_id", "x", "x_0", "v", "acceleration", "m_V", "m", "density", "pressure", "material", "is_dynamic",
        "color", "grid_ids", "grid_particles_num", "dfsph_factor", "density_adv")]


def _lib(f64: bool):
    key = bool(f64)
    if key not in _LIBS:
        build()
        lib = C.CDLL(os.path.join(_HERE, "liboracle_f64.so" if f64 else "liboracle_f32.so"))
        assert lib.oracle_real_bytes() == (8 if f64 else 4)
        lib.oracle_neighbor_build.restype = C.c_int
        lib.oracle_step.restype = C.c_int
        lib.oracle_dfsph_density_error.restype = C.c_double
        _LIBS[key] = lib
    return _LIBS[key]


class OracleSim:
    """CPU restatement of ParticleSystem + WCSPHSolver on numpy arrays."""

    def __init__(self, config, f64: bool = False, threads: int | None = None):
        from sph_taichi_b200.config_builder import SimConfig
        from sph_taichi_b200.scene import assemble_particles

        if not isinstance(config, SimConfig):
            config = SimConfig(config)
        self.cfg = config
        self.f64 = f64
        self.real = np.float64 if f64 else np.float32
        self.creal = C.c_double if f64 else C.c_float
        self.lib = _lib(f64)
        if threads:
            set_threads(threads)

        # derived constants, all folded in double like the reference's Python scope
        # (particle_system.py:16-46, sph_base.py:11-21, WCSPH.py:8-16)
        self.domain_start = np.array(config.get_cfg("domainStart"), dtype=np.float64)
        self.domain_size = np.array(config.get_cfg("domainEnd"), dtype=np.float64) - self.domain_start
        self.dim = len(self.domain_size)
        self.particle_radius = config.get_cfg("particleRadius")
        self.particle_diameter = 2 * self.particle_radius
        self.support_radius = self.particle_radius * 4.0
        self.m_V0 = 0.8 * self.particle_diameter ** self.dim
        self.grid_num = np.ceil(self.domain_size / self.support_radius).astype(int)
        self.method = config.get_cfg("simulationMethod")
        if self.method not in (0, 4):
            raise NotImplementedError("oracle restates WCSPH (0) and DFSPH (4) only")

        arrays, self.object_collection, self.object_id_rigid_body, counts = assemble_particles(
            config, self.dim, self.particle_diameter)
        n = counts["total"]
        self.n = n
        r = self.real
        # add_particle (particle_system.py:224-235): the fields are f32; m = f32(m_V0) * density
        self.object_id = arrays["object_id"].astype(np.int32)
        self.x = arrays["x"].astype(r)
        self.x_0 = self.x.copy()
        self.v = arrays["v"].astype(r)
        self.acceleration = np.zeros((n, 3), r)
        self.m_V = np.full(n, r(self.m_V0), r)
        self.density = arrays["density"].astype(r)
        self.m = (r(self.m_V0) * self.density).astype(r)
        self.pressure = arrays["pressure"].astype(r)
        self.material = arrays["material"].astype(np.int32)
        self.is_dynamic = arrays["is_dynamic"].astype(np.int32)
        self.color = np.ascontiguousarray(arrays["color"].astype(np.int32))
        self.grid_ids = np.zeros(n, np.int32)
        self.C = int(np.prod(self.grid_num))
        self.grid_particles_num = np.zeros(self.C, np.int32)
        self.dfsph_factor = np.zeros(n, r)
        self.density_adv = np.zeros(n, r)

        h = self.support_radius
        k = 8 / np.pi / h ** self.dim
        PS = _params_struct(self.creal)
        self.P = PS()
        self.P.n = n
        self.P.grid_num = (C.c_int32 * 3)(*[int(v) for v in self.grid_num])
        self.P.h = h
        self.P.diameter = self.particle_diameter
        self.P.m_V0 = self.m_V0
        self.P.density0 = config.get_cfg("density0")
        self.P.stiffness = config.get_cfg("stiffness")
        self.P.exponent = config.get_cfg("exponent")
        self.P.viscosity = 0.01
        self.P.surface_tension = 0.01
        self.P.dt = config.get_cfg("timeStepSize")
        self.P.g = (self.creal * 3)(*config.get_cfg("gravitation"))
        self.P.domain_size = (self.creal * 3)(*[float(v) for v in self.domain_size])
        self.P.k_w = k
        self.P.k_dw = 6.0 * k
        self.P.visc_eps = 0.01 * h ** 2
        self.dyn_ids = sorted(i for i in self.object_id_rigid_body if self.object_collection[i]["isDynamic"])
        self.rest_cm = {}

    # -- plumbing -------------------------------------------------------------------
    def _state(self):
        s = _State()
        for k, _ in _State._fields_:
            if k in ("dfsph_factor", "density_adv") and self.method != 4:
                setattr(s, k, None)
                continue
            a = getattr(self, k)
            assert a.flags["C_CONTIGUOUS"]
            setattr(s, k, a.ctypes.data)
        return s

    def _call(self, name, *extra):
        s = self._state()
        return getattr(self.lib, name)(C.byref(self.P), C.byref(s), *extra)

    # -- reference surface ----------------------------------------------------------
    def initialize_particle_system(self):
        rc = self._call("oracle_neighbor_build")
        if rc:
            raise RuntimeError(f"oracle_neighbor_build failed ({rc}): particle outside the grid")

    def compute_com(self, object_id):
        cm = (self.creal * 3)()
        self._call("oracle_compute_com", C.c_int(object_id), cm)
        return np.array(cm[:], dtype=self.real)

    def compute_static_boundary_volume(self):
        self._call("oracle_boundary_volume", C.c_int(0))

    def compute_moving_boundary_volume(self):
        self._call("oracle_boundary_volume", C.c_int(1))

    def initialize(self):
        self.initialize_particle_system()
        for oid in sorted(self.object_id_rigid_body):
            self.rest_cm[oid] = self.compute_com(oid)
        self.compute_static_boundary_volume()
        self.compute_moving_boundary_volume()

    def compute_densities(self):
        self._call("oracle_compute_densities")

    def compute_non_pressure_forces(self):
        self._call("oracle_compute_non_pressure_forces")

    def compute_pressure_forces(self):
        self._call("oracle_compute_pressure_forces")

    def advect(self):
        self._call("oracle_advect")

    def enforce_boundary_3D(self, particle_type):
        self._call("oracle_enforce_boundary_3D", C.c_int(particle_type))

    def solve_constraints(self, object_id):
        R = (self.creal * 9)()
        rc = (self.creal * 3)(*[float(v) for v in self.rest_cm[object_id]])
        self._call("oracle_solve_constraints", C.c_int(object_id), rc, R)
        return np.array(R[:], dtype=self.real).reshape(3, 3)

    def substep(self):
        if self.method == 4:
            return self.dfsph_substep()
        self.compute_densities()
        self.compute_non_pressure_forces()
        self.compute_pressure_forces()
        self.advect()

    # -- DFSPH (DFSPH.py; host loops restated from :236-276 and :314-352) -------------------------
    m_max_iterations_v = 100
    m_max_iterations = 100
    max_error_V = 0.1
    max_error = 0.05

    def compute_DFSPH_factor(self):
        self._call("oracle_dfsph_compute_factor")

    def compute_density_change(self):
        self._call("oracle_dfsph_density_change", C.c_int(0))

    def compute_density_adv(self):
        self._call("oracle_dfsph_density_change", C.c_int(1))

    def compute_density_error(self, offset):
        return float(self._call("oracle_dfsph_density_error", C.c_double(offset)))

    def multiply_time_step_factor(self, ts):
        self._call("oracle_dfsph_multiply_factor", C.c_double(ts))

    def divergence_solver_iteration_kernel(self):
        self._call("oracle_dfsph_iteration", C.c_int(0))

    def pressure_solve_iteration_kernel(self):
        self._call("oracle_dfsph_iteration", C.c_int(1))

    def predict_velocity(self):
        self._call("oracle_dfsph_predict_velocity")

    def dfsph_advect(self):
        self._call("oracle_dfsph_advect")

    def divergence_solve(self):
        dt = float(self.P.dt)
        rho0 = float(self.P.density0)
        n_fluid = int((self.material == 1).sum())
        self.compute_density_change()
        self.multiply_time_step_factor(1 / dt)
        it = 0
        avg = 0.0
        while it < 1 or it < self.m_max_iterations_v:
            self.divergence_solver_iteration_kernel()
            self.compute_density_change()
            avg = self.compute_density_error(0.0) / n_fluid
            eta = 1.0 / dt * self.max_error_V * 0.01 * rho0
            if avg <= eta:
                break
            it += 1
        self.multiply_time_step_factor(dt)
        self.last_iterations_v, self.last_err_v = it, avg
        return it

    def pressure_solve(self):
        dt = float(self.P.dt)
        rho0 = float(self.P.density0)
        n_fluid = int((self.material == 1).sum())
        self.compute_density_adv()
        self.multiply_time_step_factor(1 / (dt * dt))
        it = 0
        avg = 0.0
        while it < 1 or it < self.m_max_iterations:
            self.pressure_solve_iteration_kernel()
            self.compute_density_adv()
            avg = self.compute_density_error(rho0) / n_fluid
            eta = self.max_error * 0.01 * rho0
            if avg <= eta:
                break
            it += 1
        self.last_iterations, self.last_err = it, avg
        return it

    def dfsph_substep(self):
        self.compute_densities()
        self.compute_DFSPH_factor()
        self.divergence_solve()
        self.compute_non_pressure_forces()
        self.predict_velocity()
        self.pressure_solve()
        self.dfsph_advect()

    def step(self):
        if self.method == 4:
            # sph_base.py:263-271 with the DFSPH substep
            self.initialize_particle_system()
            self.compute_moving_boundary_volume()
            self.dfsph_substep()
            for oid in self.dyn_ids:
                self.solve_constraints(oid)
                self.enforce_boundary_3D(0)
            self.enforce_boundary_3D(1)
            return
        ids = (C.c_int32 * max(1, len(self.dyn_ids)))(*self.dyn_ids)
        rc = (self.creal * max(3, 3 * len(self.dyn_ids)))()
        for b, oid in enumerate(self.dyn_ids):
            for k in range(3):
                rc[3 * b + k] = float(self.rest_cm[oid][k])
        err = self._call("oracle_step", C.c_int(len(self.dyn_ids)), ids, rc)
        if err:
            raise RuntimeError(f"oracle_step failed ({err}): particle outside the grid")

    def dump(self, obj_id):
        mask = self.object_id == obj_id
        return {"position": self.x[mask].copy(), "velocity": self.v[mask].copy()}
