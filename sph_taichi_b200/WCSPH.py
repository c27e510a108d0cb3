"""``WCSPHSolver`` (reference ``WCSPH.py:5-156``): weakly compressible SPH with Tait EOS,
cohesion surface tension, Monaghan-type viscosity and Akinci rigid coupling -- shell over the
CUDA engine; the per-method entry points are the un-fused kernels, ``step()`` the fused path."""
from __future__ import annotations

from .sph_base import SPHBase


class WCSPHSolver(SPHBase):
    def __init__(self, particle_system):
        super().__init__(particle_system)
        self.exponent = self.ps.cfg.get_cfg("exponent")
        self.stiffness = self.ps.cfg.get_cfg("stiffness")
        self.surface_tension = 0.01
        self.dt[None] = self.ps.cfg.get_cfg("timeStepSize")

    def _call(self, fn):
        self.ps._push()
        fn()
        self.ps._after_engine()

    def compute_densities(self):
        self._call(self.ps._engine.compute_densities)

    def compute_non_pressure_forces(self):
        self._call(self.ps._engine.compute_non_pressure_forces)

    def compute_pressure_forces(self):
        self._call(self.ps._engine.compute_pressure_forces)

    def advect(self):
        self._call(self.ps._engine.advect)

    def substep(self):
        self.compute_densities()
        self.compute_non_pressure_forces()
        self.compute_pressure_forces()
        self.advect()

    def _fused_step_ok(self):
        return type(self).substep is WCSPHSolver.substep
