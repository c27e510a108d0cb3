"""The CUDA engine against the golden vectors written by the reference's OWN source (tests/golden/ref_*.npz, see
tests/test_golden_reference.py and DESIGN.md section 7.2): same scenes, same number of steps, through the public
Python surface.  Particle order and integer data must be identical, floats within the parity tolerances."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(f[4:-4] for f in os.listdir(GOLD) if f.startswith("ref_") and f.endswith(".npz") and "_body" not in f)
REL = 2e-5


def _rel(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max()) / max(float(np.abs(b).max()), 1e-30)


@pytest.mark.parametrize("name", NAMES)
def test_engine_reproduces_the_reference_source(name):
    from sph_taichi_b200 import ParticleSystem, SimConfig
    z = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    scene = json.loads(str(z["scene"]))
    for b in scene.get("RigidBodies", []):
        b["voxelizedPointsFile"] = os.path.join(GOLD, b["voxelizedPointsFile"])
    ps = ParticleSystem(SimConfig(scene))
    solver = ps.build_solver()
    solver.initialize()
    dfsph = scene["Configuration"]["simulationMethod"] == 4
    for stage in ("init_", "final_"):
        if stage == "final_":
            its = []
            for _ in range(int(z["steps"])):
                solver.step()
                if dfsph:
                    its.append((solver.last_iterations_v, solver.last_iterations))
            if dfsph:  # a convergence test sitting on its threshold may flip by one sweep (tests/test_gpu_dfsph.py)
                assert max(abs(a - int(b)) for (a, _), b in zip(its, z["dfsph_iterations_v"])) <= 1
                assert max(abs(a - int(b)) for (_, a), b in zip(its, z["dfsph_iterations"])) <= 1
        assert ps._engine.check_status() == 0
        if stage + "x" not in z.files:  # the 8 K cube keeps the final state only
            continue
        # Shape matching sums in a different order than the reference's serial loops (fp64 moments on the GPU), so a
        # body particle within 1e-7 of a cell face may sort into the neighbouring cell: for scenes with dynamic
        # bodies the particles are matched by their immutable (object id, x_0) key after the first step.
        got = {f: getattr(ps, f).to_numpy() for f in ("object_id", "material", "is_dynamic", "grid_ids", "x_0", "x", "v",
                                                       "m_V", "density", "pressure", "acceleration")
               if stage + f in z.files}
        want = {f: z[stage + f] for f in got}
        # ... and, on any scene, a particle within an ulp of a cell face may do the same once approximate rsqrt / rcp
        # are in play: the init stage is always exact, later stages fall back to the key match if the order differs
        loose = stage == "final_" and (any(b["isDynamic"] for b in scene.get("RigidBodies", []))
                                       or not np.array_equal(got["x_0"], want["x_0"]))
        if loose:
            kg = np.lexsort((got["x_0"][:, 2], got["x_0"][:, 1], got["x_0"][:, 0], got["object_id"]))
            kw = np.lexsort((want["x_0"][:, 2], want["x_0"][:, 1], want["x_0"][:, 0], want["object_id"]))
            got = {f: a[kg] for f, a in got.items()}
            want = {f: a[kw] for f, a in want.items()}
            assert np.mean(got["grid_ids"] == want["grid_ids"]) > 0.95  # lattice bodies start exactly on cell faces
        else:
            assert np.array_equal(got["grid_ids"], want["grid_ids"]), stage
            assert np.array_equal(ps.grid_particles_num.to_numpy(), z[stage + "grid_particles_num"]), stage
        for f in ("object_id", "material", "is_dynamic", "x_0"):
            if f in got:
                assert np.array_equal(got[f], want[f]), (stage, f)
        tol = 50 * REL if dfsph else REL
        for f, k in (("x", 1), ("v", 5), ("m_V", 1), ("density", 1), ("pressure", 10), ("acceleration", 10)):
            if f in got:
                err = _rel(got[f], want[f])
                assert err < k * tol, f"{name} {stage}{f}: {err:.3e}"
