"""On-disk formats of the headless driver (reference run_simulation.py:96-112)."""
import numpy as np


def test_ascii_ply_round_trip(tmp_path):
    from run_simulation import write_ply_ascii
    rng = np.random.default_rng(3)
    pos = rng.uniform(0.0, 5.0, size=(257, 3)).astype(np.float32)
    path = tmp_path / "particle_object_0_000000.ply"
    write_ply_ascii(str(path), pos)
    lines = path.read_text().splitlines()
    assert lines[0] == "ply" and lines[1] == "format ascii 1.0"
    end = lines.index("end_header")
    header = lines[:end]
    assert "element vertex 257" in header
    assert [ln for ln in header if ln.startswith("property")] == ["property float x", "property float y",
                                                                 "property float z"]
    back = np.array([[float(v) for v in ln.split()] for ln in lines[end + 1:]], dtype=np.float32)
    assert back.shape == (257, 3)
    assert np.abs(back - pos).max() <= 1e-6 * 5.0  # 7 significant digits


def test_posed_mesh_follows_the_rigid_transform():
    """The OBJ export poses the rest mesh with the body's R and centre of mass (sph_base.py:253-257)."""
    from sph_taichi_b200.sph_base import SPHBase

    class _PS:  # the two attributes _update_mesh touches
        pass

    rest = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
    obj = {"restPosition": rest, "restCenterOfMass": rest.mean(axis=0)}
    th = 0.3
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    cm = np.array([2.0, 0.5, 1.0])

    class _Eng:
        def rigid_state(self, idx):
            assert idx == 0
            return R, cm

    ps = _PS()
    ps.object_collection = {7: obj}
    ps._engine = _Eng()
    ps._body_index = {7: 0}
    solver = SPHBase.__new__(SPHBase)
    solver.ps = ps
    solver._update_mesh(7)
    want = cm + (R @ (rest - rest.mean(axis=0)).T).T
    assert np.allclose(obj["meshVertices"], want)
    # rigid: pairwise distances preserved
    d0 = np.linalg.norm(rest[:, None] - rest[None], axis=-1)
    d1 = np.linalg.norm(obj["meshVertices"][:, None] - obj["meshVertices"][None], axis=-1)
    assert np.allclose(d0, d1)


# ---- the same formats written by a real run on the GPU ----------------------
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _export_scene():
    from sph_taichi_b200 import scene
    sc = scene.dam_break_box([16, 12, 10], domain_end=[0.8, 0.8, 0.6], start=[0.08, 0.08, 0.08])
    sc["Configuration"].update(exportPly=True, exportObj=True, numberOfStepsPerRenderUpdate=2)
    sc["RigidBodies"] = [{
        "objectId": 1, "geometryFile": os.path.join(ROOT, "tests", "golden", "cube.obj"),
        "translation": [0.5, 0.3, 0.2], "rotationAxis": [0, 1, 0], "rotationAngle": 30, "scale": [0.12, 0.12, 0.12],
        "velocity": [0.0, -1.0, 0.0], "density": 800.0, "color": [255, 255, 255], "isDynamic": True}]
    return sc


def _read_ply(path):
    lines = open(path).read().splitlines()
    assert lines[0] == "ply" and lines[1] == "format ascii 1.0"
    end = lines.index("end_header")
    n = int([ln for ln in lines[:end] if ln.startswith("element vertex")][0].split()[-1])
    pts = np.array([[float(v) for v in ln.split()] for ln in lines[end + 1:]], dtype=np.float64)
    assert pts.shape == (n, 3)
    return pts


@pytest.mark.gpu
def test_run_simulation_exports_ply_series_and_posed_obj(tmp_path):
    """run_simulation.py as a user runs it (scene file, exportPly + exportObj): the PLY series of object 0 and
    the posed rigid mesh equal the engine state of an identical in-process run (reference
    run_simulation.py:96-112, sph_base.py:253-257)."""
    from sph_taichi_b200 import ParticleSystem, SimConfig
    sc = _export_scene()
    scene_file = tmp_path / "export_case.json"
    scene_file.write_text(json.dumps(sc))
    frames = 81  # output_interval = int(0.016 / 4e-4) = 40 render updates -> exports at update 0, 40, 80
    res = subprocess.run([sys.executable, os.path.join(ROOT, "run_simulation.py"), "--scene_file", str(scene_file),
                          "--frames", str(frames), "--quiet"], cwd=tmp_path, capture_output=True, text=True,
                         timeout=600, env=dict(os.environ, PYTHONPATH=ROOT))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    out = tmp_path / "export_case_output"
    names = sorted(os.listdir(out))
    assert [n for n in names if n.endswith(".ply")] == [f"particle_object_0_{k:06}.ply" for k in range(3)]
    assert [n for n in names if n.endswith(".obj")] == [f"obj_1_{k:06}.obj" for k in range(3)]

    ps = ParticleSystem(SimConfig(sc))
    solver = ps.build_solver()
    solver.initialize()
    rest = ps.object_collection[1]["restPosition"]
    assert rest.shape == (8, 3)
    done = 0
    for k, updates in enumerate((1, 41, 81)):  # run_simulation exports after the step of render update 0, 40, 80
        solver.step(2 * updates - done)
        done = 2 * updates
        want = ps.dump(0)["position"].astype(np.float64)
        got = _read_ply(out / f"particle_object_0_{k:06}.ply")
        assert got.shape == (ps.fluid_particle_num, 3) == want.shape
        assert np.abs(got - want).max() <= 2e-6 * max(1.0, np.abs(want).max())  # %.7g
        R, cm = ps._engine.rigid_state(ps._body_index[1])
        posed = cm + (R @ (rest - ps.object_collection[1]["restCenterOfMass"]).T).T
        lines = (out / f"obj_1_{k:06}.obj").read_text().splitlines()
        v = np.array([[float(t) for t in ln.split()[1:]] for ln in lines if ln.startswith("v ")])
        f = [ln for ln in lines if ln.startswith("f ")]
        assert v.shape == (8, 3) and len(f) == 12
        assert np.abs(v - posed).max() < 1e-5
        if k:
            assert np.abs(R - np.eye(3)).max() < 0.5 and cm[1] < 0.3 + 0.06  # the body has fallen, still a rotation
            assert abs(np.linalg.det(R) - 1.0) < 1e-4
