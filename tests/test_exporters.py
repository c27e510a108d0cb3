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
