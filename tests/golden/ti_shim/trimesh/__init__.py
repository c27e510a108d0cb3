"""Stand-in for trimesh (TEST INFRASTRUCTURE): `load()` returns a synthetic "mesh" whose voxelisation is a small
lattice block, so that the reference's rigid-BODY path (registration, rest centre of mass, shape matching) can be
executed without the real mesh library.  geometryFile = "lattice:nx,ny,nz,ix,iy,iz": counts and the integer lattice
offset of the block; the voxel centres are (offset + index) * pitch, the convention of sph_taichi_b200's fixtures."""
import numpy as np


class _Vox:
    def __init__(self, pts):
        self.points = pts

    def fill(self):
        return self


class _Mesh:
    def __init__(self, spec):
        v = [int(t) for t in spec.split(":", 1)[1].split(",")]
        self.counts = v[:3]
        self.offset = np.array(v[3:6])
        self.vertices = np.array([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]])  # only .mean(axis=0) / += offset are used

    def apply_scale(self, s):
        pass

    def apply_transform(self, m):
        pass

    def copy(self):
        return self

    def voxelized(self, pitch):
        g = np.stack(np.meshgrid(*[np.arange(c) for c in self.counts], indexing="ij"), -1).reshape(-1, 3)
        return _Vox(((g + self.offset) * pitch).astype(np.float64))


def load(path):
    if not str(path).startswith("lattice:"):
        raise NotImplementedError("the trimesh stand-in only understands lattice:nx,ny,nz,ox,oy,oz")
    return _Mesh(str(path))


class transformations:
    @staticmethod
    def rotation_matrix(angle, direction, point=None):
        return np.eye(4)


class repair:
    @staticmethod
    def fill_holes(mesh):
        return True
