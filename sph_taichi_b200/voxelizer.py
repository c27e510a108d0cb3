"""Triangle-mesh ingestion: OBJ reader, rigid transform, solid voxeliser.

The reference obtains rigid-body particles from trimesh
(``particle_system.py:421-447``): load mesh -> ``apply_scale`` -> rotate about the
vertex mean by ``rotationAngle`` (degrees, with pi taken as 3.1415926) around
``rotationAxis`` -> translate -> ``voxelized(pitch=2r).fill().points``.
trimesh is not available offline, so this module restates that pipeline with
numpy/scipy only:

* trimesh's default voxeliser subdivides every face by repeated midpoint splits
  until all edges are shorter than ``pitch / 2`` and marks the lattice sites
  ``round(vertex / pitch)``.  ``k`` midpoint splits of a triangle produce exactly
  the regular barycentric grid with ``2**k`` segments per edge, which is what
  ``surface_lattice`` samples (vectorised, grouped by ``k``).
* ``fill()`` defaults to hole filling of the dense occupancy grid
  (``scipy.ndimage.binary_fill_holes``).
* ``points`` are ``lattice_index * pitch``.

The result is a set of lattice indices; it is not guaranteed to be bit-identical to
trimesh (vertex rounding on lattice boundaries), which is acceptable because the
engine and the oracle always consume the same particle set.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

REFERENCE_PI = 3.1415926  # particle_system.py:427


def load_obj(path, with_multiplicity=False):
    """Minimal Wavefront OBJ reader: returns (vertices float64 [V,3], faces int64 [F,3]).

    ``with_multiplicity=True`` also returns, per vertex, how many copies of it trimesh's loader keeps: trimesh splits
    a vertex per distinct (position, uv, normal) VALUE combination it is used with (``unmerge_faces`` on the index
    triples, then ``merge_vertices`` by value with ``merge_tex = merge_norm = False``) and drops vertices no face
    references, so ``mesh.vertices.mean()`` -- the pivot of the reference's rotation and its rest centre of mass
    (particle_system.py:428, 436) -- is this weighted mean.  For both meshes the reference ships every vertex has
    exactly one uv / normal value (weights all 1: tests/test_voxelizer.py); restated from trimesh's source from
    memory, the package is not installable offline."""
    verts, uvs, normals = [], [], []
    faces, corners = [], []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("vt "):
                uvs.append(tuple(float(t) for t in line.split()[1:3]))
            elif line.startswith("vn "):
                normals.append(tuple(float(t) for t in line.split()[1:4]))
            elif line.startswith("f "):
                toks = [tok.split("/") for tok in line.split()[1:]]
                idx = [int(t[0]) for t in toks]
                for k in range(1, len(idx) - 1):  # fan-triangulate polygons
                    faces.append((idx[0], idx[k], idx[k + 1]))
                if with_multiplicity:
                    corners.extend(toks)
    v = np.asarray(verts, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    f = np.where(f < 0, f + len(v), f - 1)  # OBJ is 1-based; negatives are relative
    if not with_multiplicity:
        return v, f

    def pick(table, tok):
        if tok is None or tok == "" or not table:
            return None
        k = int(tok)
        return table[k - 1 if k > 0 else k]

    combos = [set() for _ in range(len(v))]
    for t in corners:
        k = int(t[0])
        vi = k - 1 if k > 0 else len(v) + k
        combos[vi].add((pick(uvs, t[1] if len(t) > 1 else None), pick(normals, t[2] if len(t) > 2 else None)))
    return v, f, np.asarray([len(c) for c in combos], dtype=np.float64)


def rotation_about_point(angle, axis, point):
    """4x4 homogeneous rotation by ``angle`` (rad) about ``axis`` through ``point``."""
    axis = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(axis)
    if n == 0.0:
        return np.eye(4)
    ax = axis / n
    c, s = np.cos(angle), np.sin(angle)
    K = np.array([[0.0, -ax[2], ax[1]], [ax[2], 0.0, -ax[0]], [-ax[1], ax[0], 0.0]])
    R = c * np.eye(3) + (1.0 - c) * np.outer(ax, ax) + s * K
    M = np.eye(4)
    M[:3, :3] = R
    p = np.asarray(point, dtype=np.float64)
    M[:3, 3] = p - R @ p
    return M


def vertex_mean(vertices, weights=None):
    """``mesh.vertices.mean(axis=0)`` of the trimesh mesh (see ``load_obj``: vertices weighted by their copies)."""
    v = np.asarray(vertices, dtype=np.float64)
    if weights is None:
        return v.mean(axis=0)
    w = np.asarray(weights, dtype=np.float64)
    return (v * w[:, None]).sum(axis=0) / w.sum()


def transform_rigid_mesh(vertices, scale, rotation_angle_deg, rotation_axis, translation, weights=None):
    """Scale, rotate about the vertex mean, translate (particle_system.py:424-431)."""
    v = np.asarray(vertices, dtype=np.float64) * np.asarray(scale, dtype=np.float64)
    angle = rotation_angle_deg / 360 * 2 * REFERENCE_PI
    M = rotation_about_point(angle, rotation_axis, vertex_mean(v, weights))
    v = v @ M[:3, :3].T + M[:3, 3]
    return v + np.asarray(translation, dtype=np.float64)


def surface_lattice(vertices, faces, pitch, max_iter=10):
    """Unique lattice sites ``round(p / pitch)`` hit by the subdivided surface."""
    tri = vertices[faces]  # [F,3,3]
    e = np.stack([tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 1], tri[:, 0] - tri[:, 2]], axis=1)
    longest = np.linalg.norm(e, axis=2).max(axis=1)
    max_edge = pitch / 2.0
    with np.errstate(divide="ignore"):
        k = np.ceil(np.log2(np.maximum(longest / max_edge, 1.0))).astype(np.int64)
    k = np.clip(k, 0, max_iter)
    hits = []
    for kk in np.unique(k):
        sel = tri[k == kk]
        n = 1 << int(kk)
        ii, jj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
        keep = (ii + jj) <= n
        b1 = ii[keep] / n
        b2 = jj[keep] / n
        b0 = 1.0 - b1 - b2
        W = np.stack([b0, b1, b2], axis=1)  # [P,3]
        # chunk to bound memory
        step = max(1, int(4_000_000 // max(1, W.shape[0])))
        for s in range(0, sel.shape[0], step):
            pts = np.einsum("pk,fkd->fpd", W, sel[s:s + step]).reshape(-1, 3)
            hits.append(np.unique(np.round(pts / pitch).astype(np.int64), axis=0))
    return np.unique(np.concatenate(hits, axis=0), axis=0)


def fill_lattice(sites):
    """Hole-fill an occupied lattice-site set; returns all occupied sites (sorted)."""
    origin = sites.min(axis=0)
    rel = sites - origin
    shape = rel.max(axis=0) + 1
    dense = np.zeros(shape, dtype=bool)
    dense[rel[:, 0], rel[:, 1], rel[:, 2]] = True
    dense = ndimage.binary_fill_holes(dense)
    filled = np.argwhere(dense) + origin  # argwhere is lexicographic (x slowest)
    return filled


def voxelize_solid(vertices, faces, pitch):
    """Lattice indices of the filled voxelisation; points are ``idx * pitch``."""
    return fill_lattice(surface_lattice(vertices, faces, pitch))


def voxelize_rigid_body(path, scale, rotation_angle_deg, rotation_axis, translation, pitch):
    """Full reference pipeline for one ``RigidBodies`` entry.

    Returns ``(lattice_idx int64 [M,3], transformed_vertices, faces, vertex_weights)``.
    """
    v, f, w = load_obj(path, with_multiplicity=True)
    v = transform_rigid_mesh(v, scale, rotation_angle_deg, rotation_axis, translation, weights=w)
    return voxelize_solid(v, f, pitch), v, f, w
