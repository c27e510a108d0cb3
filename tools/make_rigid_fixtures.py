"""Generate the committed rigid-body lattice fixtures and scene JSON files.

Run in the authoring container (needs the reference's mesh assets, which are NOT
copied into this repo):

    python tools/make_rigid_fixtures.py [/root/reference]

For every ``RigidBodies`` entry of the bath scenes it runs ``voxelizer.py`` on the
reference mesh with the scene's transform and stores the filled lattice indices
(int16) plus the pitch in ``sph_taichi_b200/data/rigid/<scene>_<objectId>.npz``.
Points are ``lattice * pitch``.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_taichi_b200 import scene, voxelizer  # noqa: E402


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = os.path.join(os.path.dirname(scene.__file__), "data", "rigid")
    os.makedirs(out, exist_ok=True)
    for name in ("dragon_bath", "armadillo_bath_dynamic"):
        sc = scene.NAMED_SCENES[name]()
        pitch = 2.0 * sc["Configuration"]["particleRadius"]
        for body in sc["RigidBodies"]:
            mesh = os.path.join(ref, body["geometryFile"])
            idx, verts, _, weights = voxelizer.voxelize_rigid_body(
                mesh, body["scale"], body["rotationAngle"], body["rotationAxis"], body["translation"], pitch)
            assert np.abs(idx).max() < 32767
            dst = os.path.join(out, os.path.basename(body["voxelizedPointsFile"]))
            np.savez_compressed(dst, lattice=idx.astype(np.int16), pitch=np.float64(pitch),
                                rest_center_of_mass=voxelizer.vertex_mean(verts, weights))
            print(f"{name} body {body['objectId']}: {idx.shape[0]} voxels -> {dst}")


if __name__ == "__main__":
    main()
