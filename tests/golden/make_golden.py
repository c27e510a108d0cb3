"""Regenerate the committed golden vectors (run from the repo root: python tests/golden/make_golden.py).

The reference cannot be executed offline (no taichi), so these vectors do NOT come from it: they are
produced by the fp64 build of the CPU oracle (oracle/sph_oracle.c; the vectors that DO come from the reference's
source are made by make_reference_golden.py) and serve as
regression anchors for BOTH the fp32 oracle and the CUDA engine:

  dam_break_10x12x10_f64_40steps.npz   x_0 (float32 key), x, v (float64) after 40 WCSPH steps of a
                                       1200-particle dam break that hits two walls
  known_answers.json                   closed-form values derived from the reference formulas
                                       
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.sph_oracle import OracleSim  # noqa: E402
from sph_taichi_b200 import scene  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_scene():
    sc = scene.dam_break_box([10, 12, 10], domain_end=[0.5, 0.5, 0.4], start=[0.05, 0.05, 0.05])
    sc["FluidBlocks"][0]["velocity"] = [-1.0, -2.0, 0.5]  # reaches the x and y walls within 40 steps
    return sc


def main():
    o = OracleSim(golden_scene(), f64=True)
    o.initialize()
    for _ in range(40):
        o.step()
    np.savez_compressed(os.path.join(HERE, "dam_break_10x12x10_f64_40steps.npz"),
                        x0=o.x_0.astype(np.float32), x=o.x, v=o.v, density=o.density)
    h = 0.04
    k = 8 / np.pi / h ** 3
    known = {"k_w": k, "W0": k, "m_V0": 0.8 * 0.02 ** 3, "m_V0_W0": 0.8 * 0.02 ** 3 * k,
             "lattice_interior_density": 799.978, "lattice_corner_density": 485.249,
             "lattice_interior_neighbours": 26, "lattice_corner_neighbours": 7,
             "dragon_bath_fluid_particles": 423500, "armadillo_fluid_particles": 1723968,
             "high_fluid_particles": 243000, "bath_grid": [125, 75, 50]}
    with open(os.path.join(HERE, "known_answers.json"), "w") as fh:
        json.dump(known, fh, indent=1)
    print("wrote golden vectors:", o.n, "particles")


if __name__ == "__main__":
    main()
