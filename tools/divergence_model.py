"""Lane-divergence model of the density pass's hit loop.

For a jittered lattice at rest density it counts, per warp of 32 consecutive sorted particles, how many loop
iterations the warp spends when the accepted hits are processed column by column (the shipped kernel: sum over
the nine (dx, dy) columns of the max over lanes) and when columns are processed in groups (max over lanes of the
per-group sums).  Pure numpy / scipy, no GPU.

    python tools/divergence_model.py [--n 28] [--jitter 0.3]
"""
import argparse

import numpy as np
from scipy.spatial import cKDTree


def hits_per_column(n1=28, jitter=0.3, seed=0, d=0.02, h=0.04):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(*[np.arange(n1)] * 3, indexing="ij"), -1).reshape(-1, 3)
    x = (0.1 + g * d + rng.uniform(-jitter, jitter, size=g.shape) * d).astype(np.float32)
    cell = (x / np.float32(h)).astype(int)
    dims = cell.max(0) + 2
    flat = (cell[:, 0] * dims[1] + cell[:, 1]) * dims[2] + cell[:, 2]
    order = np.argsort(flat, kind="stable")
    x, cell = x[order], cell[order]
    pairs = cKDTree(x).query_pairs(h * 0.999999, output_type="ndarray")
    i = np.concatenate([pairs[:, 0], pairs[:, 1]])
    j = np.concatenate([pairs[:, 1], pairs[:, 0]])
    dc = cell[j] - cell[i]
    col = (dc[:, 0] + 1) * 3 + (dc[:, 1] + 1)
    H = np.zeros((len(x), 9), int)
    np.add.at(H, (i, col), 1)
    W = len(x) // 32
    Hw = H[: W * 32].reshape(W, 32, 9)
    return Hw[Hw.sum(2).min(1) >= 25]  # warps whose particles all have full neighbourhoods


def iterations(Hw, groups):
    return float(sum(Hw[:, :, list(g)].sum(2).max(1) for g in groups).mean())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=28)
    ap.add_argument("--jitter", type=float, default=0.3)
    a = ap.parse_args()
    Hw = hits_per_column(a.n, a.jitter)
    print(f"{len(Hw)} interior warps, {Hw.sum(2).mean():.1f} hits per particle")
    for name, gs in [
        ("column by column (shipped kernel)", [(c,) for c in range(9)]),
        ("three groups by dx", [(0, 1, 2), (3, 4, 5), (6, 7, 8)]),
        ("opposite pairs + centre", [(0, 8), (1, 7), (2, 6), (3, 5), (4,)]),
        ("natural order, two groups", [(0, 1, 2, 3, 4), (5, 6, 7, 8)]),
        ("corners + centre, edges (v9 order)", [(0, 2, 6, 8, 4), (1, 3, 5, 7)]),
        ("all nine in one loop", [tuple(range(9))]),
    ]:
        print(f"{iterations(Hw, gs):6.1f} iterations per warp  {name}")
