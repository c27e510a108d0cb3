"""Independent all-pairs numpy restatement of the WCSPH pair sums (TEST INFRASTRUCTURE ONLY).

Purpose: pin ``sph_oracle.c`` against a second, structurally different implementation of the
same reference formulas -- dense O(N^2) distance matrices instead of the grid walk, vectorised
instead of per-pair loops.  Small N only.  Formulas follow ``sph_base.py:23-68`` (W, grad W),
``WCSPH.py:33-43`` (density), ``WCSPH.py:88-140`` (cohesion + viscosity), ``WCSPH.py:70-85``
(Tait EOS + pressure gradient, Akinci boundary mirror) of the reference.
"""
from __future__ import annotations

import numpy as np


def w_cubic(r, h):
    k = 8.0 / np.pi / h ** 3
    q = r / h
    inner = k * (6.0 * q ** 3 - 6.0 * q ** 2 + 1.0)
    outer = 2.0 * k * (1.0 - q) ** 3
    return np.where(q <= 0.5, inner, np.where(q <= 1.0, outer, 0.0))


def grad_w_cubic(rvec, h):
    k = 6.0 * 8.0 / np.pi / h ** 3
    r = np.linalg.norm(rvec, axis=-1)
    q = r / h
    safe = np.where(r > 1e-5, r, 1.0)
    scal = np.where(q <= 0.5, k * q * (3.0 * q - 2.0), -k * (1.0 - q) ** 2)
    scal = np.where((r > 1e-5) & (q <= 1.0), scal, 0.0)
    return (scal / (safe * h))[..., None] * rvec


def _pairs(x, h):
    rvec = x[:, None, :] - x[None, :, :]
    r = np.linalg.norm(rvec, axis=-1)
    nb = r < h
    np.fill_diagonal(nb, False)
    return rvec, r, nb


def densities(x, m_V, material, h, rho0):
    """rho_i for fluid i (others returned unchanged as NaN)."""
    _, r, nb = _pairs(x, h)
    W = np.where(nb, w_cubic(r, h), 0.0)
    rho = (m_V * w_cubic(0.0, h) + W @ m_V) * rho0
    return np.where(material == 1, rho, np.nan)


def non_pressure_acc(x, v, m, density, material, is_dynamic, h, d, g, sigma=0.01, nu=0.01):
    rvec, r, nb = _pairs(x, h)
    fl = material == 1
    ff = nb & fl[:, None] & fl[None, :]
    Wc = np.where(r * r > d * d, w_cubic(r, h), w_cubic(d, h))
    coh = -(sigma / m[:, None] * m[None, :] * Wc)[..., None] * rvec
    vxy = np.einsum("ijk,ijk->ij", v[:, None, :] - v[None, :, :], rvec)
    s = 10.0 * nu * (m / density)[None, :] * vxy / (r * r + 0.01 * h * h)
    visc = s[..., None] * grad_w_cubic(rvec, h)
    acc = np.where(ff[..., None], coh + visc, 0.0).sum(axis=1) + np.asarray(g)[None, :]
    static_rigid = (material == 0) & (is_dynamic == 0)
    dyn_rigid = (material == 0) & (is_dynamic != 0)
    acc[static_rigid] = 0.0
    acc[dyn_rigid] = np.asarray(g)
    return acc


def eos(density, material, rho0, stiffness, exponent):
    rho = np.where(material == 1, np.maximum(density, rho0), density)
    p = stiffness * ((rho / rho0) ** exponent - 1.0)
    return rho, p


def pressure_acc(x, m_V, rho_clamped, p, material, is_dynamic, body_density, h, rho0):
    """Pressure acceleration added to fluid particles and reaction added to dynamic rigid ones."""
    rvec, r, nb = _pairs(x, h)
    gW = grad_w_cubic(rvec, h)
    fl = material == 1
    dp = np.where(fl, p / rho_clamped ** 2, 0.0)
    dpi = dp[:, None]
    dpj = np.where(fl[None, :], dp[None, :], (p / rho0 ** 2)[:, None])
    coef = -rho0 * m_V[None, :] * (dpi + dpj)
    f = np.where((nb & fl[:, None])[..., None], coef[..., None] * gW, 0.0)  # f[i, j]: on fluid i from j
    acc = f.sum(axis=1)
    dyn = (material == 0) & (is_dynamic != 0)
    react = np.where(dyn[None, :, None], -f * (rho0 / body_density)[None, :, None], 0.0).sum(axis=0)
    return acc, react
