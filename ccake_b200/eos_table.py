"""Synthetic 4-D equation-of-state tables in the reference's ``thermo.dat`` format.

The lattice-QCD table the reference runs with (EoS/Houston/thermo.dat, Zenodo 6829115) is not part of
its tree and cannot be downloaded here, so configs 3 and 4 run on a table generated from a smooth
analytic pressure whose first and second derivatives are known in closed form:

    p(T, mu) = c0 * S(T) * X^2,   X = T^2 + sum_ij k_ij mu_i mu_j,   S(T) = d + (1-d)/2 (1 + tanh((T-Tc)/Ts))

(conformal at high T, a crossover around Tc, positive-definite susceptibilities with small cross terms).

File format (reference src/interpolatorND.cpp:125-190, src/eos_table.cpp:14-122): first line = the four grid
sizes (T, muB, muQ, muS); then one row per node, T slowest and muS fastest (InterpolatorND::indexer,
include/interpolatorND.h:77-83): T muB muQ muS [MeV] followed by the 17 fields
p s B S Q e cs2 chiBB chiQQ chiSS chiBQ chiBS chiQS chiTB chiTQ chiTS chiTT, either as dimensionless ratios
(p/T^4, s/T^3, chi/T^2; ``normalize_by_T: true``, the reference default) or in MeV^n.

The table is READ by the library itself (``ccb_load_eos_table_dat``), with the same parsing and rescaling
arithmetic as the reference, so both sides hold bit-identical node values.
"""
from __future__ import annotations

import os

import numpy as np

HBARC_MEVFM = 197.3269804  # reference include/constants.h:16

# symmetric "stiffness" of the chemical potentials, order (B, Q, S) = the table's axis order
_K = np.array([[0.110, 0.012, -0.020],
               [0.012, 0.160, 0.015],
               [-0.020, 0.015, 0.140]])
_C0 = 4.2
_TC, _TS, _D = 170.0, 45.0, 0.12  # MeV


def pressure_and_derivatives(T, muB, muQ, muS):
    """p, its gradient (s, nB, nQ, nS) and Hessian [T,B,Q,S]x[T,B,Q,S] in MeV units (MeV^4, MeV^3, MeV^2)."""
    T, muB, muQ, muS = np.broadcast_arrays(*(np.asarray(a, dtype=np.float64) for a in (T, muB, muQ, muS)))
    mu = np.stack([muB, muQ, muS], -1)
    kmu = mu @ _K.T
    X = T * T + np.einsum("...i,...i->...", mu, kmu)
    th = np.tanh((T - _TC) / _TS)
    S = _D + (1.0 - _D) * 0.5 * (1.0 + th)
    S1 = (1.0 - _D) * 0.5 * (1.0 - th * th) / _TS
    S2 = -2.0 * th * S1 / _TS
    p = _C0 * S * X * X
    g = np.empty(T.shape + (4,))
    g[..., 0] = _C0 * (S1 * X * X + S * 4.0 * X * T)
    for i in range(3):
        g[..., 1 + i] = _C0 * S * 4.0 * X * kmu[..., i]
    H = np.empty(T.shape + (4, 4))
    H[..., 0, 0] = _C0 * (S2 * X * X + 8.0 * S1 * X * T + S * (8.0 * T * T + 4.0 * X))
    for i in range(3):
        H[..., 0, 1 + i] = H[..., 1 + i, 0] = _C0 * (S1 * 4.0 * X * kmu[..., i] + S * 8.0 * T * kmu[..., i])
        for j in range(3):
            H[..., 1 + i, 1 + j] = _C0 * S * (8.0 * kmu[..., i] * kmu[..., j] + 4.0 * X * _K[i, j])
    return p, g, H


def synthetic_table(T_axis, mu_axis, normalize_by_T=True):
    """Rows of a thermo.dat table on the product grid T_axis x mu_axis^3 (all in MeV)."""
    T_axis, mu_axis = np.asarray(T_axis, dtype=np.float64), np.asarray(mu_axis, dtype=np.float64)
    T, B, Q, S = np.meshgrid(T_axis, mu_axis, mu_axis, mu_axis, indexing="ij")
    T, B, Q, S = T.ravel(), B.ravel(), Q.ravel(), S.ravel()
    p, g, H = pressure_and_derivatives(T, B, Q, S)
    s, nB, nQ, nS = g[:, 0], g[:, 1], g[:, 2], g[:, 3]
    e = T * s + B * nB + Q * nQ + S * nS - p
    w = e + p
    cs2 = np.einsum("ni,ni->n", g, np.linalg.solve(H, g[..., None])[..., 0]) / w  # n.H^-1.n / w at fixed s/n_X
    f = np.stack([p, s, nB, nS, nQ, e, cs2, H[:, 1, 1], H[:, 2, 2], H[:, 3, 3], H[:, 1, 2], H[:, 1, 3], H[:, 2, 3],
                  H[:, 0, 1], H[:, 0, 2], H[:, 0, 3], H[:, 0, 0]], 1)
    if normalize_by_T:
        powers = np.array([4, 3, 3, 3, 3, 4, 0] + [2] * 10)
        f = f / T[:, None] ** powers[None, :]
    grid = np.stack([T, B, Q, S], 1)
    return dict(sizes=(T_axis.size, mu_axis.size, mu_axis.size, mu_axis.size), grid=grid, fields=f,
                normalize_by_T=bool(normalize_by_T))


def write_thermo_dat(directory, table):
    """``<directory>/thermo.dat`` with 17 significant digits (values round-trip exactly)."""
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, "thermo.dat")
    rows = np.concatenate([table["grid"], table["fields"]], 1)
    with open(path, "w") as f:
        f.write(" ".join(str(int(n)) for n in table["sizes"]) + "\n")
        np.savetxt(f, rows, fmt="%.17g")
    return path


def default_axes(coarse=False):
    """Default synthetic grid. The T axis must contain the reference's 580 MeV seed (src/sph_workstation.cpp:1291).
    fine: T 30..800 step 5 MeV, mu -450..450 step 50 MeV (155 x 19^3 = 1.06 M nodes);
    coarse (tests): T 40..640 step 20, mu -450..450 step 150 (31 x 7^3 = 10 633 nodes)."""
    if coarse:
        return np.arange(40.0, 640.0 + 1e-9, 20.0), np.arange(-450.0, 450.0 + 1e-9, 150.0)
    return np.arange(30.0, 800.0 + 1e-9, 5.0), np.arange(-450.0, 450.0 + 1e-9, 50.0)


def ensure_table(directory, coarse=True, normalize_by_T=True):
    """Write the synthetic table once; returns the directory (the settings' ``eos_path``)."""
    path = os.path.join(directory, "thermo.dat")
    if not os.path.exists(path):
        T_axis, mu_axis = default_axes(coarse)
        write_thermo_dat(directory, synthetic_table(T_axis, mu_axis, normalize_by_T))
    return directory
