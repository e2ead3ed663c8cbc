"""Pre-inverted equation-of-state tables (``Settings::online_inverter_enabled = false``; reference ``EoS_Interpolator``,
``include/eos_interpolator.h``, and its generator ``utilities/generate_eos_table.cpp``).

The reference keeps these tables in HDF5 (absent from this image); here they are plain arrays -- handed to the CUDA library
through ``ccb_set_preinverted_table`` / ``ccb_set_preinverted_flat`` -- with ``numpy.savez`` files as the on-disk format.
The generator runs on the device: every node of a decade tile is located with the run's own inversion kernel
(``ccb_generate_preinverted_tile``), which turns the reference's multi-hour OpenMP job into seconds.
"""
from __future__ import annotations

import math

import numpy as np

HBARC_MEVFM = 197.3269804  # reference include/constants.h:16
PRE_VARS = ["T", "muB", "muQ", "muS", "e", "p", "cs2", "dwds", "dwdB", "dwdQ", "dwdS"]  # ccake::eos_variables (include/eos_interpolator.h:36-43)
TILE_RANGES = [range(-3, 2), range(-3, 1), range(-3, 1), range(-3, 1)]  # decades of s, rhoB, rhoS, rhoQ (generate_eos_table.cpp:391-395)


def tile_axes(exponent, N=20):
    """Axes of the decade tile `exponent` = (es, eB, eS, eQ) (Inverter::set_range, utilities/generate_eos_table.cpp:189-203):
    x_min = 10^e, x_max = 10^(e + 1), N nodes x_min + i dx with dx = x_max / N."""
    mn = [10.0 ** e for e in exponent]
    mx = [10.0 ** (e + 1) for e in exponent]
    return mn, mx, [m / N for m in mx]


def generate_tile(ws, exponent, N=20):
    """One tile on the device (``ws``: an SPHWorkstation whose EoS is the forward EoS to invert). Returns the dict the
    setters take: exponent, min, max, step, vars (11, N, N, N, N), eos (N, N, N, N) = answering member of chosen_EOSs (-1: no root).
    NB the header the reference's generator writes (max = 10^(e + 1)) makes its loader count N + 1 nodes per axis
    (get_N, src/eos_interpolator.cpp:118-125); ``max`` here is the last node, so that N is read back."""
    vars_, eos = ws.generate_preinverted_tile(exponent, N)
    mn, mx, dx = tile_axes(exponent, N)
    last = [mn[k] + (N - 1) * dx[k] for k in range(4)]
    return dict(exponent=tuple(int(e) for e in exponent), min=mn, max=last, step=dx, vars=vars_, eos=eos)


def generate_table(ws, exponents=None, N=20):
    """All 5 x 4 x 4 x 4 decade tiles (or the listed ones)."""
    if exponents is None:
        exponents = [(a, b, c, d) for a in TILE_RANGES[0] for b in TILE_RANGES[1] for c in TILE_RANGES[2] for d in TILE_RANGES[3]]
    return [generate_tile(ws, e, N) for e in exponents]


def save_tiles(path, tiles):
    np.savez(path, **{"%s_%d" % (k, i): np.asarray(t[k]) for i, t in enumerate(tiles) for k in ("exponent", "min", "max", "step", "vars")})


def load_tiles(path):
    z = np.load(path)
    n = len([k for k in z.files if k.startswith("vars_")])
    return [dict(exponent=tuple(int(e) for e in z["exponent_%d" % i]), min=list(z["min_%d" % i]), max=list(z["max_%d" % i]),
                 step=list(z["step_%d" % i]), vars=z["vars_%d" % i]) for i in range(n)]


def conformal_flat_table(Ns=121, NB=81, s_range=(1e-3, 1e3), nB_range=(1e-7, 10.0), c=None):
    """A FLAT (s, nB) table (the second on-disk format of EoS_Interpolator, src/eos_interpolator.cpp:262-521) of the conformal EoS
    p = c (T^2 + muB^2)^2 at muS = muQ = 0, in the units the file format uses (MeV^n), with the chi tensor the loader derives
    dw/ds, dw/dB from. Closed form: R = T^2 + mu^2 = ((s^2 + nB^2) / 16 c^2)^(1/3), T = s / 4cR, mu = nB / 4cR."""
    if c is None:
        c = math.pi ** 2 * (2.0 * (3.0 ** 2 - 1.0) + 3.5 * 3.0 * 2.5) / 90.0
    s = np.logspace(math.log10(s_range[0]), math.log10(s_range[1]), Ns)
    nB = np.logspace(math.log10(nB_range[0]), math.log10(nB_range[1]), NB)
    S, B = np.meshgrid(s, nB, indexing="ij")
    R = ((S * S + B * B) / (16.0 * c * c)) ** (1.0 / 3.0)
    T, mu = S / (4 * c * R), B / (4 * c * R)
    p = c * R * R
    hc = HBARC_MEVFM
    z = np.zeros_like(S)
    return dict(s=S * hc ** 3, nB=B * hc ** 3, T=T * hc, muB=mu * hc, muS=z, muQ=z, e=3 * p * hc ** 4, p=p * hc ** 4, cs2=z + 1.0 / 3.0,
                chiTT=4 * c * (3 * T * T + mu * mu) * hc ** 2, chiTB=8 * c * T * mu * hc ** 2, chiBB=4 * c * (T * T + 3 * mu * mu) * hc ** 2)


def save_ccbpre(path, tiles=None, flat=None):
    """The binary file the drop-in ``EoS_Interpolator(fs::path)`` reads (``include/ccake_dropin/eos_interpolator.h``; HDF5 is absent
    from this image): magic ``CCBPRE01``, int32 flat, int32 ntiles, then either Ns, NB + the twelve (Ns, NB) arrays of the flat table
    in ccb_preinverted_flat order, or per tile slot[4], N[4], log_spacing[4] (int32), min[4], step[4] (float64) and vars (11, N0..N3)."""
    import struct

    with open(path, "wb") as f:
        f.write(b"CCBPRE01")
        if flat is not None:
            Ns, NB = flat["s"].shape
            f.write(struct.pack("<iiii", 1, 0, Ns, NB))
            for k in ("s", "nB", "T", "muB", "muS", "muQ", "e", "p", "cs2", "chiTT", "chiTB", "chiBB"):
                f.write(np.ascontiguousarray(flat[k], dtype=np.float64).tobytes())
        else:
            f.write(struct.pack("<ii", 0, len(tiles)))
            for t in tiles:
                v = np.ascontiguousarray(t["vars"], dtype=np.float64)
                f.write(struct.pack("<4i", *[int(e) + 3 for e in t["exponent"]]))   # slot = decade exponent + 3 (ccb_preinverted_tile)
                f.write(struct.pack("<4i", *v.shape[1:]))
                f.write(struct.pack("<4i", *[int(x) for x in t.get("log_spacing", (0, 0, 0, 0))]))
                f.write(struct.pack("<4d", *[float(x) for x in t["min"]]))
                f.write(struct.pack("<4d", *[float(x) for x in t["step"]]))
                f.write(v.tobytes())
