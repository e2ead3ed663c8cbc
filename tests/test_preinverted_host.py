"""Pre-inverted EoS lookup (reference EoS_Interpolator, src/eos_interpolator.cpp) on the CPU: the arithmetic of the CUDA kernels
(ccake_b200/csrc/eos_preinverted.cuh, compiled for the host by tests/host_harness.cpp) against the reference's own
fill_thermodynamics / interpolate4D / query_thermo_host / invert_e_to_s_host / efreeze_at_T, which the oracle build runs unmodified
on tables handed over through the in-memory HDF5 stand-in. No GPU involved."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from common import charged_settings, dp, gubser_ic, harness, have_ref, make_ref

HC = 197.3269804
# reference thermo_info columns of the 11 interpolated fields, in ccake::eos_variables order (T muB muQ muS e p cs2 dwds dwdB dwdQ dwdS)
REF_COLS = [0, 1, 3, 2, 4, 9, 10, 12, 13, 15, 14]


def _tiles_from_reference(exponents, N):
    """Decade tiles whose nodes the REFERENCE's own inversion fills (what utilities/generate_eos_table.cpp does per node)."""
    from ccake_b200 import eos_preinverted as P

    settings = charged_settings()
    sim = make_ref(settings, gubser_ic(dx=0.5))
    tiles = []
    for ex in exponents:
        mn, mx, dx = P.tile_axes(ex, N)
        ax = [mn[k] + dx[k] * np.arange(N) for k in range(4)]
        S, B, Ss, Q = np.meshgrid(*ax, indexing="ij")
        dens = np.stack([S.ravel(), B.ravel(), Ss.ravel(), Q.ravel()], 1)
        seed = np.tile(np.array([750.0 / 197.0, 0.0, 0.0, 0.0]), (dens.shape[0], 1))
        th, ok = sim.eos_update_s(dens, seed)
        assert ok.all()
        v = np.stack([th[:, c] for c in REF_COLS], 0).reshape(11, N, N, N, N)
        tiles.append(dict(exponent=ex, min=mn, max=[a[-1] for a in ax], step=dx, vars=v))
    return tiles


def _hh_fill(tiles, dens, flat=False, log=None):
    L = harness()
    n = dens.shape[0]
    ints = np.zeros((len(tiles), 12), dtype=np.int32)
    dbl = np.zeros((len(tiles), 8))
    ptrs = (C.POINTER(C.c_double) * len(tiles))()
    keep = []
    for k, t in enumerate(tiles):
        v = np.ascontiguousarray(t["vars"], dtype=np.float64)
        keep.append(v)
        ints[k, 0:4] = [e + 3 for e in t["exponent"]] if not flat else [0, 0, 0, 0]
        ints[k, 4:8] = v.shape[1:5]
        ints[k, 8:12] = t.get("log_spacing", (0, 0, 0, 0))
        dbl[k, 0:4], dbl[k, 4:8] = t["min"], t["step"]
        ptrs[k] = dp(v)
    out = np.zeros((n, 11)); ok = np.zeros(n, dtype=np.int32)
    d = np.ascontiguousarray(dens, dtype=np.float64)
    L.hh_pre_fill(len(tiles), ints.ctypes.data_as(C.POINTER(C.c_int)), dp(dbl), ptrs, int(flat), C.c_int64(n), dp(d), dp(out),
                  ok.ctypes.data_as(C.POINTER(C.c_int)))
    return out, ok.astype(bool)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_tiled_lookup_matches_the_reference_interpolator():
    from oracle import ref

    exps = [(0, -1, -2, -1), (0, -1, -1, -1), (1, -1, -2, -1), (0, 0, -2, -1), (-1, -3, -3, -3), (-3, -3, -3, -3)]
    tiles = _tiles_from_reference(exps, N=6)
    path = os.path.join(tempfile.mkdtemp(prefix="ccb_pre_"), "tiles.h5")
    # the reference loads ALL 320 tiles: missing decades are registered as copies of the first tile (never addressed below)
    allt = []
    have = {t["exponent"]: t for t in tiles}
    for a in range(-3, 2):
        for b in range(-3, 1):
            for c in range(-3, 1):
                for d in range(-3, 1):
                    allt.append(have.get((a, b, c, d), dict(tiles[0], exponent=(a, b, c, d))))
    ref.h5_put_tiles(path, allt)
    rng = np.random.default_rng(5)
    dens = []
    for ex in exps:   # points inside every real tile, both signs of the charges
        lo = np.array([10.0 ** e for e in ex]); hi = 9.3 * lo   # below the last node (9.33 lo at N = 6): beyond it the reference reads past the array (idx + 1 = N)
        x = lo + (hi - lo) * rng.uniform(0, 1, (200, 4))
        x[:, 1:] *= rng.choice([-1.0, 1.0], (200, 3))
        dens.append(x)
    dens.append(np.array([[2e-4, 3e-4, -5e-4, 1e-4]]))   # everything below 1e-3: clamped onto the first node of the first tile
    dens = np.concatenate(dens)
    want = ref.interp_fill(path, dens)
    got, ok = _hh_fill(tiles, dens)
    assert ok.all()
    for k, c in enumerate(REF_COLS):
        assert np.max(np.abs(got[:, k] - want[:, c])) <= 1e-13 * max(np.max(np.abs(want[:, c])), 1.0), (k, c)
    assert np.max(np.abs(want[:, 11] - (got[:, 4] + got[:, 5]))) < 1e-12   # w = e + p
    # a decade without a tile is reported, not read
    _, ok2 = _hh_fill(tiles, np.array([[50.0, 0.5, 0.5, 0.5]]))
    assert not ok2[0]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_flat_table_lookup_and_init_queries_match_the_reference():
    from ccake_b200 import eos_preinverted as P
    from oracle import ref

    flat = P.conformal_flat_table(Ns=61, NB=41)
    path = os.path.join(tempfile.mkdtemp(prefix="ccb_pre_"), "flat.h5")
    ref.h5_put_flat(path, flat)
    # the loader's derived arrays (ccb_set_preinverted_flat does the same on the host; here through numpy for the harness)
    hc = HC
    s_ax, nB_ax = flat["s"][:, 0] / hc ** 3, flat["nB"][0, :] / hc ** 3
    T, muB = flat["T"] / hc, flat["muB"] / hc
    cTT, cTB, cBB = flat["chiTT"] / hc ** 2, flat["chiTB"] / hc ** 2, flat["chiBB"] / hc ** 2
    EPS = 1e-30
    S, B = np.meshgrid(s_ax, nB_ax, indexing="ij")
    dwds = T + S / (cTT - cTB * cTB / (cBB + EPS)) + B / (cTB - cTT * cBB / (cTB + EPS))
    dwdB = muB + S / (cTB - cBB * cTT / (cTB + EPS)) + B / (cBB - cTB * cTB / (cTT + EPS))
    e, p = flat["e"] / hc ** 4, flat["p"] / hc ** 4
    z = np.zeros_like(T)
    vars_ = np.stack([T, muB, z, z, e, p, flat["cs2"], dwds, dwdB, z, z], 0)[:, :, :, None, None]
    ls, lb = np.log10(s_ax), np.log10(nB_ax)
    tile = dict(exponent=(-3, -3, -3, -3), min=[ls[0], lb[0], 0.0, 0.0], step=[(ls[-1] - ls[0]) / (ls.size - 1), (lb[-1] - lb[0]) / (lb.size - 1), 0.0, 0.0],
                vars=vars_, log_spacing=(1, 1, 0, 0))
    rng = np.random.default_rng(6)
    n = 600
    dens = np.stack([10 ** rng.uniform(-2.5, 2.5, n), rng.choice([-1.0, 1.0], n) * 10 ** rng.uniform(-6, 0.5, n), rng.normal(0, 0.1, n), rng.normal(0, 0.1, n)], 1)
    want = ref.interp_fill(path, dens)
    got, ok = _hh_fill([tile], dens, flat=True)
    assert ok.all()
    for k, c in enumerate(REF_COLS):
        assert np.max(np.abs(got[:, k] - want[:, c])) <= 1e-12 * max(np.max(np.abs(want[:, c])), 1.0), (k, c)
    # the interpolated table reproduces the closed form it was built from (bilinear in log-log on a 10-nodes-per-decade grid: per cent)
    c0 = np.pi ** 2 * (2.0 * 8.0 + 3.5 * 3.0 * 2.5) / 90.0
    R = ((dens[:, 0] ** 2 + dens[:, 1] ** 2) / (16 * c0 * c0)) ** (1.0 / 3.0)
    assert np.max(np.abs(got[:, 0] / (dens[:, 0] / (4 * c0 * R)) - 1.0)) < 5e-2
    # initialisation queries
    L = harness()
    arrs = [np.ascontiguousarray(a) for a in (s_ax, nB_ax, T, muB, z, z, e, p, flat["cs2"], dwds, dwdB)]
    ptrs = (C.POINTER(C.c_double) * 11)(*[dp(a) for a in arrs])
    for energy in (False, True):
        a = 10 ** rng.uniform(-3.5, 3.5, n) if not energy else 10 ** rng.uniform(-4, 5, n)
        rb = rng.choice([-1.0, 1.0], n) * 10 ** rng.uniform(-8, 1.2, n)
        want_q, ok_q, ef = ref.interp_query(path, a, rb, energy=energy)
        out = np.zeros((n, 12)); okh = np.zeros(n, dtype=np.int32); efh = C.c_double()
        L.hh_pre_flat_query(s_ax.size, nB_ax.size, ptrs, int(energy), C.c_int64(n), dp(np.ascontiguousarray(a)), dp(np.ascontiguousarray(rb)), dp(out),
                            okh.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(150.0 / HC), C.byref(efh))
        assert np.array_equal(okh.astype(bool), ok_q) and ok_q.all()
        for k in range(12):
            assert np.max(np.abs(out[:, k] - want_q[:, k])) <= 1e-12 * max(np.max(np.abs(want_q[:, k])), 1.0), (energy, k)
        assert abs(efh.value - ef) <= 1e-13 * abs(ef)


def test_ccbpre_file_layout(tmp_path):
    """eos_preinverted.save_ccbpre writes what include/ccake_dropin/eos_interpolator.h reads: magic, flat flag, tile count, then the
    flat arrays in ccb_preinverted_flat order or per tile slot (= exponent + 3), N, log_spacing, min, step, vars."""
    import struct

    from ccake_b200 import eos_preinverted as P

    flat = P.conformal_flat_table(Ns=5, NB=4)
    p = tmp_path / "f.ccbpre"
    P.save_ccbpre(str(p), flat=flat)
    b = open(p, "rb").read()
    assert b[:8] == b"CCBPRE01" and struct.unpack("<iiii", b[8:24]) == (1, 0, 5, 4) and len(b) == 24 + 12 * 5 * 4 * 8
    assert np.array_equal(np.frombuffer(b[24:24 + 160]), flat["s"].ravel()) and np.array_equal(np.frombuffer(b[-160:]), flat["chiBB"].ravel())
    tiles = [dict(exponent=(-1, 0, -2, -3), min=[1, 2, 3, 4], step=[.1, .2, .3, .4], vars=np.arange(11 * 16, dtype=float).reshape(11, 2, 2, 2, 2))]
    P.save_ccbpre(str(p), tiles=tiles)
    b = open(p, "rb").read()
    assert struct.unpack("<ii", b[8:16]) == (0, 1) and struct.unpack("<4i", b[16:32]) == (2, 3, 1, 0) and struct.unpack("<4i", b[32:48]) == (2, 2, 2, 2)
    assert struct.unpack("<4d", b[64:96]) == (1.0, 2.0, 3.0, 4.0) and len(b) == 16 + 48 + 64 + 11 * 16 * 8
