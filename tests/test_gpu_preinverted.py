"""Pre-inverted equation of state on the GPU (Settings::online_inverter_enabled = false, the reference's default mode):
the table generator (utilities/generate_eos_table.cpp) run with the device inversion kernel, the phase-04 lookup
(EoS_Interpolator::fill_thermodynamics) through the C ABI, and a whole run on the flat (s, nB) table, all against the
reference build (oracle/_ref), which runs its own EoS_Interpolator on the same arrays."""
import os
import tempfile

import numpy as np
import pytest

from common import charged_settings, gubser_ic, have_ref, make_ref, rel_err, viscous_settings
from test_gpu_parity import TOL, _cmp, _ext_shv5, _table
from test_preinverted_host import REF_COLS

pytestmark = [pytest.mark.gpu]
HC = 197.3269804


@pytest.mark.parametrize("eos", ["conformal"])
def test_generator_matches_the_reference_inversion(eos):
    """ccb_generate_preinverted_tile: every node of a decade tile against EquationOfState::update_s from the generator's seed."""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    from ccake_b200 import SPHWorkstation, eos_preinverted as P

    # (the generator's seed, T = 750 / 197 fm^-1, lies above the 640 MeV the coarse test table covers, where the reference exits, and the
    # reference segfaults on a one-member tanh_conformal chain when a node has no root: the chain path of the same inversion kernel is
    # compared with the reference in test_gpu_parity.py::test_eos_queries_match_reference instead)
    settings = charged_settings(eos_type=eos)
    ws = SPHWorkstation(settings)
    sim = make_ref(settings, gubser_ic(dx=0.5))
    N = 6
    for ex in [(0, -1, -2, -1), (1, -1, -1, -2), (-1, -2, -3, -3)]:
        t = P.generate_tile(ws, ex, N)
        mn, mx, dx = P.tile_axes(ex, N)
        ax = [mn[k] + dx[k] * np.arange(N) for k in range(4)]
        S, B, Ss, Q = np.meshgrid(*ax, indexing="ij")
        dens = np.stack([S.ravel(), B.ravel(), Ss.ravel(), Q.ravel()], 1)
        th, ok = sim.eos_update_s(dens, np.tile(np.array([750.0 / 197.0, 0.0, 0.0, 0.0]), (dens.shape[0], 1)))
        got_ok = t["eos"].ravel() >= 0
        assert np.array_equal(got_ok, ok), (eos, ex, int(ok.sum()), int(got_ok.sum()))
        v = t["vars"].reshape(11, -1)
        sc_pt = np.max(np.abs(th[ok][:, 0:4]))
        for k, c in enumerate(REF_COLS):
            sc = sc_pt if k < 4 else max(np.max(np.abs(th[ok][:, c])), 1e-12)
            # the reference stores a hybrids iterate whose densities miss the node by up to sum |f| < 1e-10 ABSOLUTE
            # (include/rootfinder.h:37), where the conformal EoS is inverted in closed form here: in the dilute decades that is
            # ~1e-9 of a chemical potential, hence the absolute floor
            assert np.max(np.abs(v[k][ok] - th[ok][:, c])) <= 1e-9 * sc + 5e-10, (eos, ex, k)


def _all_tiles(tiles):
    have = {tuple(t["exponent"]): t for t in tiles}
    out = []
    for a in range(-3, 2):
        for b in range(-3, 1):
            for c in range(-3, 1):
                for d in range(-3, 1):
                    out.append(have.get((a, b, c, d), dict(tiles[0], exponent=(a, b, c, d))))
    return out


def test_tiled_lookup_through_the_c_abi():
    """ccb_set_preinverted_table + ccb_update_thermodynamics against EoS_Interpolator::fill_thermodynamics on the same tiles
    (generated on the device)."""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    from ccake_b200 import SPHWorkstation, eos_preinverted as P
    from oracle import ref

    gen = SPHWorkstation(charged_settings())
    exps = [(0, -1, -2, -1), (0, -1, -1, -1), (1, -1, -2, -1), (0, 0, -2, -1), (-3, -3, -3, -3)]
    tiles = P.generate_table(gen, exps, N=8)
    path = os.path.join(tempfile.mkdtemp(prefix="ccb_pre_"), "tiles.h5")
    ref.h5_put_tiles(path, _all_tiles(tiles))
    rng = np.random.default_rng(8)
    dens = []
    for ex in exps:
        lo = np.array([10.0 ** e for e in ex])
        x = lo + (8.7 * lo - lo) * rng.uniform(0, 1, (400, 4))   # below the last node (8.75 lo at N = 8)
        x[:, 1:] *= rng.choice([-1.0, 1.0], (400, 3))
        dens.append(x)
    dens = np.concatenate(dens)
    n = dens.shape[0]
    ws = SPHWorkstation(charged_settings(online_inverter_enabled=False, preinverted_tiles=tiles))
    ws.allocate(n)
    th = np.zeros((n, 16)); th[:, 5:9] = dens
    ws.set("thermo", th)
    ws.update_all_particle_thermodynamics()
    ws.synchronize()
    got = ws.get("thermo")
    want = ref.interp_fill(path, dens)
    for c in (0, 1, 2, 3, 4, 9, 10, 11, 12, 13, 14, 15):
        assert np.max(np.abs(got[:, c] - want[:, c])) <= 1e-12 * max(np.max(np.abs(want[:, c])), 1.0), c
    assert np.array_equal(got[:, 5:9], dens)   # s and the charge densities stay what the field gather stored


def test_whole_run_on_the_flat_preinverted_table():
    """online_inverter_enabled = false end to end: initialisation through invert_e_to_s_host, efcheck through efreeze_at_T, phase 04
    through fill_thermodynamics, two RK2 steps -- against the reference running its own EoS_Interpolator on the same flat table."""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    from ccake_b200 import SPHWorkstation, eos_preinverted as P
    from oracle import ref

    flat = P.conformal_flat_table()
    path = os.path.join(tempfile.mkdtemp(prefix="ccb_pre_"), "flat.h5")
    ref.h5_put_flat(path, flat)
    base = viscous_settings(online_inverter_enabled=False, baryon_charge_enabled=True)
    ic = gubser_ic(dx=0.1, shear=True, charges=True)
    ic["rhoS"], ic["rhoQ"] = None, None
    sim = make_ref(dict(base, preinverted_eos_path=path), ic)
    sim.initialize()
    ws = SPHWorkstation(dict(base, preinverted_flat=flat))
    ws.upload_initial_conditions(ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"], ic["u"], ic["bulk"], ic["shv6"])
    n = sim.n
    th, th_ref = ws.get("thermo"), sim.member("thermo")
    for c in (0, 1, 4, 5, 6, 9, 10, 11, 12, 13):
        _cmp("flat init", "thermo%d" % c, th[:, c], th_ref[:, c], scale=max(np.max(np.abs(th_ref[:, c])), 1e-12))
    assert np.array_equal(ws.get_int("freeze"), sim.ints("freeze"))
    _cmp("flat init", "extensive", ws.get("extensive"), sim.member("extensive")[:, [0, 2, 3, 4]])
    for step in range(2):
        sim.advance(base["dt"], 2)
        ws.advance_timestep(base["dt"])
        ws.synchronize()
        tol = TOL * (step + 1)
        _cmp("flat run", "position", ws.get("position"), sim.member("position"), tol)
        _cmp("flat run", "u", ws.get("u"), sim.get("hydro_vector", "u"), tol)
        _cmp("flat run", "extensive", ws.get("extensive"), sim.member("extensive")[:, [0, 2, 3, 4]], tol)
        _cmp("flat run", "extensive_shv", ws.get("extensive_shv"), _ext_shv5(sim), tol, scale=max(np.max(np.abs(_ext_shv5(sim))), 1e-6))
    th, th_ref = ws.get("thermo"), sim.member("thermo")
    for c in (0, 1, 4, 9, 10, 11, 12, 13):
        _cmp("flat run", "thermo%d" % c, th[:, c], th_ref[:, c], 3 * TOL, scale=max(np.max(np.abs(th_ref[:, c])), 1e-12))
