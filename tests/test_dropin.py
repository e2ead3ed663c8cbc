"""The drop-in boundary: the reference's UNMODIFIED src/BSQHydro.cpp -- constructor, its own text readers,
initialize_hydrodynamics and the run loop with conservation / evolution / freeze-out output -- compiled against the headers of
include/ccake_dropin/ (tests/cpp/dropin_bsqhydro.cpp -> tests/libdropin_bsqhydro.so, built where /root/reference exists) and run on
the GPU, against the reference build of the same classes (oracle/_ref) on identical Settings and the identical IC file."""
import ctypes as C
import os

import numpy as np
import pytest

import common
from common import blob_ic_3d, charged_settings, gubser_ic, have_ref, make_ref, viscous_settings
from ccake_b200 import workloads as W

TOL = 1e-10
ROW = 24
COLS = {"r": (0, 3), "u": (3, 6), "T": (6, 7), "e": (7, 8), "p": (8, 9), "s": (9, 10), "rho": (10, 13), "bulk": (13, 14), "shv": (14, 19),
        "sigma": (19, 20), "gamma": (20, 21), "extensive_s": (22, 23), "smoothed_s": (23, 24)}


def _lib():
    if not os.path.exists(common.DROPIN_LIB):
        pytest.skip("tests/libdropin_bsqhydro.so not built (needs /root/reference at build time)")
    L = C.CDLL(common.DROPIN_LIB)
    L.dropin_bsqhydro_run.argtypes = [C.c_char_p, C.c_long, C.POINTER(C.c_long), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int),
                                      C.c_char_p, C.c_int]
    return L


def test_dropin_library_is_built_from_the_reference_translation_unit():
    """CPU: the library loads, and its symbols are the reference's own BSQHydro<D,TEOM> instantiations (src/BSQHydro.cpp:9-14)."""
    L = _lib()
    assert L.dropin_row_width() == ROW
    import subprocess

    syms = subprocess.run(["nm", "-DC", "--defined-only", common.DROPIN_LIB], capture_output=True, text=True).stdout
    for inst in ("ccake::BSQHydro<2u, ccake::EoM_default>::run()", "ccake::BSQHydro<3u, ccake::EoM_cartesian>::run()",
                 "ccake::BSQHydro<2u, ccake::EoM_default>::read_ccake()", "ccake::BSQHydro<2u, ccake::EoM_default>::read_ICCING()"):
        assert inst in syms, inst


def _run_dropin(settings, ic, tmp_path, **extra):
    L = _lib()
    s = dict(settings)
    s.update(extra)
    path = tmp_path / "ic.dat"
    W.write_ccake(str(path), s, ic)
    out = tmp_path / "results"
    out.mkdir()
    s.update(IC_type="ccake", IC_file=str(path), results_directory=str(out))
    keys = common.ref_module().SETTING_KEYS | {"IC_type", "IC_file", "results_directory", "txt_evolution", "evolution_stride", "get_neighbors",
                                                 "input_initial_diffusion"}
    kv = "\n".join("%s=%s" % (k, ("true" if v else "false") if isinstance(v, bool) else repr(v) if isinstance(v, float) else v)
                   for k, v in s.items() if k in keys or k.startswith("kappa_")).encode()
    n = ic["r"].shape[0]
    rows = np.zeros((n, ROW))
    n_out, t_out, steps = C.c_long(), C.c_double(), C.c_int()
    err = C.create_string_buffer(512)
    rc = L.dropin_bsqhydro_run(kv, n, C.byref(n_out), C.byref(t_out), rows.ctypes.data_as(C.POINTER(C.c_double)), C.byref(steps), err, 512)
    assert rc == 0, err.value.decode()
    return rows[: n_out.value], t_out.value, steps.value, out


def _ref_rows(sim):
    th = sim.member("thermo")
    shv = sim.get("hydro_spacetime_matrix", "shv")
    r = np.zeros((sim.n, 3)); u = np.zeros((sim.n, 3))
    pos, uu = sim.member("position"), sim.get("hydro_vector", "u")
    D = int(sim.settings.get("dim", 2))
    r[:, :D], u[:, :D] = pos[:, :D], uu[:, :D]
    cols = [r, u, th[:, 0:1], th[:, 4:5], th[:, 9:10], th[:, 5:6], th[:, 6:9], sim.get("hydro_scalar", "bulk")[:, None],
            np.stack([shv[:, 0, 0], shv[:, 1, 1], shv[:, 2, 2], shv[:, 1, 2], shv[:, 3, 3]], 1), sim.get("hydro_scalar", "sigma")[:, None],
            sim.get("hydro_scalar", "gamma")[:, None], sim.ints("freeze")[:, None].astype(float), sim.member("extensive")[:, 0:1],
            sim.member("smoothed")[:, 0:1]]
    return np.concatenate(cols, 1)


def _cases():
    D3 = dict(dim=3, hT=0.5, hEta=0.5, stepx=0.25, stepy=0.25, stepEta=0.25, xmin=-1.375, ymin=-1.375, etamin=-1.375)
    return {
        "viscous_2d_milne": (viscous_settings(dt=0.05, max_tau=1.12), gubser_ic(dx=0.1, shear=True)),
        "charged_2d_diffusion": (charged_settings(dt=0.05, max_tau=1.07), gubser_ic(dx=0.1, shear=True, charges=True)),
        "viscous_2d_rk4": (viscous_settings(dt=0.05, max_tau=1.07, rk_order=4), gubser_ic(dx=0.1, shear=True)),   # Evolver::advance_timestep_rk4
        "viscous_3d_cartesian": (viscous_settings(dt=0.05, max_tau=1.07, coordinate_system="cartesian", baryon_charge_enabled=True,
                                                  strange_charge_enabled=True, electric_charge_enabled=True, **D3), blob_ic_3d(n_side=12)),
    }


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(_cases()))
def test_unmodified_bsqhydro_runs_on_the_dropin_headers(case, tmp_path):
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    settings, ic = _cases()[case]
    rows, t, steps, out = _run_dropin(settings, ic, tmp_path, txt_evolution=True, print_conservation_status=True)
    sim = make_ref(settings, ic)
    sim.initialize()
    k = 0
    while sim.t <= settings["max_tau"]:   # SPHWorkstation::continue_evolution, src/sph_workstation.cpp:1993-2024
        sim.advance(settings["dt"], int(settings.get("rk_order", 2)))
        k += 1
    assert steps == k and abs(t - sim.t) < 1e-12 and rows.shape[0] == sim.n
    ref = _ref_rows(sim)
    assert np.array_equal(rows[:, 21], ref[:, 21]), "Freeze"
    pmax = float(np.max(np.abs(ref[:, 8])))
    for name, (lo, hi) in COLS.items():
        scale = pmax if name in ("bulk", "shv") else None   # Pi and pi^{mu nu} on the scale of the pressure
        assert common.rel_err(rows[:, lo:hi], ref[:, lo:hi], scale) <= TOL * (k + 1) * (4 if settings.get("rk_order", 2) == 4 else 1), (case, name)
    # the run loop's own output (src/BSQHydro.cpp:680-700, 770-797, 860-866): one evolution file per step + the initial one,
    # conservation.dat with a header and k + 1 lines
    assert sorted(os.listdir(out)) == sorted(["conservation.dat"] + ["system_state_%d.dat" % i for i in range(k + 1)])
    cons = open(out / "conservation.dat").read().splitlines()
    assert cons[0].split() == ["t", "Eloss", "S", "Btotal", "Stotal", "Qtotal"] and len(cons) == k + 2
    S0, S1 = float(cons[1].split()[2]), float(cons[-1].split()[2])
    assert abs(S1 / S0 - 1.0) < 0.05   # entropy production of a few viscous steps is small
    last = open(out / ("system_state_%d.dat" % k)).read().splitlines()
    assert abs(float(last[0]) - sim.t) < 1e-6 and len(last) == sim.n + 1
    col = np.array([[float(x) for x in ln.split()[:11]] for ln in last[1:]])   # id t x y eta p T muB muS muQ e  (src/output.cpp:86-167)
    hc = 0.1973269804
    D = int(settings.get("dim", 2))
    assert np.allclose(col[:, 2:2 + D], ref[:, 0:D], atol=2e-6) and np.allclose(col[:, 5], ref[:, 8] * hc, rtol=2e-6) and np.allclose(col[:, 6], ref[:, 6] * hc, rtol=2e-6)


@pytest.mark.gpu
def test_unmodified_bsqhydro_with_particlization_writes_the_hypersurface(tmp_path):
    """particlization.enabled: the run loop ends when every particle is frozen out (continue_evolution), print_freeze_out buffers the
    hypersurface elements of every step and flush_freeze_out writes them; the element count equals the reference's."""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    settings = viscous_settings(particlization_enabled=True, Freeze_Out_Temperature=300.0, dt=0.05, max_tau=1.72)
    ic = gubser_ic(dx=0.1, shear=True)
    rows, t, steps, out = _run_dropin(settings, ic, tmp_path)
    sim = make_ref(settings, ic)
    sim.initialize()
    k, printed = 0, 0
    while sim.t <= settings["max_tau"] and int((sim.ints("freeze") == 4).sum()) < sim.n:
        sim.advance(settings["dt"], 2)
        _, pr = sim.freeze_results()
        printed += int(pr.sum())
        k += 1
    assert steps == k and printed > 20
    assert np.array_equal(rows[:, 21], sim.ints("freeze").astype(float))
    fo = open(out / "freeze_out.dat").read().splitlines()
    assert fo[0].startswith("#") and len(fo) - 1 == printed


@pytest.mark.gpu
def test_unmodified_bsqhydro_on_a_preinverted_table(tmp_path):
    """eos.online_inverter_enabled: false (the reference's default): SPHWorkstation::initialize constructs EoS_Interpolator from
    Settings::preinverted_eos_path -- the drop-in one reads the CCBPRE01 file eos_preinverted.save_ccbpre writes (no HDF5 here) -- and the
    run initialises and evolves on the flat (s, n_B) table; against the reference running its own EoS_Interpolator on the same table."""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    from ccake_b200 import eos_preinverted as P
    from oracle import ref

    flat = P.conformal_flat_table()
    P.save_ccbpre(str(tmp_path / "flat.ccbpre"), flat=flat)
    h5 = str(tmp_path / "flat.h5")
    ref.h5_put_flat(h5, flat)
    settings = viscous_settings(online_inverter_enabled=False, baryon_charge_enabled=True, dt=0.05, max_tau=1.07)
    ic = gubser_ic(dx=0.1, shear=True, charges=True)
    ic["rhoS"], ic["rhoQ"] = None, None
    rows, t, steps, out = _run_dropin(settings, ic, tmp_path, preinverted_eos_path=str(tmp_path / "flat.ccbpre"))
    sim = make_ref(dict(settings, preinverted_eos_path=h5), ic)
    sim.initialize()
    k = 0
    while sim.t <= settings["max_tau"]:
        sim.advance(settings["dt"], 2)
        k += 1
    assert steps == k == 2 and rows.shape[0] == sim.n
    ref_rows = _ref_rows(sim)
    assert np.array_equal(rows[:, 21], ref_rows[:, 21])
    pmax = float(np.max(np.abs(ref_rows[:, 8])))
    for name in ("r", "u", "T", "e", "p", "s", "bulk", "shv", "sigma", "gamma", "extensive_s"):
        lo, hi = COLS[name]
        scale = pmax if name in ("bulk", "shv") else None   # Pi and pi^{mu nu} on the scale of the pressure (Pi is ~1e-12 here)
        assert common.rel_err(rows[:, lo:hi], ref_rows[:, lo:hi], scale) <= 3 * TOL * (k + 1), name


@pytest.mark.gpu
def test_unmodified_bsqhydro_freezein_initial_conditions(tmp_path):
    """IC_type "freezein": the reference's own read_freezein (unmodified BSQHydro.cpp) stages the raw T^{tau mu}, j_a^mu of every cell,
    the drop-in SPHWorkstation::freeze_in_match closes them with the EoS in batched device queries, and the run proceeds; against the
    reference's freeze_in_match (one host root search per cell and iterate) fed the same cells, and against the fluid the file was
    made from (conformal EoS: P = eps / 3, so T^{tau tau} = (4 gamma^2 - 1) eps / 3, T^{tau i} = 4 eps gamma u^i / 3, j^tau = n gamma)."""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    hc = 0.1973269804
    ic = gubser_ic(dx=0.1, shear=False, charges=True)
    n = ic["r"].shape[0]
    eps, u = ic["e"], ic["u"]                       # GeV/fm^3, u^x u^y
    gamma = np.sqrt(1.0 + np.sum(u * u, 1))
    Ttau = np.zeros((n, 4)); Ttau[:, 0] = (4.0 * gamma ** 2 - 1.0) * eps / 3.0; Ttau[:, 1:3] = (4.0 * eps * gamma / 3.0)[:, None] * u
    jB, jS, jQ = (np.zeros((n, 4)) for _ in range(3))
    jB[:, 0], jS[:, 0], jQ[:, 0] = ic["rhoB"] * gamma, ic["rhoS"] * gamma, ic["rhoQ"] * gamma
    path = tmp_path / "freezein.dat"
    with open(path, "w") as f:
        f.write("#0 0.1 0.1 0.1 0 -2.0 -2.0 -0.1\n# raw (un-diagonalized) contravariant T^{mu nu}\n")
        rows = np.concatenate([ic["r"], np.zeros((n, 1)), Ttau, np.zeros((n, 6)), jB, jS, jQ], 1)
        np.savetxt(f, rows, fmt="%.17g")
    settings = charged_settings(dt=0.05, max_tau=1.07, IC_type="freezein", e_cutoff=0.05)
    L = _lib()
    out = tmp_path / "results"; out.mkdir()
    s = dict(settings, freezein_input_file=str(path), results_directory=str(out))
    keys = common.ref_module().SETTING_KEYS | {"results_directory", "freezein_input_file"}
    kv = "\n".join("%s=%s" % (k, ("true" if v else "false") if isinstance(v, bool) else repr(v) if isinstance(v, float) else v)
                   for k, v in s.items() if k in keys or k.startswith("kappa_")).encode()
    rows_out = np.zeros((n, ROW))
    n_out, t_out, steps = C.c_long(), C.c_double(), C.c_int()
    err = C.create_string_buffer(512)
    rc = L.dropin_bsqhydro_run(kv, n, C.byref(n_out), C.byref(t_out), rows_out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(steps), err, 512)
    assert rc == 0, err.value.decode()
    rows_out = rows_out[: n_out.value]
    sim = common.ref_module().RefSim({k: v for k, v in settings.items() if k in common.ref_module().SETTING_KEYS or k.startswith("kappa_")})
    sim.add_freezein(ic["r"], Ttau, jB, jS, jQ)
    sim.initialize()
    assert sim.n == n_out.value and 0 < sim.n <= n
    # every cell of this file is above the cut-off in its rest frame: the matching keeps them all, on both sides
    assert sim.n == int(np.sum(eps > settings["e_cutoff"])) == n
    k = 0
    while sim.t <= settings["max_tau"]:
        sim.advance(settings["dt"], 2)
        k += 1
    assert steps.value == k
    ref_rows = _ref_rows(sim)
    assert np.array_equal(rows_out[:, 21], ref_rows[:, 21])
    pmax = float(np.max(np.abs(ref_rows[:, 8])))
    for name in ("r", "u", "T", "e", "p", "s", "rho", "sigma", "gamma", "extensive_s"):
        lo, hi = COLS[name]
        # the two sides stop the v-iteration at the same iterate, but their EoS roots differ within the 1e-10 root tolerance
        assert common.rel_err(rows_out[:, lo:hi], ref_rows[:, lo:hi]) <= 1e-8, name
