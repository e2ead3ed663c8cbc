"""CPU checks of the library's per-particle physics headers (compiled for the host
by tests/host_harness.cpp) against the reference build oracle/_ref: EoS + thermodynamic
derivatives, pi^{mu nu} rebuild, transport coefficients and the equations of motion,
fed with the reference's own intermediate state after one get_time_derivatives().

Tolerance: 1e-10 field-scale-relative per derivative evaluation (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest

from common import (EomArrays, blob_ic_3d, charged_settings, dp, gubser_ic, harness, have_ref, ideal_settings, make_ref,
                    ref_dalpha12, ref_thermo16, rel_err, viscous_settings)

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
TOL = 1e-10


def _config(settings):
    from ccake_b200 import to_config

    return to_config(settings)


def _run_case(settings, ic):
    sim = make_ref(settings, ic)
    sim.initialize()
    sim.advance(settings["dt"], 2)  # a generic state: non-zero bulk, evolved pi^{ij} and q^i
    sim.get_time_derivatives()
    return sim


CASES = {
    "ideal_2d": (lambda: ideal_settings(), lambda: gubser_ic(dx=0.2)),
    "viscous_2d": (lambda: viscous_settings(), lambda: gubser_ic(dx=0.2, shear=True)),
    "viscous_2d_vorticity": (lambda: viscous_settings(use_vorticity=True, zetaMode="cs2_dependent"), lambda: gubser_ic(dx=0.2, shear=True)),
    "charged_2d_diffusion": (lambda: charged_settings(), lambda: gubser_ic(dx=0.2, shear=True, charges=True)),
    "viscous_3d_milne": (lambda: viscous_settings(dim=3, hT=0.5, hEta=0.5, stepx=0.25, stepy=0.25, stepEta=0.25, xmin=-1.5, ymin=-1.5, etamin=-1.5),
                         lambda: blob_ic_3d()),
    "charged_3d_milne_diffusion": (lambda: charged_settings(dim=3, hT=0.5, hEta=0.5, stepx=0.25, stepy=0.25, stepEta=0.25, xmin=-1.5,
                                                            ymin=-1.5, etamin=-1.5, use_vorticity=True), lambda: blob_ic_3d()),
    "viscous_3d_cartesian": (lambda: viscous_settings(dim=3, coordinate_system="cartesian", hT=0.5, hEta=0.5, stepx=0.25, stepy=0.25,
                                                     stepEta=0.25, xmin=-1.5, ymin=-1.5, etamin=-1.5),
                             lambda: blob_ic_3d()),
    "charged_3d_cartesian_diffusion": (lambda: charged_settings(dim=3, coordinate_system="cartesian", hT=0.5, hEta=0.5, stepx=0.25,
                                                                stepy=0.25, stepEta=0.25, xmin=-1.5, ymin=-1.5, etamin=-1.5),
                                       lambda: blob_ic_3d()),
    "ideal_2d_cartesian": (lambda: ideal_settings(coordinate_system="cartesian"), lambda: gubser_ic(dx=0.2)),
}


@pytest.fixture(scope="module", params=sorted(CASES))
def case(request):
    mk_s, mk_ic = CASES[request.param]
    settings = mk_s()
    sim = _run_case(settings, mk_ic())
    return request.param, settings, sim


def test_eos_and_derivatives(case):
    name, settings, sim = case
    H = harness()
    cfg = _config(settings)
    th_ref = ref_thermo16(sim)
    n = sim.n
    sBSQ = np.ascontiguousarray(th_ref[:, 5:9])
    th = np.zeros((n, 16))
    da = np.zeros((n, 12))
    assert H.hh_eos(C.byref(cfg), C.c_int64(n), dp(sBSQ), dp(th), dp(da)) == 0
    names = "T muB muS muQ e s rhoB rhoS rhoQ p cs2 w dwds dwdB dwdS dwdQ".split()
    for k, nm in enumerate(names):
        sc = max(np.max(np.abs(th_ref[:, k])), 1e-12)
        assert rel_err(th[:, k], th_ref[:, k], sc) < TOL, (name, nm)
    da_ref = ref_dalpha12(sim)
    if np.any(np.abs(th_ref[:, 6:9]) > 0):
        for k in range(12):
            sc = max(np.max(np.abs(da_ref[:, k])), 1e-12)
            assert rel_err(da[:, k], da_ref[:, k], sc) < 1e-8, (name, "dalpha", k)


def test_local_update(case):
    name, settings, sim = case
    H = harness()
    cfg = _config(settings)
    n = sim.n
    u = np.ascontiguousarray(sim.get("hydro_vector", "u"))
    gamma = np.ascontiguousarray(sim.get("hydro_scalar", "gamma"))
    sigma = np.ascontiguousarray(sim.get("hydro_scalar", "sigma"))
    es = sim.get("hydro_shear_aux_vector", "extensive_shv")  # (n,2,3)
    ext_shv = np.ascontiguousarray(np.stack([es[:, 0, 0], es[:, 0, 1], es[:, 0, 2], es[:, 1, 1], es[:, 1, 2]], 1))
    ext_q = np.ascontiguousarray(sim.get("hydro_space_matrix", "extensive_diffusion").reshape(n, 9))
    th = np.ascontiguousarray(ref_thermo16(sim))
    shv = np.zeros((n, 16)); q = np.zeros((n, 12)); tr = np.zeros((n, 13))
    # the reference evaluates these phases at the time of the derivative evaluation
    t = sim.t
    H.hh_local(C.byref(cfg), C.c_double(t), C.c_int64(n), dp(u), dp(gamma), dp(sigma), dp(ext_shv), dp(ext_q), dp(th), dp(shv), dp(q), dp(tr))
    shv_ref = sim.get("hydro_spacetime_matrix", "shv").reshape(n, 16)
    sc = max(np.max(np.abs(shv_ref)), 1e-12)
    assert rel_err(shv, shv_ref, sc) < TOL, name
    q_ref = sim.member("hydro_diffusion").reshape(n, 12)
    assert rel_err(q, q_ref, max(np.max(np.abs(q_ref)), 1e-12)) < TOL, name
    hs = sim.member("hydro_scalar")
    from oracle.ref import HYDRO_SCALAR as HS
    order = ["eta_pi", "tau_pi", "zeta_Pi", "tau_Pi", "delta_PiPi", "delta_pipi", "lambda_piPi", "lambda_Pipi", "phi1", "phi3", "phi6", "phi7", "tau_pipi"]
    for k, nm in enumerate(order):
        r = hs[:, HS.index(nm)]
        assert rel_err(tr[:, k], r, max(np.max(np.abs(r)), 1e-12)) < TOL, (name, nm)


def test_equations_of_motion(case):
    name, settings, sim = case
    H = harness()
    cfg = _config(settings)
    n = sim.n
    g = lambda m, f: np.ascontiguousarray(sim.get(m, f))
    A = EomArrays()
    keep = {}

    def put(k, a):
        keep[k] = np.ascontiguousarray(a, dtype=np.float64)
        setattr(A, k, dp(keep[k]))

    put("u", g("hydro_vector", "u")); put("v", g("hydro_vector", "v"))
    put("gamma", g("hydro_scalar", "gamma")); put("sigma", g("hydro_scalar", "sigma")); put("sigma_lab", g("hydro_scalar", "sigma_lab"))
    put("mass_s", sim.get("sph_mass", "s")); put("bulk", g("hydro_scalar", "bulk"))
    put("shv", sim.get("hydro_spacetime_matrix", "shv").reshape(n, 16)); put("q", sim.member("hydro_diffusion").reshape(n, 12))
    put("thermo", ref_thermo16(sim)); put("dalpha", ref_dalpha12(sim))
    for f in ("gradP", "gradBulk", "gradshear", "divshear", "div_qa", "grad_qa"):
        put(f, g("hydro_vector", f))
    put("gradV", sim.get("hydro_space_matrix", "gradV").reshape(n, 9))
    put("grad_alpha", sim.get("hydro_space_matrix", "grad_alpha_a").reshape(n, 9))
    for k, w in (("du_dt", 3), ("d_ext", 4), ("d_bulk", 1), ("d_shv", 5), ("d_q", 9), ("theta", 1), ("shv_nabla_u", 1), ("dEz", 1)):
        put(k, np.zeros((n, w)))
    H.hh_eom(C.byref(cfg), C.c_double(sim.t), C.c_int64(n), C.byref(A))

    def check(got, ref, what, tol=TOL):
        sc = max(np.max(np.abs(ref)), 1e-12)
        e = rel_err(got, ref, sc)
        assert e < tol, (name, what, e)

    check(keep["du_dt"], sim.get("hydro_vector", "du_dt"), "du_dt")
    dd = sim.member("d_dt_extensive")
    check(keep["d_ext"][:, 0], dd[:, 0], "d_ext_s")
    for c, k in enumerate((2, 3, 4)):
        check(keep["d_ext"][:, 1 + c], dd[:, k], "dN%d" % c)
    check(keep["d_bulk"][:, 0], sim.get("hydro_scalar", "d_extensive_bulk_dt"), "dbulk")
    check(keep["theta"][:, 0], sim.get("hydro_scalar", "theta"), "theta")
    check(keep["dEz"][:, 0], sim.member("contribution_to_total_dEz"), "dEz")
    from ccake_b200 import derived_switches
    sw = derived_switches(cfg_settings(settings))
    if sw["using_shear"]:
        ds = sim.get("hydro_shear_aux_vector", "d_extensive_shv_dt")
        ref5 = np.stack([ds[:, 0, 0], ds[:, 0, 1], ds[:, 0, 2], ds[:, 1, 1], ds[:, 1, 2]], 1)
        check(keep["d_shv"], ref5, "dshv")
        check(keep["shv_nabla_u"][:, 0], sim.get("hydro_scalar", "shv_nabla_u"), "shv_nabla_u")
    if sw["using_diffusion"]:
        check(keep["d_q"], sim.get("hydro_space_matrix", "d_extensive_q_dt").reshape(n, 9), "dq")


def cfg_settings(settings):
    from ccake_b200 import DEFAULTS

    s = dict(DEFAULTS)
    s.update(settings)
    return s
