"""CPU tests (no GPU): the oracle against every still-valid known answer of the
reference's own unit tests, the C-ABI library's exports, and the
host-side settings logic."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from common import ROOT, charged_settings, gubser_ic, harness, have_ref, ideal_settings, make_ref, viscous_settings

needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
PI = math.pi


# ---- kernel known answers, unit_tests/kernel_test.cpp:7-322 ---------------------------
KERNEL_GOLD = [
    (1, 0.5, 23. / 48.), (1, 1.0, 1. / 6.), (1, 1.5, 1. / 48.), (1, 2.0, 0.0),
    (2, 0.5, 115. / (112. * PI)), (2, 1.0, 5. / (14. * PI)), (2, 1.5, 5. / (112. * PI)), (2, 2.0, 0.0),
    (3, 0.5, 23. / (32. * PI)), (3, 1.0, 1. / (4. * PI)), (3, 1.5, 1. / (32. * PI)), (3, 2.0, 0.0),
]


@pytest.mark.parametrize("D,q,val", KERNEL_GOLD)
def test_kernel_values(D, q, val):
    H = harness()
    H.hh_kernel.restype = C.c_double
    assert abs(H.hh_kernel(C.c_int(D), C.c_double(q), C.c_double(1.0)) - val) < 1e-14


@pytest.mark.parametrize("h", [1.0, 0.5])
@pytest.mark.parametrize("q,val", [(0.5, -75. / (56. * PI)), (1.0, -15. / (14. * PI)), (1.5, -15. / (56. * PI))])
def test_grad_kernel_2d(h, q, val):
    """grad W along +x at distance q*h: magnitude val / h^3 (kernel_test.cpp:117-322)."""
    H = harness()
    H.hh_grad_kernel_factor.restype = C.c_double
    r = q * h
    g = H.hh_grad_kernel_factor(C.c_int(2), C.c_double(r), C.c_double(h)) * r
    assert abs(g - val / h ** 3) < 1e-13 / h ** 3


# ---- the reference build reproduces the reference's own known answers --------------------
@needs_ref
def test_ref_smoothing_sum():
    """unit_tests/smoothing_test.cpp:33-147."""
    from oracle import ref

    s = ref.RefSim(dict(dim=2, t0=1.0, hT=1.0, hEta=1.0, xmin=-4, ymin=-6, etamin=-2))
    s.allocate_raw(4)
    pos = np.zeros((4, 3)); pos[:, :2] = [[0, 0], [0, .5], [.8, .6], [1.5, 0]]
    s.set_member("position", pos)
    m = np.zeros((4, 5)); m[:, 0] = 1; m[:, 2] = 2; m[:, 4] = 3; m[:, 3] = 4
    s.set_member("sph_mass", m); s.set_member("extensive", m)
    s.put("hydro_scalar", "gamma", np.ones(4))
    s.phase("neighbour_list"); s.phase("smooth_fields")
    sm = s.member("smoothed")
    g = 20. / 7. / PI
    assert abs(sm[0, 0] - g) < 1e-14 and abs(sm[0, 2] - 4 * g) < 1e-13
    assert abs(sm[0, 4] - 9 * g) < 1e-13 and abs(sm[0, 3] - 16 * g) < 1e-13
    assert abs(s.get("hydro_scalar", "sigma_lab")[0] - g) < 1e-14


@needs_ref
def test_ref_gradP():
    """unit_tests/grad_smoothing_test.cpp:25-70 (with sigma_lab = 1)."""
    from oracle import ref

    s = ref.RefSim(dict(dim=2, t0=1.0, hT=1.0, hEta=1.0, xmin=-4, ymin=-6, etamin=-2))
    s.allocate_raw(5)
    pos = np.zeros((5, 3)); pos[:, 0] = [0, .5, 1.0, 1.5, -1.5]; pos[4, 1] = 1.5
    s.set_member("position", pos)
    m = np.ones((5, 5))
    s.set_member("sph_mass", m); s.set_member("extensive", m)
    s.put("hydro_scalar", "sigma_lab", np.ones(5))
    s.put("hydro_scalar", "gamma", np.ones(5))
    th = s.member("thermo"); th[:, ref.THERMO.index("p")] = [1, .5, 1.5, 2.5, 1]
    s.set_member("thermo", th)
    s.phase("neighbour_list"); s.phase("smooth_gradients")
    gp = s.get("hydro_vector", "gradP")
    assert abs(gp[0, 0] - 45. / 8. / PI) < 1e-13 and abs(gp[0, 1]) < 1e-15


@needs_ref
def test_ref_pi_reconstruction():
    """unit_tests/eom_default_test.cpp:185-210."""
    from oracle import ref

    s = ref.RefSim(dict(dim=2, t0=2.0, hT=1.0, hEta=1.0, xmin=-4, ymin=-6, etamin=-2))
    s.allocate_raw(1)
    hv = s.member("hydro_vector"); hv[0, ref.HYDRO_VECTOR.index("u"), :2] = [3, 4]
    s.set_member("hydro_vector", hv)
    s.put("hydro_scalar", "sigma", np.ones(1))
    es = s.member("hydro_shear_aux_vector")
    k = ref.HYDRO_SHEAR_AUX_VECTOR.index("extensive_shv")
    es[0, k, 0, 0], es[0, k, 0, 1], es[0, k, 1, 1] = 3, 4, 5
    s.set_member("hydro_shear_aux_vector", es)
    s.phase("gamma_velocity"); s.phase("reset_pi_tensor")
    shv = s.get("hydro_spacetime_matrix", "shv")[0]
    g = math.sqrt(26.)
    assert abs(s.get("hydro_scalar", "gamma")[0] - g) < 1e-14
    assert abs(shv[0, 1] - 25. / g) < 1e-13 and abs(shv[0, 2] - 32. / g) < 1e-13
    assert abs(shv[0, 0] - 203. / 26.) < 1e-13 and abs(shv[3, 3] + 5. / 104.) < 1e-14


@needs_ref
def test_ref_neighbour_list_against_brute_force():
    """The shim's linked-cell VerletList (Cabana semantics) against an O(N^2) search."""
    settings = ideal_settings()
    ic = gubser_ic(dx=0.2)
    sim = make_ref(settings, ic)
    sim.initialize()
    off, idx, cell, nc = sim.neighbours()
    pos = sim.member("position")
    d2 = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1)
    rs = (2 * settings["hT"]) ** 2
    for i in range(0, sim.n, 7):
        want = set(np.nonzero(d2[i] <= rs)[0]) - {i}
        assert set(idx[off[i]:off[i + 1]]) == want


@needs_ref
def test_ref_ideal_gubser_tracks_analytic_solution():
    """Physics-level pin (misc/Gubser_checks, utilities/ideal_gubser_ic_generator.py): after a short ideal
    evolution the energy density at the centre follows the closed-form Gubser solution."""
    from common import gubser_profile, HBARC

    settings = ideal_settings(dt=0.01, hT=0.3)
    ic = gubser_ic(dx=0.1, xmax=3.0)
    sim = make_ref(settings, ic)
    sim.initialize()
    pos0 = sim.member("position")
    e0_num = sim.get("thermo", "e")[np.argmin((pos0[:, :2] ** 2).sum(1))] * HBARC
    assert abs(e0_num / 80.0 - 1) < 5e-2
    for _ in range(10):
        sim.advance(settings["dt"], 2)
    pos = sim.member("position")
    c = np.argmin((pos[:, :2] ** 2).sum(1))
    e_num = sim.get("thermo", "e")[c] * HBARC
    e_exact = gubser_profile(np.array([0.0]), np.array([0.0]), tau0=sim.t)[0][0]
    # the decay of the central energy density over the evolved interval (SPH smoothing offsets the
    # absolute value by O(h^2); it converges with resolution: 3.5% at dx=0.1, 0.9% at dx=0.05)
    assert abs((e_num / e0_num) / (e_exact / 80.0) - 1) < 3e-2


# ---- the C-ABI library ---------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    from ccake_b200 import _capi

    assert os.path.exists(_capi.LIB_PATH), "run __graft_entry__.build() first"
    L = C.CDLL(_capi.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "ccake_b200.h")).read()
    import re

    declared = sorted(set(re.findall(r"\b(ccb_[a-z_]+)\s*\(", header)))
    assert set(declared) == set(_capi.EXPORTED), set(declared) ^ set(_capi.EXPORTED)
    for name in declared:
        assert hasattr(L, name), name


def test_field_widths_match_binding():
    from ccake_b200 import _capi

    L = _capi.lib()
    for name, (fid, w) in _capi.FIELDS.items():
        assert L.ccb_field_width(fid) == w, name
    assert L.ccb_field_width(len(_capi.FIELDS)) == -1


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from ccake_b200 import CCAKEError, SPHWorkstation, _capi

    with pytest.raises(CCAKEError) as ei:
        SPHWorkstation(ideal_settings())
    assert ei.value.code == _capi.CCB_ERR_CUDA


# ---- host logic -----------------------------------------------------------------------------
def test_derived_switches_follow_input_cpp():
    from ccake_b200 import DEFAULTS, derived_switches

    s = dict(DEFAULTS)
    s.update(etaMode="constant", constant_eta_over_s=0.0)
    assert derived_switches(s)["using_shear"] is False
    s.update(constant_eta_over_s=1e-3)
    assert derived_switches(s)["using_shear"] is False  # `<= 1.E-3` disables (src/input.cpp:404)
    s.update(constant_eta_over_s=0.08)
    assert derived_switches(s)["using_shear"] is True
    s.update(etaMode="disabled")
    assert derived_switches(s)["using_shear"] is False
    s.update(diffusionMode="constant_over_T2")
    assert derived_switches(s)["using_diffusion"] is True


def test_yaml_reader_on_reference_inputs(tmp_path):
    from ccake_b200 import from_yaml, to_config

    y = tmp_path / "visc.yaml"
    y.write_text("""
initial_conditions: {type: ccake, file: x.dat, t0: 1.0, dimension: 2, input_as_entropy: false}
parameters: {dt: 0.001, h_T: 0.1, h_eta: 0.1, rk_order: 2}
eos: {online_inverter_enabled: true, type: conformal}
particlization: {enabled: true, T: 153.334025}
hydro:
  baryon_charge_enabled: true
  viscous_parameters:
    diffusion: {mode: constant_over_T2, constant_kappa_over_T2: [[0.1, 0, 0], [0, 0.2, 0], [0, 0, 0.3]], relaxation_mode: constant_over_T}
    bulk: {mode: constant, constant_zeta_over_s: 0.0, delta_PiPi_mode: israel-stewart, lambda_Pipi_mode: default, phi1_mode: disabled}
    shear: {input_initial_shear: true, constant_eta_over_s: 0.2, mode: constant, relaxation_mode: Gubser, delta_pipi_mode: default, tau_pipi_mode: disabled}
""")
    s = from_yaml(str(y))
    assert s["hT"] == 0.1 and s["etaMode"] == "constant" and s["shearRelaxMode"] == "Gubser"
    c = to_config(s)
    assert c.using_shear == 1 and c.using_diffusion == 1 and c.shear_relax_mode == 2
    assert c.delta_PiPi_mode == 1 and c.tau_pipi_mode == 1 and c.phi1_mode == 1
    assert abs(c.constant_kappa_over_T2[4] - 0.2) < 1e-15
    assert abs(c.conformal_c - math.pi ** 2 * (16 + 26.25) / 90.0) < 1e-14


def test_ic_text_format_round_trip(tmp_path):
    """ccake_b200.workloads.read_ccake parses the reference's IC text format (BSQHydro::read_ccake, src/BSQHydro.cpp:207-425):
    header -> settings, unit handling left to the library, the energy cut-off applied row by row, optional q^mu_a columns."""
    from ccake_b200 import workloads as W

    ic = gubser_ic(dx=0.25, shear=True, charges=True)
    n = ic["r"].shape[0]
    ic["diffusion"] = np.random.default_rng(1).standard_normal((n, 3, 4))
    s = charged_settings(stepx=0.25, stepy=0.25, e_cutoff=0.0, input_initial_diffusion=True)
    path = str(tmp_path / "ic.dat")
    W.write_ccake(path, s, ic)
    s2, ic2 = W.read_ccake(path, dict(s, stepx=0.0, xmin=0.0))
    assert (s2["stepx"], s2["stepy"], s2["xmin"], s2["ymin"]) == (0.25, 0.25, -2.0, -2.0)
    for k in ("r", "e", "rhoB", "rhoS", "rhoQ", "u", "shv6", "diffusion"):
        assert np.array_equal(np.asarray(ic2[k]).reshape(n, -1), np.asarray(ic[k]).reshape(n, -1)), k
    cut = float(np.median(ic["e"]))
    _, ic3 = W.read_ccake(path, dict(s, e_cutoff=cut))
    assert ic3["e"].size == int((ic["e"] / 0.1973269804 > cut / 0.1973269804).sum()) and np.all(ic3["e"] > cut)
    # the sample header of the reference's own documentation parses too
    with open(path, "w") as f:
        f.write("#0 .025 .025 .1 0 -5 -5 -.1\n0 0 0 10 0 0 0 0 0 0 0 0 0 0 0 0 0\n")
    s4, ic4 = W.read_ccake(path, ideal_settings(e_cutoff=0.0))
    assert s4["stepEta"] == 0.1 and s4["etamin"] == -0.1 and ic4["e"][0] == 10.0


def test_yaml_reader_diffusion_block(tmp_path):
    """Key paths of the diffusion block of src/input.cpp:596-700 (regulator, critical-point suppression, C_B, kappa matrix) and the
    defaults of include/defaults.h:66-72, 119-123 when a key is missing."""
    from ccake_b200 import settings as S

    y = tmp_path / "in.yaml"
    y.write_text("""
initial_conditions: {dimension: 3, t0: 0.6, coordinate_system: hyperbolic}
parameters: {dt: 0.02, h_T: 0.25, h_eta: 0.2, rk_order: 4}
eos: {type: tanh_conformal}
hydro:
  viscous_parameters:
    diffusion:
      mode: kappaB_coth
      relaxation_mode: cb_over_T
      C_B: 0.4
      critical_scaling: true
      critical_point: {T: 140.0, muB: 420.0}
      critical_gaussian_width: {T: 25.0}
      constant_kappa_over_T2: [[0.1, 0.0, 0.0], [0.0, 0.2, 0.0], [0.0, 0.0, 0.3]]
      regulator: {enabled: true, chi0: 5.0, e0: 0.2}
""")
    s = S.from_yaml(str(y))
    assert (s["dim"], s["rk_order"], s["eos_type"], s["hT"], s["hEta"]) == (3, 4, "tanh_conformal", 0.25, 0.2)
    assert s["diffusionMode"] == "kappaB_coth" and s["diffusionRelaxMode"] == "cb_over_T" and s["C_B"] == 0.4
    assert s["critical_scaling_diffusion"] is True and s["critical_point_T"] == 140.0 and s["critical_point_muB"] == 420.0
    assert s["critical_gaussian_width_T"] == 25.0 and s["critical_gaussian_width_muB"] == 50.0      # default kept
    assert s["diffusion_regulator_enabled"] is True and s["diffusion_regulator_chi0"] == 5.0 and s["diffusion_regulator_e0"] == 0.2
    assert s["diffusion_regulator_xi0"] == 0.01 and s["diffusion_regulator_rq_max"] == 1.0          # defaults kept
    assert (s["kappa_00"], s["kappa_11"], s["kappa_22"], s["kappa_01"]) == (0.1, 0.2, 0.3, 0.0)
    c = S.to_config(s)
    assert c.rk_order == 4 and c.eos_type == 2 and c.diffusion_mode == 2 and c.diffusion_relax_mode == 1
    assert c.diffusion_regulator_enabled == 1 and c.critical_scaling_diffusion == 1 and c.using_diffusion == 1


@needs_ref
def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path on a bounded window of the workload) prints ONE JSON line with the keys the
    driver reads; no GPU involved."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-scale", "0.05"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))   # as torchrun exports it
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_steps_per_s" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["config"]["workload"] == "config4_milne_bsq_table_3p1d" and d["scaling"] == "strong"   # the configuration the metric is quoted on
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] == os.cpu_count()   # all host cores, whatever OMP_NUM_THREADS the launcher exported
    # per-phase seconds under the reference's CCAKE_PROFILE names, and the step with phase 04 OpenMP-parallelised next to it
    assert {"01_neighbour_list", "03_smooth_fields", "04_thermodynamics", "09_smooth_gradients", "10_time_derivatives"} <= set(d["cpu_baseline"]["per_phase"])
    assert d["cpu_baseline"]["phase04_openmp"]["value"] > 0 and "04_thermodynamics_openmp" in d["cpu_baseline"]["phase04_openmp"]["per_phase"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
