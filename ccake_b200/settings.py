"""Settings of a run: the reference's ``Settings`` fields (include/settings.h) and
their mapping to ``ccb_config``.

A settings dict uses the reference's member names (``hT``, ``etaMode``,
``constant_eta_over_s`` ...), so the same dict drives the oracle
(``oracle/ref.py``) and this library.  :func:`from_yaml` reads the reference's YAML
input format (key paths of ``src/input.cpp:96-912``, defaults of ``include/defaults.h``).
"""
from __future__ import annotations

import math

from . import _capi

# defaults the reference applies to missing YAML keys (include/defaults.h:12-128)
DEFAULTS = dict(
    dim=2, coordinate_system="hyperbolic", t0=1.0, dt=0.05, hT=0.3, hEta=0.3, rk_order=2, factorized_kernel=False,
    e_cutoff=0.15, input_as_entropy=False, stepx=0.0, stepy=0.0, stepEta=0.0, xmin=0.0, ymin=0.0, etamin=0.0,
    max_tau=50.0, eos_type="conformal", eos_path="", Freeze_Out_Temperature=150.0,
    baryon_charge_enabled=True, strange_charge_enabled=True, electric_charge_enabled=True,
    etaMode="constant", constant_eta_over_s=0.08, shearRelaxMode="default", use_vorticity=False,
    delta_pipi_mode="israel-stewart", tau_pipi_mode="disabled", lambda_piPi_mode="disabled", phi6_mode="disabled",
    phi7_mode="disabled", input_initial_shear=False,
    zetaMode="cs2_dependent", constant_zeta_over_s=0.005, cs2_dependent_zeta_A=1.67552, cs2_dependent_zeta_p=2.0,
    bulk_from_trace=False, bulkRelaxMode="default", modulate_zeta_with_tanh=True, critical_scaling_bulk=False,
    # defaults.h:60-63 declares these four as `bool` initialised from string literals (always true),
    # so a missing key leaves the string "1" -> setup_parameters falls to its "unknown -> default" branch.
    delta_PiPi_mode="default", lambda_Pipi_mode="default", phi1_mode="default", phi3_mode="default",
    diffusionMode="disabled", diffusionRelaxMode="default", C_B=0.0, prohibit_unstable_cs2=True,
    prohibit_acausal_cs2=True, restrict_mu_T_ratios=False, normalize_by_T=True,
    # include/defaults.h:67-71
    critical_scaling_diffusion=False, critical_point_T=150.0, critical_point_muB=350.0, critical_gaussian_width_T=30.0,
    critical_gaussian_width_muB=50.0,
    # include/defaults.h:119-123
    diffusion_regulator_enabled=False, diffusion_regulator_chi0=10.0, diffusion_regulator_e0=0.1, diffusion_regulator_xi0=0.01,
    diffusion_regulator_rq_max=1.0,
)

_SHEAR = {"default": 0, "constant": 1, "JakiParam": 2, "LinearMus": 3, "interpolate": 4}
_TAU_PI = {"default": 0, "minVal": 1, "Gubser": 2}
_ZETA = {"default": 0, "constant": 1, "DNMR": 2, "interpolate": 3, "cs2_dependent": 4}
_TAU_BULK = {"default": 0, "DNMR": 1}
_KAPPA = {"constant_over_T2": 0, "DNMR": 1, "kappaB_coth": 2}
_TAU_Q = {"constant_over_T": 0, "cb_over_T": 1}
_DELTA = {"default": 0, "israel-stewart": 1, "disabled": 2}
_ONOFF = {"default": 0, "disabled": 1}
_EOS = {"conformal": _capi.EOS_CONFORMAL, "table": _capi.EOS_TABLE, "tanh_conformal": _capi.EOS_TANH_CONFORMAL,
        "conformal_diagonal": _capi.EOS_CONFORMAL_DIAGONAL}


def derived_switches(s: dict) -> dict:
    """using_shear / using_diffusion / using_vorticity as Input::decode_settings derives them
    (src/input.cpp:390-408, 419-421, 577-582; Settings::check_consistency is never called)."""
    if s["etaMode"] == "disabled":
        shear = False
    elif s["etaMode"] == "constant":
        shear = not (s["constant_eta_over_s"] <= 1.0e-3)
    else:
        shear = False  # Settings::using_shear keeps its default for every other mode (include/settings.h:28)
    return dict(using_shear=shear, using_diffusion=s["diffusionMode"] != "disabled",
                using_vorticity=bool(s["use_vorticity"]))


def _mode(table, value, what):
    # Unknown strings select the reference's default branch (src/transport_coefficients.cpp:9-279 leaves
    # the enum unset / falls to "default"); we make that explicit instead of reading an uninitialised int.
    return table.get(value, 0)


def to_config(settings: dict, device: int = 0, rank: int = 0, nranks: int = 1) -> _capi.Config:
    s = dict(DEFAULTS)
    s.update(settings)
    if s["eos_type"] not in _EOS:
        raise NotImplementedError("eos_type %r: conformal, table, tanh_conformal and conformal_diagonal are available in this build" % s["eos_type"])
    c = _capi.Config()
    c.dim = int(s["dim"])
    c.coordinate_system = _capi.COORD_CARTESIAN if s["coordinate_system"] == "cartesian" else _capi.COORD_HYPERBOLIC
    c.rk_order = int(s["rk_order"])
    c.factorized_kernel = int(bool(s["factorized_kernel"]))
    for k in ("t0", "hT", "hEta", "xmin", "ymin", "etamin", "stepx", "stepy", "stepEta"):
        setattr(c, k, float(s[k]))
    d = derived_switches(s)
    c.using_shear, c.using_diffusion, c.using_vorticity = int(d["using_shear"]), int(d["using_diffusion"]), int(d["using_vorticity"])
    c.input_as_entropy = int(bool(s["input_as_entropy"]))
    c.input_initial_shear = int(bool(s["input_initial_shear"]))
    c.bulk_from_trace = int(bool(s["bulk_from_trace"]))
    c.shear_mode = _mode(_SHEAR, s["etaMode"], "etaMode")
    c.shear_relax_mode = _mode(_TAU_PI, s["shearRelaxMode"], "shearRelaxMode")
    c.zeta_mode = _mode(_ZETA, s["zetaMode"], "zetaMode")
    c.bulk_relax_mode = _mode(_TAU_BULK, s["bulkRelaxMode"], "bulkRelaxMode")
    c.diffusion_mode = _mode(_KAPPA, s["diffusionMode"], "diffusionMode")
    c.diffusion_relax_mode = _mode(_TAU_Q, s["diffusionRelaxMode"], "diffusionRelaxMode")
    c.delta_pipi_mode = _mode(_DELTA, s["delta_pipi_mode"], "delta_pipi_mode")
    c.tau_pipi_mode = _mode(_ONOFF, s["tau_pipi_mode"], "tau_pipi_mode")
    c.lambda_piPi_mode = _mode(_ONOFF, s["lambda_piPi_mode"], "lambda_piPi_mode")
    c.phi6_mode = _mode(_ONOFF, s["phi6_mode"], "phi6_mode")
    c.phi7_mode = _mode(_ONOFF, s["phi7_mode"], "phi7_mode")
    c.delta_PiPi_mode = _mode(_DELTA, s["delta_PiPi_mode"], "delta_PiPi_mode")
    c.lambda_Pipi_mode = _mode(_ONOFF, s["lambda_Pipi_mode"], "lambda_Pipi_mode")
    c.phi1_mode = _mode(_ONOFF, s["phi1_mode"], "phi1_mode")
    c.phi3_mode = _mode(_ONOFF, s["phi3_mode"], "phi3_mode")
    c.modulate_zeta_with_tanh = int(bool(s["modulate_zeta_with_tanh"]))
    c.critical_scaling_bulk = int(bool(s["critical_scaling_bulk"]))
    c.constant_eta_over_s = float(s["constant_eta_over_s"])
    c.constant_zeta_over_s = float(s["constant_zeta_over_s"])
    c.cs2_dependent_zeta_A = float(s["cs2_dependent_zeta_A"])
    c.cs2_dependent_zeta_p = float(s["cs2_dependent_zeta_p"])
    c.C_B = float(s["C_B"])
    for a in range(3):
        for b in range(3):
            c.constant_kappa_over_T2[a * 3 + b] = float(s.get("kappa_%d%d" % (a, b), 0.0))
    c.critical_scaling_diffusion = int(bool(s["critical_scaling_diffusion"]))
    for k in ("critical_point_T", "critical_point_muB", "critical_gaussian_width_T", "critical_gaussian_width_muB"):
        setattr(c, k, float(s[k]))
    c.diffusion_regulator_enabled = int(bool(s["diffusion_regulator_enabled"]))
    for k in ("chi0", "e0", "xi0", "rq_max"):
        setattr(c, "diffusion_regulator_" + k, float(s["diffusion_regulator_" + k]))
    c.eos_type = _EOS[s["eos_type"]]
    # EquationOfState::set_up_chosen_EOSs, src/eos_initialization.cpp:216-231
    Nc, Nf = 3.0, 2.5
    c.conformal_c = math.pi * math.pi * (2.0 * (Nc * Nc - 1.0) + (7.0 / 2.0) * Nc * Nf) / 90.0
    c.conformal_T0 = c.conformal_muB0 = c.conformal_muS0 = c.conformal_muQ0 = 1.0
    c.freeze_out_temperature_MeV = float(s["Freeze_Out_Temperature"])
    c.prohibit_unstable_cs2 = int(bool(s["prohibit_unstable_cs2"]))
    c.prohibit_acausal_cs2 = int(bool(s["prohibit_acausal_cs2"]))
    c.restrict_mu_T_ratios = int(bool(s["restrict_mu_T_ratios"]))
    c.device, c.rank, c.nranks = int(device), int(rank), int(nranks)
    c.particlization_enabled = int(bool(s.get("particlization_enabled", False)))
    c.dt = float(s["dt"])
    c.online_inverter_enabled = int(bool(s.get("online_inverter_enabled", True)))
    c.slab_axis = int(s.get("slab_axis", 0))
    return c


def _get(node, path, default):
    for k in path:
        if not isinstance(node, dict) or k not in node:
            return default
        node = node[k]
    return node


def from_yaml(path: str) -> dict:
    """Read a reference YAML input (``input/*.yaml``) into a settings dict."""
    import yaml

    with open(path) as f:
        y = yaml.safe_load(f)
    s = dict(DEFAULTS)
    vp = ("hydro", "viscous_parameters")
    table = [
        ("dim", ("initial_conditions", "dimension")), ("t0", ("initial_conditions", "t0")),
        ("coordinate_system", ("initial_conditions", "coordinate_system")),
        ("input_as_entropy", ("initial_conditions", "input_as_entropy")),
        ("dt", ("parameters", "dt")), ("hT", ("parameters", "h_T")), ("hEta", ("parameters", "h_eta")),
        ("rk_order", ("parameters", "rk_order")), ("factorized_kernel", ("parameters", "factorized_kernel")),
        ("e_cutoff", ("parameters", "energy_cutoff")), ("max_tau", ("parameters", "max_tau")),
        ("eos_type", ("eos", "type")), ("eos_path", ("eos", "path")),
        ("prohibit_unstable_cs2", ("eos", "prohibit_unstable_cs2")), ("prohibit_acausal_cs2", ("eos", "prohibit_acausal_cs2")),
        ("restrict_mu_T_ratios", ("eos", "restrict_mu_T_ratios")), ("normalize_by_T", ("eos", "normalize_by_T")),
        ("Freeze_Out_Temperature", ("particlization", "T")),
        ("baryon_charge_enabled", ("hydro", "baryon_charge_enabled")),
        ("strange_charge_enabled", ("hydro", "strange_charge_enabled")),
        ("electric_charge_enabled", ("hydro", "electric_charge_enabled")),
        ("etaMode", vp + ("shear", "mode")), ("constant_eta_over_s", vp + ("shear", "constant_eta_over_s")),
        ("shearRelaxMode", vp + ("shear", "relaxation_mode")), ("use_vorticity", vp + ("shear", "use_vorticity")),
        ("delta_pipi_mode", vp + ("shear", "delta_pipi_mode")), ("tau_pipi_mode", vp + ("shear", "tau_pipi_mode")),
        ("lambda_piPi_mode", vp + ("shear", "lambda_piPi_mode")), ("phi6_mode", vp + ("shear", "phi6_mode")),
        ("phi7_mode", vp + ("shear", "phi7_mode")), ("input_initial_shear", vp + ("shear", "input_initial_shear")),
        ("zetaMode", vp + ("bulk", "mode")), ("constant_zeta_over_s", vp + ("bulk", "constant_zeta_over_s")),
        ("cs2_dependent_zeta_A", vp + ("bulk", "cs2_dependent_zeta_A")),
        ("cs2_dependent_zeta_p", vp + ("bulk", "cs2_dependent_zeta_p")),
        ("bulk_from_trace", vp + ("bulk", "bulk_from_trace")), ("bulkRelaxMode", vp + ("bulk", "relaxation_mode")),
        ("modulate_zeta_with_tanh", vp + ("bulk", "modulate_with_tanh")),
        ("critical_scaling_bulk", vp + ("bulk", "critical_scaling_enabled")),
        ("delta_PiPi_mode", vp + ("bulk", "delta_PiPi_mode")), ("lambda_Pipi_mode", vp + ("bulk", "lambda_Pipi_mode")),
        ("phi1_mode", vp + ("bulk", "phi1_mode")), ("phi3_mode", vp + ("bulk", "phi3_mode")),
        ("diffusionMode", vp + ("diffusion", "mode")), ("diffusionRelaxMode", vp + ("diffusion", "relaxation_mode")),
        ("C_B", vp + ("diffusion", "C_B")),
        ("critical_scaling_diffusion", vp + ("diffusion", "critical_scaling")),
        ("critical_point_T", vp + ("diffusion", "critical_point", "T")),
        ("critical_point_muB", vp + ("diffusion", "critical_point", "muB")),
        ("critical_gaussian_width_T", vp + ("diffusion", "critical_gaussian_width", "T")),
        ("critical_gaussian_width_muB", vp + ("diffusion", "critical_gaussian_width", "muB")),
        ("diffusion_regulator_enabled", vp + ("diffusion", "regulator", "enabled")),
        ("diffusion_regulator_chi0", vp + ("diffusion", "regulator", "chi0")),
        ("diffusion_regulator_e0", vp + ("diffusion", "regulator", "e0")),
        ("diffusion_regulator_xi0", vp + ("diffusion", "regulator", "xi0")),
        ("diffusion_regulator_rq_max", vp + ("diffusion", "regulator", "rq_max")),
    ]
    for key, p in table:
        s[key] = _get(y, p, s[key])
    mat = _get(y, vp + ("diffusion", "constant_kappa_over_T2"), None)
    if mat is not None:
        for a in range(3):
            for b in range(3):
                s["kappa_%d%d" % (a, b)] = float(mat[a][b])
    return s
