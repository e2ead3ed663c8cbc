"""Shared fixtures of the test-suite: synthetic initial conditions (the configs of
BASELINE.json / SURVEY.md 8d at test sizes), settings dicts, and the two checkers
(oracle/_ref = the reference compiled through shims; the host harness = the
library's own physics headers compiled for the CPU)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HBARC = 0.1973269804


# ---------------------------------------------------------------------------------
# initial conditions
def gubser_profile(x, y, tau0=1.0, eps0=80.0, q=1.0):
    """Ideal Gubser flow (reference utilities/ideal_gubser_ic_generator.py:22-37): eps [GeV/fm^3], u^x, u^y."""
    r = np.hypot(x, y)
    eps = (eps0 / tau0 ** (4.0 / 3.0)) * (2 * q) ** (8.0 / 3.0) / \
          (1 + 2 * q ** 2 * (tau0 ** 2 + r ** 2) + (q ** 2 * (tau0 ** 2 - r ** 2)) ** 2) ** (4.0 / 3.0)
    kap = np.arctanh(2 * q ** 2 * tau0 * r / (1 + (q * tau0) ** 2 + (q * r) ** 2))
    with np.errstate(invalid="ignore", divide="ignore"):
        ux = np.where(r > 0, np.sinh(kap) * x / r, 0.0)
        uy = np.where(r > 0, np.sinh(kap) * y / r, 0.0)
    return eps, ux, uy


def gubser_ic(dx=0.1, xmax=2.0, shear=False, charges=False, seed=1):
    """2+1D lattice IC; optional synthetic pi^{ij} and charge densities (config 2 flavour)."""
    xs = np.arange(-xmax, xmax + dx / 2, dx)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    x, y = X.ravel(), Y.ravel()
    eps, ux, uy = gubser_profile(x, y)
    n = x.size
    ic = dict(r=np.stack([x, y], 1), e=eps, u=np.stack([ux, uy], 1), rhoB=None, rhoS=None, rhoQ=None, bulk=None, shv6=None)
    if shear:
        rng = np.random.default_rng(seed)
        amp = 0.05 * eps
        pixx = amp * np.cos(1.3 * x) * (1 + 0.1 * rng.standard_normal(n))
        piyy = amp * np.sin(0.7 * y + 0.3)
        pixy = 0.5 * amp * np.sin(x * y)
        ic["shv6"] = np.stack([pixx, pixy, np.zeros(n), piyy, np.zeros(n), np.zeros(n)], 1)
    if charges:
        g = np.exp(-(x ** 2 + y ** 2) / 4.0)
        ic["rhoB"] = 0.30 * g * (1 + 0.2 * np.cos(x))
        ic["rhoS"] = -0.10 * g * (1 + 0.2 * np.sin(y))
        ic["rhoQ"] = 0.12 * g
    return ic


def blob_ic_3d(n_side=12, dx=0.25, sigma=1.0, milne=True, seed=3):
    """3+1D Gaussian blob on a cubic lattice (configs 4/5 flavour)."""
    xs = (np.arange(n_side) - (n_side - 1) / 2.0) * dx
    X, Y, Z = np.meshgrid(xs, xs, xs, indexing="ij")
    x, y, z = X.ravel(), Y.ravel(), Z.ravel()
    r2 = x * x + y * y + z * z
    eps = 20.0 * np.exp(-r2 / (2 * sigma ** 2)) + 0.5
    rng = np.random.default_rng(seed)
    u = 0.15 * np.stack([x, y, 0.5 * z], 1) / (1 + r2[:, None]) + 0.01 * rng.standard_normal((x.size, 3))
    shv6 = 0.03 * eps[:, None] * np.stack([np.cos(x), 0.5 * np.sin(x * y), 0.3 * np.sin(z), np.sin(y + 0.2), 0.2 * np.cos(z * x),
                                           np.zeros_like(x)], 1)
    rhoB = 0.2 * np.exp(-r2 / 3.0) * (1 + 0.3 * np.tanh(z))
    return dict(r=np.stack([x, y, z], 1), e=eps, u=u, rhoB=rhoB, rhoS=-0.3 * rhoB, rhoQ=0.4 * rhoB, bulk=None, shv6=shv6)


def base_settings(**kw):
    s = dict(dim=2, t0=1.0, dt=0.001, hT=0.3, hEta=0.3, stepx=0.1, stepy=0.1, stepEta=0.1, xmin=-2.0, ymin=-2.0,
             etamin=-0.1, eos_type="conformal", etaMode="constant", constant_eta_over_s=0.0, shearRelaxMode="default",
             zetaMode="constant", constant_zeta_over_s=0.0, bulkRelaxMode="default",
             delta_pipi_mode="default", tau_pipi_mode="default", lambda_piPi_mode="default", phi6_mode="default",
             phi7_mode="default", delta_PiPi_mode="default", lambda_Pipi_mode="default", phi1_mode="default",
             phi3_mode="default", baryon_charge_enabled=False, strange_charge_enabled=False,
             electric_charge_enabled=False, diffusionMode="disabled", diffusionRelaxMode="constant_over_T",
             modulate_zeta_with_tanh=False, input_initial_shear=False, print_conservation_status=False)
    s.update(kw)
    return s


def ideal_settings(**kw):
    return base_settings(**kw)


def viscous_settings(**kw):
    s = base_settings(constant_eta_over_s=0.2, shearRelaxMode="Gubser", constant_zeta_over_s=0.02, input_initial_shear=True)
    s.update(kw)
    return s


def charged_settings(**kw):
    s = viscous_settings(baryon_charge_enabled=True, strange_charge_enabled=True, electric_charge_enabled=True,
                         diffusionMode="constant_over_T2", kappa_00=0.05, kappa_11=0.04, kappa_22=0.03, kappa_01=0.01,
                         kappa_10=0.01)
    s.update(kw)
    return s


# ---------------------------------------------------------------------------------
# checkers
def ref_module():
    from oracle import ref

    return ref


def have_ref():
    from oracle import ref

    return ref.available()


def make_ref(settings, ic):
    ref = ref_module()
    keys = ref.SETTING_KEYS
    st = {k: v for k, v in settings.items() if k in keys or k.startswith("kappa_")}
    sim = ref.RefSim(st)
    sim.add_particles(ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"], ic["u"], ic["bulk"], ic["shv6"], ic.get("diffusion"))
    return sim


HARNESS_SRC = os.path.join(ROOT, "tests", "host_harness.cpp")
HARNESS_LIB = os.path.join(ROOT, "tests", "libhost_harness.so")


def harness():
    """The library's physics headers compiled for the host (test infrastructure)."""
    srcs = [HARNESS_SRC] + [os.path.join(ROOT, "ccake_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "ccake_b200", "csrc"))
                            if f.endswith((".cuh", ".h"))]
    if not os.path.exists(HARNESS_LIB) or any(os.path.getmtime(s) > os.path.getmtime(HARNESS_LIB) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-x", "c++", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-o", HARNESS_LIB, HARNESS_SRC])
    return C.CDLL(HARNESS_LIB)


DROPIN_SRC = os.path.join(ROOT, "tests", "cpp", "dropin_bsqhydro.cpp")
DROPIN_LIB = os.path.join(ROOT, "tests", "libdropin_bsqhydro.so")
REFERENCE = "/root/reference"


def build_dropin():
    """The reference's unmodified src/BSQHydro.cpp compiled against include/ccake_dropin/ (tests/cpp/dropin_bsqhydro.cpp).
    Needs /root/reference; on the GPU box the prebuilt tests/libdropin_bsqhydro.so is used. The include path is an overlay
    directory: the drop-in headers + links to the reference headers they do not replace (a quoted #include looks next to the
    including file first, so the reference's own include/ cannot simply come second on the path)."""
    if not os.path.isdir(os.path.join(REFERENCE, "src")):
        return os.path.exists(DROPIN_LIB)
    dropin = os.path.join(ROOT, "include", "ccake_dropin")
    lib = os.path.join(ROOT, "ccake_b200", "libccake_b200.so")
    srcs = [DROPIN_SRC, lib, os.path.join(ROOT, "include", "ccake_b200.h"), os.path.join(ROOT, "oracle", "settings_kv.h")] + \
        [os.path.join(dropin, f) for f in os.listdir(dropin)]
    if os.path.exists(DROPIN_LIB) and all(os.path.getmtime(x) <= os.path.getmtime(DROPIN_LIB) for x in srcs):
        return True
    overlay = os.path.join(ROOT, "tests", "cpp", "build", "include")
    if os.path.isdir(overlay):
        for f in os.listdir(overlay):
            os.unlink(os.path.join(overlay, f))
    os.makedirs(overlay, exist_ok=True)
    ours = set(os.listdir(dropin))
    for f in os.listdir(os.path.join(REFERENCE, "include")):
        if f not in ours:
            os.symlink(os.path.join(REFERENCE, "include", f), os.path.join(overlay, f))
    for f in ours:
        os.symlink(os.path.join(dropin, f), os.path.join(overlay, f))
    src_link = os.path.join(ROOT, "tests", "cpp", "build", "src")   # vector.h / matrix.h include "../src/vector.cpp"
    if not os.path.islink(src_link):
        os.symlink(os.path.join(REFERENCE, "src"), src_link)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-I", overlay, "-I", os.path.join(REFERENCE, "src"),
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "oracle", "shim"),
                           DROPIN_SRC, os.path.join(REFERENCE, "src", "formatted_output.cpp"), os.path.join(REFERENCE, "src", "particle.cpp"),
                           "-o", DROPIN_LIB, "-L", os.path.dirname(lib), "-lccake_b200", "-Wl,-rpath,$ORIGIN/../ccake_b200"])
    return True


class EomArrays(C.Structure):
    _fields_ = [(k, C.POINTER(C.c_double)) for k in
                ("u", "v", "gamma", "sigma", "sigma_lab", "mass_s", "bulk", "shv", "q", "thermo", "dalpha", "gradP",
                 "gradBulk", "gradshear", "divshear", "gradV", "div_qa", "grad_qa", "grad_alpha", "du_dt", "d_ext",
                 "d_bulk", "d_shv", "d_q", "theta", "shv_nabla_u", "dEz")]


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def rel_err(a, b, scale=None):
    """Field-scale-relative error: max|a-b| / max(scale, tiny) (SURVEY.md section 7, 1e-10 tolerance note)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b)) if scale is None else scale
    return float(np.max(np.abs(a - b)) / max(s, 1e-300))


def ref_thermo16(sim):
    """thermo block of the reference in the library's 16-column order."""
    th = sim.member("thermo")
    return th[:, :16].copy()


def ref_dalpha12(sim):
    return sim.member("thermo")[:, 16:28].copy()
