"""Synthetic initial conditions and settings of the BASELINE.json configs.

Everything is generated from closed forms with fixed seeds: no file of the reference tree is
read at run time (``/root/reference`` does not exist on the GPU box).  Settings dicts use the
reference's ``Settings`` member names, so the same dict drives the CUDA library and the oracle.
"""
from __future__ import annotations

import math

import numpy as np

HBARC_GEVFM = 0.1973269804  # reference include/constants.h:15
CONFORMAL_C = math.pi ** 2 * (2.0 * (3.0 ** 2 - 1.0) + 3.5 * 3.0 * 2.5) / 90.0  # src/eos_initialization.cpp:216-231


def gubser_energy_u(x, y, tau0=1.0, eps0=80.0, q=1.0):
    """Ideal Gubser flow at tau0 (reference utilities/ideal_gubser_ic_generator.py:22-37):
    energy density [GeV/fm^3] and u^x, u^y."""
    r = np.hypot(x, y)
    eps = (eps0 / tau0 ** (4.0 / 3.0)) * (2 * q) ** (8.0 / 3.0) / \
        (1 + 2 * q ** 2 * (tau0 ** 2 + r ** 2) + (q ** 2 * (tau0 ** 2 - r ** 2)) ** 2) ** (4.0 / 3.0)
    kap = np.arctanh(2 * q ** 2 * tau0 * r / (1 + (q * tau0) ** 2 + (q * r) ** 2))
    with np.errstate(invalid="ignore", divide="ignore"):
        ux = np.where(r > 0, np.sinh(kap) * x / r, 0.0)
        uy = np.where(r > 0, np.sinh(kap) * y / r, 0.0)
    return eps, ux, uy


def gubser_charge(x, y, tau0=1.0, q=1.0):
    """chargeGubser(tau, r) of reference misc/Gubser_checks/plot_gubser_profiles.py:46-50."""
    r2 = x * x + y * y
    return (1.0 / tau0 ** 3) * (4.0 * q ** 2 * tau0 ** 2) / (1.0 + 2.0 * q ** 2 * (tau0 ** 2 + r2) + q ** 4 * (tau0 ** 2 - r2) ** 2)


def lattice_2d(dx, xmax):
    xs = np.arange(-xmax, xmax + dx / 2, dx)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    return X.ravel(), Y.ravel()


_COMMON = dict(dim=2, coordinate_system="hyperbolic", t0=1.0, rk_order=2, eos_type="conformal", input_as_entropy=False,
               e_cutoff=0.0, print_conservation_status=False)


def config1_ideal_gubser(dx=0.025, xmax=5.0):
    """BASELINE config 1: input/input_parameters_ideal_gubser_checks.yaml (h = 0.4, dt = 0.001, ideal fluid, no charges), 401^2 lattice."""
    x, y = lattice_2d(dx, xmax)
    eps, ux, uy = gubser_energy_u(x, y)
    settings = dict(_COMMON, dt=0.001, hT=0.4, hEta=0.4, stepx=dx, stepy=dx, stepEta=0.1, xmin=-xmax, ymin=-xmax, etamin=-0.1,
                    etaMode="constant", constant_eta_over_s=0.0, shearRelaxMode="default", zetaMode="constant",
                    constant_zeta_over_s=0.0, bulkRelaxMode="default", diffusionMode="disabled",
                    baryon_charge_enabled=False, strange_charge_enabled=False, electric_charge_enabled=False,
                    modulate_zeta_with_tanh=False, input_initial_shear=False)
    ic = dict(r=np.stack([x, y], 1), e=eps, u=np.stack([ux, uy], 1), rhoB=None, rhoS=None, rhoQ=None, bulk=None, shv6=None)
    return settings, ic


def tile_along_x(settings, ic, ntiles, dx):
    """Weak-scaling workload: `ntiles` copies of an IC side by side along x (period = extent + dx, so the lattice
    spacing is uniform across the seams and neighbouring copies interact through their edge particles), centred on
    x = 0 because the reference's neighbour grid is symmetric about the origin (src/system_state.cpp:656-692)."""
    if ntiles == 1:
        return settings, ic
    x = ic["r"][:, 0]
    period = (x.max() - x.min()) + dx
    shifts = (np.arange(ntiles) - (ntiles - 1) / 2.0) * period
    out = {}
    for k, v in ic.items():
        if v is None:
            out[k] = None
        elif k == "r":
            out[k] = np.concatenate([v + np.array([sh] + [0.0] * (v.shape[1] - 1)) for sh in shifts])
        else:
            out[k] = np.concatenate([v] * ntiles)
    s = dict(settings)
    s["xmin"] = float(out["r"][:, 0].min())
    return s, out


def config2_viscous_gubser_bsq(dx=0.025, xmax=5.0, seed=7):
    """BASELINE config 2: the settings of input/config_visc_gubser.yaml (h = 0.1, dt = 0.001, eta/s = 0.2 with the
    Gubser relaxation time, conformal EoS at finite mu_B, mu_S, mu_Q, diffusion code path on with kappa = 0,
    initial shear read from the IC). The IC file that YAML names is external to the reference tree, so the
    profile is synthetic: ideal-Gubser e and u^i, Gubser charge profiles with mu_B/T ~ 0.3 at the centre, and a
    smooth pi^{ij} of 2 % of e. NB the YAML spells `relaxation_mode: gubser`; the reference only recognises
    "Gubser" (src/transport_coefficients.cpp:42) and otherwise leaves the enum unset: the intended mode is used."""
    x, y = lattice_2d(dx, xmax)
    eps, ux, uy = gubser_energy_u(x, y)
    e_fm = eps / HBARC_GEVFM
    s = 4.0 * CONFORMAL_C * (e_fm / (3.0 * CONFORMAL_C)) ** 0.75  # conformal s(e) at mu = 0
    g = gubser_charge(x, y)
    g = g / g.max()
    rhoB, rhoS, rhoQ = 0.30 * s.max() * g, -0.10 * s.max() * g, 0.12 * s.max() * g
    rng = np.random.default_rng(seed)
    amp = 0.02 * eps
    pixx = amp * np.cos(1.3 * x) * (1 + 0.05 * rng.standard_normal(x.size))
    piyy = amp * np.sin(0.7 * y + 0.3)
    pixy = 0.5 * amp * np.sin(x * y)
    z = np.zeros_like(x)
    settings = dict(_COMMON, dt=0.001, hT=0.1, hEta=0.1, stepx=dx, stepy=dx, stepEta=0.1, xmin=-xmax, ymin=-xmax, etamin=-0.1,
                    Freeze_Out_Temperature=153.334025,
                    baryon_charge_enabled=True, strange_charge_enabled=True, electric_charge_enabled=True,
                    etaMode="constant", constant_eta_over_s=0.2, shearRelaxMode="Gubser", use_vorticity=False,
                    delta_pipi_mode="default", tau_pipi_mode="disabled", lambda_piPi_mode="disabled", phi6_mode="disabled",
                    phi7_mode="disabled", input_initial_shear=True,
                    zetaMode="constant", constant_zeta_over_s=0.0, cs2_dependent_zeta_A=1.67552, cs2_dependent_zeta_p=2.0,
                    bulk_from_trace=False, bulkRelaxMode="default", modulate_zeta_with_tanh=False,
                    delta_PiPi_mode="israel-stewart", lambda_Pipi_mode="default", phi1_mode="disabled", phi3_mode="disabled",
                    diffusionMode="constant_over_T2", diffusionRelaxMode="constant_over_T")
    ic = dict(r=np.stack([x, y], 1), e=eps, u=np.stack([ux, uy], 1), rhoB=rhoB, rhoS=rhoS, rhoQ=rhoQ, bulk=None,
              shv6=np.stack([pixx, pixy, z, piyy, z, z], 1))
    return settings, ic


def viscous_gubser_check(initial):
    """The reference's own physics check (misc/Gubser_checks/ac, README.txt there): Israel-Stewart hydrodynamics in Gubser flow from
    the semi-analytic tau = 1 fm profile on the 201 x 201, dx = 0.05 fm grid, eta/s = 0.2, tau_R = 5 eta / (e + P), conformal e = 3 P, no
    charges. `initial` = rows x y T[GeV] u^x u^y pi^xx pi^yy pi^xy tau^2 pi^etaeta [GeV/fm^3] (tests/golden/viscous_gubser_ac.npz);
    converted to the ccake IC columns exactly as utilities/gubser_ic_converter.py does. Settings: input/config_visc_gubser.yaml
    without its charges (h = 0.1 fm)."""
    a = np.asarray(initial, dtype=np.float64)
    x, y, T = a[:, 0], a[:, 1], a[:, 2]
    eps = HBARC_GEVFM * 3.0 * CONFORMAL_C * T ** 4 / HBARC_GEVFM ** 4
    z = np.zeros_like(x)
    dx = 0.05
    settings = dict(_COMMON, dt=0.01, hT=0.1, hEta=0.1, stepx=dx, stepy=dx, stepEta=0.1, xmin=-5.0, ymin=-5.0, etamin=-0.1,
                    Freeze_Out_Temperature=153.334025, baryon_charge_enabled=False, strange_charge_enabled=False, electric_charge_enabled=False,
                    etaMode="constant", constant_eta_over_s=0.2, shearRelaxMode="Gubser", use_vorticity=False,
                    delta_pipi_mode="default", tau_pipi_mode="disabled", lambda_piPi_mode="disabled", phi6_mode="disabled",
                    phi7_mode="disabled", input_initial_shear=True,
                    zetaMode="constant", constant_zeta_over_s=0.0, cs2_dependent_zeta_A=1.67552, cs2_dependent_zeta_p=2.0,
                    bulk_from_trace=False, bulkRelaxMode="default", modulate_zeta_with_tanh=False,
                    delta_PiPi_mode="israel-stewart", lambda_Pipi_mode="default", phi1_mode="disabled", phi3_mode="disabled",
                    diffusionMode="disabled", diffusionRelaxMode="constant_over_T")
    ic = dict(r=np.stack([x, y], 1), e=eps, u=np.stack([a[:, 3], a[:, 4]], 1), rhoB=None, rhoS=None, rhoQ=None, bulk=None,
              shv6=np.stack([a[:, 5], a[:, 7], z, a[:, 6], z, a[:, 8]], 1))
    return settings, ic


def config5_cartesian_blob(n_target=1_000_000, sigma=3.0, seed=11):
    """BASELINE config 5: 3+1D Cartesian Gaussian blob on a cubic lattice inside radius 3 sigma, h = 2 dx
    (N_ngb ~ 268), conformal EoS, ideal fluid at rest."""
    R = 3.0 * sigma
    dx = (4.0 / 3.0 * math.pi * R ** 3 / n_target) ** (1.0 / 3.0)
    m = int(math.floor(R / dx))
    xs = np.arange(-m, m + 1) * dx
    X, Y, Z = np.meshgrid(xs, xs, xs, indexing="ij")
    keep = (X * X + Y * Y + Z * Z) <= R * R
    x, y, z = X[keep], Y[keep], Z[keep]
    r2 = x * x + y * y + z * z
    eps = 30.0 * np.exp(-r2 / (2 * sigma ** 2)) + 0.3
    lo = -(m + 1) * dx
    settings = dict(_COMMON, dim=3, coordinate_system="cartesian", dt=0.01, hT=2 * dx, hEta=2 * dx, stepx=dx, stepy=dx,
                    stepEta=dx, xmin=lo, ymin=lo, etamin=lo, etaMode="constant", constant_eta_over_s=0.0,
                    shearRelaxMode="default", zetaMode="constant", constant_zeta_over_s=0.0, bulkRelaxMode="default",
                    diffusionMode="disabled", baryon_charge_enabled=False, strange_charge_enabled=False,
                    electric_charge_enabled=False, modulate_zeta_with_tanh=False, input_initial_shear=False)
    ic = dict(r=np.stack([x, y, z], 1), e=eps, u=np.zeros((x.size, 3)), rhoB=None, rhoS=None, rhoQ=None, bulk=None, shv6=None)
    return settings, ic


def _nucleon_positions(n_nucleons, rng, R=6.62, a=0.546):
    """Nucleon positions of one nucleus sampled from a Woods-Saxon density (rejection sampling), projected on x-y."""
    out = np.empty((0, 3))
    while out.shape[0] < n_nucleons:
        r = rng.uniform(0.0, R + 8 * a, 4 * n_nucleons)
        keep = rng.uniform(0.0, 1.0, r.size) * (R + 8 * a) ** 2 < r * r / (1.0 + np.exp((r - R) / a))
        r = r[keep]
        ct, ph = rng.uniform(-1.0, 1.0, r.size), rng.uniform(0.0, 2 * math.pi, r.size)
        st = np.sqrt(1.0 - ct * ct)
        out = np.concatenate([out, np.stack([r * st * np.cos(ph), r * st * np.sin(ph), r * ct], 1)])
    return out[:n_nucleons, :2]


def trento_like_profile(x, y, seed=12345, impact=0.0, sigma=0.5, e_max=90.0):
    """Synthetic TRENTo-like Pb-Pb transverse energy-density profile [GeV/fm^3]: the geometric mean (TRENTo p = 0) of the two
    nuclear thickness functions, each a sum of 208 Gaussians of width `sigma` at Woods-Saxon nucleon positions, normalised
    to `e_max` at its maximum. x, y: 1-D arrays of lattice coordinates; returns the 2-D array [len(x), len(y)]."""
    rng = np.random.default_rng(seed)
    T = []
    for sign in (-1.0, 1.0):
        pos = _nucleon_positions(208, rng)
        pos[:, 0] += sign * impact / 2
        gx = np.exp(-(x[:, None] - pos[None, :, 0]) ** 2 / (2 * sigma ** 2))   # separable Gaussians: [nx, 208] @ [208, ny]
        gy = np.exp(-(y[:, None] - pos[None, :, 1]) ** 2 / (2 * sigma ** 2))
        T.append(gx @ gy.T)
    e = np.sqrt(T[0] * T[1])
    return e * (e_max / e.max())


_TABLE_COMMON = dict(eos_type="table", normalize_by_T=True, Freeze_Out_Temperature=150.0)


def config3_trento_pbpb(n_target=1_000_000, xmax=12.0, window=1.0, seed=12345, dx=None):
    """BASELINE config 3: TRENTo-like Pb-Pb 2+1D with the settings of input/input_trento.yaml (t0 = 0.6, dt = 0.05,
    h_T = 0.3, e_cutoff = 0.15 GeV/fm^3, no conserved charges, eta/s = 0.08) plus the cs2-dependent bulk viscosity of
    include/defaults.h:53-56, on the tabulated EoS (T-only inversion of the 4-D table). `window` < 1 keeps
    the central |x|, |y| <= window * xmax part of the SAME event at the same spacing (the CPU arms' bounded sample).
    The lattice spacing is chosen so that ~`n_target` particles survive the cut-off: the area above 0.15 GeV/fm^3 is
    201.5 fm^2, i.e. 503 650 particles at SURVEY's dx = 0.02 and 1.0e6 at dx = 0.0142 (N_ngb = pi (2h)^2 / dx^2 ~ 5600)."""
    if dx is None:
        xc = np.arange(-240, 241) * 0.05
        area = float((trento_like_profile(xc, xc, seed=seed) > 0.15).sum()) * 0.05 ** 2
        dx = round(math.sqrt(area / n_target), 4)
    m = int(round(xmax * window / dx))
    xs = np.arange(-m, m + 1) * dx
    e2 = trento_like_profile(xs, xs, seed=seed)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    keep = e2 > 0.15                      # BSQHydro::read_ccake drops rows below e_cutoff (src/BSQHydro.cpp:300-306)
    x, y, eps = X[keep], Y[keep], e2[keep]
    lo = -(m * dx)
    settings = dict(_COMMON, **_TABLE_COMMON, t0=0.6, dt=0.05, hT=0.3, hEta=0.1, stepx=dx, stepy=dx, stepEta=0.1, xmin=lo, ymin=lo,
                    etamin=-0.1, e_cutoff=0.15, baryon_charge_enabled=False, strange_charge_enabled=False,
                    electric_charge_enabled=False, etaMode="constant", constant_eta_over_s=0.08, shearRelaxMode="default",
                    use_vorticity=False, delta_pipi_mode="israel-stewart", tau_pipi_mode="disabled", lambda_piPi_mode="disabled",
                    phi6_mode="disabled", phi7_mode="disabled", input_initial_shear=False,
                    zetaMode="cs2_dependent", constant_zeta_over_s=0.005, cs2_dependent_zeta_A=1.67552, cs2_dependent_zeta_p=2.0,
                    bulk_from_trace=False, bulkRelaxMode="default", modulate_zeta_with_tanh=False,
                    delta_PiPi_mode="default", lambda_Pipi_mode="default", phi1_mode="default", phi3_mode="default",
                    diffusionMode="disabled")
    ic = dict(r=np.stack([x, y], 1), e=eps, u=np.zeros((x.size, 2)), rhoB=None, rhoS=None, rhoQ=None, bulk=None, shv6=None)
    return settings, ic


def config4_milne_bsq(n_target=8_000_000, window=1.0, seed=12345):
    """BASELINE config 4: 3+1D Milne run with B, S, Q charges on the tabulated EoS (full 4-D inversion), settings of
    input/input_ampt.yaml (t0 = 1, dt = 0.05, eta/s = 0.12, constant zeta/s = 0.08, diffusion path on with kappa = 0,
    `source` disabled). Profile: the config-3 transverse event x a plateau (|eta| < 2) with Gaussian wings (sigma = 1);
    rho_B from forward/backward shifted Gaussians in eta, rho_Q = 0.4 rho_B, and a sign-changing strangeness density
    rho_S = 0.1 rho_B sin(0.7 x + 0.3) (net ~ 0; SURVEY's rho_S = 0 EXACTLY makes the reference's own output depend on
    round-off: EquationOfState::dwdS switches on |S| > 1e-25 evaluated at the root-finder's iterate, src/eos.cpp:307-318). The lattice spacing d
    (dx = dy = d fm, d_eta = d) is chosen so that ~`n_target` particles survive the energy cut-off, and h_T = h_eta = 2 d
    keeps the neighbour count of input_ampt.yaml's h / spacing = 2 (N_ngb ~ 268). `window` < 1: central part of the same
    event at the same spacing."""
    # volume above the cut-off measured on a coarse lattice, then the spacing that gives n_target
    def profile(xs, es):
        e2 = trento_like_profile(xs, xs, seed=seed)
        ae = np.abs(es)
        fe = np.where(ae < 2.0, 1.0, np.exp(-(ae - 2.0) ** 2 / 2.0))
        return e2[:, :, None] * fe[None, None, :]
    dc = 0.3
    xc, ec = np.arange(-40, 41) * dc, np.arange(-20, 21) * dc
    vol = float((profile(xc, ec) > 0.15).sum()) * dc ** 3
    d = (vol / n_target) ** (1.0 / 3.0)
    mx, me = int(round(12.0 * window / d)), int(round(6.0 * window / d))
    xs, es = np.arange(-mx, mx + 1) * d, np.arange(-me, me + 1) * d
    e3 = profile(xs, es)
    keep = e3 > 0.15
    ix, iy, ie = np.nonzero(keep)
    x, y, eta, eps = xs[ix], xs[iy], es[ie], e3[keep]
    nB = 0.15 * (eps / 0.1973269804) ** 0.75 * (np.exp(-(eta - 1.5) ** 2 / 2.0) + np.exp(-(eta + 1.5) ** 2 / 2.0))  # fm^-3
    settings = dict(_COMMON, **_TABLE_COMMON, dim=3, t0=1.0, dt=0.05, hT=2 * d, hEta=2 * d, stepx=d, stepy=d, stepEta=d,
                    xmin=-(mx * d), ymin=-(mx * d), etamin=-(me * d), e_cutoff=0.15,
                    baryon_charge_enabled=True, strange_charge_enabled=True, electric_charge_enabled=True,
                    etaMode="constant", constant_eta_over_s=0.12, shearRelaxMode="default", use_vorticity=False,
                    delta_pipi_mode="default", tau_pipi_mode="disabled", lambda_piPi_mode="disabled", phi6_mode="disabled",
                    phi7_mode="disabled", input_initial_shear=False,
                    zetaMode="constant", constant_zeta_over_s=0.08, cs2_dependent_zeta_A=1.67552, cs2_dependent_zeta_p=2.0,
                    bulk_from_trace=False, bulkRelaxMode="default", modulate_zeta_with_tanh=False,
                    delta_PiPi_mode="israel-stewart", lambda_Pipi_mode="default", phi1_mode="disabled", phi3_mode="disabled",
                    diffusionMode="constant_over_T2", diffusionRelaxMode="constant_over_T")
    ic = dict(r=np.stack([x, y, eta], 1), e=eps, u=np.zeros((x.size, 3)), rhoB=nB, rhoS=0.1 * nB * np.sin(0.7 * x + 0.3), rhoQ=0.4 * nB,
              bulk=None, shv6=None)
    return settings, ic


WORKLOADS = {
    "config1_ideal_gubser_2p1d": config1_ideal_gubser,
    "config2_viscous_gubser_bsq_2p1d": config2_viscous_gubser_bsq,
    "config3_trento_pbpb_table_2p1d": config3_trento_pbpb,
    "config4_milne_bsq_table_3p1d": config4_milne_bsq,
    "config5_cartesian_blob_3p1d": config5_cartesian_blob,
}


def read_ccake(path, settings, fmt="ccake"):
    """The reference's initial-condition text files through the library's parallel reader (ccb_ic_read, csrc/ic_reader.cpp):
    `ccake` format (BSQHydro::read_ccake, src/BSQHydro.cpp:207-425) or `iccing` (BSQHydro::read_ICCING, :84-145). 8 M rows parse in
    seconds. Returns (settings updated with the header, ic dict for SPHWorkstation.upload_initial_conditions)."""
    import ctypes as C

    from . import _capi

    L = _capi.lib()
    s = dict(settings)
    h = C.c_void_p()
    rc = L.ccb_ic_read(str(path).encode(), 0 if fmt == "ccake" else 1, int(bool(s.get("input_as_entropy", False))), float(s.get("e_cutoff", 0.0)),
                       int(bool(s.get("baryon_charge_enabled", True))), int(bool(s.get("strange_charge_enabled", True))),
                       int(bool(s.get("electric_charge_enabled", True))), int(bool(s.get("input_initial_diffusion", False))), C.byref(h))
    if rc:
        raise ValueError("ccb_ic_read(%s): cannot open or parse the file (status %d)" % (path, rc))
    try:
        hdr = np.zeros(6)
        L.ccb_ic_file_header(h, hdr.ctypes.data_as(C.POINTER(C.c_double)))
        s["stepx"], s["stepy"], s["stepEta"], s["xmin"], s["ymin"], s["etamin"] = [float(x) for x in hdr]
        v = L.ccb_ic_file_view(h).contents
        n = int(v.n)

        def arr(ptr, w):
            return np.ctypeslib.as_array(ptr, shape=(n * w,)).reshape(n, w).copy() if n else np.zeros((0, w))

        D = int(s.get("dim", 2))
        ic = dict(r=arr(v.r, 3)[:, :D].copy(), e=arr(v.e_or_s, 1)[:, 0], rhoB=arr(v.rhoB, 1)[:, 0], rhoS=arr(v.rhoS, 1)[:, 0], rhoQ=arr(v.rhoQ, 1)[:, 0],
                  u=arr(v.u, 3)[:, :D].copy(), bulk=arr(v.bulk, 1)[:, 0], shv6=arr(v.shv6, 6))
        if bool(v.diffusion):
            ic["diffusion"] = arr(v.diffusion, 12).reshape(n, 3, 4)
    finally:
        L.ccb_ic_file_free(h)
    return s, ic


def read_ccake_python(path, settings):
    """Line-by-line Python twin of read_ccake (the cross-check of the parallel reader in the tests; unusable at scale).
    The reference's `ccake` initial-condition text format (BSQHydro::read_ccake, src/BSQHydro.cpp:207-425): one header line
    `# <ignored> stepx stepy stepEta <ignored> xmin ymin etamin`, then rows
    `x y eta e|s rhoB rhoS rhoQ ux uy ueta bulk pixx pixy pixeta piyy piyeta pietaeta [qB0..3 qS0..3 qQ0..3]`
    (the twelve currents when settings["input_initial_diffusion"]). Rows below the energy cut-off are dropped exactly as the
    reference drops them (for entropy input through its ideal pion-gas estimate). Returns (settings updated with the header,
    ic dict for SPHWorkstation.upload_initial_conditions)."""
    s = dict(settings)
    with_q = bool(s.get("input_initial_diffusion", False))
    rows = []
    header_done = False
    with open(path) as f:
        for line in f:
            if not header_done:
                tok = line.split("#", 1)[1].split() if "#" in line else line.split()
                s["stepx"], s["stepy"], s["stepEta"] = float(tok[1]), float(tok[2]), float(tok[3])
                s["xmin"], s["ymin"], s["etamin"] = float(tok[5]), float(tok[6]), float(tok[7])
                header_done = True
                continue
            t = line.strip()
            if not t or t.startswith("#"):
                continue
            rows.append([float(v) for v in t.split()[: 29 if with_q else 17]])
    a = np.array(rows, dtype=np.float64).reshape(len(rows), 29 if with_q else 17)
    e_or_s = a[:, 3]
    cut = float(s.get("e_cutoff", 0.0)) / HBARC_GEVFM
    if s.get("input_as_entropy", False):
        C = 3.0 * math.pi ** 2 / 90.0
        keep = (3.0 * C) * (1.0 / (4.0 * C)) ** (4.0 / 3.0) * e_or_s ** (4.0 / 3.0) > cut
    else:
        keep = e_or_s / HBARC_GEVFM > cut
    a = a[keep]
    D = int(s.get("dim", 2))
    ic = dict(r=np.ascontiguousarray(a[:, :3] if D == 3 else a[:, :2]), e=a[:, 3].copy(), rhoB=a[:, 4].copy(), rhoS=a[:, 5].copy(),
              rhoQ=a[:, 6].copy(), u=np.ascontiguousarray(a[:, 7:10] if D == 3 else a[:, 7:9]), bulk=a[:, 10].copy(),
              shv6=np.ascontiguousarray(a[:, 11:17]))
    if with_q:
        ic["diffusion"] = np.ascontiguousarray(a[:, 17:29]).reshape(-1, 3, 4)
    return s, ic


def write_ccake(path, settings, ic):
    """Write an IC dict in the same format (used by the tests and to hand a synthetic workload to the reference binary)."""
    n = ic["r"].shape[0]
    z = np.zeros(n)
    r3 = np.zeros((n, 3)); r3[:, : ic["r"].shape[1]] = ic["r"]
    u3 = np.zeros((n, 3))
    if ic.get("u") is not None:
        u3[:, : ic["u"].shape[1]] = ic["u"]
    cols = [r3, ic["e"][:, None]] + [(z if ic.get(k) is None else ic[k])[:, None] for k in ("rhoB", "rhoS", "rhoQ")] + [u3]
    cols.append((z if ic.get("bulk") is None else ic["bulk"])[:, None])
    cols.append(np.zeros((n, 6)) if ic.get("shv6") is None else ic["shv6"])
    if ic.get("diffusion") is not None:
        cols.append(np.asarray(ic["diffusion"]).reshape(n, 12))
    with open(path, "w") as f:
        f.write("#0 %.17g %.17g %.17g 0 %.17g %.17g %.17g\n" % (settings["stepx"], settings["stepy"], settings["stepEta"], settings["xmin"],
                                                              settings["ymin"], settings["etamin"]))
        np.savetxt(f, np.concatenate(cols, 1), fmt="%.17g")
