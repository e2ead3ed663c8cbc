"""TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.

ctypes front-end of ``oracle/_ref/libccake_ref.so``: the UNMODIFIED CCAKE
reference classes compiled against the shim headers in ``oracle/shim`` (see
``oracle/Makefile`` and ``oracle/ref_driver.cpp``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.

Member blocks follow ``CabanaParticle`` (reference ``include/particle.h:54-100``) and
the enums of ``include/hydrodynamic_info.h:166-309``,
``include/thermodynamic_info.h:35-71`` and ``include/densities.h:24-33``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libccake_ref.so")

# particle_info::particle_data (include/particle.h:79-103): name -> (index, per-particle shape)
MEMBERS = {
    "hydro_space_matrix": (0, (25, 3, 3)),
    "hydro_vector": (1, (20, 3)),
    "hydro_diffusion_aux_matrix": (2, (2, 3, 9)),
    "hydro_scalar": (3, (43,)),
    "hydro_shear_aux_vector": (4, (4, 2, 3)),
    "hydro_shear_aux_matrix": (5, (3, 2, 9)),
    "thermo": (6, (31,)),
    "hydro_spacetime_matrix": (7, (2, 4, 4)),
    "hydro_diffusion": (8, (1, 3, 4)),
    "input": (9, (5,)),
    "smoothed": (10, (5,)),
    "extensive": (11, (5,)),
    "d_dt_extensive": (12, (5,)),
    "sph_mass": (13, (5,)),
    "position": (14, (3,)),
    "efcheck": (15, ()),
    "contribution_to_total_E": (16, ()),
    "contribution_to_total_dEz": (17, ()),
    "contribution_to_total_Ez": (18, ()),
}

HYDRO_SCALAR = ["t", "causality", "bulk", "extensive_bulk", "a", "rho_Q_ext", "rho_S_ext", "rho_B_ext", "tau_Pi",
                "tau_pi", "zeta_Pi", "sigma_lab", "sigma", "gamma", "theta", "delta_PiPi", "delta_pipi",
                "lambda_Pipi", "phi1", "phi3", "phi6", "phi7", "eta_pi", "tau_pipi", "lambda_piPi", "tmunu_trace",
                "div_u", "d_extensive_bulk_dt", "F_extensive_bulk", "F_shv_nabla_u", "F_extensive_entropy",
                "shv_nabla_u", "j0_ext", "shear_knudsen", "bulk_knudsen", "inverse_reynolds_shear",
                "inverse_reynolds_bulk", "knudsen_diffusion_B", "knudsen_diffusion_S", "knudsen_diffusion_Q",
                "inverse_reynolds_diffusion_B", "inverse_reynolds_diffusion_S", "inverse_reynolds_diffusion_Q"]
HYDRO_VECTOR = ["v", "u", "gradP", "gradE", "gradBulk", "divshear", "gradshear", "div_qa", "grad_qa",
                "M_extensive_bulk", "M_shv_nabla_u", "M_extensive_entropy", "R_extensive_entropy",
                "R_extensive_bulk", "F_u", "F_extensive_N", "F_0i_shear", "j_ext", "du_dt", "F_q0_a"]
HYDRO_SPACE_MATRIX = ["gradV", "grad_alpha_a", "M_u", "M_extensive_N", "R_extensive_N", "R_u", "R_0i_shear",
                      "M_0i_shear", "tau_q", "kappa_q", "delta_qq", "l_qpi", "l_qPi", "tau_qpi", "tau_qPi",
                      "lambda_qq", "lambda_qpi", "lambda_qPi", "phi_4", "phi_5", "M_q0_j_a", "R_q0_a_b",
                      "F_extensive_diffusion_ib", "d_extensive_q_dt", "extensive_diffusion"]
HYDRO_SPACETIME_MATRIX = ["shv", "sigma_tensor"]
HYDRO_SHEAR_AUX_VECTOR = ["F_extensive_shear", "F_sigma_tensor", "extensive_shv", "d_extensive_shv_dt"]
HYDRO_SHEAR_AUX_MATRIX = ["M_extensive_shear", "M_sigma_tensor", "R_extensive_shear"]
THERMO = ["T", "muB", "muS", "muQ", "e", "s", "rhoB", "rhoS", "rhoQ", "p", "cs2", "w", "dwds", "dwdB", "dwdS",
          "dwdQ", "dalpha_Bds", "dalpha_BdB", "dalpha_BdS", "dalpha_BdQ", "dalpha_Sds", "dalpha_SdB",
          "dalpha_SdS", "dalpha_SdQ", "dalpha_Qds", "dalpha_QdB", "dalpha_QdS", "dalpha_QdQ", "chiBB", "chiBB0",
          "eos_type_is_table"]
DENSITIES = ["s", "e", "rhoB", "rhoS", "rhoQ"]

_SUB = {
    "hydro_scalar": HYDRO_SCALAR, "hydro_vector": HYDRO_VECTOR, "hydro_space_matrix": HYDRO_SPACE_MATRIX,
    "hydro_spacetime_matrix": HYDRO_SPACETIME_MATRIX, "hydro_shear_aux_vector": HYDRO_SHEAR_AUX_VECTOR,
    "hydro_shear_aux_matrix": HYDRO_SHEAR_AUX_MATRIX, "thermo": THERMO, "input": DENSITIES, "smoothed": DENSITIES,
    "extensive": DENSITIES, "d_dt_extensive": DENSITIES, "sph_mass": DENSITIES,
}

PHASES = {"neighbour_list": 1, "gamma_velocity": 2, "smooth_fields": 3, "thermodynamics": 4, "reset_pi_tensor": 5,
          "regulate_diffusion": 6, "viscosities": 8, "smooth_gradients": 9, "time_derivatives": 10}


# Settings members the driver's key=value parser accepts (oracle/ref_driver.cpp: parse_settings)
SETTING_KEYS = frozenset([
    "dim", "coordinate_system", "t0", "dt", "hT", "hEta", "rk_order", "factorized_kernel", "e_cutoff", "input_as_entropy",
    "stepx", "stepy", "stepEta", "xmin", "ymin", "etamin", "max_tau", "eos_type", "eos_path", "normalize_by_T",
    "Freeze_Out_Temperature", "baryon_charge_enabled", "strange_charge_enabled", "electric_charge_enabled", "etaMode",
    "constant_eta_over_s", "shearRelaxMode", "use_vorticity", "delta_pipi_mode", "tau_pipi_mode", "lambda_piPi_mode",
    "phi6_mode", "phi7_mode", "input_initial_shear", "zetaMode", "constant_zeta_over_s", "cs2_dependent_zeta_A",
    "cs2_dependent_zeta_p", "bulk_from_trace", "bulkRelaxMode", "modulate_zeta_with_tanh", "delta_PiPi_mode",
    "lambda_Pipi_mode", "phi1_mode", "phi3_mode", "diffusionMode", "diffusionRelaxMode", "print_conservation_status",
    "prohibit_unstable_cs2", "prohibit_acausal_cs2", "restrict_mu_T_ratios", "C_B", "critical_scaling_bulk", "critical_scaling_diffusion",
    "critical_point_T", "critical_point_muB", "critical_gaussian_width_T", "critical_gaussian_width_muB", "diffusion_regulator_enabled",
    "diffusion_regulator_chi0", "diffusion_regulator_e0", "diffusion_regulator_xi0", "diffusion_regulator_rq_max",
    "online_inverter_enabled", "preinverted_eos_path", "particlization_enabled", "IC_type", "freezein_tol", "freezein_max_iter"])


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long)
        L.ccake_ref_create.restype = C.c_void_p
        L.ccake_ref_create.argtypes = [C.c_char_p]
        L.ccake_ref_destroy.argtypes = [C.c_void_p]
        L.ccake_ref_add_particles.argtypes = [C.c_void_p, C.c_long] + [dp] * 9
        L.ccake_ref_add_freezein.argtypes = [C.c_void_p, C.c_long] + [dp] * 5
        for f in ("initialize", "get_time_derivatives"):
            getattr(L, "ccake_ref_" + f).argtypes = [C.c_void_p]
        L.ccake_ref_phase.argtypes = [C.c_void_p, C.c_int]
        L.ccake_ref_allocate_raw.argtypes = [C.c_void_p, C.c_long]
        L.ccake_ref_advance.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.ccake_ref_advance_par04.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.ccake_ref_h5_reset.argtypes = [C.c_char_p]
        L.ccake_ref_h5_reset.restype = None
        L.ccake_ref_h5_put_dataset.argtypes = [C.c_char_p, C.c_char_p, dp, C.c_int, lp]
        L.ccake_ref_h5_put_attr.argtypes = [C.c_char_p, C.c_char_p, C.c_double]
        L.ccake_ref_interp_fill.argtypes = [C.c_char_p, C.c_long, dp, dp]
        L.ccake_ref_interp_query.argtypes = [C.c_char_p, C.c_int, C.c_long, dp, dp, dp, ip, C.c_double, dp]
        L.ccake_ref_eos_update_s.argtypes = [C.c_void_p, C.c_long, dp, dp, dp, ip]
        L.ccake_ref_freeze_out.argtypes = [C.c_void_p]
        L.ccake_ref_freeze_results.argtypes = [C.c_void_p, dp, ip]
        L.ccake_ref_set_threads.argtypes = [C.c_int]
        L.ccake_ref_set_threads.restype = None
        L.ccake_ref_profile_report.argtypes = [C.c_char_p, C.c_long]
        L.ccake_ref_profile_report.restype = C.c_long
        L.ccake_ref_n.restype = C.c_long
        L.ccake_ref_n.argtypes = [C.c_void_p]
        L.ccake_ref_time.restype = C.c_double
        L.ccake_ref_time.argtypes = [C.c_void_p]
        L.ccake_ref_set_time.argtypes = [C.c_void_p, C.c_double]
        L.ccake_ref_member_doubles.restype = C.c_long
        L.ccake_ref_member_doubles.argtypes = [C.c_void_p, C.c_int]
        L.ccake_ref_get_member.argtypes = [C.c_void_p, C.c_int, dp]
        L.ccake_ref_set_member.argtypes = [C.c_void_p, C.c_int, dp]
        L.ccake_ref_get_ints.argtypes = [C.c_void_p, C.c_int, ip]
        L.ccake_ref_set_ints.argtypes = [C.c_void_p, C.c_int, ip]
        L.ccake_ref_neighbours.restype = C.c_long
        L.ccake_ref_neighbours.argtypes = [C.c_void_p, lp, ip, ip, ip]
        L.ccake_ref_totals.argtypes = [C.c_void_p, dp]
        L.ccake_ref_omp_threads.restype = C.c_int
        L.ccake_ref_eos_names.argtypes = [C.c_void_p, ip]
        L.ccake_ref_eos_thermo.argtypes = [C.c_void_p, C.c_char_p, C.c_long, dp, dp]
        _lib = L
    return _lib


def omp_threads() -> int:
    """Threads the reference's parallel loops run on (omp_get_max_threads of the build)."""
    return int(lib().ccake_ref_omp_threads())


def set_threads(n: int) -> None:
    """omp_set_num_threads for the reference's parallel loops (torchrun pins OMP_NUM_THREADS=1 in its workers)."""
    lib().ccake_ref_set_threads(int(n))


def profile_table() -> dict:
    """The reference's own CCAKE_PROFILE table (include/phase_profiler.h) parsed into {phase: (total_s, calls)}.
    Empty unless CCAKE_PROFILE was set in the environment before this library was loaded."""
    need = lib().ccake_ref_profile_report(None, 0)
    buf = C.create_string_buffer(int(need) + 8)
    lib().ccake_ref_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        f = line.split()
        if len(f) == 5 and f[4].endswith("%"):
            try:
                out[f[0]] = (float(f[1]), int(f[2]))
            except ValueError:
                pass
    return out


# ---- pre-inverted EoS tables for the reference's EoS_Interpolator (in-memory HDF5 stand-in, oracle/shim/H5Cpp.h) ----------
PRE_VARS = ["T", "muB", "muQ", "muS", "e", "p", "cs2", "dwds", "dwdB", "dwdQ", "dwdS"]   # ccake::eos_variables order


def _put(path, name, arr):
    a = np.ascontiguousarray(arr, dtype=np.float64)
    dims = (C.c_long * a.ndim)(*a.shape)
    assert lib().ccake_ref_h5_put_dataset(path.encode(), name.encode(), _dp(a), a.ndim, dims) == 0


def h5_put_tiles(path, tiles):
    """Register a TILED pre-inverted table (the groups utilities/generate_eos_table.cpp writes: table_<es>_<eB>_<eS>_<eQ> with the
    axis attributes and one dataset per field) under `path`. tiles: list of dicts exponent[4], min[4], max[4], step[4], vars[11, N0..N3]."""
    lib().ccake_ref_h5_reset(path.encode())
    for t in tiles:
        g = "table_%d_%d_%d_%d" % tuple(t["exponent"])
        for k, ax in enumerate(("s", "rhoB", "rhoS", "rhoQ")):
            for what, key in (("%s_min" % ax, "min"), ("%s_max" % ax, "max"), (("ds" if ax == "s" else "d" + ax), "step")):
                assert lib().ccake_ref_h5_put_attr(path.encode(), ("%s@%s" % (g, what)).encode(), float(t[key][k])) == 0
            assert lib().ccake_ref_h5_put_attr(path.encode(), ("%s@%s_Axis" % (g, ax)).encode(), float(k)) == 0
        for v, name in enumerate(PRE_VARS):
            _put(path, g + "/" + name, t["vars"][v])


def h5_put_flat(path, flat):
    """Register a FLAT (s, nB) table (src/eos_interpolator.cpp:262-340): /dimensions + /data/<field> 2-D datasets, MeV units."""
    lib().ccake_ref_h5_reset(path.encode())
    Ns, NB = flat["s"].shape
    _put(path, "dimensions", np.array([Ns, NB], dtype=np.float64))
    for k in ("s", "nB", "T", "muB", "muS", "muQ", "e", "p", "cs2", "chiTT", "chiTB", "chiBB"):
        _put(path, "data/" + k, flat[k])


def interp_fill(path, dens):
    """EoS_Interpolator::fill_thermodynamics on probe particles: dens (n, 4) = s, rhoB, rhoS, rhoQ -> thermo (n, 31)."""
    d = np.ascontiguousarray(dens, dtype=np.float64).reshape(-1, 4)
    out = np.zeros((d.shape[0], 31))
    lib().ccake_ref_interp_fill(path.encode(), d.shape[0], _dp(d), _dp(out))
    return out


def interp_query(path, a, rhoB, energy=False, T_fo=150.0 / 197.3269804):
    """query_thermo_host (a = s) / invert_e_to_s_host (a = e) of the flat table -> ((n, 12) T muB muS muQ e p cs2 s rhoB w dwds dwdB,
    ok (n,), efreeze_at_T(T_fo))."""
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(rhoB, dtype=np.float64)
    out = np.zeros((a.size, 12)); ok = np.zeros(a.size, dtype=np.int32); ef = C.c_double()
    lib().ccake_ref_interp_query(path.encode(), int(bool(energy)), a.size, _dp(a), _dp(b), _dp(out), ok.ctypes.data_as(C.POINTER(C.c_int)),
                                 float(T_fo), C.byref(ef))
    return out, ok.astype(bool), ef.value


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefSim:
    """One reference simulation (SystemState + SPHWorkstation) built from a settings dict."""

    def __init__(self, settings: dict):
        kv = "\n".join(f"{k}={_fmt(v)}" for k, v in settings.items())
        self.h = lib().ccake_ref_create(kv.encode())
        if not self.h:
            raise ValueError("ccake_ref_create rejected the settings")
        self.settings = dict(settings)

    def close(self):
        if self.h:
            lib().ccake_ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_particles(self, r, e_or_s, rhoB=None, rhoS=None, rhoQ=None, u=None, bulk=None, shv6=None, diffusion=None):
        r = np.ascontiguousarray(r, dtype=np.float64)
        n = r.shape[0]
        r3 = np.zeros((n, 3))
        r3[:, : r.shape[1]] = r
        z = np.zeros(n)
        u3 = np.zeros((n, 3))
        if u is not None:
            u = np.asarray(u, dtype=np.float64)
            u3[:, : u.shape[1]] = u
        arrs = [r3, np.ascontiguousarray(e_or_s, dtype=np.float64)]
        for a in (rhoB, rhoS, rhoQ):
            arrs.append(z if a is None else np.ascontiguousarray(a, dtype=np.float64))
        arrs.append(u3)
        arrs.append(z if bulk is None else np.ascontiguousarray(bulk, dtype=np.float64))
        arrs.append(np.zeros((n, 6)) if shv6 is None else np.ascontiguousarray(shv6, dtype=np.float64))
        q = None if diffusion is None else np.ascontiguousarray(np.asarray(diffusion, dtype=np.float64).reshape(n, 12))
        lib().ccake_ref_add_particles(self.h, n, *[_dp(a) for a in arrs], None if q is None else _dp(q))

    def add_freezein(self, r, Ttau, jB, jS, jQ):
        """IC_type "freezein": per cell the position and the raw T^{tau mu} [GeV/fm^3], j_B^mu, j_S^mu, j_Q^mu [1/fm^3] (n x 4 each), as
        BSQHydro::read_freezein stages them; initialize() then runs SPHWorkstation::freeze_in_match."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        n = r.shape[0]
        r3 = np.zeros((n, 3)); r3[:, : r.shape[1]] = r
        arrs = [r3] + [np.ascontiguousarray(a, dtype=np.float64).reshape(n, 4) for a in (Ttau, jB, jS, jQ)]
        lib().ccake_ref_add_freezein(self.h, n, *[_dp(a) for a in arrs])

    def initialize(self):
        lib().ccake_ref_initialize(self.h)

    def allocate_raw(self, n):
        """n blank particles; inject state with set_member/put, then run phases."""
        lib().ccake_ref_allocate_raw(self.h, int(n))

    def phase(self, name_or_id):
        k = PHASES[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        if lib().ccake_ref_phase(self.h, k):
            raise ValueError(f"unknown phase {name_or_id}")

    def get_time_derivatives(self):
        lib().ccake_ref_get_time_derivatives(self.h)

    def advance(self, dt, rk_order=2):
        lib().ccake_ref_advance(self.h, float(dt), int(rk_order))

    def advance_par04(self, dt, rk_order=2):
        """The same step with phase 04 OpenMP-parallelised (ref_driver.cpp: thermodynamics_openmp); everything else is the
        unmodified reference."""
        lib().ccake_ref_advance_par04(self.h, float(dt), int(rk_order))

    @property
    def n(self):
        return lib().ccake_ref_n(self.h)

    @property
    def t(self):
        return lib().ccake_ref_time(self.h)

    @t.setter
    def t(self, v):
        lib().ccake_ref_set_time(self.h, float(v))

    def member(self, name):
        idx, shape = MEMBERS[name]
        out = np.empty((self.n,) + shape, dtype=np.float64)
        assert out.size == lib().ccake_ref_member_doubles(self.h, idx), name
        lib().ccake_ref_get_member(self.h, idx, _dp(out))
        return out

    def set_member(self, name, arr):
        idx, shape = MEMBERS[name]
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape((self.n,) + shape)
        lib().ccake_ref_set_member(self.h, idx, _dp(a))

    def get(self, member, field):
        """e.g. get('hydro_scalar', 'sigma_lab') -> (n,), get('hydro_vector', 'gradP') -> (n,3)."""
        return self.member(member)[:, _SUB[member].index(field)]

    def put(self, member, field, values):
        a = self.member(member)
        a[:, _SUB[member].index(field)] = values
        self.set_member(member, a)

    def ints(self, which):
        k = {"id": 0, "btrack": 1, "freeze": 2}[which]
        out = np.empty(self.n, dtype=np.int32)
        lib().ccake_ref_get_ints(self.h, k, out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def set_ints(self, which, vals):
        k = {"id": 0, "btrack": 1, "freeze": 2}[which]
        a = np.ascontiguousarray(vals, dtype=np.int32)
        lib().ccake_ref_set_ints(self.h, k, a.ctypes.data_as(C.POINTER(C.c_int)))

    def neighbours(self):
        """CSR neighbour list of the last rebuild: (offsets[n+1], idx[nnz], cell_of[n], ncell[3])."""
        n = self.n
        nnz = lib().ccake_ref_neighbours(self.h, None, None, None, None)
        off = np.empty(n + 1, dtype=np.int64)
        idx = np.empty(max(nnz, 0), dtype=np.int32)
        cell = np.empty(n, dtype=np.int32)
        nc = np.empty(3, dtype=np.int32)
        ip = C.POINTER(C.c_int)
        lib().ccake_ref_neighbours(self.h, off.ctypes.data_as(C.POINTER(C.c_long)), idx.ctypes.data_as(ip),
                                   cell.ctypes.data_as(ip), nc.ctypes.data_as(ip))
        return off, idx, cell, nc

    def eos_index(self):
        """Per particle (id order): which member of the EoS chain it last used (0 table, 1 tanh_conformal,
        2 conformal, 3 conformal_diagonal; -1 = "default"/unset)."""
        out = np.full(self.n, -1, dtype=np.int32)
        lib().ccake_ref_eos_names(self.h, out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def eos_thermo(self, name, points):
        """EoS_base::get_full_thermo of the named chain member at points (T, muB, muQ, muS) [fm^-1]: n x 17."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 4)
        out = np.zeros((pts.shape[0], 17))
        rc = lib().ccake_ref_eos_thermo(self.h, name.encode(), pts.shape[0], _dp(pts), _dp(out))
        if rc:
            raise KeyError(name)
        return out

    def eos_update_s(self, dens, seed):
        """EquationOfState::tbqs(seed, "default") + update_s(dens) + set_thermo per row, as utilities/generate_eos_table.cpp:295-311
        does per node: dens (n, 4) = s, rhoB, rhoS, rhoQ; seed (n, 4) = T, muB, muQ, muS -> (thermo (n, 31), ok (n,))."""
        d = np.ascontiguousarray(dens, dtype=np.float64).reshape(-1, 4)
        sd = np.ascontiguousarray(seed, dtype=np.float64).reshape(-1, 4)
        out = np.zeros((d.shape[0], 31)); ok = np.zeros(d.shape[0], dtype=np.int32)
        lib().ccake_ref_eos_update_s(self.h, d.shape[0], _dp(d), _dp(sd), _dp(out), ok.ctypes.data_as(C.POINTER(C.c_int)))
        return out, ok.astype(bool)

    def freeze_out(self):
        """SPHWorkstation::freeze_out_particles (advance() calls it itself when particlization_enabled)."""
        lib().ccake_ref_freeze_out(self.h)

    def freeze_results(self):
        """FreezeOut::results: (n, 64) rows in ResultsTypes member order (= the CCB_FO_* offsets), print flags (n,)."""
        out = np.zeros((self.n, 64)); pr = np.zeros(self.n, dtype=np.int32)
        lib().ccake_ref_freeze_results(self.h, _dp(out), pr.ctypes.data_as(C.POINTER(C.c_int)))
        return out, pr.astype(bool)

    def totals(self):
        out = np.zeros(6)
        lib().ccake_ref_totals(self.h, _dp(out))
        return dict(zip(["S", "B", "Sq", "Q", "E", "Ez"], out))


def _fmt(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, float):
        return repr(v)
    return str(v)
