"""ctypes view of ``include/ccake_b200.h`` (the C ABI of the CUDA library).

This is the only place where the shared library is loaded.  There is no fallback:
if ``libccake_b200.so`` is missing the import of :func:`lib` raises, and every
compute entry point of the library itself fails with ``CCB_ERR_CUDA`` when no
device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libccake_b200.so")

CCB_OK, CCB_ERR_INVALID, CCB_ERR_CUDA, CCB_ERR_NEG_ENTROPY, CCB_ERR_EOS_NO_ROOT, CCB_ERR_OUT_OF_GRID, CCB_ERR_STATE, CCB_ERR_COMM = range(8)
COORD_HYPERBOLIC, COORD_CARTESIAN = 0, 1
EOS_CONFORMAL, EOS_TABLE, EOS_TANH_CONFORMAL, EOS_CONFORMAL_DIAGONAL = 0, 1, 2, 3


class Config(C.Structure):
    """``ccb_config`` (include/ccake_b200.h)."""

    _fields_ = [
        ("dim", C.c_int), ("coordinate_system", C.c_int), ("rk_order", C.c_int), ("factorized_kernel", C.c_int),
        ("t0", C.c_double), ("hT", C.c_double), ("hEta", C.c_double),
        ("xmin", C.c_double), ("ymin", C.c_double), ("etamin", C.c_double),
        ("stepx", C.c_double), ("stepy", C.c_double), ("stepEta", C.c_double),
        ("using_shear", C.c_int), ("using_diffusion", C.c_int), ("using_vorticity", C.c_int),
        ("input_as_entropy", C.c_int), ("input_initial_shear", C.c_int), ("bulk_from_trace", C.c_int),
        ("shear_mode", C.c_int), ("shear_relax_mode", C.c_int), ("zeta_mode", C.c_int), ("bulk_relax_mode", C.c_int),
        ("diffusion_mode", C.c_int), ("diffusion_relax_mode", C.c_int),
        ("delta_pipi_mode", C.c_int), ("tau_pipi_mode", C.c_int), ("lambda_piPi_mode", C.c_int), ("phi6_mode", C.c_int),
        ("phi7_mode", C.c_int),
        ("delta_PiPi_mode", C.c_int), ("lambda_Pipi_mode", C.c_int), ("phi1_mode", C.c_int), ("phi3_mode", C.c_int),
        ("modulate_zeta_with_tanh", C.c_int), ("critical_scaling_bulk", C.c_int),
        ("constant_eta_over_s", C.c_double), ("constant_zeta_over_s", C.c_double), ("cs2_dependent_zeta_A", C.c_double),
        ("cs2_dependent_zeta_p", C.c_double), ("C_B", C.c_double),
        ("constant_kappa_over_T2", C.c_double * 9),
        ("critical_scaling_diffusion", C.c_int),
        ("critical_point_T", C.c_double), ("critical_point_muB", C.c_double),
        ("critical_gaussian_width_T", C.c_double), ("critical_gaussian_width_muB", C.c_double),
        ("diffusion_regulator_enabled", C.c_int),
        ("diffusion_regulator_chi0", C.c_double), ("diffusion_regulator_e0", C.c_double),
        ("diffusion_regulator_xi0", C.c_double), ("diffusion_regulator_rq_max", C.c_double),
        ("eos_type", C.c_int),
        ("conformal_c", C.c_double), ("conformal_T0", C.c_double), ("conformal_muB0", C.c_double),
        ("conformal_muS0", C.c_double), ("conformal_muQ0", C.c_double),
        ("freeze_out_temperature_MeV", C.c_double),
        ("prohibit_unstable_cs2", C.c_int), ("prohibit_acausal_cs2", C.c_int), ("restrict_mu_T_ratios", C.c_int),
        ("device", C.c_int), ("rank", C.c_int), ("nranks", C.c_int), ("particlization_enabled", C.c_int), ("dt", C.c_double), ("online_inverter_enabled", C.c_int), ("slab_axis", C.c_int),
    ]


class EosTable(C.Structure):
    """``ccb_eos_table``."""

    _fields_ = [("n", C.c_int32 * 4), ("grid", C.POINTER(C.c_double)), ("fields", C.POINTER(C.c_double)), ("normalize_by_T", C.c_int)]


class PreTile(C.Structure):
    """``ccb_preinverted_tile``."""

    _fields_ = [("slot", C.c_int32 * 4), ("min", C.c_double * 4), ("step", C.c_double * 4), ("N", C.c_int32 * 4),
                ("log_spacing", C.c_int32 * 4), ("vars", C.POINTER(C.c_double))]


class PreFlat(C.Structure):
    """``ccb_preinverted_flat``."""

    _fields_ = [("Ns", C.c_int32), ("NB", C.c_int32)] + [(k, C.POINTER(C.c_double)) for k in
                                                           ("s", "nB", "T", "muB", "muS", "muQ", "e", "p", "cs2", "chiTT", "chiTB", "chiBB")]


class IC(C.Structure):
    """``ccb_ic``."""

    _fields_ = [("n", C.c_int64)] + [(k, C.POINTER(C.c_double)) for k in
                                       ("r", "e_or_s", "rhoB", "rhoS", "rhoQ", "u", "bulk", "shv6")] + [("id", C.POINTER(C.c_int32)),
                                                                                           ("diffusion", C.POINTER(C.c_double))]


class Totals(C.Structure):
    _fields_ = [("S", C.c_double), ("B", C.c_double), ("Sq", C.c_double), ("Q", C.c_double), ("E", C.c_double),
                ("Ez", C.c_double), ("n_frozen", C.c_int64)]


# ccb_field, in header order: name -> (id, width)
_FIELD_NAMES = [
    ("position", 3), ("u", 3), ("extensive", 4), ("sph_mass", 4), ("extensive_bulk", 1), ("extensive_shv", 5),
    ("extensive_diffusion", 9), ("Ez", 1), ("tseed", 4), ("v", 3), ("gamma", 1), ("sigma_lab", 1), ("sigma", 1),
    ("smoothed", 4), ("thermo", 16), ("dalpha", 12), ("bulk", 1), ("shv", 16), ("diffusion", 12), ("transport", 13),
    ("gradP", 3), ("gradE", 3), ("gradBulk", 3), ("gradshear", 3), ("divshear", 3), ("gradV", 9), ("div_qa", 3),
    ("grad_qa", 3), ("grad_alpha", 9), ("du_dt", 3), ("d_extensive_dt", 4), ("d_bulk_dt", 1), ("d_shv_dt", 5),
    ("d_diffusion_dt", 9), ("theta", 1), ("shv_nabla_u", 1), ("dEz", 1), ("chiBB", 2),
]
FIELDS = {name: (i, w) for i, (name, w) in enumerate(_FIELD_NAMES)}
INT_FIELDS = {"id": 0, "btrack": 1, "freeze": 2, "cell": 3, "eos": 4}
PHASE_NAMES = ["01_neighbour_list", "02_gamma_velocity", "03_smooth_fields", "04_thermodynamics",
               "05_06_08_local_update", "09_smooth_gradients", "10_time_derivatives", "other", "01a_slab_migration", "01b_slab_column_tables"]

_lib = None


def lib():
    """Load ``libccake_b200.so`` (built by ``__graft_entry__.build()``) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` first "
                          "(the CUDA library is the only implementation of the hot path)")
    L = C.CDLL(LIB_PATH)
    ctx = C.c_void_p
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.ccb_create.argtypes = [C.POINTER(ctx), C.POINTER(Config)]
    L.ccb_destroy.argtypes = [ctx]
    L.ccb_last_error.argtypes = [ctx]
    L.ccb_last_error.restype = C.c_char_p
    L.ccb_field_width.argtypes = [C.c_int]
    L.ccb_allocate.argtypes = [ctx, C.c_int64]
    L.ccb_set_eos_table.argtypes = [ctx, C.POINTER(EosTable)]
    L.ccb_load_eos_table_dat.argtypes = [ctx, C.c_char_p, C.c_int]
    L.ccb_eos_chain_parameters.argtypes = [ctx, dp]
    L.ccb_eos_full_thermo.argtypes = [ctx, C.c_int, C.c_int64, dp, dp]
    L.ccb_eos_locate.argtypes = [ctx, C.c_int, C.c_int64, dp, dp, ip, dp, ip]
    L.ccb_upload_ic.argtypes = [ctx, C.POINTER(IC)]
    L.ccb_set_field.argtypes = [ctx, C.c_int, dp]
    L.ccb_get_field.argtypes = [ctx, C.c_int, dp]
    L.ccb_set_int_field.argtypes = [ctx, C.c_int, ip]
    L.ccb_get_int_field.argtypes = [ctx, C.c_int, ip]
    L.ccb_n.argtypes = [ctx]
    L.ccb_n.restype = C.c_int64
    L.ccb_time.argtypes = [ctx]
    L.ccb_time.restype = C.c_double
    L.ccb_set_time.argtypes = [ctx, C.c_double]
    L.ccb_grid_shape.argtypes = [ctx, ip]
    for f in ("build_cells", "smooth_fields", "update_thermodynamics", "local_update", "smooth_gradients",
              "time_derivatives", "get_time_derivatives", "synchronize"):
        getattr(L, "ccb_" + f).argtypes = [ctx]
    L.ccb_step.argtypes = [ctx, C.c_double]
    L.ccb_reduce.argtypes = [ctx, C.POINTER(Totals)]
    L.ccb_kernel_launches.argtypes = [ctx]
    L.ccb_kernel_launches.restype = C.c_int64
    L.ccb_profile_enable.argtypes = [ctx, C.c_int]
    L.ccb_profile_get.argtypes = [ctx, C.c_int, dp, C.POINTER(C.c_int64)]
    L.ccb_state_width.argtypes = [ctx]
    L.ccb_upload_state.argtypes = [ctx, dp]
    L.ccb_download_state.argtypes = [ctx, dp]
    L.ccb_measure_dfma_peak.argtypes = [C.c_int, dp]
    L.ccb_comm_unique_id.argtypes = [C.c_char_p]
    L.ccb_comm_init.argtypes = [ctx, C.c_char_p]
    L.ccb_comm_init_local.argtypes = [ctx, C.c_int]
    L.ccb_set_slab.argtypes = [ctx, C.c_int32, C.c_int32]
    L.ccb_grid_geometry.argtypes = [ctx, dp, dp, ip]
    L.ccb_migration_counts.argtypes = [ctx, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.ccb_get_field_device_order.argtypes = [ctx, C.c_int, dp]
    L.ccb_get_int_field_device_order.argtypes = [ctx, C.c_int, ip]
    L.ccb_set_preinverted_table.argtypes = [ctx, C.c_int32, C.POINTER(PreTile)]
    L.ccb_set_preinverted_flat.argtypes = [ctx, C.POINTER(PreFlat)]
    L.ccb_generate_preinverted_tile.argtypes = [ctx, ip, C.c_int32, dp, ip]
    L.ccb_ic_read.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.ccb_ic_file_view.argtypes = [C.c_void_p]
    L.ccb_ic_file_view.restype = C.POINTER(IC)
    L.ccb_ic_file_header.argtypes = [C.c_void_p, dp]
    L.ccb_ic_file_free.argtypes = [C.c_void_p]
    L.ccb_ic_file_free.restype = None
    L.ccb_efcheck.argtypes = [ctx, dp]
    L.ccb_profile_phase_name.argtypes = [C.c_int]
    L.ccb_profile_phase_name.restype = C.c_char_p
    L.ccb_freeze_out.argtypes = [ctx, C.POINTER(C.c_int64)]
    L.ccb_freeze_out_results.argtypes = [ctx, dp, ip]
    L.ccb_debug_neighbour_list.argtypes = [ctx, C.POINTER(C.c_int64), ip, C.c_int64]
    L.ccb_stream.argtypes = [ctx]
    L.ccb_stream.restype = C.c_void_p
    L.ccb_device_field.argtypes = [ctx, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    _lib = L
    return L


# Symbols include/ccake_b200.h declares; tests check that the shared library exports each of them.
EXPORTED = ["ccb_create", "ccb_destroy", "ccb_last_error", "ccb_field_width", "ccb_allocate", "ccb_upload_ic",
            "ccb_set_field", "ccb_get_field", "ccb_set_int_field", "ccb_get_int_field", "ccb_n", "ccb_time",
            "ccb_set_time", "ccb_grid_shape", "ccb_build_cells", "ccb_smooth_fields", "ccb_update_thermodynamics",
            "ccb_local_update", "ccb_smooth_gradients", "ccb_time_derivatives", "ccb_get_time_derivatives", "ccb_step",
            "ccb_synchronize", "ccb_reduce", "ccb_kernel_launches", "ccb_profile_enable", "ccb_profile_get",
            "ccb_stream", "ccb_device_field", "ccb_state_width", "ccb_upload_state", "ccb_download_state",
            "ccb_measure_dfma_peak", "ccb_comm_unique_id", "ccb_comm_init", "ccb_comm_init_local", "ccb_set_slab",
            "ccb_grid_geometry", "ccb_migration_counts", "ccb_get_field_device_order", "ccb_get_int_field_device_order",
            "ccb_debug_neighbour_list", "ccb_profile_phase_count", "ccb_profile_phase_name", "ccb_ic_read", "ccb_ic_file_view", "ccb_ic_file_header", "ccb_ic_file_free", "ccb_efcheck", "ccb_freeze_out", "ccb_freeze_out_results", "ccb_set_preinverted_table", "ccb_set_preinverted_flat", "ccb_generate_preinverted_tile", "ccb_set_eos_table", "ccb_load_eos_table_dat", "ccb_eos_chain_parameters", "ccb_eos_full_thermo", "ccb_eos_locate"]
