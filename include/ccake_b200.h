/* ccake_b200.h -- C ABI of the B200-native CCAKE SPH hot path.
 *
 * CCAKE has no plugin/FFI boundary of its own: the per-timestep path is ordinary
 * C++ templates (SystemState<D>, SPHWorkstation<D,TEOM>, Evolver<D>) linked into
 * one binary. This header is the seam a maintainer binds those classes to; every
 * entry point names the reference member function(s) it replaces (file:line under
 * the reference tree). INTEGRATION.md shows the C++ glue.
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every call returns a ccb_status;
 *     ccb_last_error(ctx) gives a library-owned, human-readable string.
 *   - one ccb_ctx per GPU; a context is NOT thread-safe; the caller is the
 *     simulation's main-loop thread.
 *   - all host buffers are caller-owned; the library owns every device allocation.
 *   - per-particle host arrays are in PARTICLE-ID order (the order of upload);
 *     on the device particles are physically re-sorted by cell every RK stage.
 *   - units and index conventions are the reference's: fm, fm^-1 ... ; charges
 *     0=B 1=S 2=Q; shv[4][4] with 0=tau 1=x 2=y 3=eta.
 */
#ifndef CCAKE_B200_H
#define CCAKE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ccb_ctx ccb_ctx;

typedef enum {
  CCB_OK = 0,
  CCB_ERR_INVALID = 1,      /* bad argument / unsupported configuration */
  CCB_ERR_CUDA = 2,         /* CUDA runtime failure (no device, launch error, OOM) */
  CCB_ERR_NEG_ENTROPY = 3,  /* reference: exit(EXIT_FAILURE), src/evolver.cpp:437-450 */
  CCB_ERR_EOS_NO_ROOT = 4,  /* reference: exit(404), src/sph_workstation.cpp:1748-1760 */
  CCB_ERR_OUT_OF_GRID = 5,  /* particle left the neighbour grid (reference: undefined behaviour, src/system_state.cpp:636-638) */
  CCB_ERR_STATE = 6,        /* call sequence error (e.g. step before upload) */
  CCB_ERR_COMM = 7          /* halo / migration transport failure (NCCL) */
} ccb_status;

enum { CCB_COORD_HYPERBOLIC = 0, CCB_COORD_CARTESIAN = 1 };
/* Settings::eos_type (EquationOfState::set_up_chosen_EOSs, src/eos_initialization.cpp:216-279): "conformal", "table",
 * "tanh_conformal" (Tc = 220 MeV, Ts = 120 MeV), "conformal_diagonal" */
enum { CCB_EOS_CONFORMAL = 0, CCB_EOS_TABLE = 1, CCB_EOS_TANH_CONFORMAL = 2, CCB_EOS_CONFORMAL_DIAGONAL = 3 };

/* Mirror of the Settings fields the hot path consumes (reference include/settings.h,
 * SURVEY.md appendix B). Mode integers take the values of the reference enums in
 * include/transport_coefficients.h:18-97. */
typedef struct {
  int dim;                 /* 2 or 3 (Settings::dim) */
  int coordinate_system;   /* CCB_COORD_* (Settings::coordinate_system) */
  int rk_order;            /* 2 or 4 (Settings::rk_order; Evolver::execute_timestep, src/evolver.cpp:98-116) */
  int factorized_kernel;   /* Settings::factorized_kernel (D == 3 only) */
  double t0, hT, hEta;
  double xmin, ymin, etamin;          /* IC header minima -> neighbour grid (src/system_state.cpp:656-692) */
  double stepx, stepy, stepEta;       /* IC lattice spacing (init only) */
  int using_shear, using_diffusion, using_vorticity; /* Settings::check_consistency */
  int input_as_entropy, input_initial_shear, bulk_from_trace;
  /* transport (src/transport_coefficients.cpp:9-279) */
  int shear_mode, shear_relax_mode, zeta_mode, bulk_relax_mode, diffusion_mode, diffusion_relax_mode;
  int delta_pipi_mode, tau_pipi_mode, lambda_piPi_mode, phi6_mode, phi7_mode;
  int delta_PiPi_mode, lambda_Pipi_mode, phi1_mode, phi3_mode;
  int modulate_zeta_with_tanh, critical_scaling_bulk;
  double constant_eta_over_s, constant_zeta_over_s, cs2_dependent_zeta_A, cs2_dependent_zeta_p, C_B;
  double constant_kappa_over_T2[9];
  /* diffusion-current regulator, SPHWorkstation::regulate_diffusion (src/sph_workstation.cpp:133-218; Settings,
   * include/settings.h:211-215; off by default, include/defaults.h:119-123). e0 and xi0 in GeV/fm^3 as in the YAML. */
  /* Gaussian suppression of kappa near the critical point (Settings::critical_scaling_diffusion ...,
   * src/transport_coefficients.cpp:940-943, 1047-1056): table-EoS particles only; T, muB and widths in MeV */
  int critical_scaling_diffusion;
  double critical_point_T, critical_point_muB, critical_gaussian_width_T, critical_gaussian_width_muB;
  int diffusion_regulator_enabled;
  double diffusion_regulator_chi0, diffusion_regulator_e0, diffusion_regulator_xi0, diffusion_regulator_rq_max;
  /* equation of state */
  int eos_type;            /* CCB_EOS_* */
  double conformal_c, conformal_T0, conformal_muB0, conformal_muS0, conformal_muQ0;
  double freeze_out_temperature_MeV;
  int prohibit_unstable_cs2, prohibit_acausal_cs2, restrict_mu_T_ratios;  /* Settings, include/settings.h:143-147 */
  /* runtime */
  int device;              /* CUDA device ordinal */
  int rank, nranks;        /* slab decomposition (1 process per GPU) */
  int particlization_enabled; /* Settings::particlization_enabled: ccb_step ends with SPHWorkstation::freeze_out_particles (src/sph_workstation.cpp:2063) */
  double dt;               /* Settings::dt: the freeze-out interpolation steps back by it for particles without neighbours (include/freeze_out.h:717) */
  int online_inverter_enabled; /* Settings::online_inverter_enabled (include/defaults.h:15): 1 = phase 04 inverts the EoS per particle
                              (conformal closed form / table root search); 0 = phase 04 is a lookup in a PRE-INVERTED table
                              (EoS_Interpolator::fill_thermodynamics), set with ccb_set_preinverted_table / ccb_set_preinverted_flat */
  int slab_axis;           /* spatial axis the slabs cut: 0 = x (default), 1 = y, 2 = eta / z (3+1D). The neighbour-grid cell key is
                              ordered with this axis slowest, so that a slab's edge column is one contiguous range of the sorted order */
} ccb_config;

/* Initial-condition rows, one entry per particle, units as in the reference's IC
 * text format (BSQHydro::read_ccake, src/BSQHydro.cpp:207-425): positions in fm,
 * e in GeV/fm^3 (or s in fm^-3 when input_as_entropy), rho in fm^-3, u^i
 * dimensionless, bulk and shv6 = (xx, xy, x-eta, yy, y-eta, eta-eta) in GeV/fm^3.
 * r and u are n x 3, row-major; unused dimensions are ignored. Any pointer except
 * r and e_or_s may be NULL (= zeros). */
typedef struct {
  int64_t n;
  const double* r;
  const double* e_or_s;
  const double* rhoB;
  const double* rhoS;
  const double* rhoQ;
  const double* u;
  const double* bulk;
  const double* shv6;
  const int32_t* id;   /* optional: global particle ids of these rows (slab decomposition); NULL = 0 .. n-1 */
  const double* diffusion; /* optional (Settings::input_initial_diffusion): n x 12 initial q^mu_a, [charge B,S,Q][mu 0..3], as the IC file
                              holds them (BSQHydro::read_ccake, src/BSQHydro.cpp:261-372); NULL = zeros */
} ccb_ic;

/* Initial-condition TEXT files at scale (host code, no device involved): BSQHydro::read_ccake (src/BSQHydro.cpp:207-425; format 0:
 * header "# <ignored> stepx stepy stepEta <ignored> xmin ymin etamin", rows x y eta e|s rhoB rhoS rhoQ ux uy ueta bulk pixx pixy pixeta
 * piyy piyeta pietaeta [+ 12 currents when with_diffusion]) and BSQHydro::read_ICCING (:84-145; format 1: header "<ignored> stepx stepy
 * <ignored> xmin ymin", rows x y e rhoB rhoS rhoQ). Parsed in parallel straight into the arrays of a ccb_ic, no Particle objects;
 * rows at or below e_cutoff [GeV/fm^3] are dropped and disabled charges zeroed as the reference does. The ccb_ic_file owns the arrays. */
typedef struct ccb_ic_file ccb_ic_file;
int ccb_ic_read(const char* path, int format, int input_as_entropy, double e_cutoff_GeV, int baryon_charge_enabled, int strange_charge_enabled,
                int electric_charge_enabled, int with_diffusion, ccb_ic_file** out);
const ccb_ic* ccb_ic_file_view(const ccb_ic_file* f);
int ccb_ic_file_header(const ccb_ic_file* f, double header[6]);   /* stepx, stepy, stepEta, xmin, ymin, etamin */
void ccb_ic_file_free(ccb_ic_file* f);

/* Per-particle fields addressable through ccb_get_field / ccb_set_field.
 * Width = doubles per particle. Names follow the reference members. */
typedef enum {
  CCB_F_POSITION = 0,      /* 3  particle_info::position */
  CCB_F_U,                 /* 3  hydro_vector_info::u */
  CCB_F_EXTENSIVE,         /* 4  extensive s, rhoB, rhoS, rhoQ */
  CCB_F_SPH_MASS,          /* 4  sph_mass  s, rhoB, rhoS, rhoQ */
  CCB_F_EXTENSIVE_BULK,    /* 1 */
  CCB_F_EXTENSIVE_SHV,     /* 5  xx, xy, x-eta, yy, y-eta */
  CCB_F_EXTENSIVE_DIFFUSION,/* 9 [icharge][idir] */
  CCB_F_EZ,                /* 1  contribution_to_total_Ez */
  CCB_F_TSEED,             /* 4  thermo T, muB, muS, muQ (EoS warm start / result) */
  CCB_F_V,                 /* 3 */
  CCB_F_GAMMA,             /* 1 */
  CCB_F_SIGMA_LAB,         /* 1 */
  CCB_F_SIGMA,             /* 1 */
  CCB_F_SMOOTHED,          /* 4  lab-frame smoothed s, rhoB, rhoS, rhoQ */
  CCB_F_THERMO,            /* 16 T muB muS muQ e s rhoB rhoS rhoQ p cs2 w dwds dwdB dwdS dwdQ */
  CCB_F_DALPHA,            /* 12 dalpha_{B,S,Q}/d{s,B,S,Q} */
  CCB_F_BULK,              /* 1 */
  CCB_F_SHV,               /* 16 full pi^{mu nu} */
  CCB_F_DIFFUSION,         /* 12 q^mu_a [icharge][mu] */
  CCB_F_TRANSPORT,         /* 13 eta_pi tau_pi zeta_Pi tau_Pi delta_PiPi delta_pipi lambda_piPi lambda_Pipi phi1 phi3 phi6 phi7 tau_pipi */
  CCB_F_GRADP,             /* 3 */
  CCB_F_GRADE,             /* 3 */
  CCB_F_GRADBULK,          /* 3 */
  CCB_F_GRADSHEAR,         /* 3 */
  CCB_F_DIVSHEAR,          /* 3 */
  CCB_F_GRADV,             /* 9 [i][j] = d_i v_j */
  CCB_F_DIV_QA,            /* 3 */
  CCB_F_GRAD_QA,           /* 3 */
  CCB_F_GRAD_ALPHA,        /* 9 [idir][icharge] */
  CCB_F_DU_DT,             /* 3 */
  CCB_F_D_EXTENSIVE_DT,    /* 4  s, rhoB, rhoS, rhoQ */
  CCB_F_D_BULK_DT,         /* 1 */
  CCB_F_D_SHV_DT,          /* 5 */
  CCB_F_D_DIFFUSION_DT,    /* 9 */
  CCB_F_THETA,             /* 1 */
  CCB_F_SHV_NABLA_U,       /* 1 */
  CCB_F_DEZ,               /* 1  contribution_to_total_dEz */
  CCB_F_CHIBB,             /* 2  thermo chiBB, chiBB0 (include/thermodynamic_info.h:64-65) */
  CCB_F_COUNT
} ccb_field;

/* CCB_I_EOS: index of the equation of state the particle last used, in the order of EquationOfState::chosen_EOSs
 * (thermodynamic_info::eos_name; table run: 0 table, 1 tanh_conformal, 2 conformal, 3 conformal_diagonal) */
typedef enum { CCB_I_ID = 0, CCB_I_BTRACK, CCB_I_FREEZE, CCB_I_CELL, CCB_I_EOS, CCB_I_COUNT } ccb_int_field;

typedef struct {
  double S, B, Sq, Q, E, Ez;  /* SystemState::conservation_{entropy,BSQ,energy}, src/system_state.cpp:739-885 */
  int64_t n_frozen;           /* SPHWorkstation::continue_evolution, src/sph_workstation.cpp:2011-2016 */
} ccb_totals;

/* --- lifetime ------------------------------------------------------------- */
int ccb_create(ccb_ctx** out, const ccb_config* cfg);   /* SystemState<D>/SPHWorkstation<D,TEOM> ctors + SPHWorkstation::initialize (src/sph_workstation.cpp:38-84) */
int ccb_destroy(ccb_ctx* ctx);
const char* ccb_last_error(ccb_ctx* ctx);
int ccb_field_width(int field);

/* --- tabulated equation of state (eos_type = CCB_EOS_TABLE) ------------------------------------------------
 * The 4-D table of EoS_table / InterpolatorND<4> (src/eos_table.cpp:14-122, src/interpolatorND.cpp:125-226) as the
 * FILE holds it: grid in MeV, T slowest and muS fastest, 17 fields per node in the order
 * p s B S Q e cs2 chiBB chiQQ chiSS chiBQ chiBS chiQS chiTB chiTQ chiTS chiTT, dimensionless ratios
 * (normalize_by_T) or MeV^n. The library applies the constructor's unit conversions, copies the table to the
 * device (a 4-field "hot" copy for the root search + the full table) and matches the fall-back chain
 * tanh_conformal -> conformal -> conformal_diagonal to it (EquationOfState::set_up_chosen_EOSs,
 * src/eos_initialization.cpp:273-455). Must be called after ccb_create and before ccb_allocate / ccb_upload_ic.
 * The caller keeps ownership of the arrays. */
typedef struct {
  int32_t n[4];          /* nodes along T, muB, muQ, muS */
  const double* grid;    /* nodes x 4, row-major */
  const double* fields;  /* nodes x 17, row-major */
  int normalize_by_T;    /* Settings::normalize_by_T */
} ccb_eos_table;
int ccb_set_eos_table(ccb_ctx* ctx, const ccb_eos_table* table);
/* same, reading `<path>` in the reference's thermo.dat text format (InterpolatorND::load_data_from_dat) */
int ccb_load_eos_table_dat(ccb_ctx* ctx, const char* path, int normalize_by_T);
/* parameters of the matched fall-back EoSs: out[0..6] tanh_conformal (c, T0, muB0, muS0, muQ0, Tc, Ts), out[7..11]
 * conformal, out[12..16] conformal_diagonal; out[17] = efcheck (src/sph_workstation.cpp:66-81) */
int ccb_eos_chain_parameters(ccb_ctx* ctx, double out[18]);

/* --- pre-inverted equation of state (online_inverter_enabled = 0; EoS_Interpolator, include/eos_interpolator.h:45-163) ----------
 * The reference reads these tables from HDF5 (src/eos_interpolator.cpp:16-521); here they arrive as arrays, and
 * ccb_generate_preinverted_tile produces them on the device with the run's own inversion kernel (utilities/generate_eos_table.cpp).
 * TILED format: one tile per decade of (s, |rhoB|, |rhoS|, |rhoQ|), slot = exponent + 3 in [0,5) x [0,4)^3; every tile has its own
 * axes (min, step, N; a log-spaced axis stores log10 of the physical value) and 11 fields in the order of ccake::eos_variables
 * (T muB muQ muS e p cs2 dwds dwdB dwdQ dwdS), each N[0] x N[1] x N[2] x N[3] with the rhoQ index fastest.
 * FLAT format (src/eos_interpolator.cpp:262-521): ONE (s, nB) table, log-spaced in both, S and Q degenerate, handed over as the
 * file holds it (MeV units; s, nB as 2-D arrays [Ns][NB]); the library applies the loader's unit conversion and derives dw/ds,
 * dw/dB from chiTT, chiTB, chiBB as the loader does. Only the flat format serves a whole run (initialisation and efcheck:
 * query_thermo_host, invert_e_to_s_host, efreeze_at_T are flat-only in the reference); tiles serve ccb_update_thermodynamics. */
typedef struct {
  int32_t slot[4];
  double min[4], step[4];
  int32_t N[4];
  int32_t log_spacing[4];
  const double* vars;      /* 11 x N[0] x N[1] x N[2] x N[3] */
} ccb_preinverted_tile;
typedef struct {
  int32_t Ns, NB;
  const double *s, *nB;                               /* [Ns][NB], MeV^3 */
  const double *T, *muB, *muS, *muQ, *e, *p, *cs2;    /* MeV, MeV^4, dimensionless */
  const double *chiTT, *chiTB, *chiBB;                /* MeV^2 */
} ccb_preinverted_flat;
int ccb_set_preinverted_table(ccb_ctx* ctx, int32_t ntiles, const ccb_preinverted_tile* tiles);
int ccb_set_preinverted_flat(ccb_ctx* ctx, const ccb_preinverted_flat* table);
/* one decade tile of N^4 nodes (the reference generator uses N = 20): vars[11][N^4] as ccb_preinverted_tile::vars wants them, eos[N^4] =
 * answering member of chosen_EOSs, -1 (and NaN fields) where the reference's generator aborts ("EoS root-finding failed") */
int ccb_generate_preinverted_tile(ccb_ctx* ctx, const int32_t exponent[4], int32_t N, double* vars, int32_t* eos);

/* --- host queries of the equation of state (EquationOfState / EoS_base surface, include/eos.h:62-210, include/eos_base.h:12-50) ---
 * Batched and evaluated on the device; all arrays are caller-owned HOST arrays, row-major; both calls synchronise.
 * ccb_eos_full_thermo: EoS_base::get_full_thermo of member `member` of chosen_EOSs (0 = the run's default EoS) at n points
 *   (T, muB, muQ, muS) [fm^-1] -> out[n][17] = p s B S Q e cs2 chiBB chiQQ chiSS chiBQ chiBS chiQS chiTB chiTQ chiTS chiTT;
 *   a point outside the member's domain gives a row of NaN.
 * ccb_eos_locate: EquationOfState::update_s (energy = 0: dens = s, rhoB, rhoS, rhoQ, located from seed = T, muB, muS, muQ through
 *   the fall-back chain, starting at member start[i]; start may be NULL) or ::s_out (energy = 1: dens = e, rhoB, rhoS, rhoQ; seed
 *   ignored, may be NULL) followed by set_thermo -> thermo[n][16] in CCB_F_THERMO order and the answering member eos[n]
 *   (-1 and a NaN row where the reference would report "no solution"). */
int ccb_eos_full_thermo(ccb_ctx* ctx, int member, int64_t n, const double* points, double* out);
int ccb_eos_locate(ccb_ctx* ctx, int energy, int64_t n, const double* dens, const double* seed, const int32_t* start, double* thermo, int32_t* eos);

/* --- state in / out ------------------------------------------------------- */
/* n blank particles (ids 0..n-1); state is then injected with ccb_set_field.
 * Replaces SystemState::allocate_cabana_particles (src/system_state.cpp:58-63). */
int ccb_allocate(ccb_ctx* ctx, int64_t n);
/* IC rows -> device + the whole of BSQHydro::initialize_hydrodynamics
 * (src/BSQHydro.cpp:569-611): e_cutoff is the caller's job. */
int ccb_upload_ic(ccb_ctx* ctx, const ccb_ic* ic);
/* SystemState::copy_host_to_device / copy_device_to_host (src/system_state.cpp:72-555), field by field */
int ccb_set_field(ccb_ctx* ctx, int field, const double* host);
int ccb_get_field(ccb_ctx* ctx, int field, double* host);
int ccb_set_int_field(ccb_ctx* ctx, int field, const int32_t* host);
int ccb_get_int_field(ccb_ctx* ctx, int field, int32_t* host);
/* the whole evolved state (reference EvolverCache member list, include/evolver.h:21-47) as ONE column-major
 * block [column][particle id] of ccb_state_width() columns: position 3, u 3, extensive 4, extensive_bulk 1,
 * extensive_shv 5, extensive_diffusion 9 (only when using_diffusion), Ez 1. One host<->device copy each way
 * (pass pinned memory for full PCIe speed). upload is asynchronous on the context's stream; download
 * synchronises and surfaces device-side errors. With a slab decomposition the block holds this rank's ccb_n()
 * particles in DEVICE order (column stride ccb_n(); ids through ccb_get_int_field_device_order), and an upload must
 * hand back the block of the preceding download. */
int ccb_state_width(ccb_ctx* ctx);
int ccb_upload_state(ccb_ctx* ctx, const double* host);
int ccb_download_state(ccb_ctx* ctx, double* host);
int64_t ccb_n(ccb_ctx* ctx);
double ccb_time(ccb_ctx* ctx);
int ccb_set_time(ccb_ctx* ctx, double t);
/* neighbour-grid shape of the last cell build: cells per axis (src/system_state.cpp:651-700) */
int ccb_grid_shape(ccb_ctx* ctx, int32_t ncell[3]);

/* --- slab decomposition over the GPUs of one box (new: the reference is single-process) ----------------------
 * One context per GPU / process, config.rank of config.nranks. Rank r owns the particles whose neighbour-grid cell
 * column along x lies in [col_lo, col_hi) (contiguous, ascending with the rank; the outermost slabs extend to the grid
 * edge). Every cell build then hands particles that left the slab to rank -+ 1 and appends the arrivals; every
 * gather reads the edge columns' packed records of its neighbours (the 2h-wide halo): straight out of their record arrays over
 * NVLink peer memory (CUDA IPC, ordered by ready / done flag words; the default between processes) or as a message appended
 * behind the own records (environment CCB_HALO_P2P=0). Transport of everything else: NCCL
 * send/recv (ccb_comm_init; rank 0 creates the id with ccb_comm_unique_id and the caller broadcasts it), or mailboxes
 * between contexts of ONE process, each driven by its own host thread (ccb_comm_init_local; used by the single-GPU tests).
 * Call order: ccb_create, ccb_comm_init*, ccb_set_slab, ccb_upload_ic (own rows + global ids). Host arrays in particle-id
 * order do not exist across slabs: read fields with the *_device_order getters together with CCB_I_ID. */
int ccb_comm_unique_id(char id[128]);
int ccb_comm_init(ccb_ctx* ctx, const char id[128]);
int ccb_comm_init_local(ccb_ctx* ctx, int group);
int ccb_set_slab(ccb_ctx* ctx, int32_t col_lo, int32_t col_hi);
int ccb_grid_geometry(ccb_ctx* ctx, double min[3], double edge[3], int32_t ncell[3]);  /* to choose the column ranges */
int ccb_migration_counts(ccb_ctx* ctx, int64_t* sent, int64_t* received);              /* running totals */
int ccb_get_field_device_order(ccb_ctx* ctx, int field, double* host);          /* n x width, current (cell-sorted) order */
int ccb_get_int_field_device_order(ccb_ctx* ctx, int field, int32_t* host);

/* debug / test: the neighbour sets the field gather visits, as the CSR list SystemState::reset_neighbour_list stores
 * (src/system_state.cpp:651-719: VerletList, full neighbour tag, self excluded): offsets[ccb_n() + 1] in particle-id order,
 * idx = neighbour ids, ascending within a particle; cap = entries idx can hold (sum of CCB_I_BTRACK). Call after
 * ccb_smooth_fields (any later phase may have run too: phase 02 is repeated to re-pack the field records, which
 * invalidates the gradient records -- run ccb_local_update before the next ccb_smooth_gradients). The product path never
 * materialises this list. */
int ccb_debug_neighbour_list(ccb_ctx* ctx, int64_t* offsets, int32_t* idx, int64_t cap);

/* --- the ten phases of SPHWorkstation::get_time_derivatives (src/sph_workstation.cpp:1543-1583) --- */
int ccb_build_cells(ccb_ctx* ctx);        /* 01 SystemState::reset_neighbour_list (src/system_state.cpp:651-719) */
int ccb_smooth_fields(ccb_ctx* ctx);      /* 02+03 TEOM::update_velocity (src/eom_default.cpp:18-38) + smooth_all_particle_fields (src/sph_workstation.cpp:495-578) */
int ccb_update_thermodynamics(ccb_ctx* ctx); /* 04 update_all_particle_thermodynamics (src/sph_workstation.cpp:1714-1772) */
int ccb_local_update(ccb_ctx* ctx);       /* 05+06+08 TEOM::reset_pi_tensor (src/eom_default.cpp:57-172) + regulate_diffusion (src/sph_workstation.cpp:133-218) + update_all_particle_viscosities (src/sph_workstation.cpp:1623-1699) */
int ccb_smooth_gradients(ccb_ctx* ctx);   /* 09 smooth_all_particle_gradients (src/sph_workstation.cpp:797-986) */
int ccb_time_derivatives(ccb_ctx* ctx);   /* 10 TEOM::evaluate_time_derivatives (src/eom_default.cpp:212-869) */
int ccb_get_time_derivatives(ccb_ctx* ctx); /* all of the above in order */

/* --- time stepping -------------------------------------------------------- */
/* SPHWorkstation::advance_timestep -> Evolver::execute_timestep (src/sph_workstation.cpp:2043-2073,
 * src/evolver.cpp:98-561): one full RK step of size dt (RK2 midpoint, or the reference's five-evaluation RK4 sequence when
 * rk_order = 4); asynchronous on the context's stream. */
int ccb_step(ccb_ctx* ctx, double dt);
int ccb_synchronize(ccb_ctx* ctx);        /* waits for the stream and surfaces device-side error flags */
int ccb_reduce(ccb_ctx* ctx, ccb_totals* out);

/* --- freeze-out (include/freeze_out.h:165-933) ---------------------------------------------------------------------------------
 * ccb_freeze_out = SPHWorkstation::freeze_out_particles (src/sph_workstation.cpp:1475-1485): FreezeOut::check_freeze_out_status (the
 * 0 -> 1 -> 3 status machine on e <= efcheck) + FreezeOut::bsqsvfreezeout (snapshot bookkeeping of the last two time steps and of the
 * back-up steps of particles that are losing their neighbours; hypersurface interpolation of every particle that reached status 3,
 * which then becomes 4). ccb_step calls it itself when particlization_enabled; ccb_upload_ic runs the initial status check
 * (SPHWorkstation::initial_smoothing, :450-452). n_freezing_out = particles frozen out by this call.
 * ccb_freeze_out_results: the reference's `results` record (ResultsTypes, include/freeze_out.h:35-49) of every particle, particle-id
 * order, CCB_FO_ROW doubles per particle at the CCB_FO_* offsets; print[id] = 1 for the particles frozen out by the LAST call (the rows
 * Output::print_freeze_out writes). Single-context runs only. */
enum {
  CCB_FO_SHEARSUB = 0, CCB_FO_DIFFOUT = 16, CCB_FO_DIVT = 28, CCB_FO_DIVE = 31, CCB_FO_DIVP = 34, CCB_FO_RSUB = 37, CCB_FO_UOUT = 40,
  CCB_FO_DIVTTEMP = 43, CCB_FO_DIVPPRESS, CCB_FO_DIVEENER, CCB_FO_GSUB, CCB_FO_BULKSUB, CCB_FO_THETASUB, CCB_FO_SWSUB, CCB_FO_SHEAR33SUB,
  CCB_FO_TLIST, CCB_FO_SFO, CCB_FO_EFLUC, CCB_FO_TFLUC, CCB_FO_MUBFLUC, CCB_FO_MUSFLUC, CCB_FO_MUQFLUC, CCB_FO_CS2FZFLUC, CCB_FO_WFZFLUC,
  CCB_FO_RHOBFLUC, CCB_FO_RHOSFLUC, CCB_FO_RHOQFLUC, CCB_FO_EOS_TYPE, CCB_FO_ROW = 64
};
int ccb_efcheck(ccb_ctx* ctx, double* efcheck);   /* SystemState::efcheck = e(T_FO, mu = 0) [fm^-4] (SPHWorkstation::initialize, src/sph_workstation.cpp:66-81) */
int ccb_freeze_out(ccb_ctx* ctx, int64_t* n_freezing_out);
int ccb_freeze_out_results(ccb_ctx* ctx, double* results, int32_t* print);

/* --- instrumentation ------------------------------------------------------ */
/* number of kernels launched by this context since creation, and device time (ms) per
 * reference phase name ("01_neighbour_list" ...) when profiling is on (CCAKE_PROFILE, include/phase_profiler.h) */
int64_t ccb_kernel_launches(ccb_ctx* ctx);
int ccb_profile_enable(ccb_ctx* ctx, int on);
int ccb_profile_get(ccb_ctx* ctx, int phase, double* ms, int64_t* calls);
int ccb_profile_phase_count(void);
const char* ccb_profile_phase_name(int phase);   /* the CCAKE_PHASE labels of src/sph_workstation.cpp:1543-1582; NULL past the last phase */
/* fp64 FMA peak (TFLOP/s) of a device from a DFMA micro-benchmark: roofline denominator of the gathers */
int ccb_measure_dfma_peak(int device, double* tflops);
void* ccb_stream(ccb_ctx* ctx);           /* cudaStream_t of the context */
/* device pointer of a field in CURRENT (cell-sorted) order, for zero-copy consumers (halo packing) */
int ccb_device_field(ccb_ctx* ctx, int field, int component, double** dptr);

#ifdef __cplusplus
}
#endif
#endif
