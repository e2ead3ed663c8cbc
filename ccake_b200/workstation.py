"""Host-side mirror of the reference's SystemState / SPHWorkstation / Evolver
surface for the hot path, on top of the C ABI (``include/ccake_b200.h``).

Method names follow the reference (``reset_neighbour_list``,
``smooth_all_particle_fields``, ``advance_timestep`` ...,
``include/sph_workstation.h:113-156``, ``include/system_state.h:40-127``) so that the
parity tests read like the reference's own unit tests.  The C++ twin of this file
is ``include/ccake_b200_host.hpp``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .settings import DEFAULTS, to_config


class CCAKEError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ccake_b200 error {code}: {msg}")
        self.code = code


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class SPHWorkstation:
    """One context (= one GPU) holding the particle system and running the phases."""

    def __init__(self, settings: dict, device: int = 0, rank: int = 0, nranks: int = 1):
        self.settings = dict(DEFAULTS)
        self.settings.update(settings)
        self._L = _capi.lib()
        self._cfg = to_config(self.settings, device=device, rank=rank, nranks=nranks)
        self._ctx = C.c_void_p()
        rc = self._L.ccb_create(C.byref(self._ctx), C.byref(self._cfg))
        if rc:
            msg = self._L.ccb_last_error(self._ctx).decode() if self._ctx else "ccb_create failed"
            if self._ctx:
                self._L.ccb_destroy(self._ctx)
                self._ctx = C.c_void_p()
            raise CCAKEError(rc, msg)
        if self.settings["eos_type"] == "table":
            # EquationOfState::init -> EoS_table(eos_path, normalize_by_T), src/eos_initialization.cpp:273-279
            data = self.settings.get("eos_table_data")
            if data is not None:
                # the same table handed over as arrays (ccb_set_eos_table): what a host that has already read
                # thermo.dat / thermo.h5 does; `data` is the dict of ccake_b200.eos_table.synthetic_table
                tab = _capi.EosTable()
                grid = np.ascontiguousarray(data["grid"], dtype=np.float64)
                fields = np.ascontiguousarray(data["fields"], dtype=np.float64)
                for k in range(4):
                    tab.n[k] = int(data["sizes"][k])
                tab.grid, tab.fields, tab.normalize_by_T = _dp(grid), _dp(fields), int(bool(data["normalize_by_T"]))
                self._check(self._L.ccb_set_eos_table(self._ctx, C.byref(tab)))
            else:
                import os

                path = os.path.join(str(self.settings["eos_path"]), "thermo.dat")
                self._check(self._L.ccb_load_eos_table_dat(self._ctx, path.encode(), int(bool(self.settings["normalize_by_T"]))))

        if not self.settings.get("online_inverter_enabled", True):
            # SPHWorkstation::initialize -> EoS_Interpolator(preinverted_eos_path), src/sph_workstation.cpp:47-58
            flat, tiles = self.settings.get("preinverted_flat"), self.settings.get("preinverted_tiles")
            if flat is not None:
                self.set_preinverted_flat(flat)
            elif tiles is not None:
                self.set_preinverted_table(tiles)

    # -- pre-inverted EoS (EoS_Interpolator) --------------------------------------------
    def set_preinverted_table(self, tiles):
        """TILED table: list of dicts exponent[4], min[4], step[4], vars (11, N0, N1, N2, N3) (ccake_b200.eos_preinverted)."""
        arr = (_capi.PreTile * len(tiles))()
        keep = []
        for k, t in enumerate(tiles):
            v = np.ascontiguousarray(t["vars"], dtype=np.float64)
            keep.append(v)
            for i in range(4):
                arr[k].slot[i] = int(t["exponent"][i]) + 3
                arr[k].min[i], arr[k].step[i] = float(t["min"][i]), float(t["step"][i])
                arr[k].N[i] = int(v.shape[1 + i])
                arr[k].log_spacing[i] = int(bool(t.get("log_spacing", (0, 0, 0, 0))[i]))
            arr[k].vars = _dp(v)
        self._check(self._L.ccb_set_preinverted_table(self._ctx, len(tiles), arr))

    def set_preinverted_flat(self, flat):
        """FLAT (s, nB) table as the file holds it (MeV units): dict of (Ns, NB) arrays s nB T muB muS muQ e p cs2 chiTT chiTB chiBB."""
        f = _capi.PreFlat()
        keep = {k: np.ascontiguousarray(flat[k], dtype=np.float64) for k in ("s", "nB", "T", "muB", "muS", "muQ", "e", "p", "cs2", "chiTT", "chiTB", "chiBB")}
        f.Ns, f.NB = keep["s"].shape
        for k, v in keep.items():
            setattr(f, k, _dp(v))
        self._check(self._L.ccb_set_preinverted_flat(self._ctx, C.byref(f)))

    def generate_preinverted_tile(self, exponent, N=20):
        """utilities/generate_eos_table.cpp for one decade tile, on the device: ((11, N, N, N, N) fields, (N, N, N, N) answering EoS)."""
        ex = (C.c_int32 * 4)(*[int(e) for e in exponent])
        v = np.empty((11, N, N, N, N), dtype=np.float64)
        eos = np.empty((N, N, N, N), dtype=np.int32)
        self._check(self._L.ccb_generate_preinverted_tile(self._ctx, ex, int(N), _dp(v), eos.ctypes.data_as(C.POINTER(C.c_int32))))
        return v, eos

    def eos_chain_parameters(self):
        """Matched parameters of the fall-back EoSs + efcheck (table runs; EquationOfState::set_up_chosen_EOSs)."""
        out = np.zeros(18)
        self._check(self._L.ccb_eos_chain_parameters(self._ctx, _dp(out)))
        return dict(tanh_conformal=out[0:7], conformal=out[7:12], conformal_diagonal=out[12:17], efcheck=out[17])

    def eos_full_thermo(self, points, member=0):
        """EoS_base::get_full_thermo at (T, muB, muQ, muS) points [fm^-1] -> (n, 17) in the table's field order."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 4)
        out = np.empty((pts.shape[0], 17))
        self._check(self._L.ccb_eos_full_thermo(self._ctx, int(member), pts.shape[0], _dp(pts), _dp(out)))
        return out

    def eos_locate(self, densities, seed=None, start=None, energy=False):
        """EquationOfState::update_s (s, rhoB, rhoS, rhoQ from `seed` = T, muB, muS, muQ) or ::s_out (energy=True: e, rhoB, rhoS, rhoQ)
        + set_thermo -> ((n, 16) thermo in CCB_F_THERMO order, (n,) answering member of chosen_EOSs; -1 = no solution)."""
        d = np.ascontiguousarray(densities, dtype=np.float64).reshape(-1, 4)
        n = d.shape[0]
        sd = None if seed is None else np.ascontiguousarray(seed, dtype=np.float64).reshape(n, 4)
        st = None if start is None else np.ascontiguousarray(start, dtype=np.int32)
        th, eos = np.empty((n, 16)), np.empty(n, dtype=np.int32)
        ip = C.POINTER(C.c_int32)
        self._check(self._L.ccb_eos_locate(self._ctx, int(bool(energy)), n, _dp(d), None if sd is None else _dp(sd),
                                           None if st is None else st.ctypes.data_as(ip), _dp(th), eos.ctypes.data_as(ip)))
        return th, eos

    # -- plumbing --------------------------------------------------------------
    def _check(self, rc):
        if rc:
            raise CCAKEError(rc, self._L.ccb_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.ccb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- SystemState -----------------------------------------------------------
    @property
    def n_particles(self):
        return self._L.ccb_n(self._ctx)

    @property
    def t(self):
        return self._L.ccb_time(self._ctx)

    @t.setter
    def t(self, v):
        self._check(self._L.ccb_set_time(self._ctx, float(v)))

    def allocate(self, n):
        """SystemState::allocate_cabana_particles for n blank particles."""
        self._check(self._L.ccb_allocate(self._ctx, int(n)))

    # -- slab decomposition ------------------------------------------------------
    def grid_geometry(self):
        """(min[3], edge[3], ncell[3]) of the neighbour grid (src/system_state.cpp:651-700)."""
        mn, ed, nc = np.zeros(3), np.zeros(3), (C.c_int32 * 3)()
        self._check(self._L.ccb_grid_geometry(self._ctx, _dp(mn), _dp(ed), nc))
        return mn, ed, np.array(list(nc), dtype=np.int64)

    @property
    def slab_axis(self):
        """Spatial axis the slab decomposition cuts (0 = x, 2 = eta): ccb_config::slab_axis."""
        return int(self._cfg.slab_axis)

    def comm_init_local(self, group=0):
        """In-process transport between the contexts of one process (each driven by its own thread)."""
        self._check(self._L.ccb_comm_init_local(self._ctx, int(group)))

    def comm_init_nccl(self):
        """NCCL transport; the unique id travels through torch.distributed (already initialised by the caller)."""
        import torch
        import torch.distributed as dist

        buf = C.create_string_buffer(128)
        if dist.get_rank() == 0:
            self._check(self._L.ccb_comm_unique_id(buf))
        t = torch.tensor(list(buf.raw), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        self._check(self._L.ccb_comm_init(self._ctx, bytes(t.cpu().tolist())))

    def set_slab(self, col_lo, col_hi):
        self._check(self._L.ccb_set_slab(self._ctx, int(col_lo), int(col_hi)))

    def migration_counts(self):
        a, b = C.c_int64(), C.c_int64()
        self._check(self._L.ccb_migration_counts(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_device_order(self, field):
        """One field of this rank's particles in the CURRENT device order; pair with get_int_device_order('id')."""
        fid, w = _capi.FIELDS[field]
        out = np.empty((self.n_particles, w), dtype=np.float64)
        self._check(self._L.ccb_get_field_device_order(self._ctx, fid, _dp(out)))
        return out[:, 0] if w == 1 else out

    def get_int_device_order(self, field):
        out = np.empty(self.n_particles, dtype=np.int32)
        self._check(self._L.ccb_get_int_field_device_order(self._ctx, _capi.INT_FIELDS[field], out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def upload_initial_conditions(self, r, e_or_s, rhoB=None, rhoS=None, rhoQ=None, u=None, bulk=None, shv6=None, ids=None,
                                  diffusion=None):
        """BSQHydro::read_ccake rows + BSQHydro::initialize_hydrodynamics, on the device."""
        r = np.asarray(r, dtype=np.float64)
        n = r.shape[0]
        r3 = np.zeros((n, 3))
        r3[:, : r.shape[1]] = r
        keep = [r3, np.ascontiguousarray(e_or_s, dtype=np.float64)]
        ic = _capi.IC()
        ic.n = n
        ic.r = _dp(keep[0])
        ic.e_or_s = _dp(keep[1])
        # BSQHydro::read_ccake drops a charge density whose `*_charge_enabled` switch is off (src/BSQHydro.cpp:341-343)
        if not self.settings["baryon_charge_enabled"]:
            rhoB = None
        if not self.settings["strange_charge_enabled"]:
            rhoS = None
        if not self.settings["electric_charge_enabled"]:
            rhoQ = None
        for name, a, w in (("rhoB", rhoB, 1), ("rhoS", rhoS, 1), ("rhoQ", rhoQ, 1), ("u", u, 3), ("bulk", bulk, 1), ("shv6", shv6, 6)):
            if a is None:
                continue
            a = np.asarray(a, dtype=np.float64)
            if w == 3:
                a3 = np.zeros((n, 3))
                a3[:, : a.shape[1]] = a
                a = a3
            a = np.ascontiguousarray(a)
            keep.append(a)
            setattr(ic, name, _dp(a))
        if diffusion is not None:
            # (n, 3, 4) initial q^mu_a; a disabled charge's row is dropped (src/BSQHydro.cpp:345-372)
            q = np.array(diffusion, dtype=np.float64).reshape(n, 3, 4)
            for c, key in enumerate(("baryon_charge_enabled", "strange_charge_enabled", "electric_charge_enabled")):
                if not self.settings[key]:
                    q[:, c, :] = 0.0
            q = np.ascontiguousarray(q)
            keep.append(q)
            ic.diffusion = _dp(q)
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32)
            keep.append(ids)
            ic.id = ids.ctypes.data_as(C.POINTER(C.c_int32))
        self._check(self._L.ccb_upload_ic(self._ctx, C.byref(ic)))

    def get(self, field):
        """copy_device_to_host for one field, particle-id order; shape (n,) or (n, width)."""
        fid, w = _capi.FIELDS[field]
        out = np.empty((self.n_particles, w), dtype=np.float64)
        self._check(self._L.ccb_get_field(self._ctx, fid, _dp(out)))
        return out[:, 0] if w == 1 else out

    def set(self, field, values):
        fid, w = _capi.FIELDS[field]
        a = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(self.n_particles, w))
        self._check(self._L.ccb_set_field(self._ctx, fid, _dp(a)))

    @property
    def state_width(self):
        return self._L.ccb_state_width(self._ctx)

    def copy_host_to_device(self, state):
        """SystemState::copy_host_to_device of the evolved state: ``state`` is a C-contiguous
        (state_width, n) float64 array in particle-id order (pinned memory for full PCIe speed)."""
        assert state.dtype == np.float64 and state.flags.c_contiguous and state.size >= self.state_width * self.n_particles
        self._check(self._L.ccb_upload_state(self._ctx, _dp(state)))

    def copy_device_to_host(self, out):
        """SystemState::copy_device_to_host of the evolved state into ``out`` (same layout); synchronises."""
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size >= self.state_width * self.n_particles
        self._check(self._L.ccb_download_state(self._ctx, _dp(out)))

    def get_int(self, field):
        out = np.empty(self.n_particles, dtype=np.int32)
        self._check(self._L.ccb_get_int_field(self._ctx, _capi.INT_FIELDS[field], out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def set_int(self, field, values):
        a = np.ascontiguousarray(values, dtype=np.int32)
        self._check(self._L.ccb_set_int_field(self._ctx, _capi.INT_FIELDS[field], a.ctypes.data_as(C.POINTER(C.c_int32))))

    def neighbours(self):
        """The neighbour sets the field gather visits as a CSR list (offsets[n + 1], ascending ids), particle-id order: the
        reference's Verlet list (SystemState::reset_neighbour_list). Debug / test only; call after smooth_all_particle_fields."""
        n = self.n_particles
        bt = self.get_int("btrack")
        total = int(np.sum(bt[bt > 0]))
        off = np.zeros(n + 1, dtype=np.int64)
        idx = np.empty(max(total, 1), dtype=np.int32)
        self._check(self._L.ccb_debug_neighbour_list(self._ctx, off.ctypes.data_as(C.POINTER(C.c_int64)),
                                                     idx.ctypes.data_as(C.POINTER(C.c_int32)), int(idx.size)))
        return off, idx[:total]

    def grid_shape(self):
        nc = (C.c_int32 * 3)()
        self._check(self._L.ccb_grid_shape(self._ctx, nc))
        return np.array(list(nc), dtype=np.int32)

    def conservation(self):
        """conservation_entropy / conservation_BSQ / conservation_energy in one device pass."""
        t = _capi.Totals()
        self._check(self._L.ccb_reduce(self._ctx, C.byref(t)))
        return dict(S=t.S, B=t.B, Sq=t.Sq, Q=t.Q, E=t.E, Ez=t.Ez, n_frozen=t.n_frozen)

    # -- SPHWorkstation phases ---------------------------------------------------
    def reset_neighbour_list(self):
        self._check(self._L.ccb_build_cells(self._ctx))

    def smooth_all_particle_fields(self):
        """calculate_gamma_and_velocities + smooth_all_particle_fields."""
        self._check(self._L.ccb_smooth_fields(self._ctx))

    def update_all_particle_thermodynamics(self):
        self._check(self._L.ccb_update_thermodynamics(self._ctx))

    def reset_pi_tensor_and_viscosities(self):
        """reset_pi_tensor + update_all_particle_viscosities."""
        self._check(self._L.ccb_local_update(self._ctx))

    def smooth_all_particle_gradients(self):
        self._check(self._L.ccb_smooth_gradients(self._ctx))

    def evaluate_time_derivatives(self):
        self._check(self._L.ccb_time_derivatives(self._ctx))

    def get_time_derivatives(self):
        self._check(self._L.ccb_get_time_derivatives(self._ctx))

    def advance_timestep(self, dt, rk_order=None):
        if rk_order is not None and rk_order != self._cfg.rk_order:
            raise ValueError("rk_order is fixed at context creation")
        self._check(self._L.ccb_step(self._ctx, float(dt)))

    def freeze_out_particles(self):
        """SPHWorkstation::freeze_out_particles: status machine + hypersurface interpolation; returns the particles frozen out now.
        (advance_timestep calls it itself when settings['particlization_enabled'].)"""
        k = C.c_int64()
        self._check(self._L.ccb_freeze_out(self._ctx, C.byref(k)))
        return k.value

    def freeze_out_results(self):
        """FreezeOut::results of every particle ((n, 64) at the CCB_FO_* offsets, particle-id order) and the print flags (n,)."""
        n = self.n_particles
        res = np.zeros((n, 64)); pr = np.zeros(n, dtype=np.int32)
        self._check(self._L.ccb_freeze_out_results(self._ctx, _dp(res), pr.ctypes.data_as(C.POINTER(C.c_int32))))
        return res, pr.astype(bool)

    def synchronize(self):
        self._check(self._L.ccb_synchronize(self._ctx))

    # -- instrumentation ---------------------------------------------------------
    @property
    def kernel_launches(self):
        return self._L.ccb_kernel_launches(self._ctx)

    def profile(self, on=True):
        self._check(self._L.ccb_profile_enable(self._ctx, int(on)))

    def profile_table(self):
        out = {}
        for k, name in enumerate(_capi.PHASE_NAMES):
            ms, calls = C.c_double(), C.c_int64()
            self._L.ccb_profile_get(self._ctx, k, C.byref(ms), C.byref(calls))
            out[name] = (ms.value, calls.value)
        return out

    @property
    def stream(self):
        return self._L.ccb_stream(self._ctx)
