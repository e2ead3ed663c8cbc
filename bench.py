#!/usr/bin/env python
"""bench.py -- particle-steps/s of one full RK2 step of the CCAKE SPH hot path on B200.

    python bench.py --gpus N --steps K --warmup W                    (own arm: the CUDA library behind the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU implementation, oracle/_ref)

One "step" = SPHWorkstation::advance_timestep (two derivative evaluations + RK2 updates) over the whole
particle system. Workload: BASELINE.json configs[3], the configuration the metric is quoted on ("3+1D Milne ... ~8M
particles, slab-split 1-8 GPUs"): 7 989 930 particles with B, S, Q charges on the tabulated EoS; it fits one B200, so
N = 1 runs it whole and N > 1 cuts the SAME system into N slabs (strong scaling). The other configs are --workload
choices. Rank 0 prints ONE JSON line. See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle_steps_per_s"
UNIT = "particle-steps/s"
DEFAULT_WORKLOAD = "config4_milne_bsq_table_3p1d"
# bounded CPU sample: fraction of the lattice extent (configs 1-4) or of the particle count (config 5) the reference's CPU
# implementation runs, sized for 10-30 s of host work at the workload's own spacing, smoothing length and neighbour count
CPU_SCALE = {"config1": 0.25, "config2": 0.5, "config3": 0.07, "config4": 0.1, "config5": 0.1}


# ---------------------------------------------------------------------------------------------
_TABLE = {}


def eos_table(as_file):
    """The synthetic fine-grid table of configs 3 and 4 (ccake_b200/eos_table.py): arrays for the CUDA library
    (ccb_set_eos_table), a thermo.dat under /tmp for the reference build, which reads files."""
    from ccake_b200 import eos_table as E

    if "data" not in _TABLE:
        T_axis, mu_axis = E.default_axes(coarse=False)
        _TABLE["data"] = E.synthetic_table(T_axis, mu_axis, normalize_by_T=True)
    if not as_file:
        return _TABLE["data"]
    d = os.path.join("/tmp", "ccb_eos_fine_%d" % os.getuid())
    if not os.path.exists(os.path.join(d, "thermo.dat")):
        E.write_thermo_dat(d + ".partial", _TABLE["data"])
        os.replace(d + ".partial", d)
    return d


def make_workload(name, scale=1.0, ntiles=1):
    """(settings, ic, description). `scale` < 1 shrinks the lattice EXTENT at fixed spacing and smoothing length
    (same neighbour count per particle): the bounded sample the CPU arms run. `ntiles` > 1: that many copies side by
    side along x (weak scaling: one copy per GPU)."""
    from ccake_b200 import workloads as W

    if ntiles > 1:
        s, ic, desc = make_workload(name, scale, 1)
        s, ic = W.tile_along_x(s, ic, ntiles, s["stepx"])
        return s, ic, desc + " x %d copies along x" % ntiles

    if name == "config2_viscous_gubser_bsq_2p1d":
        xmax = 5.0 * scale
        xmax = round(xmax / 0.025) * 0.025
        s, ic = W.config2_viscous_gubser_bsq(dx=0.025, xmax=xmax)
        desc = "viscous Gubser 2+1D, conformal BSQ EoS, shear + diffusion path, h=0.1, dx=0.025, |x|,|y|<=%.3f" % xmax
    elif name == "config1_ideal_gubser_2p1d":
        xmax = round(5.0 * scale / 0.025) * 0.025
        s, ic = W.config1_ideal_gubser(dx=0.025, xmax=xmax)
        desc = "ideal Gubser 2+1D, conformal EoS, h=0.4, dx=0.025, |x|,|y|<=%.3f" % xmax
    elif name.startswith("config3_trento_pbpb_table_2p1d"):
        n_target = int(float(name.split(":")[1])) if ":" in name else 1_000_000
        s, ic = W.config3_trento_pbpb(n_target=n_target, window=scale)
        desc = ("TRENTo-like Pb-Pb 2+1D, tabulated EoS (synthetic 155x19^3 thermo.dat grid, T-only inversion), shear + cs2-dependent "
                "bulk, h=0.3, dx=%.4f, |x|,|y|<=%.2f" % (s["stepx"], 12.0 * scale))
    elif name.startswith("config4_milne_bsq_table_3p1d"):
        n_target = int(float(name.split(":")[1])) if ":" in name else 8_000_000
        s, ic = W.config4_milne_bsq(n_target=n_target, window=scale)
        desc = ("3+1D Milne with B,S,Q charges, tabulated EoS (synthetic 155x19^3 grid, full 4-D inversion), shear + bulk + diffusion "
                "path, h=2d, d=%.4f, central %.0f%% window" % (s["stepx"], 100 * scale))
    elif name.startswith("config5_cartesian_blob_3p1d"):
        n_target = int(float(name.split(":")[1])) if ":" in name else 1_000_000
        s, ic = W.config5_cartesian_blob(n_target=int(n_target * scale))
        desc = "3+1D Cartesian Gaussian blob, conformal EoS, ideal, h=2dx, ~%d particles" % ic["r"].shape[0]
    else:
        raise SystemExit("unknown workload %r" % name)
    return s, ic, desc


@contextlib.contextmanager
def stdout_to_stderr():
    """The reference build prints progress to stdout; keep stdout for the single JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smmax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smmax = float(f[1]); power.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smmax, "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
def time_reference(workload, scale, steps, warmup, par04_steps=0):
    """The reference's own CPU implementation (oracle/_ref: its unmodified translation units compiled against the
    shims of oracle/shim, OpenMP over ALL host cores) on a bounded sample of the workload. Per-phase seconds come from
    the reference's own CCAKE_PROFILE table (include/phase_profiler.h). `par04_steps` > 0 adds a second leg in which
    phase 04 -- a serial host loop in the reference as shipped (src/sph_workstation.cpp:1737-1762) -- runs under OpenMP
    (oracle/ref_driver.cpp: thermodynamics_openmp), everything else unchanged."""
    os.environ["CCAKE_PROFILE"] = "1"                    # read once, when the PhaseProfiler singleton is first touched
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())  # torchrun exports 1 to its workers
    from oracle import ref

    if not ref.available():
        raise SystemExit("oracle/_ref/libccake_ref.so is missing: run __graft_entry__.build() where /root/reference exists")
    ref.set_threads(os.cpu_count())
    settings, ic, desc = make_workload(workload, scale)
    if settings["eos_type"] == "table":
        settings["eos_path"] = eos_table(as_file=True)

    def phases(before, after):
        out = {}
        for k, (tot, calls) in after.items():
            t0, c0 = before.get(k, (0.0, 0))
            if calls > c0:
                out[k] = {"seconds": round(tot - t0, 4), "calls": calls - c0, "ms_per_call": round(1e3 * (tot - t0) / (calls - c0), 3)}
        return out

    with stdout_to_stderr():
        keys = ref.SETTING_KEYS
        sim = ref.RefSim({k: v for k, v in settings.items() if k in keys or k.startswith("kappa_")})
        sim.add_particles(ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"], ic["u"], ic["bulk"], ic["shv6"])
        sim.initialize()
        n = sim.n
        dt = settings["dt"]
        tw = time.perf_counter()
        for _ in range(warmup):
            sim.advance(dt, 2)
        tw = (time.perf_counter() - tw) / max(warmup, 1)
        if steps is None:  # cpu_baseline leg of the own arm: about 15 s of host work, 2..20 steps
            steps = int(min(20, max(2, round(15.0 / max(tw, 1e-3)))))
        p0 = ref.profile_table()
        t0 = time.perf_counter()
        for _ in range(steps):
            sim.advance(dt, 2)
        el = time.perf_counter() - t0
        p1 = ref.profile_table()
        par = None
        if par04_steps > 0:
            sim.advance_par04(dt, 2)        # builds the per-thread EquationOfState pool (table load): untimed
            p2 = ref.profile_table()
            t0 = time.perf_counter()
            for _ in range(par04_steps):
                sim.advance_par04(dt, 2)
            el2 = time.perf_counter() - t0
            par = {"value": n * par04_steps / el2, "unit": UNIT, "ms_per_step": 1e3 * el2 / par04_steps, "steps": par04_steps,
                   "per_phase": phases(p2, ref.profile_table()),
                   "what": "the same reference step with phase 04 (update_all_particle_thermodynamics, serial as shipped) run under "
                           "OpenMP, one EquationOfState per thread; results bit-identical to the serial loop"}
    cores = ref.omp_threads()
    return dict(value=n * steps / el, seconds=el, n=n, cores=cores, desc=desc, steps=steps, per_phase=phases(p0, p1), phase04_openmp=par)


def reference_arm(args, rank):
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    r = time_reference(args.workload, args.cpu_scale, steps, warmup, par04_steps=max(2, min(steps, 5)))
    out = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * r["seconds"] / steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "description": r["desc"], "particles": r["n"], "rk_order": 2},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                         "sample": "%d of the workload's particles (same spacing, smoothing length and neighbour count; lattice extent x%.2f), "
                                   "%d RK2 steps after %d warm-up" % (r["n"], args.cpu_scale, steps, warmup),
                         "per_phase": r["per_phase"], "phase04_openmp": r["phase04_openmp"],
                         "note": "value = the reference AS SHIPPED: every phase OpenMP over `cores` threads except 04_thermodynamics, which is a "
                                 "serial host loop in the reference; phase04_openmp = the same step with that loop parallelised"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------
# Algorithmic (compulsory) bytes per particle per launch of each kernel: every input column read once, every
# output column written once, 8 B per f64 (DESIGN.md "Kernels and rooflines").
def algorithmic_bytes(D, shear, diffusion):
    rec_g = 36 if diffusion else 20
    grad_out = 5 * D + D * D + ((3 + 3 + 3 * D) if diffusion else 0)
    persistent = 3 + 3 + 4 + 4 + 1 + 5 + (9 if diffusion else 0) + 1 + 4
    evolved = 3 + 3 + 4 + 1 + 5 + (9 if diffusion else 0) + 1
    return {
        # keys + sort + permute of the persistent columns and the RK snapshot. Averaged over the two stages of an RK2 step: the first
        # sort of a step WRITES y0 (the sorted evolved columns stored twice), the second permutes it (read + write)
        "01_neighbour_list": 8 * (3 + 2 * persistent + 1.5 * evolved) + 4 * (2 + 4 * 2 * 3 + 8),
        "02_gamma_velocity": 8 * ((D + 3 + 8) + (1 + D + 8)),
        "03_smooth_fields": 8 * (8 + 1 + 10) + 8,
        "04_thermodynamics": 8 * (4 + D + 3 + 1 + 5 + (9 if diffusion else 0) + 3 + 4) + 8 * (16 + 12 + 4 + 1 + 16 + 12 + 13 + rec_g),
        "09_smooth_gradients": 8 * (rec_g + grad_out),
        "10_time_derivatives": 8 * (28 + 2 * D + 4 + 1 + 16 + (12 + 6 if diffusion else 0) + grad_out + evolved) + 8 * (evolved + D + 4 + 1 + 5 + 3),
    }


PAIR_FLOP = {"03_smooth_fields": 31.0, "09_smooth_gradients": None}  # gradients: computed from D and the switches


def gradient_pair_flop(D, shear, diffusion):
    """fp64 flop per interacting pair in smooth_gradients AS THE REFERENCE WRITES IT (src/sph_workstation.cpp:833-984):
    distance + gradKernel ~ 20, then the listed sums (mul + add each)."""
    f = 20 + 2 * D + 4 + 2 * D * 3 + 3 * D * D
    if shear:
        f += 2 * D + 3 * D + 4 * D * D
    if diffusion:
        f += 3 * (3 + 4 * D + 3 * D)
    return float(f)


def gradient_pair_flop_executed(D, shear, diffusion, currents_zero=False):
    """fp64 flop per pair the CUDA kernel executes (moment form, gather.cuh GradPolicy; FMA = 2): exact distance 3D - 1,
    rsqrt + spline 19, grad W and A_b 2D, sum A_b D, then one FMA per moment. `currents_zero`: the run keeps q^mu_a = 0
    (kappa = 0, no initial diffusion), so the kernel skips the q^0_a and q^i_a moments (DeviceView::qflag)."""
    fma = 3 * D + D * D
    if shear or diffusion:
        fma += D  # K.v_a
    if shear:
        fma += D + D * D
    if diffusion:
        fma += 3 * D + (0 if currents_zero else 3 + 3 * D)
    return float(3 * D - 1 + 19 + 3 * D + (1 if (shear or diffusion) else 0) + 2 * fma)


def parity_check(rank, world, local_rank, table, slab_axis=0):
    """N > 1, inside the driver's own scaling run: the slab-decomposed path exactly as the timed region uses it (per-stage
    migration and column tables over NCCL; the 2h halo once by CUDA-IPC peer memory over NVLink, once by NCCL messages) against a
    SINGLE-context run of the same system on rank 0's GPU. A 3e5-particle instance of the bench workload (same h / spacing, same
    table, same switches), 4 RK2 steps. Integers (ids, neighbour counts) must be equal, the evolved state within 1e-11."""
    import torch.distributed as dist

    from ccake_b200 import SPHWorkstation, slabs, workloads as W

    settings, ic = W.config4_milne_bsq(n_target=300_000)
    settings["eos_table_data"] = table
    settings["slab_axis"] = slab_axis   # the axis the timed run cuts
    dt, steps = settings["dt"], 4
    fields = ("position", "u", "extensive", "extensive_bulk", "extensive_shv")
    n_global = ic["r"].shape[0]
    one = None
    if rank == 0:
        ws1 = SPHWorkstation(settings, device=local_rank)
        ws1.upload_initial_conditions(ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"], ic["u"], ic["bulk"], ic["shv6"])
        for _ in range(steps):
            ws1.advance_timestep(dt)
        ws1.synchronize()
        one = {f: ws1.get(f) for f in fields}
        one["btrack"] = ws1.get_int("btrack")
        ws1.close()
    out = {"workload": "config4_milne_bsq_table_3p1d:3e5", "particles": n_global, "steps": steps, "tolerance": 1e-11, "slab_axis": slab_axis, "halo_modes": {}}
    saved = os.environ.get("CCB_HALO_P2P")
    for mode, env in (("peer_memory", "1"), ("nccl_message", "0")):
        os.environ["CCB_HALO_P2P"] = env
        res = {"ok": False}
        try:
            ws = SPHWorkstation(settings, device=local_rank, rank=rank, nranks=world)
            mn, ed, nc = ws.grid_geometry()
            ax = ws.slab_axis
            cols = slabs.column_of(ic["r"][:, ax], mn[ax], ed[ax], int(nc[ax]))
            ranges = slabs.partition_columns(cols, int(nc[ax]), world)
            part = slabs.split_initial_conditions(ic, cols, ranges[rank])
            ws.comm_init_nccl()
            ws.set_slab(*ranges[rank])
            ws.upload_initial_conditions(part["r"], part["e"], part["rhoB"], part["rhoS"], part["rhoQ"], part["u"], part["bulk"], part["shv6"],
                                         ids=part["id"])
            for _ in range(steps):
                ws.advance_timestep(dt)
            ws.synchronize()
            mine = {f: ws.get_device_order(f) for f in fields}
            mine["id"], mine["btrack"], mine["mig"] = ws.get_int_device_order("id"), ws.get_int_device_order("btrack"), ws.migration_counts()
            ws.close()
            err = None
        except Exception as e:  # a rank that cannot run a mode (no CUDA IPC ...) reports it; the others are told by the gather
            mine, err = None, repr(e)
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, err))
        if rank == 0:
            errs = [g[1] for g in gathered if g[1]]
            if errs:
                res = {"ok": False, "error": errs[0]}
            else:
                ids = np.concatenate([g[0]["id"] for g in gathered])
                ids_ok = bool(np.array_equal(np.sort(ids), np.arange(n_global)))
                worst = 0.0
                for f in fields:
                    got = np.full_like(one[f], np.nan)
                    for g in gathered:
                        got[g[0]["id"]] = g[0][f]
                    worst = max(worst, float(np.max(np.abs(got - one[f])) / max(np.max(np.abs(one[f])), 1e-300)))
                bt = np.full(n_global, -999, dtype=np.int64)
                for g in gathered:
                    bt[g[0]["id"]] = g[0]["btrack"]
                bt_ok = bool(np.array_equal(bt, one["btrack"]))
                res = {"ok": bool(ids_ok and bt_ok and worst < 1e-11), "worst_rel_err": worst, "btrack_equal": bt_ok, "ids_complete": ids_ok,
                       "migrated_out_in": [list(g[0]["mig"]) for g in gathered]}
        out["halo_modes"][mode] = res
    if saved is None:
        os.environ.pop("CCB_HALO_P2P", None)
    else:
        os.environ["CCB_HALO_P2P"] = saved
    if rank == 0:
        out["ok"] = all(m.get("ok") for m in out["halo_modes"].values())
        out["worst_rel_err"] = max((m.get("worst_rel_err", float("inf")) for m in out["halo_modes"].values()), default=None)
        out["btrack_equal"] = all(m.get("btrack_equal", False) for m in out["halo_modes"].values())
    return out


def committed_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture of this
    workload (profiles/traffic.json, written from the .ncu-rep by scripts/ncu_summary.py --traffic); None if not captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(workload, {}).get(kernel)
    except (OSError, ValueError):
        return None


def own_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from ccake_b200 import SPHWorkstation, _capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    # weak scaling (default): one copy of the workload per GPU, side by side along x; strong: the ONE workload cut into slabs
    settings, ic, desc = make_workload(args.workload, 1.0, world if args.scaling == "weak" else 1)
    if settings["eos_type"] == "table":
        settings["eos_table_data"] = eos_table(as_file=False)
    dt = settings["dt"]
    if world > 1 and args.scaling == "strong" and "slab_axis" not in settings and os.environ.get("CCB_SLAB_AXIS", "auto") == "auto":
        # cut along the axis that balances best (config 4 is long and flat in eta: largest slab 1.02x the mean there, 1.09x along x)
        from ccake_b200 import slabs as _slabs

        D_ = ic["r"].shape[1]
        settings["slab_axis"], _ = _slabs.best_axis(ic["r"], [2.0 * settings["hT"]] * D_, world, axes=sorted({0, D_ - 1}))   # x or eta: the two axes the NCCL tests cut
    elif os.environ.get("CCB_SLAB_AXIS", "auto") != "auto":
        settings["slab_axis"] = int(os.environ["CCB_SLAB_AXIS"])
    ws = SPHWorkstation(settings, device=local_rank, rank=rank, nranks=world)
    n_global = ic["r"].shape[0]
    if world > 1:
        # slab decomposition along x: every rank derives the same cut from the (synthetic) IC and uploads its own rows
        from ccake_b200 import slabs

        mn, ed, nc = ws.grid_geometry()
        ax = ws.slab_axis
        cols = slabs.column_of(ic["r"][:, ax], mn[ax], ed[ax], int(nc[ax]))
        # strong scaling: cut at quantiles of the estimated cost (the gathers' work grows with the local density), not of the count
        cost = slabs.particle_cost(ic["r"], mn, ed) if args.scaling == "strong" and os.environ.get("CCB_SLAB_CUT", "cost") == "cost" else None
        ranges = slabs.partition_columns(cols, int(nc[ax]), world, weights=cost, halo_cost=float(os.environ.get("CCB_SLAB_HALO_COST", "0.33")) if args.scaling == "strong" else 0.0)
        part = slabs.split_initial_conditions(ic, cols, ranges[rank])
        ws.comm_init_nccl()
        ws.set_slab(*ranges[rank])
        ws.upload_initial_conditions(part["r"], part["e"], part["rhoB"], part["rhoS"], part["rhoQ"], part["u"], part["bulk"], part["shv6"],
                                     ids=part["id"])
    else:
        ws.upload_initial_conditions(ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"], ic["u"], ic["bulk"], ic["shv6"])
    n = ws.n_particles
    stream = torch.cuda.ExternalStream(ws.stream, device=local_rank)
    D = int(settings["dim"])
    from ccake_b200 import DEFAULTS, derived_switches
    full = dict(DEFAULTS); full.update(settings)
    sw = derived_switches(full)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        ws.advance_timestep(dt)
    ws.synchronize()

    # ---- device-resident timed region ------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ws.profile(True)
    launches0 = ws.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize(); ws.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        ws.advance_timestep(dt)
    e1.record(stream)
    ws.synchronize(); torch.cuda.synchronize(); barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ws.kernel_launches - launches0
    phases = ws.profile_table()
    ws.profile(False)
    clocks = sampler.stop()
    # neighbour counts of the last timed stage (the pair count of the roofline): read here, before the end-to-end leg evolves the
    # system further (an expanding fireball loses neighbours: 245 -> 220 per particle over 40 more steps of config 4)
    bt = ws.get_int_device_order("btrack") if world > 1 else ws.get_int("btrack")

    # ---- end to end through the public API with HOST buffers --------------------------------
    w = ws.state_width
    host = torch.empty(w * (n + n // 4 + 8192), dtype=torch.float64, pin_memory=True).numpy()  # room for arrivals (slabs)
    ws.copy_device_to_host(host)
    for _ in range(2):
        ws.copy_host_to_device(host); ws.advance_timestep(dt); ws.copy_device_to_host(host)
    barrier(); torch.cuda.synchronize()
    e0.record(stream)
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        ws.copy_host_to_device(host)   # SystemState::copy_host_to_device
        ws.advance_timestep(dt)        # SPHWorkstation::advance_timestep
        ws.copy_device_to_host(host)   # SystemState::copy_device_to_host (synchronises)
    e1.record(stream)
    ws.synchronize(); torch.cuda.synchronize()
    t_host = time.perf_counter() - t_host0
    barrier()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), 1e3 * t_host))
    assert np.all(np.isfinite(host[: w * ws.n_particles]))
    del host
    verify = None
    if world > 1 and not args.no_verify:
        verify = parity_check(rank, world, local_rank, eos_table(as_file=False), slab_axis=int(settings.get("slab_axis", 0)))
    # per-rank view of the timed region: device time in the phases that do not wait for a neighbour (compute), and in the two slab
    # exchanges at the head of every stage, where a rank that finished its stage early waits for its neighbours (load imbalance)
    per_rank = None
    if world > 1:
        mine_ = {"particles": int(ws.n_particles),
                 "compute_ms_per_step": sum(v[0] for k, v in phases.items() if not k.startswith(("01a", "01b", "other"))) / args.steps
                 - sum(v[0] for k, v in phases.items() if k.startswith(("01a", "01b"))) / args.steps,
                 "slab_exchange_ms_per_step": sum(v[0] for k, v in phases.items() if k.startswith(("01a", "01b"))) / args.steps}
        gathered_ = [None] * world
        dist.all_gather_object(gathered_, mine_)
        per_rank = gathered_

    if rank != 0:
        return
    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json") if "hbm_gbs" in peaks else (6650.0, "fallback")
    tf = _capi.C.c_double()
    _capi.lib().ccb_measure_dfma_peak(local_rank, _capi.C.byref(tf))
    alg = algorithmic_bytes(D, sw["using_shear"], sw["using_diffusion"])
    per_kernel = {}
    for name, (pms, calls) in phases.items():
        if calls:
            per_kernel[name] = {"ms_per_launch": pms / calls, "launches": calls, "share": pms / ms}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_launch"] * per_kernel[k]["launches"])
    t_dom = per_kernel[dom]["ms_per_launch"] * 1e-3
    achieved = alg.get(dom, 0) * n / t_dom / 1e9
    n = ws.n_particles
    # which member of the EoS fall-back chain located each particle at the end of the run (rank 0's particles): a population that
    # drifts away from the table would show here (EquationOfState::chosen_EOSs order, src/eos_initialization.cpp:249-455)
    eos_members = None
    if settings["eos_type"] == "table":
        em = ws.get_int_device_order("eos") if world > 1 else ws.get_int("eos")
        cnt = np.bincount(np.clip(em, 0, 3), minlength=4)
        eos_members = {k: int(c) for k, c in zip(("table", "tanh_conformal", "conformal", "conformal_diagonal"), cnt)}
    pairs = float(np.sum(bt[bt > 0])) + n
    flop_ref = PAIR_FLOP.get(dom) or gradient_pair_flop(D, sw["using_shear"], sw["using_diffusion"])
    currents_zero = all(float(settings.get("kappa_%d%d" % (a, b), 0.0)) == 0.0 for a in range(3) for b in range(3))
    flop_pair = PAIR_FLOP.get(dom) or gradient_pair_flop_executed(D, sw["using_shear"], sw["using_diffusion"], currents_zero)
    fp64_achieved = pairs * flop_pair / t_dom / 1e12 if dom in ("03_smooth_fields", "09_smooth_gradients") else None
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": committed_traffic(args.workload, dom) if world == 1 else None, "peak_source": peak_src, "algorithmic_bytes_per_particle": alg.get(dom),
                "ms_per_launch": per_kernel[dom]["ms_per_launch"],
                "fp64": {"achieved": fp64_achieved, "peak": tf.value, "unit": "TFLOP/s",
                         "frac": (fp64_achieved / tf.value) if (fp64_achieved and tf.value) else None,
                         "pairs_per_launch": pairs, "flop_per_pair": flop_pair, "flop_per_pair_reference_expression": flop_ref,
                         "peak_source": "ccb_measure_dfma_peak (this run)"},
                "per_kernel": {k: {"ms_per_launch": round(v["ms_per_launch"], 5), "launches": v["launches"], "share": round(v["share"], 4),
                                   "GBps": round(alg.get(k, 0) * n / (v["ms_per_launch"] * 1e-3) / 1e9, 1)} for k, v in per_kernel.items()}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = time_reference(args.workload, args.cpu_scale, None, 1, par04_steps=2)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
               "sample": "%d particles (lattice extent x%.2f at the workload's spacing, smoothing length and neighbour count), %d RK2 steps "
                         "after 1 warm-up, %.1f s" % (r["n"], args.cpu_scale, r["steps"], r["seconds"]),
               "per_phase": r["per_phase"], "phase04_openmp": r["phase04_openmp"],
               "note": "value = the reference as shipped (phase 04 is a serial host loop); phase04_openmp = the same step with that loop under OpenMP"}

    bytes_state = 8 * w * n_global
    out = {
        "metric": METRIC, "value": n_global * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "per_rank": per_rank,
        "config": {"workload": args.workload, "description": desc, "particles": n_global, "particles_rank0": n, "rk_order": 2,
                   "mean_neighbours": float(np.mean(bt[bt >= 0])), "eos_chain_members_rank0": eos_members,
                   "parallelism": "single GPU" if world == 1 else "%d slabs along %s (%s), 2h halo + per-stage migration over NCCL" % (
                       world, "xyz"[ws.slab_axis] if ws.slab_axis < 2 or D == 2 else "eta", "one copy of the workload per GPU" if args.scaling == "weak" else "the one workload cut into contiguous cell-column ranges that minimise the largest slab"),
                   "l2": "no flush: per-step working set %.0f MB of particle columns per GPU > 126 MB L2" % (8e-6 * n * 190)},
        "e2e": {"value": n_global * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": bytes_state,
                "d2h_bytes_per_step": bytes_state, "ms_per_step": ms_e2e / args.steps,
                "api": "copy_host_to_device(pinned state) + advance_timestep + copy_device_to_host via the C ABI"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    if verify is not None:
        out["parity_check"] = verify
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--cpu-scale", type=float, default=None, help="lattice-extent factor of the bounded CPU sample (default: per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong = the ONE workload cut into N slabs (default: the north star's 1 -> 8 GPU curve of the 3+1D config), "
                         "weak = one copy of the workload per GPU side by side along x")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the slab-vs-single-context parity check")
    args = ap.parse_args()

    if args.cpu_scale is None:
        args.cpu_scale = next(v for k, v in CPU_SCALE.items() if args.workload.startswith(k))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        own_arm(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
