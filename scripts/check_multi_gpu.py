#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the slab-decomposed run over NCCL -- per-stage migration and
column tables, the 2h halo by CUDA-IPC peer memory (CCB_HALO_P2P=1) or by NCCL messages (CCB_HALO_P2P=0) -- against a
single-context run of the same system on rank 0's GPU, after steps in which particles migrate. Prints one JSON line on rank 0;
exit code 1 on mismatch.

    torchrun --nproc-per-node 2 scripts/check_multi_gpu.py [--workload config2|config4] [--axis 0|1|2] [--steps K]

tests/test_multi_gpu.py runs it on min(2, device_count) GPUs; bench.py --gpus N runs the same comparison inside the driver's
scaling run (parity_check in its JSON line)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ccake_b200 import SPHWorkstation, slabs, workloads as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config2", choices=["config2", "config4"])
ap.add_argument("--axis", type=int, default=0)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if args.workload == "config2":
    settings, ic = W.config2_viscous_gubser_bsq(dx=0.05, xmax=5.0)
    settings = dict(settings, dt=0.01)
else:
    from ccake_b200 import eos_table as E

    settings, ic = W.config4_milne_bsq(n_target=150_000)
    T_axis, mu_axis = E.default_axes(coarse=True)
    settings["eos_table_data"] = E.synthetic_table(T_axis, mu_axis, normalize_by_T=True)
settings["slab_axis"] = args.axis
steps = args.steps
ws = SPHWorkstation(settings, device=local, rank=rank, nranks=world)
mn, ed, nc = ws.grid_geometry()
ax = ws.slab_axis
cols = slabs.column_of(ic["r"][:, ax], mn[ax], ed[ax], int(nc[ax]))
ranges = slabs.partition_columns(cols, int(nc[ax]), world)
part = slabs.split_initial_conditions(ic, cols, ranges[rank])
ws.comm_init_nccl()
ws.set_slab(*ranges[rank])
ws.upload_initial_conditions(part["r"], part["e"], part["rhoB"], part["rhoS"], part["rhoQ"], part["u"], part["bulk"], part["shv6"], ids=part["id"])
for _ in range(steps):
    ws.advance_timestep(settings["dt"])
ws.synchronize()
n_global = ic["r"].shape[0]
mine = {f: ws.get_device_order(f) for f in ("position", "u", "extensive", "extensive_shv", "gradP")}
ids = ws.get_int_device_order("id")
bt = ws.get_int_device_order("btrack")
gathered = [None] * world
dist.all_gather_object(gathered, (ids, bt, mine, ws.migration_counts()))
ok = True
if rank == 0:
    one = SPHWorkstation(settings, device=local)
    one.upload_initial_conditions(ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"], ic["u"], ic["bulk"], ic["shv6"])
    for _ in range(steps):
        one.advance_timestep(settings["dt"])
    one.synchronize()
    allids = np.sort(np.concatenate([g[0] for g in gathered]))
    ok &= bool(np.array_equal(allids, np.arange(n_global)))
    worst = 0.0
    for f in mine:
        ref = one.get(f)
        got = np.full_like(ref, np.nan)
        for g in gathered:
            got[g[0]] = g[2][f]
        err = float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300))
        worst = max(worst, err)
        ok &= err < 1e-11
    btg = np.full(n_global, -999)
    for g in gathered:
        btg[g[0]] = g[1]
    bt_ok = bool(np.array_equal(btg, one.get_int("btrack")))
    ok &= bt_ok
    print(json.dumps({"check": "multi_gpu", "world": world, "workload": args.workload, "slab_axis": ax, "halo_p2p": os.environ.get("CCB_HALO_P2P", "default"),
                      "particles": n_global, "steps": steps, "migrated_out_in": [list(g[3]) for g in gathered], "worst_rel_err": worst,
                      "btrack_equal": bt_ok, "ok": bool(ok)}), flush=True)
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(1 if flag.item() else 0)
