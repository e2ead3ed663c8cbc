#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the slab-decomposed run over NCCL (and, with CCB_HALO_P2P=1, the
peer-memory halo over CUDA IPC) against a single-context run of the same system on rank 0's GPU, after steps in which particles
migrate. Prints one line; exit code 1 on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ccake_b200 import SPHWorkstation, slabs, workloads as W  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
settings, ic = W.config2_viscous_gubser_bsq(dx=0.05, xmax=5.0)
settings = dict(settings, dt=0.01)
steps = 5
ws = SPHWorkstation(settings, device=local, rank=rank, nranks=world)
mn, ed, nc = ws.grid_geometry()
cols = slabs.column_of(ic["r"][:, 0], mn[0], ed[0], int(nc[0]))
ranges = slabs.partition_columns(cols, int(nc[0]), world)
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
        ok &= err < 1e-12
    btg = np.full(n_global, -999)
    for g in gathered:
        btg[g[0]] = g[1]
    ok &= bool(np.array_equal(btg, one.get_int("btrack")))
    print("multi-gpu check: world=%d p2p=%s particles=%d migrated(out,in)=%s worst rel err %.2e btrack equal %s -> %s" % (
        world, os.environ.get("CCB_HALO_P2P", "0"), n_global, [g[3] for g in gathered], worst, bool(np.array_equal(btg, one.get_int("btrack"))),
        "OK" if ok else "MISMATCH"), flush=True)
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(1 if flag.item() else 0)
