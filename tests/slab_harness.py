"""N slab contexts inside one process, one host thread each, linked by the library's in-process transport
(ccb_comm_init_local): the whole multi-GPU logic on a single GPU."""
from __future__ import annotations

import itertools
import threading

import numpy as np

_group = itertools.count(1)


class SlabRun:
    def __init__(self, settings, ic, nranks, device=0):
        from ccake_b200 import SPHWorkstation, slabs

        self.nranks, self.settings = nranks, settings
        group = next(_group)
        self.ws = [SPHWorkstation(settings, device=device, rank=r, nranks=nranks) for r in range(nranks)]
        mn, ed, nc = self.ws[0].grid_geometry()
        cols = slabs.column_of(ic["r"][:, 0], mn[0], ed[0], int(nc[0]))
        self.ranges = slabs.partition_columns(cols, int(nc[0]), nranks)
        parts = [slabs.split_initial_conditions(ic, cols, rg) for rg in self.ranges]
        self.n_global = ic["r"].shape[0]
        for r, w in enumerate(self.ws):
            w.comm_init_local(group)
            w.set_slab(*self.ranges[r])
        self.each(lambda r, w: w.upload_initial_conditions(parts[r]["r"], parts[r]["e"], parts[r]["rhoB"], parts[r]["rhoS"], parts[r]["rhoQ"],
                                                            parts[r]["u"], parts[r]["bulk"], parts[r]["shv6"], ids=parts[r]["id"]))

    def each(self, fn):
        """Run fn(rank, workstation) on every rank concurrently (the ranks talk to each other inside the calls)."""
        out, errs = [None] * self.nranks, []

        def run(r):
            try:
                out[r] = fn(r, self.ws[r])
            except BaseException as e:  # noqa: BLE001
                errs.append((r, e))

        th = [threading.Thread(target=run, args=(r,)) for r in range(self.nranks)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0][1]
        return out

    def gather(self, field):
        """A field of ALL particles in global particle-id order."""
        parts = [(w.get_int_device_order("id"), w.get_device_order(field)) for w in self.ws]
        first = parts[0][1]
        out = np.full((self.n_global,) + first.shape[1:], np.nan)
        for ids, vals in parts:
            out[ids] = vals
        return out

    def gather_int(self, field):
        out = np.full(self.n_global, -999, dtype=np.int64)
        for w in self.ws:
            out[w.get_int_device_order("id")] = w.get_int_device_order(field)
        return out

    def counts(self):
        return [w.n_particles for w in self.ws]
