"""Slab decomposition of a particle system over the GPUs of one box (host-side logic).

The neighbour grid is cut along x into contiguous ranges of cell COLUMNS, one per rank, at particle-count
quantiles.  Everything here is plain numpy, so it runs (and is tested) without a GPU; the device
side -- ghost columns, migration, NCCL -- is in ``csrc/ccake_b200.cu`` behind ``ccb_set_slab`` / ``ccb_comm_init``.
"""
from __future__ import annotations

import numpy as np


def column_of(x, grid_min_x, edge_x, ncols):
    """Cell column of every particle: the library's own formula (cell_key_kernel), clamped to the grid."""
    c = np.floor((np.asarray(x, dtype=np.float64) - grid_min_x) * (1.0 / edge_x)).astype(np.int64)
    return np.clip(c, 0, ncols - 1)


def partition_columns(columns, ncols, nranks):
    """Column ranges [lo, hi) per rank with (nearly) equal particle counts; every rank gets at least one column
    and the ranges tile [0, ncols)."""
    if nranks > ncols:
        raise ValueError("more ranks (%d) than cell columns (%d)" % (nranks, ncols))
    counts = np.bincount(columns, minlength=ncols)
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, nranks):
        target = total * r / nranks
        b = int(np.searchsorted(cum, target, side="left"))
        b = max(b, bounds[-1] + 1)            # at least one column per rank
        b = min(b, ncols - (nranks - r))      # leave one column for each remaining rank
        bounds.append(b)
    bounds.append(ncols)
    return [(bounds[r], bounds[r + 1]) for r in range(nranks)]


def split_initial_conditions(ic, columns, col_range):
    """Rows of an IC dict (``r``, ``e``, ``u`` ... as in ccake_b200.workloads) that fall into one slab, + their global ids."""
    lo, hi = col_range
    sel = np.nonzero((columns >= lo) & (columns < hi))[0]
    out = {}
    for k, v in ic.items():
        out[k] = None if v is None else np.ascontiguousarray(np.asarray(v)[sel])
    out["id"] = sel.astype(np.int32)
    return out
