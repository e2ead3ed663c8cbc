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


def particle_cost(r, grid_min, edge, gather_share=0.7):
    """Relative cost of a particle per step: the two neighbour gathers (`gather_share` of a step at mean density) scan the
    particles of the 3^D cells around the target's cell, so their cost grows with the local cell occupancy; everything else
    is per particle. r: (n, D) positions; grid_min, edge: the neighbour grid (ccb_grid_geometry). Mean cost = 1."""
    r = np.asarray(r, dtype=np.float64)
    D = r.shape[1]
    c = np.floor((r - np.asarray(grid_min[:D])) / np.asarray(edge[:D])).astype(np.int64)
    c -= c.min(0)
    dims = c.max(0) + 1
    key = np.ravel_multi_index(tuple(c[:, d] for d in range(D)), tuple(int(x) for x in dims))
    occ = np.bincount(key)[key].astype(np.float64)
    return (1.0 - gather_share) + gather_share * occ / occ.mean()


def partition_columns(columns, ncols, nranks, weights=None, halo_cost=0.0):
    """Column ranges [lo, hi) per rank with (nearly) equal particle counts -- or, with per-particle `weights` (particle_cost),
    equal summed cost; every rank gets at least one column and the ranges tile [0, ncols). `halo_cost` > 0 charges a slab for each
    ghost column it has to receive and gather against, as that fraction of the ghost column's own load (measured on eight B200s:
    an interior slab of config 4 costs what an edge slab with 12 % more particles costs, i.e. ~0.33 of a column per side)."""
    if nranks > ncols:
        raise ValueError("more ranks (%d) than cell columns (%d)" % (nranks, ncols))
    counts = np.bincount(columns, weights=weights, minlength=ncols).astype(np.float64)
    cum = np.concatenate([[0.0], np.cumsum(counts)])
    # contiguous ranges that minimise the LARGEST load (the step runs at the pace of the fullest slab): dynamic programme over
    # (ranks used, columns covered); a column is the unit, so with few columns per rank the plain quantile cut can be 5-10 % worse
    INF = float("inf")
    best = np.full((nranks + 1, ncols + 1), INF)
    arg = np.zeros((nranks + 1, ncols + 1), dtype=np.int64)
    best[0, 0] = 0.0
    for r in range(1, nranks + 1):
        for j in range(r, ncols - (nranks - r) + 1):
            i = np.arange(r - 1, j)                      # the last range is [i, j): at least one column
            load = cum[j] - cum[i]
            if halo_cost > 0.0:
                load = load + halo_cost * (np.where(i > 0, counts[np.maximum(i - 1, 0)], 0.0) + (counts[j] if j < ncols else 0.0))
            cand = np.maximum(best[r - 1, i], load)
            k = int(np.argmin(cand))
            best[r, j], arg[r, j] = cand[k], i[k]
    bounds = [ncols]
    for r in range(nranks, 0, -1):
        bounds.append(int(arg[r, bounds[-1]]))
    bounds = bounds[::-1]
    return [(bounds[r], bounds[r + 1]) for r in range(nranks)]


def best_axis(r, edge, nranks, halo_cost=0.33, axes=None):
    """The spatial axis whose cell columns cut into `nranks` slabs with the smallest largest load, a slab's load being its
    particles + halo_cost x the particles of the ghost columns it gathers against (an event that is long and flat in rapidity
    balances better along eta by count, but its eta columns are thicker): (axis, largest load / mean slab). r: (n, D); edge: cell
    edge per axis."""
    r = np.asarray(r, dtype=np.float64)
    out = None
    for ax in (range(r.shape[1]) if axes is None else axes):
        lo = r[:, ax].min() - 1e-9 * edge[ax]
        ncols = int(np.floor((r[:, ax].max() - lo) / edge[ax])) + 1
        if ncols < nranks:
            continue
        cols = column_of(r[:, ax], lo, edge[ax], ncols)
        counts = np.bincount(cols, minlength=ncols)
        ranges = partition_columns(cols, ncols, nranks, halo_cost=halo_cost)
        worst = max(counts[a:b].sum() + halo_cost * ((counts[a - 1] if a > 0 else 0) + (counts[b] if b < ncols else 0))
                    for a, b in ranges) * nranks / float(r.shape[0])
        if out is None or worst < out[1] - 1e-3:
            out = (ax, float(worst))
    if out is None:
        raise ValueError("no axis has %d cell columns" % nranks)
    return out


def split_initial_conditions(ic, columns, col_range):
    """Rows of an IC dict (``r``, ``e``, ``u`` ... as in ccake_b200.workloads) that fall into one slab, + their global ids."""
    lo, hi = col_range
    sel = np.nonzero((columns >= lo) & (columns < hi))[0]
    out = {}
    for k, v in ic.items():
        out[k] = None if v is None else np.ascontiguousarray(np.asarray(v)[sel])
    out["id"] = sel.astype(np.int32)
    return out
