"""Slab decomposition. CPU: the partition logic (also as a 2-process gloo job, the way the ranks agree on it).
GPU: 2 and 3 slabs on one GPU through the in-process transport against the single-context run and the reference."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import ROOT, charged_settings, gubser_ic, have_ref, make_ref, rel_err, viscous_settings, blob_ic_3d


def test_partition_columns_balanced_and_tiling():
    from ccake_b200 import slabs

    rng = np.random.default_rng(0)
    x = rng.normal(0.0, 2.0, 100000)
    ncols = 64
    cols = slabs.column_of(x, -8.0, 0.25, ncols)
    for nranks in (1, 2, 3, 8):
        rg = slabs.partition_columns(cols, ncols, nranks)
        assert rg[0][0] == 0 and rg[-1][1] == ncols
        assert all(a[1] == b[0] for a, b in zip(rg, rg[1:])) and all(hi > lo for lo, hi in rg)
        n = [np.count_nonzero((cols >= lo) & (cols < hi)) for lo, hi in rg]
        assert sum(n) == x.size and max(n) < 1.35 * x.size / nranks


def test_partition_degenerate_cases():
    from ccake_b200 import slabs

    cols = np.zeros(1000, dtype=np.int64)  # everything in column 0: the other ranks still get (empty) columns
    rg = slabs.partition_columns(cols, 4, 4)
    assert rg == [(0, 1), (1, 2), (2, 3), (3, 4)]
    with pytest.raises(ValueError):
        slabs.partition_columns(cols, 2, 3)
    ic = dict(r=np.arange(10.0)[:, None] * np.ones((1, 2)), e=np.arange(10.0), u=None)
    part = slabs.split_initial_conditions(ic, np.arange(10), (3, 6))
    assert list(part["id"]) == [3, 4, 5] and part["u"] is None and np.array_equal(part["e"], [3.0, 4.0, 5.0])


_GLOO_JOB = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from ccake_b200 import slabs
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(5)
x = rng.normal(0.0, 1.5, 20000)                      # every rank holds the same IC and derives the same cut
cols = slabs.column_of(x, -6.0, 0.2, 60)
rg = slabs.partition_columns(cols, 60, world)
mine = slabs.split_initial_conditions(dict(r=np.stack([x, x], 1), e=x), cols, rg[rank])
n = torch.tensor([mine["id"].size], dtype=torch.int64)
dist.all_reduce(n)
ids = [None] * world
dist.all_gather_object(ids, mine["id"].tolist())
allids = sorted(i for part in ids for i in part)
assert n.item() == x.size and allids == list(range(x.size)), "slabs must tile the particle set"
lo = torch.tensor([rg[rank][0]]); his = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(his, torch.tensor([rg[rank][1]], dtype=torch.int64))
if rank > 0: assert his[rank - 1].item() == rg[rank][0]
# the strong-scaling bench also chooses the slab axis and a halo-aware cut from the IC: every rank must arrive at the same answer
r3 = np.stack([x, rng.normal(0.0, 1.5, x.size), rng.uniform(-5.0, 5.0, x.size)], 1)
ax, worst = slabs.best_axis(r3, [0.2, 0.2, 0.2], world, axes=[0, 2])
rg2 = slabs.partition_columns(slabs.column_of(r3[:, ax], r3[:, ax].min() - 1e-9, 0.2, 80), 80, world, halo_cost=0.33)
answers = [None] * world
dist.all_gather_object(answers, (int(ax), round(float(worst), 12), rg2))
assert all(a == answers[0] for a in answers), answers
dist.destroy_process_group()
print("OK", rank)
"""


def test_ranks_agree_on_the_partition_gloo_world2(tmp_path):
    script = tmp_path / "job.py"
    script.write_text(_GLOO_JOB % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("OK") == 2


# ---------------------------------------------------------------------------------------------
D3 = dict(dim=3, hT=0.5, hEta=0.5, stepx=0.25, stepy=0.25, stepEta=0.25, xmin=-1.5, ymin=-1.5, etamin=-1.5)
SLAB_CASES = {
    "viscous_2d": (lambda: viscous_settings(dt=0.02), lambda: gubser_ic(dx=0.1, shear=True)),
    "charged_2d": (lambda: charged_settings(dt=0.02), lambda: gubser_ic(dx=0.1, shear=True, charges=True)),
    "viscous_3d_cartesian": (lambda: viscous_settings(coordinate_system="cartesian", dt=0.02, **D3), lambda: blob_ic_3d(n_side=14)),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(SLAB_CASES))
@pytest.mark.parametrize("nranks", [2, 3])
def test_slabs_match_single_context_and_reference(case, nranks):
    """Same physics whatever the number of slabs: every field after 4 RK2 steps (large dt, so that particles do migrate)
    agrees with the single-context run to rounding and with the reference build to 1e-10 per step; ids stay a permutation."""
    from ccake_b200 import SPHWorkstation
    from slab_harness import SlabRun

    mk_s, mk_ic = SLAB_CASES[case]
    settings, ic = mk_s(), mk_ic()
    one = SPHWorkstation(settings)
    one.upload_initial_conditions(ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"], ic["u"], ic["bulk"], ic["shv6"])
    run = SlabRun(settings, ic, nranks)
    assert sum(run.counts()) == one.n_particles
    assert rel_err(run.gather("sigma_lab"), one.get("sigma_lab")) < 1e-13
    assert rel_err(run.gather("extensive"), one.get("extensive")) < 1e-13
    steps, dt = 4, settings["dt"]
    for _ in range(steps):
        one.advance_timestep(dt)
        run.each(lambda r, w: w.advance_timestep(dt))
    one.synchronize()
    run.each(lambda r, w: w.synchronize())
    assert sum(run.counts()) == one.n_particles
    ids = np.sort(np.concatenate([w.get_int_device_order("id") for w in run.ws]))
    assert np.array_equal(ids, np.arange(one.n_particles))
    for f in ("position", "u", "extensive", "extensive_shv", "extensive_bulk", "sigma_lab", "gradP", "du_dt"):
        ref = one.get(f)
        assert rel_err(run.gather(f), ref, max(np.max(np.abs(ref)), 1e-12)) < 1e-12, (case, nranks, f)
    assert np.array_equal(run.gather_int("btrack"), one.get_int("btrack"))
    tot1 = one.conservation()
    totn = run.each(lambda r, w: w.conservation())
    for k in ("S", "B", "Sq", "Q", "E"):
        assert abs(totn[0][k] - tot1[k]) <= 1e-11 * max(abs(tot1[k]), 1.0), (k, totn[0][k], tot1[k])
        assert all(t[k] == totn[0][k] for t in totn)      # every rank holds the same totals
    if have_ref():
        sim = make_ref(settings, ic)
        sim.initialize()
        for _ in range(steps):
            sim.advance(dt, 2)
        assert rel_err(run.gather("u"), sim.get("hydro_vector", "u")) < steps * 1e-10
        assert rel_err(run.gather("extensive")[:, 0], sim.member("extensive")[:, 0]) < steps * 1e-10
    moved = [w.migration_counts() for w in run.ws]
    print(case, nranks, "ranges", run.ranges, "counts", run.counts(), "migrated (out, in)", moved)


@pytest.mark.gpu
def test_slabs_rk4_accumulators_follow_the_particles():
    """rk_order 4 with two slabs: the k accumulator columns migrate with their particles, so the slab run equals the
    single-context run to rounding after steps in which particles change slab."""
    from ccake_b200 import SPHWorkstation
    from slab_harness import SlabRun

    settings, ic = viscous_settings(dt=0.02, rk_order=4), gubser_ic(dx=0.1, shear=True)
    one = SPHWorkstation(settings)
    one.upload_initial_conditions(ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"], ic["u"], ic["bulk"], ic["shv6"])
    run = SlabRun(settings, ic, 2)
    for _ in range(3):
        one.advance_timestep(settings["dt"])
        run.each(lambda r, w: w.advance_timestep(settings["dt"]))
    one.synchronize()
    run.each(lambda r, w: w.synchronize())
    assert sum(m[0] for m in (w.migration_counts() for w in run.ws)) > 0
    for f in ("position", "u", "extensive", "extensive_shv"):
        ref = one.get(f)
        assert rel_err(run.gather(f), ref, max(np.max(np.abs(ref)), 1e-12)) < 1e-12, f


@pytest.mark.gpu
def test_peer_memory_halo_matches_halo_messages(tmp_path):
    """CCB_HALO_P2P=1 with two in-process slabs: the boundary tasks read the neighbour's record array directly, ordered by the
    ready / done flags, and give the single-context answer like the message path does. Run in a child process: the switch is read
    at communicator creation, and the spinning wait kernels want every stream on its own hardware queue."""
    job = tmp_path / "p2p_job.py"
    job.write_text(r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from common import viscous_settings, charged_settings, gubser_ic, rel_err
from ccake_b200 import SPHWorkstation
from slab_harness import SlabRun
for settings, ic in ((viscous_settings(dt=0.02), gubser_ic(dx=0.1, shear=True)), (charged_settings(dt=0.02), gubser_ic(dx=0.1, shear=True, charges=True))):
    one = SPHWorkstation(settings)
    one.upload_initial_conditions(ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"], ic["u"], ic["bulk"], ic["shv6"])
    run = SlabRun(settings, ic, 2)
    for _ in range(4):
        one.advance_timestep(settings["dt"]); run.each(lambda r, w: w.advance_timestep(settings["dt"]))
    one.synchronize(); run.each(lambda r, w: w.synchronize())
    for f in ("position", "u", "extensive", "gradP", "du_dt"):
        ref = one.get(f)
        assert rel_err(run.gather(f), ref, max(np.max(np.abs(ref)), 1e-12)) < 1e-12, f
    assert np.array_equal(run.gather_int("btrack"), one.get_int("btrack"))
print("P2P OK")
""" % (ROOT, os.path.join(ROOT, "tests")))
    env = dict(os.environ, CCB_HALO_P2P="1", CUDA_DEVICE_MAX_CONNECTIONS="32")
    out = subprocess.run([sys.executable, str(job)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "P2P OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_min_max_partition_and_axis_choice():
    """partition_columns minimises the largest slab over all contiguous cuts (brute force on a small case), charges ghost columns when
    asked to, and best_axis prefers the axis along which the system is flat."""
    import itertools

    from ccake_b200 import slabs

    rng = np.random.default_rng(3)
    counts = rng.integers(1, 50, 12)
    cols = np.repeat(np.arange(12), counts)
    for nr in (2, 3, 4):
        rg = slabs.partition_columns(cols, 12, nr)
        got = max(counts[a:b].sum() for a, b in rg)
        best = min(max(counts[a:b].sum() for a, b in zip((0,) + cut, cut + (12,))) for cut in itertools.combinations(range(1, 12), nr - 1))
        assert got == best and rg[0][0] == 0 and rg[-1][1] == 12 and all(rg[i][1] == rg[i + 1][0] for i in range(nr - 1))
    # ghost columns charged: the interior slab of three gives way to the edge slabs
    flat = np.repeat(np.arange(12), 10)
    assert slabs.partition_columns(flat, 12, 3) == [(0, 4), (4, 8), (8, 12)]
    a, b = slabs.partition_columns(flat, 12, 3, halo_cost=1.0)[1]
    assert b - a < 4
    # a system that is a flat slab along z and a narrow Gaussian along x cuts better along z
    r = np.stack([rng.normal(0.0, 0.5, 200000), rng.normal(0.0, 0.5, 200000), rng.uniform(-4.0, 4.0, 200000)], 1)
    ax, worst = slabs.best_axis(r, [0.25, 0.25, 0.25], 8, halo_cost=0.0)
    assert ax == 2 and worst < 1.1
    w = slabs.particle_cost(r, r.min(0), [0.25, 0.25, 0.25])
    assert abs(w.mean() - 1.0) < 1e-12 and w.min() >= 0.3
