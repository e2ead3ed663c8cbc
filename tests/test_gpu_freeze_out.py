"""Freeze-out on the GPU (reference FreezeOut<D>, include/freeze_out.h; SPHWorkstation::freeze_out_particles): the 0 -> 1 -> 3 -> 4
status machine, the back-up snapshots of particles that lose their neighbours, and the hypersurface records, step by step against
the reference build -- integers bit-exact, hypersurface fields at the per-step bar."""
import numpy as np
import pytest

from common import blob_ic_3d, gubser_ic, have_ref, make_ref, viscous_settings
from test_gpu_parity import D3, TOL, _make_ws

pytestmark = [pytest.mark.gpu]


def _dense_2d():
    s = viscous_settings(particlization_enabled=True, Freeze_Out_Temperature=300.0, dt=0.05)
    return s, gubser_ic(dx=0.1, shear=True), 14


def _sparse_2d():
    """h = 0.15 on a jittered 0.27 fm lattice: 0 .. 5 neighbours per particle, so that the few-neighbour (0 < btrack <= 3) and
    no-neighbour (btrack = 0 -> frozen at once, btrack = -1) branches of FreezeOut::freeze_out_evo run"""
    s = viscous_settings(particlization_enabled=True, Freeze_Out_Temperature=300.0, dt=0.05, hT=0.15)
    ic = gubser_ic(dx=0.27, xmax=2.0, shear=True)
    ic["r"] = ic["r"] + np.random.default_rng(4).uniform(-0.06, 0.06, ic["r"].shape)
    return s, ic, 10


def _blob_3d(cartesian):
    s = viscous_settings(particlization_enabled=True, Freeze_Out_Temperature=250.0, dt=0.05, **D3)
    if cartesian:
        s["coordinate_system"] = "cartesian"
    return s, blob_ic_3d(n_side=12), 10


CASES = {"dense_2d": _dense_2d, "sparse_2d": _sparse_2d, "blob_3d_milne": lambda: _blob_3d(False), "blob_3d_cartesian": lambda: _blob_3d(True)}
# hypersurface fields and the natural scale each is compared on
VEC = {"shearsub": (0, 16), "diffout": (16, 28), "divT": (28, 31), "divE": (31, 34), "divP": (34, 37), "rsub": (37, 40), "uout": (40, 43)}
SCALARS = ["divTtemp", "divPpress", "divEener", "gsub", "bulksub", "thetasub", "swsub", "shear33sub", "tlist", "sFO", "Efluc", "Tfluc", "muBfluc",
           "muSfluc", "muQfluc", "cs2fzfluc", "wfzfluc", "rhoBfluc", "rhoSfluc", "rhoQfluc", "eos_type"]


@pytest.mark.parametrize("case", sorted(CASES))
def test_freeze_out_follows_the_reference(case):
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    settings, ic, steps = CASES[case]()
    sim = make_ref(settings, ic)
    sim.initialize()
    ws = _make_ws(settings, ic)
    assert np.array_equal(ws.get_int("freeze"), sim.ints("freeze")), "initial check_freeze_out_status"
    frozen_total = 0
    for step in range(steps):
        sim.advance(settings["dt"], 2)
        ws.advance_timestep(settings["dt"])   # ends with freeze_out_particles (particlization_enabled)
        ws.synchronize()
        assert np.array_equal(ws.get_int("freeze"), sim.ints("freeze")), (case, step)
        assert np.array_equal(ws.get_int("btrack"), sim.ints("btrack")), (case, step)
        res, pr = ws.freeze_out_results()
        res_ref, pr_ref = sim.freeze_results()
        assert np.array_equal(pr, pr_ref), (case, step)
        frozen_total += int(pr.sum())
        if not pr.any():
            continue
        a, b = res[pr], res_ref[pr]
        tol = TOL * (step + 2)
        for name, (lo, hi) in VEC.items():
            sc = max(np.max(np.abs(b[:, lo:hi])), 1e-12)
            assert np.max(np.abs(a[:, lo:hi] - b[:, lo:hi])) <= tol * sc, (case, step, name)
        for k, name in enumerate(SCALARS):
            sc = max(np.max(np.abs(b[:, 43 + k])), 1e-12)
            assert np.max(np.abs(a[:, 43 + k] - b[:, 43 + k])) <= tol * sc, (case, step, name)
    assert frozen_total > 20, "the scenario must actually freeze particles out"
    assert ws.conservation()["n_frozen"] == int((sim.ints("freeze") >= 4).sum())
