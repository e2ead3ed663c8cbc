"""Parity against committed golden vectors (tests/golden/*.npz = outputs of the reference's own translation units,
written by scripts/make_golden.py where /root/reference exists).

  * CPU: a present oracle/_ref still reproduces the fixtures (they are pinned to the reference, not to this library);
  * GPU: the CUDA path, through the C ABI, reproduces them with the bars of BASELINE.json's north star -- cell ids and
    neighbour counts bit-exact, fields 1e-10 (field-scale-relative) per derivative evaluation and per step. These tests do
    NOT need oracle/_ref on the GPU box.
"""
import importlib.util
import os

import numpy as np
import pytest

from common import ROOT, have_ref, rel_err

TOL = 1e-10
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "scripts", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)
CASE_NAMES = ["viscous_2d", "charged_2d_diffusion", "charged_3d_milne_diffusion", "viscous_3d_cartesian", "table_charged_2d_diffusion"]


def _golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def test_fixtures_are_committed():
    for name in CASE_NAMES:
        g = _golden(name)
        assert g["init_position"].shape[0] > 400 and "step3_u" in g.files and "deriv_du_dt" in g.files


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["viscous_2d", "table_charged_2d_diffusion"])
def test_reference_build_reproduces_the_fixtures(name):
    settings, ic = mg.cases()[name]
    fresh, g = mg.generate(name, settings, ic), _golden(name)
    assert set(fresh) == set(g.files)
    for k in g.files:
        if g[k].dtype.kind == "i":
            assert np.array_equal(fresh[k], g[k]), k
        else:  # the reference's sums run through OpenMP atomics: the order, hence the last bits, vary from run to run
            assert rel_err(fresh[k], g[k], max(np.max(np.abs(g[k])), 1e-12)) < 1e-11, k


def _cmp(name, what, got, ref, tol=TOL, scale=None):
    sc = max(np.max(np.abs(ref)), 1e-12) if scale is None else scale
    e = rel_err(got, ref, sc)
    assert e < tol, (name, what, e)


def _cmp_thermo(name, th, th_ref):
    """As tests/test_gpu_parity.py: the phase-diagram point on its common scale; densities to the reference's root tolerance."""
    sc_pt = max(np.max(np.abs(th_ref[:, 0:4])), 1e-12)
    sc_dw = max(np.max(np.abs(th_ref[:, 12:16])), 1e-12)
    for k in range(16):
        sc = sc_pt if k < 4 else sc_dw if k >= 12 else None
        if 5 <= k <= 8:
            assert np.max(np.abs(th[:, k] - th_ref[:, k])) < 1e-10 + TOL * np.max(np.abs(th_ref[:, k])), (name, "thermo%d" % k)
        else:
            _cmp(name, "thermo%d" % k, th[:, k], th_ref[:, k], scale=sc)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASE_NAMES)
def test_cuda_path_reproduces_the_fixtures(name):
    from ccake_b200 import DEFAULTS, SPHWorkstation, derived_switches

    settings, ic = mg.cases()[name]
    g = _golden(name)
    ws = SPHWorkstation(settings)
    ws.upload_initial_conditions(ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"], ic["u"], ic["bulk"], ic["shv6"])
    n = ws.n_particles
    assert n == g["init_position"].shape[0]
    # BSQHydro::initialize_hydrodynamics
    _cmp(name, "init extensive", ws.get("extensive"), g["init_extensive"])
    _cmp(name, "init sigma_lab", ws.get("sigma_lab"), g["init_sigma_lab"])
    _cmp(name, "init extensive_shv", ws.get("extensive_shv"), g["init_extensive_shv"], scale=max(np.max(np.abs(g["init_extensive_shv"])), 1e-6))
    _cmp_thermo(name, ws.get("thermo"), g["init_thermo"])
    # one SPHWorkstation::get_time_derivatives
    ws.get_time_derivatives()
    ws.synchronize()
    assert np.array_equal(ws.grid_shape(), g["deriv_ncell"])
    assert np.array_equal(ws.get_int("cell"), g["deriv_cell"])
    assert np.array_equal(ws.get_int("btrack"), g["deriv_btrack"])
    for f in ("gamma", "v", "sigma_lab", "smoothed", "shv"):
        _cmp(name, f, ws.get(f).reshape(g["deriv_" + f].shape), g["deriv_" + f])
    th_ref = g["deriv_thermo"]
    _cmp_thermo(name, ws.get("thermo"), th_ref)
    if np.any(np.abs(th_ref[:, 6:9]) > 0):
        _cmp(name, "dalpha", ws.get("dalpha"), g["deriv_dalpha"], tol=1e-8)
    _cmp(name, "bulk", ws.get("bulk"), g["deriv_bulk"], scale=max(np.max(np.abs(th_ref[:, 9])), 1e-12))
    pscale = np.max(np.abs(th_ref[:, 9])) / settings["hT"]
    for f in ("gradP", "gradBulk", "gradshear", "divshear"):
        _cmp(name, f, ws.get(f), g["deriv_" + f], scale=pscale)
    _cmp(name, "gradE", ws.get("gradE"), g["deriv_gradE"], scale=3 * pscale)
    _cmp(name, "gradV", ws.get("gradV"), g["deriv_gradV"], scale=1.0 / settings["hT"])
    _cmp(name, "du_dt", ws.get("du_dt"), g["deriv_du_dt"])
    _cmp(name, "d_extensive_dt", ws.get("d_extensive_dt"), g["deriv_d_extensive_dt"], scale=max(np.max(np.abs(g["deriv_d_extensive_dt"])), 1e-12))
    _cmp(name, "d_bulk_dt", ws.get("d_bulk_dt"), g["deriv_d_bulk_dt"])
    _cmp(name, "theta", ws.get("theta"), g["deriv_theta"])
    full = dict(DEFAULTS); full.update(settings)
    sw = derived_switches(full)
    if sw["using_shear"]:
        _cmp(name, "d_shv_dt", ws.get("d_shv_dt"), g["deriv_d_shv_dt"])
    if sw["using_diffusion"]:
        _cmp(name, "d_diffusion_dt", ws.get("d_diffusion_dt"), g["deriv_d_diffusion_dt"])
    # three Evolver RK2 steps
    for s in (1, 2, 3):
        ws.advance_timestep(settings["dt"])
        ws.synchronize()
        tol = TOL * s
        _cmp(name, "position", ws.get("position"), g["step%d_position" % s], tol)
        _cmp(name, "u", ws.get("u"), g["step%d_u" % s], tol)
        _cmp(name, "extensive", ws.get("extensive"), g["step%d_extensive" % s], tol)
        _cmp(name, "extensive_bulk", ws.get("extensive_bulk"), g["step%d_extensive_bulk" % s], tol,
             scale=max(np.max(np.abs(g["step%d_extensive_bulk" % s])), 1e-6))
        _cmp(name, "extensive_shv", ws.get("extensive_shv"), g["step%d_extensive_shv" % s], tol,
             scale=max(np.max(np.abs(g["step%d_extensive_shv" % s])), 1e-6))
    tot = ws.conservation()
    for k, ref in zip(("S", "B", "Sq", "Q", "E", "Ez"), g["step3_totals"]):
        assert abs(tot[k] - ref) <= 1e-9 * max(abs(ref), 1.0), (name, k, tot[k], ref)
    if settings["eos_type"] == "table":  # index into chosen_EOSs: 0 table, 1 tanh_conformal, 2 conformal, 3 conformal_diagonal
        assert np.array_equal(ws.get_int("eos"), g["eos_index"])
