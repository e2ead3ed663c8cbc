#!/usr/bin/env python
"""Golden fixtures: outputs of the REFERENCE ITSELF (oracle/_ref = its unmodified translation units, oracle/Makefile) on
small seeded initial conditions, written to tests/golden/*.npz. Run where /root/reference exists:

    python scripts/make_golden.py

The GPU parity tests of tests/test_golden_fixtures.py compare the CUDA path with these files, so that parity does not
depend on oracle/_ref having travelled to the GPU box; a CPU test checks that a present oracle/_ref still reproduces them.
The inputs are not stored: they are regenerated from the same seeded functions of tests/common.py (CASES below)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")

from common import blob_ic_3d, charged_settings, gubser_ic, make_ref, ref_dalpha12, ref_thermo16, viscous_settings  # noqa: E402

D3 = dict(dim=3, hT=0.5, hEta=0.5, stepx=0.25, stepy=0.25, stepEta=0.25, xmin=-1.5, ymin=-1.5, etamin=-1.5)


def table_dir():
    """The coarse synthetic table (ccake_b200/eos_table.py): deterministic, so it is regenerated, not stored."""
    from ccake_b200 import eos_table

    return eos_table.ensure_table(os.path.join("/tmp", "ccb_eos_golden_%d" % os.getuid()), coarse=True)


def cases():
    tab = dict(eos_type="table", eos_path=table_dir(), normalize_by_T=True, Freeze_Out_Temperature=150.0)
    return {
        "viscous_2d": (viscous_settings(stepx=0.2, stepy=0.2), gubser_ic(dx=0.2, shear=True)),
        "charged_2d_diffusion": (charged_settings(stepx=0.2, stepy=0.2), gubser_ic(dx=0.2, shear=True, charges=True)),
        "charged_3d_milne_diffusion": (charged_settings(**D3), blob_ic_3d(n_side=9)),
        "viscous_3d_cartesian": (viscous_settings(coordinate_system="cartesian", **D3), blob_ic_3d(n_side=9)),
        "table_charged_2d_diffusion": (charged_settings(stepx=0.2, stepy=0.2, **tab), gubser_ic(dx=0.2, shear=True, charges=True)),
    }


def ext_shv5(sim, field="extensive_shv"):
    es = sim.get("hydro_shear_aux_vector", field)
    return np.stack([es[:, 0, 0], es[:, 0, 1], es[:, 0, 2], es[:, 1, 1], es[:, 1, 2]], 1)


def snapshot_derivatives(sim):
    n = sim.n
    off, idx, cell, nc = sim.neighbours()
    return dict(
        cell=cell.astype(np.int32), ncell=np.asarray(nc, dtype=np.int32), btrack=np.diff(off).astype(np.int32),
        gamma=sim.get("hydro_scalar", "gamma"), v=sim.get("hydro_vector", "v"), sigma_lab=sim.get("hydro_scalar", "sigma_lab"),
        smoothed=sim.member("smoothed")[:, [0, 2, 3, 4]], thermo=ref_thermo16(sim), dalpha=ref_dalpha12(sim),
        bulk=sim.get("hydro_scalar", "bulk"), shv=sim.get("hydro_spacetime_matrix", "shv").reshape(n, 16),
        gradP=sim.get("hydro_vector", "gradP"), gradE=sim.get("hydro_vector", "gradE"), gradBulk=sim.get("hydro_vector", "gradBulk"),
        gradshear=sim.get("hydro_vector", "gradshear"), divshear=sim.get("hydro_vector", "divshear"),
        gradV=sim.get("hydro_space_matrix", "gradV").reshape(n, 9), du_dt=sim.get("hydro_vector", "du_dt"),
        d_extensive_dt=sim.member("d_dt_extensive")[:, [0, 2, 3, 4]], d_bulk_dt=sim.get("hydro_scalar", "d_extensive_bulk_dt"),
        d_shv_dt=ext_shv5(sim, "d_extensive_shv_dt"), theta=sim.get("hydro_scalar", "theta"),
        d_diffusion_dt=sim.get("hydro_space_matrix", "d_extensive_q_dt").reshape(n, 9))


def snapshot_state(sim):
    return dict(position=sim.member("position"), u=sim.get("hydro_vector", "u"), extensive=sim.member("extensive")[:, [0, 2, 3, 4]],
                extensive_bulk=sim.get("hydro_scalar", "extensive_bulk"), extensive_shv=ext_shv5(sim),
                totals=np.array([sim.totals()[k] for k in ("S", "B", "Sq", "Q", "E", "Ez")]))


def generate(name, settings, ic, steps=3):
    sim = make_ref(settings, ic)
    sim.initialize()
    out = {"init_" + k: v for k, v in snapshot_state(sim).items()}
    out["init_thermo"] = ref_thermo16(sim)
    out["init_sigma_lab"] = sim.get("hydro_scalar", "sigma_lab")
    sim.get_time_derivatives()
    out.update({"deriv_" + k: v for k, v in snapshot_derivatives(sim).items()})
    for s in range(steps):
        sim.advance(settings["dt"], 2)
        out.update({"step%d_%s" % (s + 1, k): v for k, v in snapshot_state(sim).items()})
    out["eos_index"] = sim.eos_index().astype(np.int32)
    return out


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    for name, (settings, ic) in cases().items():
        data = generate(name, settings, ic)
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **data)
        print(name, ic["r"].shape[0], "particles ->", path, "%.0f kB" % (os.path.getsize(path) / 1e3), file=sys.stderr)
