"""A small run that touches every kernel family, for compute-sanitizer (memcheck / racecheck): 3-D Milne with charges, shear and
diffusion on the conformal EoS and on the table (hit masks, both sparsity twins, the Newton search), freeze-out, RK4, and a
two-slab in-process run with both halo paths."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from common import blob_ic_3d, charged_settings, gubser_ic, viscous_settings  # noqa: E402
from ccake_b200 import SPHWorkstation  # noqa: E402

D3 = dict(dim=3, hT=0.5, hEta=0.5, stepx=0.25, stepy=0.25, stepEta=0.25, xmin=-1.375, ymin=-1.375, etamin=-1.375)


def run(settings, ic, steps=2):
    ws = SPHWorkstation(settings)
    ws.upload_initial_conditions(ic["r"], ic["e"], ic.get("rhoB"), ic.get("rhoS"), ic.get("rhoQ"), ic["u"], ic.get("bulk"), ic.get("shv6"))
    for _ in range(steps):
        ws.advance_timestep(settings["dt"])
    ws.synchronize()
    assert np.all(np.isfinite(ws.get("u")))
    ws.close()


run(charged_settings(dt=0.02, **D3), blob_ic_3d(n_side=8))                                      # diffusion currents non-zero (QNZ twin)
run(charged_settings(dt=0.02, kappa_00=0.0, kappa_11=0.0, kappa_22=0.0, kappa_01=0.0, kappa_10=0.0, **D3), blob_ic_3d(n_side=8))   # zero currents
run(viscous_settings(dt=0.02, rk_order=4), gubser_ic(dx=0.2, shear=True))
run(viscous_settings(dt=0.05, particlization_enabled=True, Freeze_Out_Temperature=300.0), gubser_ic(dx=0.2, shear=True), steps=4)
from test_gpu_parity import _table  # noqa: E402

run(charged_settings(dt=0.02, **_table()), gubser_ic(dx=0.2, shear=True, charges=True))
from slab_harness import SlabRun  # noqa: E402

for p2p in ("1", "0"):
    os.environ["CCB_HALO_P2P"] = p2p
    sr = SlabRun(charged_settings(dt=0.02), gubser_ic(dx=0.2, shear=True, charges=True), 2)
    sr.each(lambda r, w: [w.advance_timestep(0.02) for _ in range(2)])
    sr.each(lambda r, w: w.synchronize())
print("sanitize_case: done")
