"""Diagnostic (GPU): phase-by-phase comparison of the mixed table / fall-back population against oracle/_ref."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import *
from ccake_b200 import eos_table, SPHWorkstation
td = eos_table.ensure_table(tempfile.mkdtemp(prefix="ccb_eos_"), coarse=True)
settings = viscous_settings(eos_type="table", eos_path=td, normalize_by_T=True, Freeze_Out_Temperature=150.0)
ic = gubser_ic(dx=0.1, shear=True)
ic["e"] = ic["e"] * np.where(np.hypot(ic["r"][:, 0], ic["r"][:, 1]) > 1.6, 2e-6, 1.0)
fd = os.dup(1); os.dup2(2, 1)
sim = make_ref(settings, ic); sim.initialize()
ws = SPHWorkstation(settings)
ws.upload_initial_conditions(ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"], ic["u"], ic["bulk"], ic["shv6"])
out = os.fdopen(fd, "w")
def rep(tag, a, b, eos):
    a = np.asarray(a).reshape(len(eos), -1); b = np.asarray(b).reshape(len(eos), -1)
    d = np.abs(a - b).max(1); i = int(np.argmax(d))
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    print("%-22s max abs %.3e (scale %.3e) at %d eos %d ref %s | max pointwise rel %.3e" % (tag, d.max(), np.abs(b).max(), i, eos[i], b[i][:3], rel.max()), file=out)
def compare(tag):
    eos = sim.eos_index()
    print("---", tag, "eos equal", np.array_equal(ws.get_int("eos"), eos), file=out)
    rep("extensive", ws.get("extensive"), sim.member("extensive")[:, [0, 2, 3, 4]], eos)
    rep("u", ws.get("u"), sim.get("hydro_vector", "u"), eos)
    rep("sigma_lab", ws.get("sigma_lab"), sim.get("hydro_scalar", "sigma_lab"), eos)
    rep("smoothed", ws.get("smoothed"), sim.member("smoothed")[:, [0, 2, 3, 4]], eos)
    th, thr = ws.get("thermo"), ref_thermo16(sim)
    for k in (0, 4, 5, 9, 10, 11, 12):
        rep("thermo%d" % k, th[:, k], thr[:, k], eos)
    rep("gradP", ws.get("gradP"), sim.get("hydro_vector", "gradP"), eos)
    rep("gradV", ws.get("gradV"), sim.get("hydro_space_matrix", "gradV").reshape(len(eos), 9), eos)
    rep("du_dt", ws.get("du_dt"), sim.get("hydro_vector", "du_dt"), eos)
    rep("d_ext", ws.get("d_extensive_dt"), sim.member("d_dt_extensive")[:, [0, 2, 3, 4]], eos)
compare("init")
sim.get_time_derivatives(); ws.get_time_derivatives(); ws.synchronize()
compare("after derivative evaluation")
sim.advance(settings["dt"], 2); ws.advance_timestep(settings["dt"]); ws.synchronize()
compare("after 1 step")
out.flush()
