#!/usr/bin/env python
"""Packs the reference's semi-analytic viscous-Gubser check (misc/Gubser_checks/ac: the tau = 1 fm initial profile on the 201 x 201,
dx = 0.05 fm grid and the y = 0 / y = x cuts at tau = 1.2, 1.5, 2.0 fm; columns x y T[GeV] u^x u^y pi^xx pi^yy pi^xy tau^2 pi^etaeta
[GeV/fm^3], README.txt there) into tests/golden/viscous_gubser_ac.npz. The data are the reference's physics fixtures, not code; this
script runs in the build container only (/root/reference does not exist on the GPU box)."""
import os
import sys

import numpy as np

REF = "/root/reference/misc/Gubser_checks/ac"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out = {"initial": np.loadtxt(os.path.join(REF, "Initial_Profile_tau=1fm.dat")).astype(np.float32)}
    for axis in ("y=0", "y=x"):
        for tau in ("1.2", "1.5", "2.0"):
            out["%s_tau%s" % ("y0" if axis == "y=0" else "yx", tau.replace(".", "p"))] = np.loadtxt(os.path.join(REF, "%s_tau=%s_SemiAnalytic.dat" % (axis, tau)))
    path = os.path.join(ROOT, "tests", "golden", "viscous_gubser_ac.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
