"""The synthetic workloads of BASELINE.json's configs (ccake_b200/workloads.py) are deterministic and have the advertised shape:
the bench and the parity tests on both sides (CUDA library, reference build) see the same particles."""
import numpy as np

from common import ROOT  # noqa: F401  (puts the repository root on sys.path)
from ccake_b200 import eos_table as E
from ccake_b200 import workloads as W


def test_config_sizes_and_neighbour_counts():
    s1, ic1 = W.config1_ideal_gubser()
    assert ic1["r"].shape == (160801, 2) and s1["hT"] == 0.4 and ic1["rhoB"] is None
    s2, ic2 = W.config2_viscous_gubser_bsq()
    assert ic2["r"].shape == (160801, 2) and s2["hT"] == 0.1 and s2["diffusionMode"] == "constant_over_T2" and ic2["shv6"].shape == (160801, 6)
    assert np.pi * (2 * s2["hT"]) ** 2 / s2["stepx"] ** 2 > 190      # ~ 201 lattice points inside the 2h disc
    s3, ic3 = W.config3_trento_pbpb(n_target=20000)
    assert ic3["r"].shape == (19978, 2) and s3["eos_type"] == "table" and ic3["e"].min() > 0.15 and abs(ic3["e"].max() - 90.0) < 1e-9
    s4, ic4 = W.config4_milne_bsq(n_target=25000)
    assert ic4["r"].shape == (25040, 3) and s4["dim"] == 3 and abs(s4["hT"] - 2 * s4["stepx"]) < 1e-15
    assert np.all(ic4["rhoB"] > 0) and np.any(ic4["rhoS"] < 0) and np.any(ic4["rhoS"] > 0) and np.allclose(ic4["rhoQ"], 0.4 * ic4["rhoB"])
    s5, ic5 = W.config5_cartesian_blob(n_target=100000)
    assert s5["coordinate_system"] == "cartesian" and abs(ic5["r"].shape[0] - 100000) < 2000 and abs(s5["hT"] - 2 * s5["stepx"]) < 1e-15


def test_workloads_are_deterministic():
    a, b = W.config3_trento_pbpb(n_target=5000), W.config3_trento_pbpb(n_target=5000)
    assert np.array_equal(a[1]["r"], b[1]["r"]) and np.array_equal(a[1]["e"], b[1]["e"])
    a, b = W.config2_viscous_gubser_bsq(dx=0.1, xmax=2.0), W.config2_viscous_gubser_bsq(dx=0.1, xmax=2.0)
    assert np.array_equal(a[1]["shv6"], b[1]["shv6"])


def test_tile_along_x_keeps_the_lattice_uniform():
    s, ic = W.config2_viscous_gubser_bsq(dx=0.1, xmax=1.0)
    s4, ic4 = W.tile_along_x(s, ic, 4, s["stepx"])
    assert ic4["r"].shape[0] == 4 * ic["r"].shape[0]
    xs = np.unique(np.round(ic4["r"][:, 0], 9))
    assert np.allclose(np.diff(xs), 0.1) and abs(xs[0] + xs[-1]) < 1e-9 and s4["xmin"] == ic4["r"][:, 0].min()


def test_synthetic_table_is_thermodynamically_consistent():
    """The synthetic thermo.dat table is built from one analytic pressure: s, n_X are its gradient and e = T s + mu n - p."""
    T_axis, mu_axis = E.default_axes(coarse=True)
    t = E.synthetic_table(T_axis, mu_axis, normalize_by_T=False)
    assert t["sizes"] == (31, 7, 7, 7) and t["fields"].shape == (31 * 343, 17)
    T, B, Q, S = t["grid"].T
    p, s, nB, nS, nQ, e = (t["fields"][:, k] for k in range(6))
    assert np.allclose(e, T * s + B * nB + Q * nQ + S * nS - p, rtol=1e-12)
    h = 1e-3
    p1, _, _ = E.pressure_and_derivatives(T + h, B, Q, S)
    p0, _, _ = E.pressure_and_derivatives(T - h, B, Q, S)
    assert np.allclose((p1 - p0) / (2 * h), s, rtol=1e-6)
    # 0 < cs2 < 1 where runs live (the corners T ~ 40 MeV, |mu| ~ 450 MeV of the model pressure are unphysical: there the
    # reference's prohibit_unstable_cs2 / prohibit_acausal_cs2 switches hand the particle to the fall-back EoS)
    core = (T >= 100.0) & (np.abs(B) <= 150.0) & (np.abs(Q) <= 150.0) & (np.abs(S) <= 150.0)
    assert np.all(t["fields"][core, 6] > 0) and np.all(t["fields"][core, 6] < 1.0)


def test_parallel_ic_reader_matches_the_line_by_line_reader(tmp_path):
    """ccb_ic_read (csrc/ic_reader.cpp, host code) against the per-line Python twin of BSQHydro::read_ccake: header, energy cut-off,
    disabled charges, the twelve diffusion columns, blank and comment lines; and read_ICCING's six-column format."""
    rng = np.random.default_rng(5)
    s, ic = W.config2_viscous_gubser_bsq(dx=0.1, xmax=2.0)
    n = ic["r"].shape[0]
    ic["diffusion"] = rng.standard_normal((n, 3, 4))
    ic["bulk"] = rng.standard_normal(n) * 0.01
    s = dict(s, e_cutoff=float(np.median(ic["e"])), input_initial_diffusion=True, strange_charge_enabled=False)
    path = tmp_path / "ic.dat"
    W.write_ccake(str(path), s, ic)
    txt = open(path).read().splitlines()
    txt.insert(5, "")
    txt.insert(9, "# a comment between rows")
    open(path, "w").write("\n".join(txt))   # no trailing newline
    s_py, ic_py = W.read_ccake_python(str(path), s)
    s_c, ic_c = W.read_ccake(str(path), s)
    assert 0 < ic_py["r"].shape[0] < n, "the cut-off must drop some rows"
    for k in ("stepx", "stepy", "stepEta", "xmin", "ymin", "etamin"):
        assert s_py[k] == s_c[k]
    for k in ("r", "e", "rhoB", "rhoQ", "u", "bulk", "shv6"):
        assert np.array_equal(ic_py[k], ic_c[k]), k
    assert np.array_equal(ic_py["diffusion"][:, [0, 2]], ic_c["diffusion"][:, [0, 2]])
    assert np.all(ic_c["rhoS"] == 0.0) and np.all(ic_c["diffusion"][:, 1, :] == 0.0)   # disabled charge (src/BSQHydro.cpp:340-372)
    # entropy input: the cut-off acts on the ideal-gas energy estimate (src/BSQHydro.cpp:383-398)
    s_e = dict(s, input_as_entropy=True, e_cutoff=0.12 * float(np.median(ic["e"])) ** (4.0 / 3.0), input_initial_diffusion=False)
    a, b = W.read_ccake_python(str(path), s_e)[1], W.read_ccake(str(path), s_e)[1]
    assert 0 < a["e"].size < n and np.array_equal(a["e"], b["e"]) and np.array_equal(a["r"], b["r"])
    # ICCING: header "<ignored> stepx stepy <ignored> xmin ymin", rows x y e rhoB rhoS rhoQ (src/BSQHydro.cpp:84-145)
    p2 = tmp_path / "iccing.dat"
    rows = np.column_stack([ic["r"], ic["e"], ic["rhoB"], ic["rhoS"], ic["rhoQ"]])
    with open(p2, "w") as f:
        f.write("0 0.1 0.1 0 -2.0 -2.0\n")
        np.savetxt(f, rows, fmt="%.17g")
    s_i, ic_i = W.read_ccake(str(p2), dict(s, e_cutoff=0.0, strange_charge_enabled=True), fmt="iccing")
    assert s_i["stepx"] == 0.1 and s_i["xmin"] == -2.0 and np.array_equal(ic_i["r"], ic["r"]) and np.array_equal(ic_i["e"], ic["e"])
    assert np.array_equal(ic_i["rhoS"], ic["rhoS"]) and np.all(ic_i["u"] == 0.0) and np.all(ic_i["shv6"] == 0.0)
