"""CPU checks of the tabulated EoS and its fall-back chain (ccake_b200/csrc/physics_eos_chain.cuh +
eos_table_host.h, compiled for the host by tests/host_harness.cpp) against the reference build oracle/_ref:
table interpolation, the matched tanh_conformal / conformal / conformal_diagonal members, the
(s,B,S,Q) -> (T,mu) location of every particle of a charged run (phase 04), the initial e -> s location, and the
hand-over to the fall-back EoS for particles the table does not cover.

Tolerances: forward evaluations 1e-12 (same arithmetic); located points: 1e-10 on the field's scale, i.e. the
reference's own root-finder tolerance (it stores a GSL hybrids iterate with sum|f| < 1e-10, include/rootfinder.h:37,
where the library runs Newton's method to the same criterion)."""
import ctypes as C
import os

import numpy as np
import pytest

from common import charged_settings, dp, gubser_ic, harness, have_ref, make_ref, rel_err

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
HC = 197.3269804
NAMES = ["T", "muB", "muS", "muQ", "e", "s", "rhoB", "rhoS", "rhoQ", "p", "cs2", "w", "dwds", "dwdB", "dwdS", "dwdQ"]


@pytest.fixture(scope="module")
def table_dir(tmp_path_factory):
    from ccake_b200 import eos_table

    return eos_table.ensure_table(str(tmp_path_factory.mktemp("eos")), coarse=True)


def _settings(table_dir, **kw):
    s = charged_settings(eos_type="table", eos_path=table_dir, normalize_by_T=True, Freeze_Out_Temperature=150.0)
    s.update(kw)
    return s


@pytest.fixture(scope="module")
def chain(table_dir):
    H = harness()
    H.hh_chain_create.restype = C.c_void_p
    H.hh_chain_create.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
    H.hh_chain_solve.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                 C.POINTER(C.c_int32)] + [C.POINTER(C.c_double)] * 3
    H.hh_chain_table_thermo.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    H.hh_chain_member_thermo.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    H.hh_chain_parameters.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    h = H.hh_chain_create(os.path.join(table_dir, "thermo.dat").encode(), 1, 150.0, 1, 1, 0)
    assert h
    yield H, C.c_void_p(h)
    H.hh_chain_destroy.argtypes = [C.c_void_p]
    H.hh_chain_destroy(C.c_void_p(h))


@pytest.fixture(scope="module")
def sim(table_dir):
    s = _settings(table_dir)
    ic = gubser_ic(dx=0.1, xmax=2.0, shear=True, charges=True)
    sim = make_ref(s, ic)
    sim.initialize()
    return s, ic, sim


def _solve(chain, energy, targets, seed, eos):
    H, h = chain
    n = targets.shape[0]
    thermo, dalpha, chi = np.zeros((n, 16)), np.zeros((n, 12)), np.zeros((n, 2))
    eos = np.ascontiguousarray(eos, dtype=np.int32)
    fails = H.hh_chain_solve(h, energy, n, dp(np.ascontiguousarray(targets)), dp(np.ascontiguousarray(seed)),
                             eos.ctypes.data_as(C.POINTER(C.c_int32)), dp(thermo), dp(dalpha), dp(chi))
    return fails, eos, thermo, dalpha, chi


def _points(rng, n, tlo=45.0, thi=630.0, mu=440.0):
    return np.stack([rng.uniform(tlo, thi, n), rng.uniform(-mu, mu, n), rng.uniform(-mu, mu, n), rng.uniform(-mu, mu, n)], 1) / HC


def test_table_interpolation_matches_reference(chain, sim):
    """4-D multilinear interpolation of all 17 fields, incl. points on nodes, on the upper faces and in the last cells."""
    H, h = chain
    _, _, ref = sim
    rng = np.random.default_rng(5)
    pts = _points(rng, 400)
    T_axis, mu_axis = np.arange(40.0, 641.0, 20.0) / HC, np.arange(-450.0, 451.0, 150.0) / HC
    special = np.array([[T_axis[3], mu_axis[2], mu_axis[4], mu_axis[1]], [T_axis[-1], mu_axis[-1], mu_axis[-1], mu_axis[-1]],
                        [T_axis[0], mu_axis[0], mu_axis[0], mu_axis[0]], [T_axis[-1] * (1 - 1e-9), 0.0, mu_axis[-1], 0.1],
                        [2.0, 0.0, 0.0, 0.0]])
    pts = np.concatenate([pts, special])
    got = np.zeros((pts.shape[0], 17))
    H.hh_chain_table_thermo(h, pts.shape[0], dp(np.ascontiguousarray(pts)), dp(got))
    want = ref.eos_thermo("table", pts)
    for k in range(17):
        assert rel_err(got[:, k], want[:, k]) < 1e-13, k


@pytest.mark.parametrize("member,name", [(1, "tanh_conformal"), (2, "conformal"), (3, "conformal_diagonal")])
def test_fallback_members_match_reference(chain, sim, member, name):
    """The three analytic EoSs with the parameters the library matched to the table (set_up_chosen_EOSs) reproduce
    the reference's get_full_thermo, including the tanh EoS's written-out speed of sound."""
    H, h = chain
    _, _, ref = sim
    pts = _points(np.random.default_rng(11 + member), 300, tlo=60.0, thi=700.0, mu=300.0)
    got = np.zeros((pts.shape[0], 17))
    H.hh_chain_member_thermo(h, member, pts.shape[0], dp(np.ascontiguousarray(pts)), dp(got))
    want = ref.eos_thermo(name, pts)
    for k in range(17):
        if name == "tanh_conformal" and k == 6:
            # cs2 = n.H^-1.n / w: the reference evaluates the cofactor expansion (~2000 flop with cancellations)
            assert np.max(np.abs(got[:, k] - want[:, k])) < 1e-9, k
        else:
            assert rel_err(got[:, k], want[:, k]) < 1e-12, (name, k)


def test_locate_sBSQ_matches_reference_phase04(chain, sim):
    """Every particle of a charged viscous run located from its previous point, as phase 04 does."""
    s, _, ref = sim
    ref.advance(s["dt"], 2)
    for ph in ("neighbour_list", "gamma_velocity", "smooth_fields"):
        ref.phase(ph)
    before = ref.member("thermo").copy()
    start = ref.eos_index()
    ref.phase("thermodynamics")
    after = ref.member("thermo")
    fails, eos, thermo, dalpha, chi = _solve(chain, 0, before[:, 5:9], before[:, 0:4], start)
    assert fails == 0
    assert np.array_equal(eos, ref.eos_index())
    for k, nm in enumerate(NAMES):
        assert rel_err(thermo[:, k], after[:, k]) < (1e-9 if nm.startswith(("mu", "rho", "dwd")) and nm != "dwds" else 1e-10), nm
    # densities reproduced to the root-finder's absolute tolerance
    assert np.max(np.abs(thermo[:, 5:9] - before[:, 5:9])) < 1e-10
    assert rel_err(dalpha, after[:, 16:28]) < 1e-9
    assert rel_err(chi, after[:, 28:30]) < 1e-10


def test_initial_location_from_energy_density(chain, sim):
    """e -> (T, mu) -> s at initialisation (EquationOfState::rootfinder_s_out) against the reference's particles."""
    s, ic, _ = sim
    ref = make_ref(s, ic)
    ref.initialize()
    inp = ref.member("input")  # s, e (host only), rhoB, rhoS, rhoQ
    assert inp.shape[0] == ic["e"].shape[0]
    e_fm = ic["e"] / 0.1973269804  # BSQHydro::read_ccake, src/BSQHydro.cpp:317
    targets = np.stack([e_fm, inp[:, 2], inp[:, 3], inp[:, 4]], 1)
    fails, eos, thermo, _, _ = _solve(chain, 1, targets, np.zeros_like(targets), np.zeros(targets.shape[0]))
    assert fails == 0 and np.all(eos == 0)
    assert rel_err(thermo[:, 5], inp[:, 0]) < 1e-10          # the entropy density the reference stored as input.s
    assert np.max(np.abs(thermo[:, 4] - e_fm)) < 1e-10       # e reproduced


def test_particles_outside_the_table_fall_back(chain, table_dir):
    """Dilute particles whose temperature lies below the table's first node are handed to tanh_conformal by the
    reference and by the library alike, and stay there (skip_failed_EoS)."""
    s = _settings(table_dir, baryon_charge_enabled=False, strange_charge_enabled=False, electric_charge_enabled=False,
                  diffusionMode="disabled")
    ic = gubser_ic(dx=0.1, xmax=2.0, shear=True)
    ic["e"] = ic["e"] * np.where(np.hypot(ic["r"][:, 0], ic["r"][:, 1]) > 1.6, 2e-6, 1.0)  # cold corona: T ~ 10-35 MeV
    ref = make_ref(s, ic)
    ref.initialize()
    inp = ref.member("input")
    ref_eos = ref.eos_index()
    assert (ref_eos == 1).sum() > 50 and (ref_eos == 0).sum() > 50
    targets = np.stack([ic["e"] / 0.1973269804, inp[:, 2], inp[:, 3], inp[:, 4]], 1)
    fails, eos, thermo, _, _ = _solve(chain, 1, targets, np.zeros_like(targets), np.zeros(targets.shape[0]))
    assert fails == 0
    assert np.array_equal(eos, ref_eos)
    # the entropy density each EoS assigned (stored by the reference as input.s). The reference stops at
    # |e - e_target| < 1e-10 fm^-4 (absolute), i.e. |ds| = |de| / T < 1e-10 / T: that bound is the tolerance.
    assert np.max(np.abs(thermo[:, 4] - targets[:, 0])) < 1e-10
    assert np.all(np.abs(thermo[:, 5] - inp[:, 0]) * thermo[:, 0] < 2e-10)
    assert rel_err(thermo[ref_eos == 0, 5], inp[ref_eos == 0, 0]) < 1e-10
