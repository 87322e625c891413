"""Checks of the CPU oracle itself (it is the checker for everything else)."""
import numpy as np
import pytest

import brute
import oracle_binding as ob


def _liq(pkg, n, **kw):
    from lammps_ucg_dev_b200 import synth
    return synth.fcc_liquid(n, **kw)


def test_ranpark_known_answers():
    # Park-Miller minimal standard: x1 = 16807*seed mod (2^31-1); 10000th value from seed 1 is 1043618065
    u = ob.ranpark(1, 10000)
    assert u[0] == pytest.approx(16807 / 2147483647.0, rel=0, abs=1e-18)
    assert round(u[-1] * 2147483647) == 1043618065


def test_ranmars_is_uniform_and_reproducible():
    a, b = ob.ranmars(48291, 200000), ob.ranmars(48291, 200000)
    assert np.array_equal(a, b)
    assert a.min() >= 0.0 and a.max() < 1.0
    assert abs(a.mean() - 0.5) < 4e-3 and abs(a.var() - 1 / 12) < 2e-3
    assert not np.array_equal(a[:100], ob.ranmars(48292, 100))
    # 24-bit lattice of Marsaglia's generator
    assert np.all(np.abs(a * 2 ** 24 - np.round(a * 2 ** 24)) < 1e-9)


def test_pair_ucgld_matches_bruteforce(pkg, fixtures):
    liq = _liq(pkg, 5)  # 500 sites, box 8.4 > 2*(2.5+0.3)
    o = ob.Oracle.single_type(liq, fixtures["table4096"])
    o.neigh_build_all()
    o.force_clear()
    o.pair_ucgld(1, 1)
    o.reverse_comm()
    a = o.get_atoms()
    tabs = {}
    for (s, t), idx in {(0, 0): 0, (0, 1): 1, (1, 0): 1, (1, 1): 3}.items():
        tabs[s, t] = dict(o.table_params(idx), e=o.table_get(idx, "e"), f=o.table_get(idx, "f"))
    f, uf, sc, E, vir = brute.ucgld_bruteforce(a["x"], liq.box_hi - liq.box_lo, a["ucgl"], a["ucgstate"], tabs,
                                               2.5 ** 2, (0.0, 0.5), 1.0)
    assert np.abs(a["f"] - f).max() < 1e-10 * np.abs(f).max()
    assert np.abs(a["ucgforce"] - uf).max() < 1e-10 * np.abs(uf).max()
    assert np.abs(a["ucgsoftmaxscores"] - sc).max() < 1e-10 * np.abs(sc).max()
    assert o.eng_vdwl() == pytest.approx(E, rel=1e-12)
    assert np.allclose(o.virial(), vir, rtol=1e-10, atol=1e-9)
    assert np.all(a["num_ucgstates"] == 2)


def test_half_and_full_lists_hold_the_same_pairs(pkg, fixtures):
    liq = _liq(pkg, 6)
    oh = ob.Oracle.single_type(liq, fixtures["table1024"], tablength=1024, full=0)
    of = ob.Oracle.single_type(liq, fixtures["table1024"], tablength=1024, full=1)
    oh.neigh_build_all()
    of.neigh_build_all()
    hi, hj = oh.neigh_pairs()
    fi, fj = of.neigh_pairs()
    assert 2 * hi.size == fi.size
    n = liq.n + 1
    half = np.sort(np.minimum(hi, hj).astype(np.int64) * n + np.maximum(hi, hj))
    full = np.sort(fi.astype(np.int64) * n + fj)
    sym = np.sort(np.concatenate([hi.astype(np.int64) * n + hj, hj.astype(np.int64) * n + hi]))
    assert np.array_equal(full, sym)
    assert np.unique(half).size == half.size


def test_newton_third_law_and_energy_conservation(pkg, fixtures):
    liq = _liq(pkg, 6)
    o = ob.Oracle.single_type(liq, fixtures["table4096"], dt=0.001)
    o.fix_ttarget(1.0)
    o.fix_nve()
    o.fix_ucgstate(mode=1)  # ld: probabilities only -> Hamiltonian dynamics in (x, lambda)
    o.setup()
    a = o.get_atoms()
    assert np.abs(a["f"].sum(0)).max() < 1e-9

    def etot():
        b = o.get_atoms()
        return o.eng_vdwl() + 0.5 * (b["v"] ** 2).sum() + 0.5 * (liq.ucgml * b["ucgvl"] ** 2).sum() \
            + 0.5 * b["ucgl"].sum()  # + mu1*lambda (mu0 = 0): the chemical-potential term of ucgforce

    e0 = etot()
    o.run(100, thermo_every=100)
    e1 = etot()
    assert o.nbuilds() >= 1
    assert abs(e1 - e0) < 2e-4 * abs(e0)


def test_wall_and_ucgstate_rules(pkg, fixtures):
    liq = _liq(pkg, 4)
    o = ob.Oracle.single_type(liq, fixtures["table1024"], tablength=1024)
    n = liq.n
    rng = np.random.default_rng(5)
    sc = rng.normal(0, 3, (n, 2))
    sc[0] = (800.0, 900.0)       # exp clamp at 700 -> p = 0.5 -> round half away from zero -> 1
    sc[1] = (50.0, -50.0)        # p clamps to 1e-6
    o.neigh_build_all(); o.force_clear(); o.pair_ucgld(0, 0)  # sets num_ucgstates
    o.set_forces(scores=sc)
    o.ucgstate_post_force(mode=0)
    a = o.get_atoms()
    p = np.clip(np.exp(np.minimum(sc[:, 1], 700)) / (np.exp(np.minimum(sc[:, 0], 700)) + np.exp(np.minimum(sc[:, 1], 700))), 1e-6, 1 - 1e-6)
    assert np.allclose(a["ucgp"], p, rtol=1e-15)
    assert a["ucgstate"][0] == 1 and a["ucgp"][1] == 1e-6
    assert np.array_equal(a["ucgstate"], np.floor(p + 0.5).astype(np.int32))
    assert np.array_equal(a["ucgl"], a["ucgp"])
