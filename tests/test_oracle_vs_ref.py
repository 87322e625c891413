"""Live cross-check of the C restatement against oracle/_ref (the reference's own sources
compiled against the LAMMPS-API shim) on inputs other than the golden ones.  Skipped when
oracle/_ref has not been built (it needs /root/reference at build time)."""
import numpy as np
import pytest

import oracle_binding as ob
import ref_binding as rb

pytestmark = pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built")


def _liq(n, **kw):
    from lammps_ucg_dev_b200 import synth
    return synth.fcc_liquid(n, **kw)


def test_half_list_is_identical(pkg, fixtures):
    liq = _liq((4, 5, 7), seed=99)
    r = rb.RefSim.single_type(liq, fixtures["table4096"], fixtures["state"])
    r.command("fix 0 all ttarget/stub 1.0")
    r.compute_once(0)
    o = ob.Oracle.single_type(liq, fixtures["table4096"])
    o.neigh_build_all()
    ri, rj = r.neigh_pairs()
    oi, oj = o.neigh_pairs()
    assert np.array_equal(ri, oi) and np.array_equal(rj, oj)       # same pairs in the same order
    assert r.nghost() == o.nghost()


@pytest.mark.parametrize("tabstyle,code,n,cut", [("linear", 1, 4096, 2.5), ("spline", 2, 700, 2.4), ("lookup", 0, 3000, 2.5),
                                                 ("bitmap", 3, 12, 2.5)])
def test_ucgld_all_table_styles(pkg, fixtures, tabstyle, code, n, cut):
    liq = _liq(6, seed=7)
    r = rb.RefSim.single_type(liq, fixtures["table4096"], fixtures["state"], tabstyle=tabstyle, tablength=n, cut=cut)
    r.command("fix 0 all ttarget/stub 0.8")
    r.compute_once(1)
    a = r.get_atoms()
    o = ob.Oracle.single_type(liq, fixtures["table4096"], tabstyle=code, tablength=n, cut=cut, kT=0.8)
    o.neigh_build_all(); o.force_clear(); o.pair_ucgld(1, 1); o.reverse_comm()
    b = o.get_atoms()
    for k in ("f", "ucgforce", "ucgsoftmaxscores"):
        assert np.array_equal(a[k], b[k]), k
    assert r.eng_vdwl() == o.eng_vdwl()
    assert np.array_equal(r.virial()[1], o.virial())


def test_langevin_wall_trajectory(pkg, fixtures):
    liq = _liq(6, seed=3)
    r = rb.RefSim.single_type(liq, fixtures["table4096"], fixtures["state"])
    for c in ("fix 1 all nve/ucgld/wall/hard", "fix 2 all ucgld/langevin 1.0 0.5 0.2 777", "fix 3 all ucgstate ld"):
        r.command(c)
    r.setup(1); r.run(60, 0)
    o = ob.Oracle.single_type(liq, fixtures["table4096"])
    o.fix_nve_wall(1, 0, 0.1); o.fix_langevin(1.0, 0.5, 0.2, 777); o.fix_ucgstate(mode=1)
    o.setup(); o.run(60)
    a, b = r.get_atoms(), o.get_atoms()
    for k in ("x", "v", "ucgl", "ucgvl", "ucgp", "ucgstate", "ucgforce"):
        assert np.array_equal(a[k], b[k]), k
    assert r.fix_scalar(1) == o.lambda_temp()
    assert r.nbuilds() == o.nbuilds()


def test_reference_quirks_are_what_the_survey_says(pkg, fixtures, tmp_path):
    liq = _liq(4)
    # Q3: virial silently zero with newton on
    r = rb.RefSim.single_type(liq, fixtures["table1024"], fixtures["state"], tablength=1024)
    r.command("fix 0 all ttarget/stub 1.0")
    r.compute_once(1)
    shipped, tally = r.virial()
    assert not shipped.any() and tally.any()
    # fix ucgstate refuses to start without a t_target provider before it
    r2 = rb.RefSim.single_type(liq, fixtures["table1024"], fixtures["state"], tablength=1024)
    r2.command("fix 1 all nve/ucgld")
    r2.command("fix 2 all ucgstate")
    with pytest.raises(RuntimeError, match="requires a thermostat fix BEFORE ITSELF"):
        r2.setup(0)
    # Q23: a one-state type cannot be given pair coefficients as shipped
    sf = tmp_path / "mixed.conf"
    sf.write_text("2 3 2\n1 1\n2 2\n2 3\n0.0 0.3\n")
    r3 = rb.RefSim()
    r3.box(liq.box_lo, liq.box_hi, 3)
    r3.atoms(liq)
    r3.command(f"pair_style table_ucgld linear 1024 {sf}")
    t = fixtures["table1024"]
    with pytest.raises(RuntimeError, match="Formal type not defined"):
        r3.command(f"pair_coeff 1 1 1 1 {t} UCG_00 2.5")
