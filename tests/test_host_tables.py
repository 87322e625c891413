"""Host-side set-up logic (table construction, state settings, pair_coeff bookkeeping)
against the oracle — bit-exact, as the tables are a precondition of force parity."""
import numpy as np
import pytest

import oracle_binding as ob


def _orc(style, n, mu=(0.0, 0.5)):
    o = ob.Oracle()
    o.pair_style(style, n)
    o.set_types(1, 2, [0, 2], [[0, 0], [1, 2]], [0.0, mu[0], mu[1]], [0.0, 1.0, 1.0])
    return o


@pytest.mark.parametrize("style,n,cut,tfile", [
    (1, 4096, 2.5, "table4096"),   # LINEAR, match=1 (file grid == table grid)
    (1, 1000, 2.4, "table4096"),   # LINEAR, re-splined
    (2, 1000, 2.4, "table4096"),   # SPLINE
    (0, 1000, 2.5, "table1024"),   # LOOKUP
    (2, 500, 2.3, "tableR"),       # SPLINE from an R-spaced file
    (1, 777, 2.5, "tableR"),
])
def test_table_construction_bit_exact(pkg, fixtures, style, n, cut, tfile):
    from lammps_ucg_dev_b200 import engine
    for key in ("UCG_00", "UCG_01", "UCG_11"):
        o = _orc(style, n)
        idx = o.table_add_file(fixtures[tfile], key, cut)
        ht = engine.HostTable.from_file(fixtures[tfile], key, cut, style, n)
        info, ref = ht.info(), o.table_params(idx)
        for k in ("innersq", "delta", "invdelta", "deltasq6", "cut", "match"):
            assert info[k] == ref[k], k
        for w in ("rsq", "e", "f", "de", "df", "e2", "f2"):
            a, b = ht.array(w), o.table_get(idx, w)
            assert a.shape == b.shape, (w, a.shape, b.shape)
            assert np.array_equal(a, b), w


def test_bitmap_table_bit_exact(pkg, fixtures):
    from lammps_ucg_dev_b200 import engine
    o = _orc(3, 10)
    idx = o.table_add_file(fixtures["table4096"], "UCG_00", 2.5)
    ht = engine.HostTable.from_file(fixtures["table4096"], "UCG_00", 2.5, 3, 10)
    assert ht.info()["nmask"] == o.table_params(idx)["nmask"]
    assert ht.info()["nshiftbits"] == o.table_params(idx)["nshiftbits"]
    for w in ("rsq", "e", "f", "de", "df", "drsq"):
        assert np.array_equal(ht.array(w), o.table_get(idx, w)), w


def test_table_errors_match_reference_texts(pkg, fixtures):
    from lammps_ucg_dev_b200 import engine
    with pytest.raises(pkg.UCGError, match="Pair table cutoff outside of table"):
        engine.HostTable.from_file(fixtures["table4096"], "UCG_00", 2.6, 1, 100)
    with pytest.raises(pkg.UCGError, match="Did not find keyword"):
        engine.HostTable.from_file(fixtures["table4096"], "NOPE", 2.5, 1, 100)
    with pytest.raises(pkg.UCGError, match="Illegal number of pair table entries"):
        engine.HostTable.from_file(fixtures["table4096"], "UCG_00", 2.5, 1, 1)


def test_single_matches_linear_rule(pkg, fixtures):
    from lammps_ucg_dev_b200 import engine
    import brute
    ht = engine.HostTable.from_file(fixtures["table4096"], "UCG_01", 2.5, 1, 4096)
    tab = dict(ht.info(), e=ht.array("e"), f=ht.array("f"))
    rsq = np.random.default_rng(0).uniform(0.3, 6.2, 200)
    u, f = brute.linear_table(tab, rsq)
    for k, r2 in enumerate(rsq):
        rc, phi, ff = ht.single(float(r2))
        assert rc == 0 and phi == pytest.approx(u[k], rel=1e-14) and ff == pytest.approx(f[k], rel=1e-14)
    assert ht.single(0.2)[0] == 1      # < table inner cutoff
    assert ht.single(6.25)[0] == 2     # > table outer cutoff


def test_statemap_file_and_coeff(pkg, fixtures, tmp_path):
    from lammps_ucg_dev_b200 import engine
    sm = engine.StateMap.from_file(fixtures["state"])
    assert sm.sizes() == (1, 2)
    sm.coeff(1, 1, 1, 1, 2, 2, [0, 1, 2, 3], [2.5, 2.5, 2.5, 2.5])
    sm.init()
    got = sm.get()
    o = _orc(1, 4096)
    idx = [o.table_add_file(fixtures["table4096"], k, 2.5) for k in ("UCG_00", "UCG_01", "UCG_01", "UCG_11")]
    o.pair_coeff(1, 1, 1, 1, 2, 2, idx)
    o.pair_init()
    ti, cs = o.get_pair_maps()
    assert np.array_equal(got["tabindex"], ti)          # incl. init_one's (2,1) <- (1,2) overwrite
    assert np.array_equal(got["cutsq"], cs)
    assert got["tabindex"][2, 1] == got["tabindex"][1, 2] == 1
    assert list(got["chem_pot"]) == [0.0, 0.0, 0.5]
    # grammar errors of read_state_settings
    bad = tmp_path / "bad.conf"
    bad.write_text("1 2 2\n1 3\n")
    with pytest.raises(pkg.UCGError, match="Only 1 or 2 states are allowed"):
        engine.StateMap.from_file(str(bad))
    bad.write_text("2 3 2\n2 1\n")
    with pytest.raises(pkg.UCGError, match="Please write orderly"):
        engine.StateMap.from_file(str(bad))
    sm2 = engine.StateMap.from_file(fixtures["state"])
    with pytest.raises(pkg.UCGError, match="All pair coeffs are not set"):
        sm2.init()
