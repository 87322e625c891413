"""The C restatement (oracle/ucg_oracle.c) and the product's host-side table code against the
golden vectors produced by the reference's own compiled sources (tests/golden/make_golden.py ->
ucg_ref_golden.npz).  Bit-exact: both are serial IEEE evaluations of the same expressions,
including the sequential RanMars streams of fix ucgld/langevin and fix ucgstate mc."""
import os

import numpy as np
import pytest

import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ucg_ref_golden.npz"))
NCELL, TABLEN, NSTEPS = 5, 1024, 25


def _liq():
    from lammps_ucg_dev_b200 import synth
    return synth.fcc_liquid(NCELL)


@pytest.mark.parametrize("name,method,pseudo,prior", [
    ("ucgld", None, None, None),
    ("bethe", 1, 0, 2),        # method bethe, pseudo yes (flag 0), prior ucgl
    ("bethe_sce", 1, 1, 2),    # pseudo no -> SCE conditionals
    ("bethe_mf", 0, 0, 0),     # mean field, prior from chemical potentials
])
def test_single_evaluation_matches_reference(pkg, fixtures, name, method, pseudo, prior):
    o = ob.Oracle.single_type(_liq(), fixtures["table1024"], tablength=TABLEN)
    o.neigh_build_all()
    o.force_clear()
    if method is None:
        o.pair_ucgld(1, 1)
    else:
        o.pair_bethe_config(method, pseudo, prior)
        o.pair_bethe(1, 1)
    o.reverse_comm()
    a = o.get_atoms()
    for k in ("f", "ucgforce", "ucgsoftmaxscores", "num_ucgstates"):
        assert np.array_equal(a[k], G[f"{name}_once_{k}"]), (name, k)
    assert o.eng_vdwl() == float(G[f"{name}_once_E"])
    assert np.array_equal(o.virial(), G[f"{name}_once_virial_tally"])
    assert not G[f"{name}_once_virial_shipped"].any()       # zero as shipped
    assert o.neigh_pairs()[0].size == int(G[f"{name}_npairs"])


def _deck(o, name):
    if name == "traj_c1":
        o.fix_ttarget(1.0); o.fix_nve(); o.fix_ucgstate(mode=0)
    elif name == "traj_wall":
        o.fix_ttarget(1.0); o.fix_nve_wall(1, 1, 0.1); o.fix_ucgstate(mode=1)
    elif name == "traj_langevin":
        o.fix_nve(); o.fix_langevin(1.0, 1.0, 0.5, 4711); o.fix_ucgstate(mode=1)
    elif name == "traj_mc":
        o.fix_ttarget(1.0); o.fix_nve(); o.fix_ucgstate(mode=2, seed=991, rate=0.3)


@pytest.mark.parametrize("name", ["traj_c1", "traj_wall", "traj_langevin", "traj_mc"])
def test_trajectories_match_reference(pkg, fixtures, name):
    o = ob.Oracle.single_type(_liq(), fixtures["table1024"], tablength=TABLEN)
    _deck(o, name)   # kT = boltz * t_target of the first provider = 1.0 in every deck
    o.setup(1, 1)
    o.run(NSTEPS, NSTEPS)
    a = o.get_atoms()
    for k in ("x", "v", "f", "ucgl", "ucgvl", "ucgp", "ucgstate", "ucgforce", "ucgsoftmaxscores"):
        assert np.array_equal(a[k], G[f"{name}_{k}"]), (name, k)
    assert o.eng_vdwl() == float(G[f"{name}_E"])
    assert o.nbuilds() == int(G[f"{name}_nbuilds"])
    if name == "traj_langevin":
        assert o.lambda_temp() == float(G["traj_langevin_lambda_temp"])


@pytest.mark.parametrize("style,code,n", [("lookup", 0, 900), ("linear", 1, 1000), ("spline", 2, 800), ("bitmap", 3, 10)])
def test_host_tables_against_reference_pair_single(pkg, fixtures, style, code, n):
    """product host code (ucgb200_host_table_*) vs Pair::single of the reference"""
    from lammps_ucg_dev_b200 import engine
    rsq = G["single_rsq"]
    ref = G[f"single_{style}"]                 # [(1,1),(1,2),(2,2)] x rsq x (phi, fforce)
    for row, key in enumerate(("UCG_00", "UCG_01", "UCG_11")):
        ht = engine.HostTable.from_file(fixtures["table1024"], key, 2.5, code, n)
        for k, r2 in enumerate(rsq):
            rc, phi, ff = ht.single(float(r2))
            assert rc == 0
            assert phi == ref[row, k, 0] and ff == ref[row, k, 1], (style, key, r2)
