"""pair_style table_ucg_bethe_density on the GPU against the reference's own source compiled with
the documented minimal repair (oracle/repair_bethe_density.py): single
domain, periodic, newton off.  Quirks kept as-is on both sides: Q13 (proximity function in the
back-force), the list length in the entropy term, reactions onto ghost images dropped."""
import numpy as np
import pytest

import ref_binding as rb
from decks import rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built")]

# reference grammar (pair_table_ucg_bethe_density.cpp:827-880); the tokens of the third line are split
# on blanks only, so the entropy keyword needs a trailing blank before the newline
BD_STATE = "1 2 2\n1 2\n1 2 {kind} {ent} \n{dens}{mu0} {mu1}\n"


def _setup(pkg, fixtures, tmp_path, liq, kind="density", ent="entropy", rho_th=12.0, r_th=1.5, mu=(0.0, 0.5), n=4096):
    from lammps_ucg_dev_b200 import engine
    sf = tmp_path / "bd.conf"
    sf.write_text(BD_STATE.format(kind=kind, ent=ent, dens=f"{rho_th} {r_th}\n" if kind == "density" else "",
                                  mu0=mu[0], mu1=mu[1]))
    ucg_sf = tmp_path / "ucg.conf"
    ucg_sf.write_text(f"1 2 2\n1 2\n1 2\n{mu[0]} {mu[1]}\n")
    t = fixtures["table4096"]
    ref = rb.RefSim()
    ref.box(liq.box_lo, liq.box_hi, 2)
    ref.atoms(liq)
    for c in ("newton off", "neighbor 0.3 bin", "timestep 0.002",
              f"pair_style table_ucg_bethe_density linear {n} {sf}",
              f"pair_coeff 1 1 2 2 {t} UCG_00 2.5 {t} UCG_01 2.5 {t} UCG_01 2.5 {t} UCG_11 2.5",
              "fix 0 all ttarget/stub 1.0"):
        ref.command(c)
    ctx = pkg.Context(0)
    engine.setup_single_type(ctx, t, str(ucg_sf), tablength=n, box=(liq.box_lo, liq.box_hi))
    ctx.pair_bethe_density_configure([0, 1 if kind == "density" else 0], [0, 1 if ent == "entropy" else 0],
                                     [0.0, rho_th], [0.0, r_th])
    engine.upload_liquid(ctx, liq)
    return ref, ctx


def _liq(n, **kw):
    from lammps_ucg_dev_b200 import synth
    return synth.fcc_liquid(n, **kw)


@pytest.mark.parametrize("kind,ent,rho_th", [("density", "entropy", 12.0), ("density", "no_entropy", 11.0),
                                             ("chempot", "no_entropy", 12.0)])
def test_bethe_density_single_evaluation(pkg, fixtures, tmp_path, kind, ent, rho_th):
    liq = _liq(7)
    ref, ctx = _setup(pkg, fixtures, tmp_path, liq, kind=kind, ent=ent, rho_th=rho_th)
    ref.compute_once(1)
    a = ref.get_atoms()
    ctx.neigh_build()
    ctx.pair_bethe_density(1, 1)
    b = ctx.atoms_download(["f", "ucgp", "ucgsoftmaxscores"])
    e, vir = ctx.pair_energy_virial()
    p0, cvf = ctx.pair_bethe_density_priors()
    assert 0.02 < p0.mean() < 0.98
    assert rel_err(b["f"], a["f"]) <= 1e-6
    assert np.abs(b["ucgp"] - a["ucgp"]).max() <= 1e-9          # posterior published through atom->ucgp
    assert np.abs(b["ucgsoftmaxscores"]).max() == 0.0          # the style keeps its scores private
    assert abs(e - ref.eng_vdwl()) <= 1e-8 * abs(ref.eng_vdwl())
    assert rel_err(vir, ref.virial()[0]) <= 1e-8                # newton off: the shipped code tallies the virial
    assert ctx.status()[0] == 0


def test_bethe_density_trajectory(pkg, fixtures, tmp_path):
    """config-3 deck: bethe_density + fix nve/ucgld + fix ucgstate, 20 steps with rebuilds"""
    liq = _liq(6)
    ref, ctx = _setup(pkg, fixtures, tmp_path, liq)
    ref.command("fix 1 all nve/ucgld")
    ref.command("fix 2 all ucgstate")
    # the style never writes atom->num_ucgstates (uninitialised in LAMMPS, quirk Q24); the device uses
    # the state count of the site's type, which is what the other UCG pair styles store there
    ref.command("num_ucgstates 2")
    ref.setup(1)
    ref.run(20, 1)
    a = ref.get_atoms()
    ctx.deck_configure(pair_style=3, nve=1, ucgstate=1, thermo_every=1)
    ctx.setup()
    ctx.run(20)
    b = ctx.atoms_download(["x", "v", "f", "ucgp", "ucgstate"])
    box = liq.box_hi - liq.box_lo
    dx = b["x"] - a["x"]
    dx -= box * np.round(dx / box)
    assert np.abs(dx).max() <= 1e-9
    assert rel_err(b["v"], a["v"]) <= 1e-8
    assert rel_err(b["f"], a["f"]) <= 1e-6
    assert np.abs(b["ucgp"] - a["ucgp"]).max() <= 1e-9
    assert np.array_equal(b["ucgstate"], a["ucgstate"])
    assert abs(ctx.thermo()[0] - ref.eng_vdwl()) <= 1e-7 * abs(ref.eng_vdwl())


def test_bethe_density_rejects_density_types_other_than_1(pkg, fixtures, tmp_path):
    """Q15: the density -> probability map exists for actual type 1 only"""
    from lammps_ucg_dev_b200 import engine
    liq = _liq(4)
    liq.type[:] = 2
    t = fixtures["table4096"]
    ctx = pkg.Context(0)
    ctx.set_units(1.0, 1.0, 1.0)
    ctx.set_box(liq.box_lo, liq.box_hi)
    idx = engine.HostTable.from_file(t, "UCG_00", 2.5, 1, 4096).upload(ctx)
    # actual type 1 = one-state (formal 1), actual type 2 = two-state (formal 2,3)
    tabindex = np.full((4, 4), idx, np.int32)
    cutsq = np.zeros((4, 4)); cutsq[1:, 1:] = 2.5 ** 2
    ctx.set_types(2, 3, [0, 1, 2], [[0, 0], [1, 0], [2, 3]], [0.0, 0.0, 0.0, 0.1], [0.0, 1.0, 1.0, 1.0])
    ctx.set_pair_maps(tabindex, cutsq)
    ctx.set_kT(1.0)
    ctx.pair_bethe_density_configure([0, 0, 1], [0, 0, 1], [0.0, 0.0, 12.0], [0.0, 0.0, 1.5])
    ctx.neigh_configure(0.3)
    engine.upload_liquid(ctx, liq)
    ctx.neigh_build()
    ctx.pair_bethe_density(1, 1)
    assert ctx.status()[0] == 4


def test_bethe_density_log_sum_switch(pkg, fixtures, tmp_path, monkeypatch):
    """UCGB200_BD_LOGSUM=0 (one logarithm per ratio, as the reference writes it, instead of one logarithm of the running
    product per lane) agrees with the default to rounding"""
    liq = _liq(6)
    out = {}
    for logsum in ("1", "0"):
        monkeypatch.setenv("UCGB200_BD_LOGSUM", logsum)
        _, ctx = _setup(pkg, fixtures, tmp_path, liq)
        ctx.neigh_build()
        ctx.pair_bethe_density(1, 1)
        out[logsum] = (ctx.atoms_download(["f"])["f"], ctx.pair_energy_virial())
    (fa, (ea, va)), (fc, (ec, vc)) = out["1"], out["0"]
    assert rel_err(fa, fc) <= 1e-13 and abs(ea - ec) <= 1e-13 * abs(ec) and rel_err(va, vc) <= 1e-13


def test_bethe_density_uniform_back_force_sweep_changes_nothing(pkg, fixtures, tmp_path, monkeypatch):
    """one threshold radius and one cutoff: the CV back-force sweep gathers one packed {x, y, z, cvf-or-0} record per
    neighbor (k_cv_back_fast) instead of position, type, type parameters and CV force; UCGB200_CV_FAST=0 selects the
    general sweep.  Same expressions in the same order: identical bits"""
    liq = _liq(6)
    out = []
    for fast in ("1", "0"):
        monkeypatch.setenv("UCGB200_CV_FAST", fast)
        _, ctx = _setup(pkg, fixtures, tmp_path, liq)
        ctx.neigh_build()
        ctx.pair_bethe_density(1, 1)
        out.append((ctx.atoms_download(["f"])["f"], ctx.pair_energy_virial()))
    (fa, (ea, va)), (fb, (eb, vb)) = out
    assert np.abs(fa).max() > 0
    assert np.array_equal(fa, fb) and ea == eb and np.array_equal(va, vb)
