"""pair_style table_rleucg_interface on the GPU against the reference's own compiled source
(oracle/_ref): single-domain, periodic, newton off, eflag on (quirk Q16)."""
import numpy as np
import pytest

import ref_binding as rb
from decks import rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built")]

RLE_STATE = "1 2\n2 density {ent}\n{rho_th} {r_th}\n{mu}\n"


def _setup(pkg, fixtures, tmp_path, liq, ent="use_entropy", rho_th=12.0, r_th=1.5, mu=0.3, tabstyle="linear", code=1, n=4096):
    from lammps_ucg_dev_b200 import engine
    sf = tmp_path / "rle.conf"
    sf.write_text(RLE_STATE.format(ent=ent, rho_th=rho_th, r_th=r_th, mu=mu))
    t = fixtures["table4096"]
    ref = rb.RefSim()
    ref.box(liq.box_lo, liq.box_hi, 2)
    ref.atoms(liq)
    for c in ("newton off", "neighbor 0.3 bin", "timestep 0.002", f"pair_style table_rleucg_interface {tabstyle} {n} {sf}",
              f"pair_coeff 1 1 {t} UCG_00 2.5", f"pair_coeff 1 2 {t} UCG_01 2.5", f"pair_coeff 2 2 {t} UCG_11 2.5",
              "fix 0 all ttarget/stub 1.0"):
        ref.command(c)
    ctx = pkg.Context(0)
    ctx.set_units(1.0, 1.0, 1.0)
    ctx.set_box(liq.box_lo, liq.box_hi)
    idx = [engine.HostTable.from_file(t, k, 2.5, code, n).upload(ctx) for k in ("UCG_00", "UCG_01", "UCG_11")]
    tabindex = np.zeros((3, 3), np.int32)
    tabindex[1, 1], tabindex[1, 2], tabindex[2, 1], tabindex[2, 2] = idx[0], idx[1], idx[1], idx[2]
    cutsq = np.zeros((3, 3)); cutsq[1:, 1:] = 2.5 ** 2
    ctx.pair_rleucg_configure(2, [0, 1, 1], 1, [0, 2], [0, 1 if ent == "use_entropy" else 0], [0.0, rho_th], [0.0, r_th],
                              [0.0, mu, 0.0], tabindex, cutsq, [0.0, 1.0, 1.0], 1.0)
    ctx.neigh_configure(0.3)
    engine.upload_liquid(ctx, liq)
    return ref, ctx


def _liq(n, **kw):
    from lammps_ucg_dev_b200 import synth
    return synth.fcc_liquid(n, **kw)


@pytest.mark.parametrize("ent,rho_th", [("use_entropy", 12.0), ("no_entropy", 11.0)])
def test_rleucg_single_evaluation(pkg, fixtures, tmp_path, ent, rho_th):
    liq = _liq(7)
    ref, ctx = _setup(pkg, fixtures, tmp_path, liq, ent=ent, rho_th=rho_th)
    ref.compute_once(1)
    a = ref.get_atoms()
    ctx.neigh_build()
    ctx.pair_rleucg(1, 1)
    b = ctx.atoms_download(["f"])
    e, vir = ctx.pair_energy_virial()
    prob, cvf = ctx.pair_rleucg_probabilities()
    assert 0.05 < prob.mean() < 0.95 and prob.std() > 0.005         # both substates populated, sites differ
    assert rel_err(b["f"], a["f"]) <= 1e-6
    assert abs(e - ref.eng_vdwl()) <= 1e-8 * abs(ref.eng_vdwl())
    assert rel_err(vir, ref.virial()[0]) <= 1e-8                    # newton off: the shipped code tallies the virial
    assert ctx.status()[0] == 0


def test_rleucg_trajectory_with_wall_integrator(pkg, fixtures, tmp_path):
    liq = _liq(6)
    ref, ctx = _setup(pkg, fixtures, tmp_path, liq)
    ref.command("fix 1 all nve/ucgld/wall/hard")
    ref.setup(1)
    ref.run(20, 1)
    a = ref.get_atoms()
    ctx.set_timestep(0.002)
    ctx.deck_configure(pair_style=2, nve=2, thermo_every=1)
    ctx.setup()
    ctx.run(20)
    b = ctx.atoms_download(["x", "v", "f"])
    box = liq.box_hi - liq.box_lo
    dx = b["x"] - a["x"]
    dx -= box * np.round(dx / box)
    assert np.abs(dx).max() <= 1e-9
    assert rel_err(b["v"], a["v"]) <= 1e-8
    assert rel_err(b["f"], a["f"]) <= 1e-6
    assert abs(ctx.thermo()[0] - ref.eng_vdwl()) <= 1e-7 * abs(ref.eng_vdwl())


def test_rleucg_rejects_density_types_other_than_1(pkg, fixtures, tmp_path):
    """Q15: the density -> probability map exists for actual type 1 only"""
    from lammps_ucg_dev_b200 import engine
    liq = _liq(4)
    liq.type[:] = 2
    t = fixtures["table4096"]
    ctx = pkg.Context(0)
    ctx.set_box(liq.box_lo, liq.box_hi)
    idx = engine.HostTable.from_file(t, "UCG_00", 2.5, 1, 4096).upload(ctx)
    tabindex = np.full((4, 4), idx, np.int32)
    cutsq = np.zeros((4, 4)); cutsq[1:, 1:] = 2.5 ** 2
    # actual type 1 = one-state (type 1), actual type 2 = two-state (types 2,3)
    ctx.pair_rleucg_configure(3, [0, 1, 2, 2], 2, [0, 1, 2], [0, 0, 1], [0.0, 0.0, 12.0], [0.0, 0.0, 1.5],
                              [0.0, 0.0, 0.1, 0.0], tabindex, cutsq, [0.0, 1.0, 1.0, 1.0], 1.0)
    ctx.neigh_configure(0.3)
    engine.upload_liquid(ctx, liq)
    ctx.neigh_build()
    ctx.pair_rleucg(1, 1)
    assert ctx.status()[0] == 4


def test_rleucg_uniform_back_force_sweep_changes_nothing(pkg, fixtures, tmp_path, monkeypatch):
    """one threshold radius and one cutoff: the CV back-force sweep gathers one packed {x, y, z, cvf-or-0} record per
    neighbor (k_cv_back_fast) instead of position, type, type parameters and CV force; UCGB200_CV_FAST=0 selects the
    general sweep.  Same expressions in the same order: identical bits"""
    liq = _liq(6)
    out = []
    for fast in ("1", "0"):
        monkeypatch.setenv("UCGB200_CV_FAST", fast)
        _, ctx = _setup(pkg, fixtures, tmp_path, liq)
        ctx.neigh_build()
        ctx.pair_rleucg(1, 1)
        out.append((ctx.atoms_download(["f"])["f"], ctx.pair_energy_virial()))
    (fa, (ea, va)), (fb, (eb, vb)) = out
    assert np.abs(fa).max() > 0
    assert np.array_equal(fa, fb) and ea == eb and np.array_equal(va, vb)
