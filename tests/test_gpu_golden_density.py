"""The density pair styles and fix cluster_switch against COMMITTED golden vectors
(tests/golden/ucg_ref_golden_density.npz, made by tests/golden/make_golden_density.py from the reference's
own compiled sources): these tests do not need oracle/_ref at run time."""
import os

import numpy as np
import pytest

from decks import rel_err

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ucg_ref_golden_density.npz"))


def _liq(n, **kw):
    from lammps_ucg_dev_b200 import synth
    return synth.fcc_liquid(n, **kw)


def test_rleucg_golden(pkg, fixtures):
    from lammps_ucg_dev_b200 import engine
    liq = _liq(5)
    t = fixtures["table4096"]
    ctx = pkg.Context(0)
    ctx.set_units(1.0, 1.0, 1.0)
    ctx.set_box(liq.box_lo, liq.box_hi)
    idx = [engine.HostTable.from_file(t, k, 2.5, 1, 4096).upload(ctx) for k in ("UCG_00", "UCG_01", "UCG_11")]
    tabindex = np.zeros((3, 3), np.int32)
    tabindex[1, 1], tabindex[1, 2], tabindex[2, 1], tabindex[2, 2] = idx[0], idx[1], idx[1], idx[2]
    cutsq = np.zeros((3, 3)); cutsq[1:, 1:] = 2.5 ** 2
    ctx.pair_rleucg_configure(2, [0, 1, 1], 1, [0, 2], [0, 1], [0.0, 12.0], [0.0, 1.5], [0.0, 0.3, 0.0], tabindex, cutsq,
                              [0.0, 1.0, 1.0], 1.0)
    ctx.neigh_configure(0.3)
    engine.upload_liquid(ctx, liq)
    ctx.neigh_build()
    ctx.pair_rleucg(1, 1)
    e, vir = ctx.pair_energy_virial()
    assert rel_err(ctx.atoms_download(["f"])["f"], GOLD["rleucg_f"]) <= 1e-6
    assert abs(e - float(GOLD["rleucg_E"])) <= 1e-8 * abs(float(GOLD["rleucg_E"]))
    assert rel_err(vir, GOLD["rleucg_virial"]) <= 1e-8


def test_bethe_density_golden(pkg, fixtures):
    from lammps_ucg_dev_b200 import engine
    liq = _liq(5)
    ctx = pkg.Context(0)
    engine.setup_single_type(ctx, fixtures["table4096"], fixtures["state"], tablength=4096, box=(liq.box_lo, liq.box_hi))
    ctx.pair_bethe_density_configure([0, 1], [0, 1], [0.0, 12.0], [0.0, 1.5])
    engine.upload_liquid(ctx, liq)
    ctx.neigh_build()
    ctx.pair_bethe_density(1, 1)
    b = ctx.atoms_download(["f", "ucgp"])
    e, vir = ctx.pair_energy_virial()
    assert rel_err(b["f"], GOLD["bethe_density_f"]) <= 1e-6
    assert np.abs(b["ucgp"] - GOLD["bethe_density_ucgp"]).max() <= 1e-9
    assert abs(e - float(GOLD["bethe_density_E"])) <= 1e-8 * abs(float(GOLD["bethe_density_E"]))
    assert rel_err(vir, GOLD["bethe_density_virial"]) <= 1e-8


def test_cluster_switch_golden(pkg, fixtures):
    import test_gpu_cluster_switch as T
    liq, half = T._system(6)
    ctx = T._gpu(pkg, liq, half, fixtures, 1.08, 5, 15123, 0.3)
    ctx.setup()
    ctx.run(13)
    b = ctx.atoms_download(["x", "type"])
    assert np.array_equal(b["type"], GOLD["cluster_type"])                 # accept decisions: exact
    assert np.array_equal(ctx.cluster_stats()[:7], GOLD["cluster_stats"])
    box = liq.box_hi - liq.box_lo
    dx = b["x"] - GOLD["cluster_x"]
    dx -= box * np.round(dx / box)
    assert np.abs(dx).max() <= 1e-9
