"""Brick decomposition check, LAMMPS style ("1 vs P ranks"): the same liquid run as 1, 2, 4 and 8
bricks must agree with itself (forces to 1e-10 relative, neighbor sets and states
exact).  All bricks live in this process as separate contexts on one GPU; the halo buffers
move through the in-process transport, the kernels are the ones a real multi-GPU run uses."""
import numpy as np
import pytest

import decks
from decks import rel_err

pytestmark = pytest.mark.gpu


def _cluster(pkg, fixtures, liq_parts, grid, box):
    import torch
    from lammps_ucg_dev_b200 import engine, multigpu
    nranks = len(liq_parts)
    bricks = {}
    for r, liq in enumerate(liq_parts):
        ctx = multigpu.make_gpu_brick(pkg, 0)
        engine.setup_single_type(ctx, fixtures["table4096"], fixtures["state"], box=box)
        ctx.halo_configure(r, nranks, grid)
        engine.upload_liquid(ctx, liq)
        bricks[r] = ctx

    def sync():
        for b in bricks.values():
            b.sync()
        torch.cuda.synchronize()

    alloc = multigpu.torch_alloc(torch.device("cuda", 0))
    return multigpu.BrickCluster(bricks, multigpu.InProcessTransport(nranks), alloc, pkg.Context.halo_record_bytes(), sync)


def _pairs(cl, n):
    out = []
    for b in cl.bricks.values():
        nl = b.neigh_download()
        ti = np.repeat(nl["tag_i"], nl["numneigh"]).astype(np.int64)
        out.append(ti * (n + 1) + nl["neigh_tags"])
    return np.sort(np.concatenate(out))


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_bricks_agree_with_single_domain(pkg, fixtures, nranks):
    from lammps_ucg_dev_b200 import multigpu, synth
    ncell = (8, 8, 8)
    grid = multigpu.procgrid_for(nranks)
    whole = synth.fcc_liquid_brick(ncell, (1, 1, 1), 0)
    parts = [synth.fcc_liquid_brick(ncell, grid, r) for r in range(nranks)]
    box = (whole.box_lo, whole.box_hi)
    deck = dict(dt=0.002, ucgstate=0)
    ref = _cluster(pkg, fixtures, [whole], (1, 1, 1), box)
    cl = _cluster(pkg, fixtures, parts, grid, box)
    ref.setup(deck)
    cl.setup(deck)
    n = whole.n
    assert sum(b.natoms()[0] for b in cl.bricks.values()) == n        # migration conserved the sites
    assert np.array_equal(_pairs(cl, n), _pairs(ref, n))              # identical neighbor sets
    nsteps = 30
    ref.run(nsteps, ev_last=True)
    cl.run(nsteps, ev_last=True)
    assert ref.nrebuilds >= 2 and cl.nrebuilds == ref.nrebuilds       # same rebuild steps
    a = ref.gather_atoms(["x", "v", "f", "ucgl", "ucgp", "ucgstate", "ucgforce"])
    b = cl.gather_atoms(["x", "v", "f", "ucgl", "ucgp", "ucgstate", "ucgforce"])
    assert np.array_equal(a["tag"], b["tag"])
    L = whole.box_hi - whole.box_lo
    dx = a["x"] - b["x"]
    dx -= L * np.round(dx / L)
    assert np.abs(dx).max() <= 1e-10
    for k in ("v", "f", "ucgl", "ucgp", "ucgforce"):
        assert rel_err(b[k], a[k]) <= 1e-10, k
    away = np.abs(a["ucgp"] - 0.5) > 1e-9
    assert np.array_equal(a["ucgstate"][away], b["ucgstate"][away])
    ea, eb = ref.energy_virial(), cl.energy_virial()
    assert rel_err(eb, ea) <= 1e-10
    assert np.array_equal(_pairs(cl, n), _pairs(ref, n))


def test_single_brick_cluster_equals_resident_run(pkg, fixtures):
    """the Python-driven cluster stepping and ucgb200_run execute the same kernels in the same order"""
    from lammps_ucg_dev_b200 import engine, synth
    liq = synth.fcc_liquid_brick((6, 6, 6), (1, 1, 1), 0)
    cl = _cluster(pkg, fixtures, [liq], (1, 1, 1), (liq.box_lo, liq.box_hi))
    cl.setup(dict(dt=0.002, ucgstate=0))
    cl.run(20)
    ctx = pkg.Context(0)
    engine.setup_single_type(ctx, fixtures["table4096"], fixtures["state"], box=(liq.box_lo, liq.box_hi))
    engine.upload_liquid(ctx, liq)
    ctx.deck_configure(pair_style=0, nve=1, ucgstate=1)
    ctx.setup()
    ctx.run(20)
    a = cl.gather_atoms(["x", "f", "ucgp"])
    b = ctx.atoms_download(["x", "f", "ucgp", "tag"])
    o = np.argsort(b["tag"])
    for k in ("x", "f", "ucgp"):
        assert np.array_equal(a[k], b[k][o]), k
