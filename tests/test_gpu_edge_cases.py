"""Edge cases of the hot path on the GPU: empty and tiny systems, ragged (very inhomogeneous) rows that
overflow the estimated row capacity, non-cubic boxes next to the minimum size, repeated re-configuration."""
import numpy as np
import pytest

import decks
from decks import rel_err

pytestmark = pytest.mark.gpu


def _liq(n, **kw):
    from lammps_ucg_dev_b200 import synth
    return synth.fcc_liquid(n, **kw)


def _subset(liq, keep):
    """the liquid restricted to `keep` (sites keep their tags)"""
    import copy
    s = copy.copy(liq)
    for k in ("x", "v", "type", "mask", "tag", "molecule", "ucgstate", "ucgl", "ucgvl", "ucgml"):
        setattr(s, k, np.ascontiguousarray(getattr(liq, k)[keep]))
    s.n = int(np.count_nonzero(keep)) if keep.dtype == bool else len(keep)
    return s


def test_empty_system(pkg, fixtures):
    liq = _liq(4)
    empty = _subset(liq, np.zeros(liq.n, bool))
    ctx = decks.gpu_single_type(pkg, empty, fixtures)
    ctx.neigh_build()
    ctx.pair_ucgld(0, 0)
    ctx.deck_configure(pair_style=0, nve=1, ucgstate=1, thermo_every=1)
    ctx.setup()
    ctx.run(3)
    assert ctx.natoms() == (0, 0)
    assert ctx.status()[0] == 0


@pytest.mark.parametrize("nsites", [1, 2, 3])
def test_tiny_systems_match_the_oracle(pkg, fixtures, nsites):
    """1-3 sites in a periodic box: rows are empty or hold only a few (image) neighbors"""
    liq = _liq(4)
    idx = np.array([0, 1, 5])[:nsites]          # nearest lattice neighbors ~1.19 apart: inside the cutoff
    sub = _subset(liq, idx)
    ctx = decks.gpu_single_type(pkg, sub, fixtures)
    ctx.neigh_build()
    ctx.pair_ucgld(1, 1)
    got = ctx.atoms_download(["f", "ucgforce", "ucgsoftmaxscores"])
    e, vir = ctx.pair_energy_virial()
    o = decks.orc_single_type(sub, fixtures)
    ref = decks.oracle_forces(o)
    if nsites == 1:
        assert np.abs(got["f"]).max() == 0.0 and e == 0.0
    else:
        assert rel_err(got["f"], ref["f"]) <= 1e-6
        assert abs(e - o.eng_vdwl()) <= 1e-8 * abs(o.eng_vdwl())
    assert rel_err(got["ucgforce"], ref["ucgforce"]) <= 1e-6
    assert ctx.status()[0] == 0


def test_ragged_rows_regrow_the_row_capacity(pkg, fixtures):
    """a droplet in an otherwise empty box: the row capacity estimated from the mean density is far too
    small for the droplet's sites (overflow -> regrow -> rebuild) while most of the box has empty cells"""
    liq = _liq(10)
    centre = 0.5 * (liq.box_lo + liq.box_hi)
    r = np.linalg.norm(liq.x - centre, axis=1)
    drop = _subset(liq, r < 4.2)
    assert 150 < drop.n < 400
    ctx = decks.gpu_single_type(pkg, drop, fixtures)
    ctx.neigh_build()
    total, maxrow, _ = ctx.neigh_stats()
    assert maxrow > 2 * total / drop.n * 0 + 40          # interior sites keep their ~78 neighbors
    ctx.pair_ucgld(1, 1)
    got = ctx.atoms_download(["f", "ucgforce"])
    o = decks.orc_single_type(drop, fixtures)
    ref = decks.oracle_forces(o)
    assert rel_err(got["f"], ref["f"]) <= 1e-6
    assert rel_err(got["ucgforce"], ref["ucgforce"]) <= 1e-6
    nl = ctx.neigh_download()
    of = decks.orc_single_type(drop, fixtures, full=1)
    of.neigh_build_all()
    fi, fj = of.neigh_pairs()
    big = int(liq.n) + 1
    assert np.array_equal(np.sort(np.repeat(nl["tag_i"], nl["numneigh"]).astype(np.int64) * big + nl["neigh_tags"]),
                          np.sort(fi.astype(np.int64) * big + fj))


def test_minimum_box_and_reconfiguration(pkg, fixtures):
    """a box barely larger than cut+skin in one direction (one cell, every neighbor is an image there), then the
    same context re-used for a second deck"""
    liq = _liq((2, 5, 4))        # 2 cells * 1.68 = 3.36 > 2.8
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.neigh_build()
    ctx.pair_ucgld(1, 1)
    got = ctx.atoms_download(["f"])
    o = decks.orc_single_type(liq, fixtures)
    ref = decks.oracle_forces(o)
    assert rel_err(got["f"], ref["f"]) <= 1e-6
    # a box shorter than the neighbor cutoff must be refused, like LAMMPS does
    small = _liq((1, 4, 4))
    c2 = decks.gpu_single_type(pkg, small, fixtures)
    with pytest.raises(pkg.UCGError):
        c2.neigh_build()
    # same context, new atom count and box
    liq2 = _liq(5)
    from lammps_ucg_dev_b200 import engine
    ctx.set_box(liq2.box_lo, liq2.box_hi)
    engine.upload_liquid(ctx, liq2)
    ctx.neigh_build()
    ctx.pair_ucgld(1, 1)
    o2 = decks.orc_single_type(liq2, fixtures)
    ref2 = decks.oracle_forces(o2)
    assert rel_err(ctx.atoms_download(["f"])["f"], ref2["f"]) <= 1e-6
