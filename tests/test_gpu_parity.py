"""GPU parity tests: the sm_100a kernels (through the C-ABI) against the CPU oracle on the
same seeded inputs.  Tolerances are the north star's: neighbor lists and deterministic state
assignments bit-exact, per-site forces 1e-6 relative, energy/virial 1e-8 relative."""
import os

import numpy as np
import pytest

import decks
from decks import rel_err

pytestmark = pytest.mark.gpu

F_TOL = 1e-6   # forces, ucgforce, scores (relative to the largest component)
E_TOL = 1e-8   # energy, virial


def _liq(n, **kw):
    from lammps_ucg_dev_b200 import synth
    return synth.fcc_liquid(n, **kw)


def _pair_sets(nl, n):
    ti = np.repeat(nl["tag_i"], nl["numneigh"]).astype(np.int64)
    return np.sort(ti * (n + 1) + nl["neigh_tags"])


# ------------------------------------------------------------------ neighbor
@pytest.mark.parametrize("ncell", [5, 8, (4, 6, 9)])
def test_neighbor_list_bit_exact(pkg, fixtures, ncell):
    liq = _liq(ncell)
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.neigh_build()
    nl = ctx.neigh_download()
    of = decks.orc_single_type(liq, fixtures, full=1)
    of.neigh_build_all()
    fi, fj = of.neigh_pairs()
    n = liq.n
    ref = np.sort(fi.astype(np.int64) * (n + 1) + fj)
    got = _pair_sets(nl, n)
    assert got.size == ref.size
    assert np.array_equal(got, ref)          # identical (tag_i, tag_j) multisets
    assert ctx.natoms() == (of.nlocal(), of.nghost())
    # the reference's half list (newton on) is the set of unordered pairs of the same rows
    oh = decks.orc_single_type(liq, fixtures, full=0)
    oh.neigh_build_all()
    hi, hj = oh.neigh_pairs()
    half_ref = np.sort(np.minimum(hi, hj).astype(np.int64) * (n + 1) + np.maximum(hi, hj))
    ti = np.repeat(nl["tag_i"], nl["numneigh"]).astype(np.int64)
    tj = nl["neigh_tags"].astype(np.int64)
    keep = ti < tj
    half_got = np.sort(ti[keep] * (n + 1) + tj[keep])
    assert np.array_equal(half_got, half_ref)


def _canonical_rows(nl):
    """rows with their entries ordered by (tag, image code): what every build variant must agree on.  The ORDER inside
    a row is a schedule choice: inner entries come in stencil order in every variant; the skin entries are grouped by
    displacement level (default build), sorted by distance (all-FP64 build) or left in stencil order (a stencil
    processed in chunks: then every displacement level visits the whole row)."""
    rows = np.repeat(np.arange(len(nl["tag_i"])), nl["numneigh"])
    key = nl["neigh_tags"].astype(np.int64) * 64 + nl["neigh_shift"]
    o = np.lexsort((key, rows))
    return key[o]


@pytest.mark.parametrize("knobs", [dict(UCGB200_BUILD_LANES="1"), dict(UCGB200_BUILD_TILED="0"), dict(UCGB200_TILE_CAP="64"), dict(UCGB200_TILE_CAP="200"),
                                   dict(UCGB200_BUILD_F32="0"), dict(UCGB200_BUILD_CULL="0"), dict(UCGB200_BUILD_F32="0", UCGB200_TILE_CAP="64"),
                                   dict(UCGB200_BUILD_F32="0", UCGB200_BUILD_DEFER_KEYS="0"),
                                   dict(UCGB200_BUILD_F32="0", UCGB200_BUILD_DEFER_KEYS="0", UCGB200_TILE_CAP="96")])
@pytest.mark.parametrize("ncell", [(5, 6, 7), 12])
def test_neighbor_build_variants_give_identical_rows(pkg, fixtures, monkeypatch, knobs, ncell):
    """the default build of a one-type system (cell-tiled, single-precision prefilter with an exact FP64 re-test inside
    the FP32 error band), its chunked path (small staging capacity), the all-FP64 cell-tiled build (sort keys of the
    skin entries taken in a dense pass per row, or inline; chunked) and the warp-per-site build must produce the same
    rows; variants that sort the skin entries by distance (un-chunked cell-tiled builds) must agree entry for entry"""
    liq = _liq(ncell)
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.neigh_build()
    a = ctx.neigh_download()
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    ctx2 = decks.gpu_single_type(pkg, liq, fixtures)
    ctx2.neigh_build()
    b = ctx2.neigh_download()
    for key in ("tag_i", "numneigh", "offsets"):
        assert np.array_equal(a[key], b[key]), key
    assert np.array_equal(_canonical_rows(a), _canonical_rows(b))
    # the inner partition (entries inside the cutoff at build time) comes first and in the same order everywhere
    pos = np.zeros((liq.n + 1, 3)); pos[liq.tag] = liq.x
    box = liq.box_hi - liq.box_lo
    ti = np.repeat(a["tag_i"], a["numneigh"])
    for nl in (a, b):
        d = pos[ti] - pos[nl["neigh_tags"]]
        d -= box * np.round(d / box)
        nl["inner"] = (d * d).sum(1) < 2.5 ** 2
    assert np.array_equal(a["inner"], b["inner"])
    assert np.array_equal(a["neigh_tags"][a["inner"]], b["neigh_tags"][b["inner"]])
    if "UCGB200_BUILD_LANES" in knobs or (ncell == 12 and knobs.get("UCGB200_BUILD_TILED") != "0" and "UCGB200_TILE_CAP" not in knobs and "UCGB200_BUILD_F32" not in knobs):
        # a lane per site (UCGB200_BUILD_LANES=1) against a warp per site (default): the same rows entry for entry, whatever the cell population;
        # 20 sites per cell: both runs hold the whole stencil at once and group the skin entries by displacement level
        # (the FP64 build sorts them by distance instead: same sets, same level boundaries, another order inside a level)
        for key in ("neigh_tags", "neigh_shift"):
            assert np.array_equal(a[key], b[key]), key


def test_neighbor_rebuild_decision_matches(pkg, fixtures):
    liq = _liq(6)
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    o = decks.orc_single_type(liq, fixtures)
    ctx.neigh_build()
    o.neigh_build_all()
    assert ctx.neigh_decide() == 0 and o.neigh_decide() == 0
    # move one atom by just under / just over skin/2 = 0.15
    for d, expect in ((0.1499999, 0), (0.1500001, 1)):
        x = liq.x.copy()
        x[17, 1] += d
        ctx.atoms_upload(liq.n, x=x)
        o.set_atoms(x, liq.v, liq.type, liq.mask, liq.tag, liq.molecule, liq.ucgstate, liq.ucgl, liq.ucgvl, liq.ucgml)
        assert ctx.neigh_decide() == expect
        assert o.neigh_decide() == expect


# ---------------------------------------------------------------- pair ucgld
def _check_pair(ctx, ref, o, tol_f=F_TOL):
    got = ctx.atoms_download(["f", "ucgforce", "ucgsoftmaxscores", "num_ucgstates"])
    e, vir = ctx.pair_energy_virial()
    assert rel_err(got["f"], ref["f"]) <= tol_f
    assert rel_err(got["ucgforce"], ref["ucgforce"]) <= tol_f
    assert rel_err(got["ucgsoftmaxscores"], ref["ucgsoftmaxscores"]) <= tol_f
    assert np.array_equal(got["num_ucgstates"], ref["num_ucgstates"])
    assert abs(e - o.eng_vdwl()) <= E_TOL * abs(o.eng_vdwl())
    assert rel_err(vir, o.virial()) <= E_TOL
    assert ctx.status()[0] == 0
    return got


@pytest.mark.parametrize("variant", ["fast_smem", "fast_global", "general"])
@pytest.mark.parametrize("lpa", [4, 8, 16])
def test_pair_ucgld_single_type(pkg, fixtures, variant, lpa, monkeypatch):
    if variant == "general" and lpa != 8:
        pytest.skip("general kernel has one schedule")
    monkeypatch.setenv("UCGB200_FORCE_GENERAL", "1" if variant == "general" else "0")
    monkeypatch.setenv("UCGB200_SMEM_TABLE", "0" if variant == "fast_global" else "1")
    monkeypatch.setenv("UCGB200_LPA", str(lpa))
    liq = _liq(8)
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.neigh_build()
    ctx.pair_ucgld(1, 1)
    o = decks.orc_single_type(liq, fixtures)
    ref = decks.oracle_forces(o)
    got = _check_pair(ctx, ref, o)
    # eflag=0 launch writes the same per-site results
    ctx.pair_ucgld(0, 0)
    again = ctx.atoms_download(["f", "ucgforce", "ucgsoftmaxscores"])
    for k in again:
        assert np.array_equal(again[k], got[k])


@pytest.mark.parametrize("knobs", [dict(UCGB200_TEX="0", UCGB200_BS="512"), dict(UCGB200_TEX="1", UCGB200_BS="768"),
                                   dict(UCGB200_TEX="3", UCGB200_BS="512"), dict(UCGB200_PF="0")])
def test_pair_ucgld_schedule_knobs_do_not_change_results(pkg, fixtures, monkeypatch, knobs):
    """texture-pipe gathers, CTA size and software pipelining are pure schedule choices: per-site results must
    equal the default schedule's bit for bit (same arithmetic, same summation order per lane group)"""
    liq = _liq(9)
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.neigh_build()
    ctx.pair_ucgld(0, 0)
    base = ctx.atoms_download(["f", "ucgforce", "ucgsoftmaxscores"])
    o = decks.orc_single_type(liq, fixtures)
    ref = decks.oracle_forces(o)
    assert rel_err(base["f"], ref["f"]) <= F_TOL
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    ctx.pair_ucgld(0, 0)
    again = ctx.atoms_download(["f", "ucgforce", "ucgsoftmaxscores"])
    for k in again:
        assert np.array_equal(again[k], base[k]), k


def test_skin_level_skipping_is_exact(pkg, fixtures, monkeypatch):
    """rows hold their skin entries in ascending build-time distance and the pair kernel skips those that cannot have
    entered the cutoff given the largest displacement since the build: trajectories over several rebuild intervals
    must be bit-identical with and without the skipping, and the skipping must actually engage"""
    liq = _liq(10, T=1.5)
    out = []
    for levels in ("1", "0"):
        monkeypatch.setenv("UCGB200_SKIN_LEVELS", levels)
        ctx = decks.gpu_single_type(pkg, liq, fixtures)
        ctx.deck_configure(pair_style=0, nve=1, langevin=1, t_start=1.5, t_stop=1.5, t_period=1.0, langevin_seed=11, ucgstate=2,
                           thermo_every=10)
        ctx.setup()
        ctx.run(60)
        out.append((ctx.atoms_download(["x", "v", "f", "ucgl", "ucgvl", "ucgforce", "ucgsoftmaxscores"]), ctx.thermo()))
    (a, ta), (b, tb) = out
    assert ta[11] >= 3                      # several rebuilds happened
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert ta[0] == tb[0]
    # the level counts of a fresh build: level 0 keeps the inner entries plus the first eighth of the skin
    ctx.neigh_build()
    total, _, _ = ctx.neigh_stats()
    assert total > 0


def test_pair_ucgld_tablength_25000_uses_l2_path(pkg, fixtures, tmp_path):
    """the reference's own usage comment quotes `linear 25000` (pair_table_ucg_bethe.cpp:752):
    1.2 MB of tables cannot sit in shared memory"""
    liq = _liq(6)
    ctx = decks.gpu_single_type(pkg, liq, fixtures, tablength=25000)
    ctx.neigh_build()
    ctx.pair_ucgld(1, 1)
    o = decks.orc_single_type(liq, fixtures, tablength=25000)
    _check_pair(ctx, decks.oracle_forces(o), o)


@pytest.mark.parametrize("tabstyle,tablength", [(0, 2000), (2, 1500), (3, 12)])
def test_pair_ucgld_other_table_styles(pkg, fixtures, tabstyle, tablength):
    liq = _liq(5)
    ctx = decks.gpu_single_type(pkg, liq, fixtures, tabstyle=tabstyle, tablength=tablength, cut=2.4)
    ctx.neigh_build()
    ctx.pair_ucgld(1, 1)
    o = decks.orc_single_type(liq, fixtures, tabstyle=tabstyle, tablength=tablength, cut=2.4)
    _check_pair(ctx, decks.oracle_forces(o), o)


def test_pair_ucgld_mixed_cg_ucg_types(pkg, fixtures):
    """scenarios 1-4 (pair_table_ucgld.cpp:219-519) with the intended sj keying of scenario 2 (Q1)"""
    liq = decks.mixed_types(_liq(6))
    ctx = decks.gpu_mixed(pkg, liq, fixtures)
    ctx.neigh_build()
    ctx.pair_ucgld(1, 1)
    o = decks.orc_mixed(liq, fixtures)
    _check_pair(ctx, decks.oracle_forces(o), o)


def test_pair_reports_table_inner_cutoff(pkg, fixtures):
    liq = _liq(4)
    liq.x[1] = liq.x[0] + np.array([0.3, 0.0, 0.0])   # r = 0.3 < table inner cutoff 0.5
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.neigh_build()
    ctx.pair_ucgld(0, 0)
    code, ti, tj, rsq = ctx.status()
    assert code == 1 and {ti, tj} == {1, 2} and rsq == pytest.approx(0.09, rel=1e-9)
    o = decks.orc_single_type(liq, fixtures)
    o.neigh_build_all()
    with pytest.raises(RuntimeError, match="Pair distance < table inner cutoff"):
        o.pair_ucgld(0, 0)


# --------------------------------------------------------------------- fixes
def test_fix_kernels_teacher_forcing(pkg, fixtures):
    liq = _liq(6)
    rng = np.random.default_rng(3)
    liq.ucgvl = rng.normal(0, 60.0, liq.n)          # dt*vl ~ 0.12: many sites leave [0,1] -> wall
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    o = decks.orc_single_type(liq, fixtures)
    o.neigh_build_all(); o.force_clear(); o.pair_ucgld(0, 0)   # sets num_ucgstates
    f = rng.normal(0, 5, (liq.n, 3)); uf = rng.normal(0, 3, liq.n); sc = rng.normal(0, 4, (liq.n, 2))
    sc[0] = (800.0, 900.0); sc[1] = (60.0, -60.0)
    for wall in (0, 1):
        ctx.atoms_upload(liq.n, x=liq.x, v=liq.v, ucgl=liq.ucgl, ucgvl=liq.ucgvl, ucgstate=liq.ucgstate,
                         f=f, ucgforce=uf, ucgsoftmaxscores=sc)
        o.set_atoms(liq.x, liq.v, liq.type, liq.mask, liq.tag, liq.molecule, liq.ucgstate, liq.ucgl, liq.ucgvl, liq.ucgml)
        o.pair_ucgld(0, 0)
        o.set_forces(f, uf, sc)
        dt = 0.002
        ctx.fix_nve_initial(dt, 0.5 * dt, 1, wall); o.nve_initial(1, wall)
        if wall:
            ctx.fix_wall_bias(0.1); o.wall_bias(0.1)
        ctx.fix_nve_final(0.5 * dt, 1, wall); o.nve_final(1, wall)
        got = ctx.atoms_download(["x", "v", "ucgl", "ucgvl", "ucgstate", "ucgforce"])
        ref = o.get_atoms()
        for k in ("x", "v", "ucgl", "ucgvl", "ucgforce"):
            assert rel_err(got[k], ref[k]) <= 1e-14, (wall, k)
        away = np.abs(ref["ucgl"] - 0.5) > 1e-12
        assert np.array_equal(got["ucgstate"][away], ref["ucgstate"][away])
    # fix ucgstate, deterministic and ld
    for mode in (0, 1):
        ctx.atoms_upload(liq.n, ucgl=liq.ucgl, ucgstate=liq.ucgstate, ucgsoftmaxscores=sc)
        o.set_atoms(liq.x, liq.v, liq.type, liq.mask, liq.tag, liq.molecule, liq.ucgstate, liq.ucgl, liq.ucgvl, liq.ucgml)
        o.pair_ucgld(0, 0)
        o.set_forces(None, None, sc)
        ctx.fix_ucgstate(mode=mode); o.ucgstate_post_force(mode=mode)
        got = ctx.atoms_download(["ucgp", "ucgl", "ucgstate"])
        ref = o.get_atoms()
        assert rel_err(got["ucgp"], ref["ucgp"]) <= 1e-10
        assert rel_err(got["ucgl"], ref["ucgl"]) <= 1e-10
        away = np.abs(ref["ucgp"] - 0.5) > 1e-9
        assert away.sum() >= liq.n - 2
        assert np.array_equal(got["ucgstate"][away], ref["ucgstate"][away])
        assert got["ucgstate"][0] == 1   # p == 0.5 rounds half away from zero


# --------------------------------------------------------------- trajectories
@pytest.mark.parametrize("deck", ["C1_det", "ld_wall_bias"])
def test_deterministic_trajectory(pkg, fixtures, deck):
    """config 1: table_ucgld + nve/ucgld + t_target provider + ucgstate, vs the oracle's Verlet loop"""
    liq = _liq(8)
    nsteps = 40
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    o = decks.orc_single_type(liq, fixtures)
    o.fix_ttarget(1.0)
    if deck == "C1_det":
        o.fix_nve(); o.fix_ucgstate(mode=0)
        ctx.deck_configure(pair_style=0, nve=1, ucgstate=1, thermo_every=nsteps)
    else:
        o.fix_nve_wall(1, 1, 0.1); o.fix_ucgstate(mode=1)
        ctx.deck_configure(pair_style=0, nve=2, wall_bias=1, wall_barrier=0.1, ucgstate=2, thermo_every=nsteps)
    ctx.setup(); o.setup()
    ctx.run(nsteps); o.run(nsteps, thermo_every=nsteps)
    got = ctx.atoms_download(["x", "v", "f", "ucgl", "ucgvl", "ucgp", "ucgstate", "ucgforce"])
    ref = o.get_atoms()
    th = ctx.thermo()
    assert int(th[11]) == o.nbuilds() and o.nbuilds() >= 1      # same rebuild steps
    # wrapped coordinates: compare modulo the box
    box = liq.box_hi - liq.box_lo
    dx = got["x"] - ref["x"]
    dx -= box * np.round(dx / box)
    assert np.abs(dx).max() <= 1e-9
    assert rel_err(got["v"], ref["v"]) <= 1e-8
    assert rel_err(got["f"], ref["f"]) <= F_TOL
    assert rel_err(got["ucgforce"], ref["ucgforce"]) <= F_TOL
    assert rel_err(got["ucgl"], ref["ucgl"]) <= 1e-8
    assert rel_err(got["ucgp"], ref["ucgp"]) <= 1e-8
    away = np.abs(ref["ucgp"] - 0.5) > 1e-7 if deck == "C1_det" else np.abs(ref["ucgl"] - 0.5) > 1e-9
    assert np.array_equal(got["ucgstate"][away], ref["ucgstate"][away])
    assert abs(th[0] - o.eng_vdwl()) <= 1e-7 * abs(o.eng_vdwl())


def test_langevin_statistics(pkg, fixtures):
    """fix ucgld/langevin: different RNG streams by design -> statistical check: the lambda
    temperature relaxes to the target (period 0.1 tau, 1500 steps) on both implementations"""
    liq = _liq(8, T=0.2)            # start the lambda DOF cold: vl ~ N(0, sqrt(0.2/ml))
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    o = decks.orc_single_type(liq, fixtures)
    o.fix_nve_wall(1, 0, 0.1); o.fix_langevin(1.0, 1.0, 0.1, 4711); o.fix_ucgstate(mode=1)
    ctx.deck_configure(pair_style=0, nve=2, langevin=1, t_start=1.0, t_stop=1.0, t_period=0.1, langevin_seed=4711,
                       ucgstate=2)
    ctx.setup(); o.setup()
    tg, to = [], []
    for _ in range(15):
        ctx.run(100); o.run(100)
        tg.append(ctx.thermo()[9]); to.append(o.lambda_temp())
    tg, to = np.array(tg[5:]), np.array(to[5:])
    # 2048 sites: relative sd of one sample ~ sqrt(2/2048) = 3.1 %; 10 samples -> ~1 %; allow 5 sigma + dt bias
    assert abs(tg.mean() - 1.0) < 0.08
    assert abs(to.mean() - 1.0) < 0.08
    assert abs(tg.mean() - to.mean()) < 0.06
    # langevin force distribution: uniform noise with variance gamma2^2/12 around the drag term
    g1 = np.array([0.0, -10.0 / 0.1, -10.0 / 0.1]); g2 = np.sqrt(10.0) * np.sqrt(24.0 / 0.1 / 0.002) * np.ones(3); g2[0] = 0
    ctx.force_clear()
    ctx.fix_langevin(g1, g2, 1.0, 99, 12345)
    got = ctx.atoms_download(["ucgforce", "ucgvl"])
    noise = (got["ucgforce"] - g1[1] * got["ucgvl"]) / g2[1]
    assert noise.min() >= -0.5 and noise.max() < 0.5
    assert abs(noise.mean()) < 5 * np.sqrt(1 / 12 / liq.n)
    assert abs(noise.var() - 1 / 12) < 0.01


def test_ucgstate_mc_statistics(pkg, fixtures):
    """fix ucgstate mc: literal rule (Q18): state 0 with probability min(ratio,1)*rate"""
    liq = _liq(10)
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    n = liq.n
    sc = np.zeros((n, 2)); sc[:, 1] = np.log(3.0)        # p = 0.75 everywhere
    st = (np.arange(n) % 2).astype(np.int32)
    ctx.atoms_upload(n, ucgsoftmaxscores=sc, ucgstate=st)
    ctx.fix_ucgstate(mode=2, seed=77, rate=0.6, step=5)
    got = ctx.atoms_download(["ucgstate", "ucgp"])
    assert np.allclose(got["ucgp"], 0.75)
    # from state 0: factor = min(p/(1-p),1)*rate = 0.6 ; from state 1: (1-p)/p*rate = 0.2
    f0 = 1.0 - got["ucgstate"][st == 0].mean()
    f1 = 1.0 - got["ucgstate"][st == 1].mean()
    sd = np.sqrt(0.25 / (n / 2))
    assert abs(f0 - 0.6) < 5 * sd and abs(f1 - 0.2) < 5 * sd


# ------------------------------------------------- full-size properties (1 M sites)
def test_full_size_properties_1M(pkg, fixtures):
    """BASELINE config 2 size: size-independent properties instead of an oracle run:
    total force vanishes (Newton's 3rd law through the full list), the energy is the same
    under two kernel schedules and through the shared-memory and L2 table paths, and every
    site has at least the lattice coordination in its row."""
    liq = _liq(63)
    assert liq.n == 1000188
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.neigh_build()
    total, maxrow, _ = ctx.neigh_stats()
    assert 70 * liq.n < total < 85 * liq.n
    res = {}
    for lpa, smem in ((8, 1), (16, 1), (4, 1), (8, 0)):
        os.environ["UCGB200_LPA"] = str(lpa); os.environ["UCGB200_SMEM_TABLE"] = str(smem)
        ctx.pair_ucgld(1, 1)
        res[lpa, smem] = (ctx.pair_energy_virial(), ctx.atoms_download(["f", "ucgforce"]))
    os.environ.pop("UCGB200_LPA"); os.environ.pop("UCGB200_SMEM_TABLE")
    (e0, v0), a0 = res[8, 1]
    fscale = np.abs(a0["f"]).max()
    assert np.abs(a0["f"].sum(0)).max() < 1e-9 * fscale * np.sqrt(liq.n)
    for key in ((16, 1), (4, 1), (8, 0)):
        (e, v), a = res[key]
        assert abs(e - e0) <= 1e-11 * abs(e0)
        assert rel_err(v, v0) <= 1e-10
        assert rel_err(a["f"], a0["f"]) <= 1e-11
    assert ctx.status()[0] == 0


def test_full_size_density_styles_1M(pkg, fixtures, monkeypatch):
    """the other pair styles at 1 M sites: the shared-memory-table kernels (persistent CTAs, transposed 16-byte row
    fetches) and the general kernels (tables through L1, one table_eval per (state, state) table) are independent
    code paths over the same rows and must agree to rounding"""
    from lammps_ucg_dev_b200 import engine
    liq = _liq(63)
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.pair_bethe_density_configure([0, 1], [0, 1], [0.0, 12.0], [0.0, 1.5])
    ctx.neigh_build()

    def both(fn, fields):
        out = []
        for general in ("0", "1"):
            monkeypatch.setenv("UCGB200_FORCE_GENERAL", general)
            fn()
            out.append((ctx.pair_energy_virial(), ctx.atoms_download(fields)))
        monkeypatch.setenv("UCGB200_FORCE_GENERAL", "0")
        return out
    for fn, fields in ((lambda: ctx.pair_bethe(1, 1, 1, 0, 2), ["f", "ucgsoftmaxscores"]),
                       (lambda: ctx.pair_bethe_density(1, 1), ["f", "ucgp"])):
        ((e0, v0), a0), ((e1, v1), a1) = both(fn, fields)
        assert abs(e0 - e1) <= 1e-11 * abs(e0)
        assert rel_err(v0, v1) <= 1e-10
        for k in fields:
            assert rel_err(a0[k], a1[k]) <= 1e-11, k
        assert np.isfinite(a0["f"]).all()
    assert ctx.status()[0] == 0
    # rleucg (state types) on its own context
    t = fixtures["table4096"]
    c2 = pkg.Context(0)
    c2.set_units(1.0, 1.0, 1.0)
    c2.set_box(liq.box_lo, liq.box_hi)
    idx = [engine.HostTable.from_file(t, k, 2.5, 1, 4096).upload(c2) for k in ("UCG_00", "UCG_01", "UCG_11")]
    tabindex = np.zeros((3, 3), np.int32)
    tabindex[1, 1], tabindex[1, 2], tabindex[2, 1], tabindex[2, 2] = idx[0], idx[1], idx[1], idx[2]
    cutsq = np.zeros((3, 3)); cutsq[1:, 1:] = 2.5 ** 2
    c2.pair_rleucg_configure(2, [0, 1, 1], 1, [0, 2], [0, 1], [0.0, 12.0], [0.0, 1.5], [0.0, 0.3, 0.0], tabindex, cutsq,
                             [0.0, 1.0, 1.0], 1.0)
    c2.neigh_configure(0.3)
    engine.upload_liquid(c2, liq)
    c2.neigh_build()
    res = []
    for general in ("0", "1"):
        monkeypatch.setenv("UCGB200_FORCE_GENERAL", general)
        c2.pair_rleucg(1, 1)
        res.append((c2.pair_energy_virial(), c2.atoms_download(["f"])))
    ((e0, v0), a0), ((e1, v1), a1) = res
    assert abs(e0 - e1) <= 1e-11 * abs(e0) and rel_err(v0, v1) <= 1e-10 and rel_err(a0["f"], a1["f"]) <= 1e-11
    # the pair part is symmetric, so the total force vanishes up to the CV back-force asymmetry of the style (Q17)
    assert c2.status()[0] == 0


# ---------------------------------------------------------------- pair bethe
@pytest.mark.parametrize("method,pseudo,prior", [(1, 0, 2), (0, 0, 0), (1, 0, 0), (1, 1, 2)])
def test_pair_bethe_single_evaluation(pkg, fixtures, method, pseudo, prior):
    """PairTable_UCG_Bethe::compute vs the (reference-pinned) oracle on the pristine liquid"""
    liq = _liq(7)
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.neigh_build()
    ctx.pair_bethe(1, 1, method=method, pseudo=pseudo, prior=prior)
    o = decks.orc_single_type(liq, fixtures)
    o.pair_bethe_config(method, pseudo, prior)
    ref = decks.oracle_forces(o, pair="bethe")
    got = ctx.atoms_download(["f", "ucgforce", "ucgsoftmaxscores", "num_ucgstates"])
    e, vir = ctx.pair_energy_virial()
    assert rel_err(got["f"], ref["f"]) <= F_TOL
    assert abs(e - o.eng_vdwl()) <= E_TOL * abs(o.eng_vdwl())
    assert rel_err(vir, o.virial()) <= E_TOL
    assert not got["ucgforce"].any() and not ref["ucgforce"].any()
    if pseudo == 0:      # SCE scores depend on the half-list row owner in the reference
        assert rel_err(got["ucgsoftmaxscores"], ref["ucgsoftmaxscores"]) <= F_TOL
    assert ctx.status()[0] == 0


def test_pair_bethe_trajectory_with_ucgstate(pkg, fixtures):
    """bethe + nve/ucgld + ucgstate (deterministic): ucgl = ucgp after every step, the regime in
    which the reference's prior rule does not depend on the list order"""
    liq = _liq(7)
    liq.ucgvl[:] = 0.0      # lambda at rest (ucgforce is 0 in this style), so ucgl == ucgp at every evaluation
    nsteps = 25
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    o = decks.orc_single_type(liq, fixtures)
    o.pair_bethe_config(1, 0, 2)
    o.fix_ttarget(1.0); o.fix_nve(); o.fix_ucgstate(mode=0)
    ctx.deck_configure(pair_style=1, bethe_method=1, bethe_pseudo=0, bethe_prior=2, nve=1, ucgstate=1, thermo_every=nsteps)
    ctx.setup(); o.setup()
    ctx.run(nsteps); o.run(nsteps, thermo_every=nsteps)
    got = ctx.atoms_download(["x", "v", "f", "ucgl", "ucgp", "ucgstate"])
    ref = o.get_atoms()
    box = liq.box_hi - liq.box_lo
    dx = got["x"] - ref["x"]
    dx -= box * np.round(dx / box)
    assert np.abs(dx).max() <= 1e-9
    assert rel_err(got["f"], ref["f"]) <= F_TOL
    assert rel_err(got["ucgp"], ref["ucgp"]) <= 1e-8
    away = np.abs(ref["ucgp"] - 0.5) > 1e-7
    assert np.array_equal(got["ucgstate"][away], ref["ucgstate"][away])
    assert abs(ctx.thermo()[0] - o.eng_vdwl()) <= 1e-7 * abs(o.eng_vdwl())


def test_pair_bethe_mixed_types(pkg, fixtures):
    liq = decks.mixed_types(_liq(6))
    ctx = decks.gpu_mixed(pkg, liq, fixtures)
    ctx.neigh_build()
    ctx.pair_bethe(1, 1, method=1, pseudo=0, prior=2)
    o = decks.orc_mixed(liq, fixtures)
    o.pair_bethe_config(1, 0, 2)
    ref = decks.oracle_forces(o, pair="bethe")
    got = ctx.atoms_download(["f", "ucgsoftmaxscores"])
    e, vir = ctx.pair_energy_virial()
    assert rel_err(got["f"], ref["f"]) <= F_TOL
    assert abs(e - o.eng_vdwl()) <= E_TOL * abs(o.eng_vdwl())
    assert rel_err(got["ucgsoftmaxscores"], ref["ucgsoftmaxscores"]) <= F_TOL


def test_speculative_pair_launch_changes_nothing(pkg, fixtures, monkeypatch):
    """the resident loop queues the next pair evaluation before the host has read the rebuild flag whenever the last
    displacement read-back says a rebuild is some steps away (csrc/run.cu); with and without it, over several rebuilds
    and thermo steps, every per-site array, the energy and the rebuild count are bit-identical"""
    from lammps_ucg_dev_b200 import synth
    liq = synth.fcc_liquid(10)
    out = []
    for spec in ("1", "0"):
        monkeypatch.setenv("UCGB200_SPECULATE", spec)
        ctx = decks.gpu_single_type(pkg, liq, fixtures)
        ctx.deck_configure(pair_style=0, nve=2, wall_bias=1, wall_barrier=0.2, langevin=1, t_start=1.5, t_stop=1.0, t_period=0.5,
                           langevin_seed=99, ucgstate=2, thermo_every=7)
        ctx.setup()
        ctx.run(40)
        ctx.run(37)
        th = ctx.thermo()
        a = ctx.atoms_download(["x", "v", "f", "ucgl", "ucgvl", "ucgp", "ucgforce", "ucgsoftmaxscores", "ucgstate"])
        out.append((a, th))
    (a, ta), (b, tb) = out
    assert ta[11] >= 3 and ta[11] == tb[11]          # rebuilds happened, on the same steps
    assert ta[0] == tb[0] and np.array_equal(ta[1:7], tb[1:7])
    for k in a:
        assert np.array_equal(a[k], b[k]), k


# ------------------------------------------------------- resident loop: error word and post_force order
def test_resident_run_stops_on_table_inner_cutoff(pkg, fixtures):
    """'Pair distance < table inner cutoff' is error->one in the reference (pair_table_ucgld.cpp:223): the resident
    loop must not run on with that pair's force missing.  The code comes back from ucgb200_setup /
    ucgb200_run, the word stays readable (peek) and one ucgb200_status call yields the pair and clears it."""
    from lammps_ucg_dev_b200 import UCGError
    liq = _liq(5)
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.deck_configure(pair_style=0, nve=1, ucgstate=1, thermo_every=0)
    ctx.setup()
    ctx.run(3)                               # a healthy run first
    assert ctx.status_peek()[0] == 0
    x = ctx.atoms_download(["x"])["x"]
    x[1] = x[0] + np.array([0.3, 0.0, 0.0])  # r = 0.3 < table inner cutoff 0.5
    ctx.atoms_upload(liq.n, x=x)
    with pytest.raises(UCGError) as ei:
        ctx.run(5)
    assert ei.value.rc == 1
    code, ti, tj, rsq = ctx.status_peek()
    assert code == 1 and {ti, tj} == {1, 2}
    assert ctx.status() == (code, ti, tj, rsq)
    assert ctx.status()[0] == 0              # cleared by the read
    # the same at setup
    liq.x[1] = liq.x[0] + np.array([0.3, 0.0, 0.0])
    ctx2 = decks.gpu_single_type(pkg, liq, fixtures)
    ctx2.deck_configure(pair_style=0, nve=1, ucgstate=1, thermo_every=0)
    with pytest.raises(UCGError) as ei:
        ctx2.setup()
    assert ei.value.rc == 1


@pytest.mark.parametrize("fused", ["1", "0"])
def test_post_force_order_follows_the_deck(pkg, fixtures, monkeypatch, fused):
    """[stock] Modify::post_force runs the fixes in definition order.  With fix nve/ucgld/wall/hard bias_potential
    defined BEFORE a non-ld fix ucgstate the bias reads the ucgl of the integrator, not the ucgp the state fix is
    about to write: the oracle registers the fixes in that order, the device loop is told through
    ucgb200_deck::post_force_order (VerletUCGB200::collect_deck records it from the deck)"""
    monkeypatch.setenv("UCGB200_FUSED_TAIL", fused)
    liq = _liq(6)
    nsteps = 12
    res = {}
    for order in (321, 123):          # wall bias first / last
        ctx = decks.gpu_single_type(pkg, liq, fixtures)
        o = decks.orc_single_type(liq, fixtures)
        o.fix_ttarget(1.0)
        if order == 321:
            o.fix_nve_wall(1, 1, 0.3); o.fix_ucgstate(mode=0)
        else:
            o.fix_ucgstate(mode=0); o.fix_nve_wall(1, 1, 0.3)
        ctx.deck_configure(pair_style=0, nve=2, wall_bias=1, wall_barrier=0.3, ucgstate=1, thermo_every=0,
                           post_force_order=order)
        ctx.setup(); o.setup()
        ctx.run(nsteps); o.run(nsteps)
        got = ctx.atoms_download(["x", "ucgl", "ucgvl", "ucgp", "ucgforce"])
        ref = o.get_atoms()
        for k, tol in (("ucgvl", 1e-9), ("ucgl", 1e-9), ("ucgp", 1e-9), ("ucgforce", F_TOL)):
            assert rel_err(got[k], ref[k]) <= tol, (order, k)
        res[order] = got
    # the two orders really differ (otherwise the test proves nothing)
    assert np.abs(res[321]["ucgvl"] - res[123]["ucgvl"]).max() > 1e-6
    with pytest.raises(Exception):
        ctx.deck_configure(pair_style=0, nve=2, wall_bias=1, ucgstate=1, post_force_order=122)
        ctx.setup(); ctx.run(1)


@pytest.mark.parametrize("mode", ["1", "2"])
@pytest.mark.parametrize("ncell", [8, (5, 6, 7)])
def test_pair_ucgld_newton_third_law_variant(pkg, fixtures, monkeypatch, ncell, mode):
    """UCGB200_N3L=1: every owned-owned pair evaluated once, the partner's side scattered with red.global.add.f64
    (pair_table_ucgld.cpp:500-502, :516, :527-529 — what the reference's half list does); UCGB200_N3L=2: the same
    with one 48-byte TMA bulk reduction (cp.reduce.async.bulk.add.f64) per pair.  Atomic order makes the
    sums non-reproducible in the last bits, so the comparison is to the north-star tolerances, against the oracle
    and against the default full-list kernel."""
    liq = _liq(ncell)
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.neigh_build()
    ctx.pair_ucgld(1, 1)
    base = ctx.atoms_download(["f", "ucgforce", "ucgsoftmaxscores"])
    e0, v0 = ctx.pair_energy_virial()
    monkeypatch.setenv("UCGB200_N3L", mode)
    o = decks.orc_single_type(liq, fixtures)
    ref = decks.oracle_forces(o)
    for ev in (1, 0):
        ctx.pair_ucgld(ev, ev)
        got = ctx.atoms_download(["f", "ucgforce", "ucgsoftmaxscores"])
        for k in got:
            assert rel_err(got[k], ref[k]) <= F_TOL, (ev, k)
            assert rel_err(got[k], base[k]) <= 1e-11, (ev, k)
    ctx.pair_ucgld(1, 1)
    e, vir = ctx.pair_energy_virial()
    assert abs(e - o.eng_vdwl()) <= E_TOL * abs(o.eng_vdwl())
    assert rel_err(vir, o.virial()) <= E_TOL
    assert abs(e - e0) <= 1e-11 * abs(e0) and rel_err(vir, v0) <= 1e-10
    assert ctx.status()[0] == 0


def test_neighbor_f32_prefilter_is_exact_at_the_thresholds(pkg, fixtures, monkeypatch):
    """pairs placed within 1e-9 .. 1e-5 of the two thresholds (cut+skin, cut), where the single-precision squared
    distance of the default build cannot decide: the FP64 re-test must give the rows of the all-FP64 build and of the
    oracle, entry for entry"""
    liq = _liq(7)
    rng = np.random.default_rng(17)
    n = liq.n
    x = liq.x.copy()
    box = liq.box_hi - liq.box_lo
    donors = rng.choice(n, 120, replace=False)
    for k, j in enumerate(donors):
        i = (j + 1 + k) % n
        u = rng.normal(size=3); u /= np.linalg.norm(u)
        target = (2.8, 2.5)[k % 2]
        eps = (1e-9, -1e-9, 3e-7, -3e-7, 1e-5, -1e-5, 0.0)[k % 7]
        x[j] = x[i] + (target + eps) * u
        x[j] = liq.box_lo + np.mod(x[j] - liq.box_lo, box)
    liq.x = x
    rows = {}
    for f32 in ("1", "0"):
        monkeypatch.setenv("UCGB200_BUILD_F32", f32)
        ctx = decks.gpu_single_type(pkg, liq, fixtures)
        ctx.neigh_build()
        rows[f32] = ctx.neigh_download()
    for key in ("tag_i", "numneigh", "offsets"):
        assert np.array_equal(rows["1"][key], rows["0"][key]), key
    assert np.array_equal(_canonical_rows(rows["1"]), _canonical_rows(rows["0"]))     # same rows; the order inside the skin part differs by design
    of = decks.orc_single_type(liq, fixtures, full=1)
    of.neigh_build_all()
    fi, fj = of.neigh_pairs()
    nl = rows["1"]
    assert np.array_equal(_pair_sets(nl, n), np.sort(fi.astype(np.int64) * (n + 1) + fj))


@pytest.mark.parametrize("deck", [dict(nve=1, langevin=1, t_start=1.0, t_stop=1.0, t_period=1.0, langevin_seed=5, ucgstate=2),
                                  dict(nve=2, wall_bias=1, wall_barrier=0.1, ucgstate=1),
                                  dict(nve=1, ucgstate=1)])
@pytest.mark.parametrize("split", ["0", "1"])
def test_step_host_equals_upload_run_download(pkg, fixtures, deck, split, monkeypatch):
    """ucgb200_step_host (results leave on a second stream while the step still runs: x under the pair kernel, f under
    the fix stages) must deliver bit for bit what upload + run(1) + download deliver, over steps that include rebuilds,
    for decks whose later stages do and do not rewrite lambda / the state.  split = 1: the opt-in variant that runs
    initial_integrate in two halves under the uploads and sends x back before the rebuild decision (again after a rebuild)"""
    monkeypatch.setenv("UCGB200_E2E_SPLIT", split)
    liq = _liq(9, T=2.0)
    n = liq.n
    ins = ("x", "v", "ucgl", "ucgvl", "ucgstate")
    outs = ("x", "v", "f", "ucgl", "ucgvl", "ucgstate", "ucgp", "ucgforce", "ucgsoftmaxscores")
    res = []
    for pipelined in (True, False):
        ctx = decks.gpu_single_type(pkg, liq, fixtures)
        ctx.deck_configure(pair_style=0, thermo_every=0, **deck)
        ctx.setup()
        H = ctx.atoms_download(list(outs))
        for step in range(25):
            inp = {k: H[k] for k in ins}
            if pipelined:
                ctx.step_host(inp, H)
            else:
                ctx.atoms_upload(n, **inp)
                ctx.run(1)
                ctx.atoms_download_into(**H)
        res.append(({k: v.copy() for k, v in H.items()}, ctx.thermo()[11]))
    (a, ra), (b, rb) = res
    assert ra == rb and ra >= 1
    for k in outs:
        assert np.array_equal(a[k], b[k]), k
