"""Drop-in check at the LAMMPS-class level: the same serial LAMMPS-like driver runs the same
input lines once with the reference's own classes (oracle/_ref) and once with the product's
GPU-backed classes of the same names (host/styles -> C-ABI -> sm_100a kernels)."""
import numpy as np
import pytest

import ref_binding as rb
from decks import rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (rb.available() and rb.hostdrv_available()), reason="oracle/_ref or _hostdrv not built")]


def _liq(n, **kw):
    from lammps_ucg_dev_b200 import synth
    return synth.fcc_liquid(n, **kw)


def _both(liq, fixtures, fixes, **kw):
    sims = []
    for cls in (rb.RefSim, rb.HostSim):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"], **kw)
        for f in fixes:
            s.command(f)
        sims.append(s)
    return sims


def test_c1_deck_trajectory(pkg, fixtures):
    liq = _liq(7)
    ref, gpu = _both(liq, fixtures, ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld", "fix 2 all ucgstate"])
    for s in (ref, gpu):
        s.setup(1)
        s.run(30, 30)
    a, b = ref.get_atoms(), gpu.get_atoms()
    assert rel_err(b["x"], a["x"]) <= 1e-10
    assert rel_err(b["v"], a["v"]) <= 1e-8
    assert rel_err(b["f"], a["f"]) <= 1e-6
    assert rel_err(b["ucgforce"], a["ucgforce"]) <= 1e-6
    assert rel_err(b["ucgsoftmaxscores"], a["ucgsoftmaxscores"]) <= 1e-6
    assert rel_err(b["ucgp"], a["ucgp"]) <= 1e-8
    away = np.abs(a["ucgp"] - 0.5) > 1e-7
    assert np.array_equal(a["ucgstate"][away], b["ucgstate"][away])
    assert np.array_equal(a["num_ucgstates"], b["num_ucgstates"])
    assert abs(gpu.eng_vdwl() - ref.eng_vdwl()) <= 1e-8 * abs(ref.eng_vdwl())
    # the drop-in reports the per-pair virial the shipped code leaves at zero (Q3)
    assert rel_err(gpu.virial()[0], ref.virial()[1]) <= 1e-8


def test_wall_bias_deck(pkg, fixtures):
    liq = _liq(6)
    ref, gpu = _both(liq, fixtures, ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld/wall/hard bias_potential 0.2",
                                     "fix 2 all ucgstate ld"])
    for s in (ref, gpu):
        s.setup(0)
        s.run(25, 0)
    a, b = ref.get_atoms(), gpu.get_atoms()
    for k, tol in (("x", 1e-10), ("ucgl", 1e-9), ("ucgvl", 1e-8), ("ucgp", 1e-8)):
        assert rel_err(b[k], a[k]) <= tol, k
    away = np.abs(a["ucgl"] - 0.5) > 1e-9
    assert np.array_equal(a["ucgstate"][away], b["ucgstate"][away])


def test_langevin_deck_statistics_and_t_target(pkg, fixtures):
    liq = _liq(8, T=0.3)
    ref, gpu = _both(liq, fixtures, ["fix 1 all nve/ucgld", "fix 2 all ucgld/langevin 1.0 1.0 0.1 4711", "fix 3 all ucgstate ld"])
    tr, tg = [], []
    for s, acc in ((ref, tr), (gpu, tg)):
        s.setup(0)
        for _ in range(12):
            s.run(100, 0)
            acc.append(s.fix_scalar(1))
    tr, tg = np.array(tr[4:]), np.array(tg[4:])
    assert abs(tr.mean() - 1.0) < 0.08 and abs(tg.mean() - 1.0) < 0.08
    assert abs(tr.mean() - tg.mean()) < 0.06


def test_error_texts(pkg, fixtures, tmp_path):
    liq = _liq(4)
    liq.x[1] = liq.x[0] + np.array([0.3, 0.0, 0.0])
    for cls in (rb.RefSim, rb.HostSim):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"])
        s.command("fix 0 all ttarget/stub 1.0")
        with pytest.raises(RuntimeError, match="Pair distance < table inner cutoff"):
            s.compute_once(0)
    for cls in (rb.RefSim, rb.HostSim):
        s = cls()
        s.box(liq.box_lo, liq.box_hi, 2)
        s.atoms(liq)
        with pytest.raises(RuntimeError, match="Unknown table style in pair_style command"):
            s.command(f"pair_style table_ucgld cubic 100 {fixtures['state']}")
        s.command(f"pair_style table_ucgld linear 4096 {fixtures['state']}")
        with pytest.raises(RuntimeError, match="Incorrect number of arguments for pair_coeff"):
            s.command(f"pair_coeff 1 1 2 2 {fixtures['table4096']} UCG_00 2.5 {fixtures['table4096']} UCG_01 2.5")
        with pytest.raises(RuntimeError, match="Fix langevin period must be > 0.0"):
            s.command("fix 9 all ucgld/langevin 1.0 1.0 0.0 5")


def test_bethe_deck(pkg, fixtures):
    liq = _liq(6)
    # lambda at rest: ucgl stays equal to ucgp between steps, the regime in which the reference's
    # prior rule (ucgl for the row owner, ucgp for the neighbor, quirk Q7) is list-order independent
    liq.ucgvl[:] = 0.0
    sims = []
    for cls in (rb.RefSim, rb.HostSim):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"], pair="table_ucg_bethe",
                            extra="method bethe pseudo yes prior ucgl")
        for f in ("fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld", "fix 2 all ucgstate"):
            s.command(f)
        s.setup(1)
        s.run(20, 20)
        sims.append(s)
    ref, gpu = sims
    a, b = ref.get_atoms(), gpu.get_atoms()
    assert rel_err(b["x"], a["x"]) <= 1e-10
    assert rel_err(b["f"], a["f"]) <= 1e-6
    assert rel_err(b["ucgp"], a["ucgp"]) <= 1e-8
    assert abs(gpu.eng_vdwl() - ref.eng_vdwl()) <= 1e-8 * abs(ref.eng_vdwl())


def _deck(cls, liq, lines):
    s = cls()
    s.box(liq.box_lo, liq.box_hi, 2)
    s.atoms(liq)
    for c in lines:
        s.command(c)
    return s


def test_rleucg_deck(pkg, fixtures, tmp_path):
    """pair_style table_rleucg_interface + fix nve/ucgld/wall/hard through the drop-in classes"""
    liq = _liq(6)
    sf = tmp_path / "rle.conf"
    sf.write_text("1 2\n2 density use_entropy\n12.0 1.5\n0.3\n")
    t = fixtures["table4096"]
    lines = ["newton off", "neighbor 0.3 bin", "timestep 0.002", f"pair_style table_rleucg_interface linear 4096 {sf}",
             f"pair_coeff 1 1 {t} UCG_00 2.5", f"pair_coeff 1 2 {t} UCG_01 2.5", f"pair_coeff 2 2 {t} UCG_11 2.5",
             "fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld/wall/hard"]
    ref, gpu = _deck(rb.RefSim, liq, lines), _deck(rb.HostSim, liq, lines)
    for s in (ref, gpu):
        s.setup(1)
        s.run(20, 1)
    a, b = ref.get_atoms(), gpu.get_atoms()
    assert rel_err(b["x"], a["x"]) <= 1e-10
    assert rel_err(b["v"], a["v"]) <= 1e-8
    assert rel_err(b["f"], a["f"]) <= 1e-6
    assert abs(gpu.eng_vdwl() - ref.eng_vdwl()) <= 1e-7 * abs(ref.eng_vdwl())
    assert rel_err(gpu.virial()[0], ref.virial()[0]) <= 1e-7
    for cls in (rb.RefSim, rb.HostSim):
        s = _deck(cls, liq, ["newton on"] + lines[1:])
        with pytest.raises(RuntimeError, match="Newton pair is turned on"):
            s.setup(1)


def test_bethe_density_deck(pkg, fixtures, tmp_path):
    """pair_style table_ucg_bethe_density (reference compiled with the documented repair) + nve/ucgld"""
    liq = _liq(6)
    sf = tmp_path / "bd.conf"
    sf.write_text("1 2 2\n1 2\n1 2 density entropy \n12.0 1.5\n0.0 0.5\n")
    t = fixtures["table4096"]
    lines = ["newton off", "neighbor 0.3 bin", "timestep 0.002", f"pair_style table_ucg_bethe_density linear 4096 {sf}",
             f"pair_coeff 1 1 2 2 {t} UCG_00 2.5 {t} UCG_01 2.5 {t} UCG_01 2.5 {t} UCG_11 2.5",
             "fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld"]
    ref, gpu = _deck(rb.RefSim, liq, lines), _deck(rb.HostSim, liq, lines)
    for s in (ref, gpu):
        s.setup(1)
        s.run(20, 1)
    a, b = ref.get_atoms(), gpu.get_atoms()
    assert rel_err(b["x"], a["x"]) <= 1e-10
    assert rel_err(b["f"], a["f"]) <= 1e-6
    assert np.abs(b["ucgp"] - a["ucgp"]).max() <= 1e-9
    assert abs(gpu.eng_vdwl() - ref.eng_vdwl()) <= 1e-7 * abs(ref.eng_vdwl())
    assert rel_err(gpu.virial()[0], ref.virial()[0]) <= 1e-7


def test_cluster_switch_deck(pkg, fixtures, tmp_path):
    """config-4 style deck through the drop-in classes: table_rleucg_interface + nve/ucgld/wall/hard +
    cluster_switch; same RanPark stream, so the switched atom types must agree exactly"""
    import os
    import test_gpu_cluster_switch as T
    liq, half = T._system(6)
    sims = []
    for cls, sub in ((rb.RefSim, "ref"), (rb.HostSim, "gpu")):
        d = tmp_path / sub
        d.mkdir()
        orig = rb.RefSim
        try:
            T.rb.RefSim = cls          # build the same deck with the other set of classes
            s = T._ref(liq, half, d, fixtures, 1.08, 5, 15123, 0.3)
        finally:
            T.rb.RefSim = orig
        cwd = os.getcwd()
        os.chdir(d)
        try:
            s.setup(1)
            s.run(17, 1)
        finally:
            os.chdir(cwd)
        sims.append(s)
    ref, gpu = sims
    a, b = ref.get_atoms(), gpu.get_atoms()
    assert np.array_equal(a["type"], b["type"]) and (a["type"] != liq.type).sum() > 0
    assert [ref.fix_vector(2, k) for k in range(7)] == [gpu.fix_vector(2, k) for k in range(7)]
    assert rel_err(b["x"], a["x"]) <= 1e-10
    assert rel_err(b["f"], a["f"]) <= 1e-6
    for name in ("cluster_assignment.log", "state_assignment.log"):
        assert (tmp_path / "ref" / name).read_text() == (tmp_path / "gpu" / name).read_text()


def _dump_table(path):
    """(header bytes, column names, rows) of a one-snapshot dump file"""
    raw = open(path, "rb").read()
    head, body = raw.split(b"ITEM: ATOMS", 1)
    lines = body.decode().split("\n")
    cols = lines[0].split()
    return head, cols, np.array([[float(w) for w in l.split()] for l in lines[1:] if l]).reshape(-1, len(cols))


def test_dump_and_read_dump_decks(pkg, fixtures, tmp_path):
    """`dump ... custom` + `dump_modify` + `read_dump` lines through the reference's patched dump_custom.cpp / read_dump.cpp
    and through the product's DumpCustomUCGB200 / ReadDumpUCGB200 classes (device selection, packing, formatting, scatter)"""
    liq = _liq(5, mol_size=4)
    fixes = ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld", "fix 2 all ucgstate"]
    ref, gpu = _both(liq, fixtures, fixes)
    cols = "id mol type mass x y z xs vx fx ucgstate ucgl ucgp"
    files = []
    for tag, s in (("ref", ref), ("gpu", gpu)):
        s.setup(1)
        s.run(5, 0)
        p = str(tmp_path / (tag + ".dump"))
        s.command("dump d all custom 5 %s %s" % (p, cols))
        s.command("dump_modify d sort id thresh ucgl >= 0.08 time yes")
        s.command("dump_write d")
        s.command("undump d")
        files.append(p)
        s.command("dump full all custom 5 %s id type x y z vx ucgstate ucgl ucgp" % (p + ".full"))
        s.command("dump_write full")
        s.command("undump full")
    (h0, c0, t0), (h1, c1, t1) = _dump_table(files[0]), _dump_table(files[1])
    assert h0 == h1 and c0 == c1 and t0.shape == t1.shape and 10 < len(t0) < liq.n   # same header, same atoms selected
    ints = [c0.index(k) for k in ("id", "mol", "type", "ucgstate")]
    assert np.array_equal(t0[:, ints], t1[:, ints])
    assert np.allclose(t0, t1, rtol=3e-6, atol=1e-9)            # 6 printed digits of a 1e-10-close trajectory
    # read the reference's file back into both, after scrambling the state
    states = []
    for s in (ref, gpu):
        n = s.nlocal()
        s.set_state(x=liq.x[::-1].copy(), ucgl=np.zeros(n), ucgstate=np.zeros(n, np.int32), v=np.zeros((n, 3)))
        s.command("read_dump %s 5 x y z vx ucgstate ucgl ucgp box yes" % (files[0] + ".full"))
        assert s.ntimestep() == 5
        states.append(s.get_atoms())
    a, b = states
    for k in ("x", "v", "ucgl", "ucgp", "ucgstate", "tag", "type"):
        assert np.array_equal(a[k], b[k]), k
    # and the run continues from the restored state
    for s in (ref, gpu):
        s.setup(0)
        s.run(3, 0)
    a, b = ref.get_atoms(), gpu.get_atoms()
    assert rel_err(b["x"], a["x"]) <= 1e-10 and rel_err(b["ucgp"], a["ucgp"]) <= 1e-8


@pytest.mark.parametrize("fixes,tols", [
    (["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld", "fix 2 all ucgstate"], "deterministic"),
    (["fix 1 all nve/ucgld/wall/hard bias_potential 0.2", "fix 2 all ucgld/langevin 1.0 0.6 0.5 4711", "fix 3 all ucgstate ld"], "langevin"),
])
def test_run_style_ucg_b200_is_the_offload_run_kept_on_the_device(pkg, fixtures, fixes, tols):
    """`run_style ucg/b200` (VerletUCGB200: deck handed to ucgb200_setup / ucgb200_run_between, host arrays refreshed on
    thermo steps and at the end) against the same classes driven call by call by the stock Verlet order (offload mode):
    the trajectory is identical; and, for the deterministic deck, equal to the reference's within the usual tolerances"""
    liq = _liq(6)
    sims = {}
    for tag, cls, resident in (("ref", rb.RefSim, False), ("offload", rb.HostSim, False), ("resident", rb.HostSim, True)):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"])
        for f in fixes:
            s.command(f)
        if resident:
            s.command("run_style ucg/b200")
        s.setup(1)
        e0 = s.eng_vdwl()
        s.run(12, 6)
        s.run(13, 0)      # a second run continues from the device state pulled at the end of the first
        sims[tag] = (s, e0)
    a, b = sims["offload"][0].get_atoms(), sims["resident"][0].get_atoms()
    # the deterministic deck is bit-identical; with the Langevin / wall stages the fused per-site pass of the resident loop
    # contracts a few multiply-adds differently from the single-stage kernels: last-bit differences
    exact = tols == "deterministic"
    for k in ("x", "v", "f", "ucgl", "ucgvl", "ucgp", "ucgforce", "ucgsoftmaxscores"):
        if exact:
            assert np.array_equal(a[k], b[k]), (k, float(np.abs(a[k] - b[k]).max()), int((a[k] != b[k]).sum()))
        else:
            assert rel_err(b[k], a[k]) <= 1e-11, (k, rel_err(b[k], a[k]))
    away = np.abs(a["ucgl"] - 0.5) > 1e-9
    assert np.array_equal(a["ucgstate"][away], b["ucgstate"][away]) and np.array_equal(a["num_ucgstates"], b["num_ucgstates"])
    assert sims["offload"][1] == sims["resident"][1]                      # setup energy
    assert sims["resident"][0].ntimestep() == 25
    if tols == "deterministic":
        r = sims["ref"][0].get_atoms()
        assert rel_err(b["x"], r["x"]) <= 1e-10 and rel_err(b["f"], r["f"]) <= 1e-6 and rel_err(b["ucgp"], r["ucgp"]) <= 1e-8
        assert abs(sims["resident"][1] - sims["ref"][1]) <= 1e-8 * abs(sims["ref"][1])


def test_bethe_noise_prior_offload_resident_and_reference(pkg, fixtures):
    """`prior chemical_potential noise L S` (pair_table_ucg_bethe.cpp:189-194, 228-233): acts on the first evaluation only
    (ucgp still -1).  The device draws one Philox number per site keyed by (seed, tag), so the resident loop and the
    offload classes give the same run bit for bit; against the reference (a sequential RanMars draw per site and per
    visit) the comparison is statistical: the spread the noise adds to the first posterior"""
    liq = _liq(6)
    fixes = ("fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld", "fix 2 all ucgstate")
    first = {}
    # the priors only enter the SCE scores (`pseudo no`); with pseudo-likelihood scores no prior changes anything
    noise, plain = "pseudo no prior chemical_potential noise 0.1 77", "pseudo no prior chemical_potential"
    for tag, cls, resident, extra in (("ref", rb.RefSim, False, noise), ("ref0", rb.RefSim, False, plain),
                                      ("offload", rb.HostSim, False, noise), ("resident", rb.HostSim, True, noise),
                                      ("gpu0", rb.HostSim, True, plain)):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"], pair="table_ucg_bethe", extra=extra)
        for f in fixes:
            s.command(f)
        if resident:
            s.command("run_style ucg/b200")
        s.setup(1)
        a = s.get_atoms()
        s.run(10, 0)
        first[tag] = (a, s.get_atoms())
    for k in ("x", "f", "ucgp", "ucgstate", "ucgsoftmaxscores"):
        assert np.array_equal(first["offload"][0][k], first["resident"][0][k]), k
        assert np.array_equal(first["offload"][1][k], first["resident"][1][k]), k
    # what the noise did to the first posterior, reference and device
    def logit_shift(a, b):
        sa, sb = first[a][0]["ucgsoftmaxscores"], first[b][0]["ucgsoftmaxscores"]
        return (sa[:, 1] - sa[:, 0]) - (sb[:, 1] - sb[:, 0])
    d_ref, d_gpu = logit_shift("ref", "ref0"), logit_shift("resident", "gpu0")
    print("noise prior: logit shift mean/std  reference %.4g %.4g   device %.4g %.4g" % (d_ref.mean(), d_ref.std(), d_gpu.mean(), d_gpu.std()))
    assert d_ref.std() > 1e-3 and d_gpu.std() > 1e-3                 # the prior is on
    n = liq.n
    assert abs(d_ref.mean() - d_gpu.mean()) < 6 * (d_ref.std() + d_gpu.std()) / np.sqrt(n)
    # same order only: the reference's SCE j-side formulas are not the mirror image of its i-side ones
    # (pair_table_ucg_bethe.cpp:583-601), so half of its visits respond differently to a prior
    # (measured: device 0.65, reference 0.41)
    assert 0.5 < d_gpu.std() / d_ref.std() < 2.0, (d_gpu.std(), d_ref.std())
    # after the first evaluation the prior is ucgl for every site: no further draws, the run stays finite and bounded
    assert np.all(np.isfinite(first["resident"][1]["x"]))


def test_langevin_bias_compute_in_the_resident_loop(pkg, fixtures):
    """fix ucgld/langevin with a bias temperature compute (fix_modify temp, fix_ucgld_langevin.cpp:174-177): the Tp_BIAS
    branch gives no random force to a site whose lambda velocity is exactly zero (:285).  With ucgvl = 0 on half of the
    sites the setup post_force is deterministic there: equal to the reference's; resident and offload agree everywhere"""
    liq = _liq(6)
    n = liq.n
    vl = np.where(np.arange(n) % 2 == 0, 0.0, 0.05 * np.cos(np.arange(n)))
    res = {}
    for tag, cls, resident, bias in (("ref", rb.RefSim, False, True), ("offload", rb.HostSim, False, True),
                                     ("resident", rb.HostSim, True, True), ("nobias", rb.HostSim, True, False)):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"])
        s.set_state(ucgvl=vl)
        s.command("fix 1 all nve/ucgld")
        s.command("fix 2 all ucgld/langevin 1.0 1.0 0.5 4711")
        s.command("fix 3 all ucgstate ld")
        if bias:
            s.command("compute tb all temp/bias_stub")
            s.command("fix_modify 2 temp tb")
        if resident:
            s.command("run_style ucg/b200")
        s.setup(1)
        a = s.get_atoms()
        s.run(8, 0)
        res[tag] = (a, s.get_atoms())
    order = {k: np.argsort(v[0]["tag"]) for k, v in res.items()}
    uf = {k: v[0]["ucgforce"][order[k]] for k, v in res.items()}
    still = vl == 0.0
    scale = np.abs(uf["ref"]).max()
    for k in ("offload", "resident"):
        assert np.abs(uf[k][still] - uf["ref"][still]).max() <= 1e-6 * scale, k        # no noise, no drag: the pair term alone
    assert np.abs(uf["nobias"][still] - uf["ref"][still]).max() > 1e-3 * scale          # without the bias compute they are kicked
    assert np.abs(uf["resident"][~still] - uf["ref"][~still]).max() > 1e-3 * scale     # moving sites are kicked either way
    for k in ("x", "v", "ucgl", "ucgvl", "ucgforce"):
        assert rel_err(res["resident"][1][k], res["offload"][1][k]) <= 1e-11, k


@pytest.mark.parametrize("style", ["rleucg", "bethe_density"])
def test_run_style_ucg_b200_density_styles(pkg, fixtures, tmp_path, style):
    """the three-sweep density styles under `run_style ucg/b200`: the resident run equals the offload run of the same
    classes (and so, by the deck tests above, the reference's)"""
    liq = _liq(6)
    t = fixtures["table4096"]
    if style == "rleucg":
        sf = tmp_path / "rle.conf"
        sf.write_text("1 2\n2 density use_entropy\n12.0 1.5\n0.3\n")
        lines = ["newton off", "neighbor 0.3 bin", "timestep 0.002", f"pair_style table_rleucg_interface linear 4096 {sf}",
                 f"pair_coeff 1 1 {t} UCG_00 2.5", f"pair_coeff 1 2 {t} UCG_01 2.5", f"pair_coeff 2 2 {t} UCG_11 2.5",
                 "fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld/wall/hard"]
    else:
        sf = tmp_path / "bd.conf"
        sf.write_text("1 2 2\n1 2\n1 2 density entropy \n12.0 1.5\n0.0 0.5\n")
        lines = ["newton off", "neighbor 0.3 bin", "timestep 0.002", f"pair_style table_ucg_bethe_density linear 4096 {sf}",
                 f"pair_coeff 1 1 2 2 {t} UCG_00 2.5 {t} UCG_01 2.5 {t} UCG_01 2.5 {t} UCG_11 2.5",
                 "fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld", "fix 2 all ucgstate"]
    out = []
    for resident in (False, True):
        s = _deck(rb.HostSim, liq, lines + (["run_style ucg/b200"] if resident else []))
        s.setup(1)
        s.run(20, 10)
        out.append((s.get_atoms(), s.eng_vdwl()))
    (a, ea), (b, eb) = out
    for k in ("x", "v", "f", "ucgl", "ucgp"):
        assert rel_err(b[k], a[k]) <= 1e-11, (k, rel_err(b[k], a[k]))
    assert abs(ea - eb) <= 1e-11 * abs(ea)


def test_run_style_ucg_b200_config4_deck(pkg, fixtures, tmp_path):
    """BASELINE config-4 style deck (table_rleucg_interface + nve/ucgld/wall/hard + cluster_switch) under
    `run_style ucg/b200`: the device loop rebuilds, labels and switches by itself and hands control back after every
    switch step for the two log lines — types, statistics and log files equal the offload run's (and the reference's)"""
    import os
    import test_gpu_cluster_switch as T
    liq, half = T._system(6)
    sims = {}
    for cls, sub, resident in ((rb.RefSim, "ref", False), (rb.HostSim, "gpu", False), (rb.HostSim, "res", True)):
        d = tmp_path / sub
        d.mkdir()
        orig = rb.RefSim
        try:
            T.rb.RefSim = cls
            s = T._ref(liq, half, d, fixtures, 1.08, 5, 15123, 0.3)
        finally:
            T.rb.RefSim = orig
        if resident:
            s.command("run_style ucg/b200")
        cwd = os.getcwd()
        os.chdir(d)
        try:
            s.setup(1)
            th = 0 if resident else 1     # the reference needs eflag on every step for this style (Q16); the device always has it
            s.run(17, th)
            s.run(9, th)
        finally:
            os.chdir(cwd)
        sims[sub] = s
    a, b, r = sims["gpu"].get_atoms(), sims["res"].get_atoms(), sims["ref"].get_atoms()
    assert np.array_equal(a["type"], b["type"]) and np.array_equal(r["type"], b["type"]) and (b["type"] != liq.type).sum() > 0
    assert [sims["gpu"].fix_vector(2, k) for k in range(7)] == [sims["res"].fix_vector(2, k) for k in range(7)]
    assert [sims["ref"].fix_vector(2, k) for k in range(7)] == [sims["res"].fix_vector(2, k) for k in range(7)]
    for k in ("x", "v", "f", "ucgl"):
        assert rel_err(b[k], a[k]) <= 1e-11, (k, rel_err(b[k], a[k]))
    assert rel_err(b["x"], r["x"]) <= 1e-10
    for name in ("cluster_assignment.log", "state_assignment.log"):
        text = (tmp_path / "res" / name).read_text()
        assert text == (tmp_path / "gpu" / name).read_text() == (tmp_path / "ref" / name).read_text()
        assert len(text.strip().split("\n")) == 6     # steps 1 6 11 16 21 26


# ------------------------------------------------------------------ rRESPA and minimiser entry points
def _cmp_state(a, b, tol_x=1e-10, tol_v=1e-8):
    assert rel_err(b["x"], a["x"]) <= tol_x
    assert rel_err(b["v"], a["v"]) <= tol_v
    assert rel_err(b["ucgl"], a["ucgl"]) <= 1e-9
    assert rel_err(b["ucgvl"], a["ucgvl"]) <= tol_v
    assert rel_err(b["f"], a["f"]) <= 1e-6
    assert rel_err(b["ucgp"], a["ucgp"]) <= 1e-8


@pytest.mark.parametrize("levels", ["respa 2 4", "respa 3 2 3", "respa 1"])
def test_respa_loop(pkg, fixtures, levels):
    """initial_integrate_respa / final_integrate_respa / post_force_respa of the drop-in classes against the
    reference's (UCG/fix_nve_ucgld.cpp:155-173, fix_ucgstate.cpp:134-136): the same restricted rRESPA loop (pair style
    at the outermost level) drives both class sets for 12 steps; per-level step sizes come from Respa::step, the
    innermost level moves x and lambda, the others only kick.
    Deck: nve/ucgld + ucgstate ld, in which no fix writes ucgl / ucgstate after the innermost level's forward_comm.
    (With a fix that does — deterministic ucgstate, the wall's reflection — the reference's ghosts keep the value
    from before that write while the owners have the new one, [stock] Respa communicates at level 0 only; a drop-in
    whose ghosts are always images of their owners cannot and should not reproduce that.  Those fixes are covered
    call by call below.)"""
    liq = _liq(6)
    fixes = ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld", "fix 2 all ucgstate ld"]
    sims = []
    for cls in (rb.RefSim, rb.HostSim):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"])
        s.command("run_style " + levels)
        for f in fixes:
            s.command(f)
        s.setup(1)
        s.run(12, 12)
        sims.append(s)
    ref, gpu = sims
    a, b = ref.get_atoms(), gpu.get_atoms()
    _cmp_state(a, b)
    assert np.array_equal(a["ucgstate"], b["ucgstate"])
    assert abs(gpu.eng_vdwl() - ref.eng_vdwl()) <= 1e-8 * abs(ref.eng_vdwl())
    if levels == "respa 1":
        # one level with loop 1 is velocity Verlet: must equal the verlet run of the same deck
        v = rb.HostSim.single_type(liq, fixtures["table4096"], fixtures["state"])
        for f in fixes:
            v.command(f)
        v.setup(1)
        v.run(12, 12)
        c = v.get_atoms()
        assert rel_err(b["x"], c["x"]) <= 1e-12 and rel_err(b["ucgl"], c["ucgl"]) <= 1e-12


@pytest.mark.parametrize("fixes", [
    ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld", "fix 2 all ucgstate"],
    ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld/wall/hard bias_potential 0.2", "fix 2 all ucgstate ld"],
    ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld/wall/hard", "fix 2 all ucgstate mc 77 0.5"],
])
def test_respa_entry_points_call_by_call(pkg, fixtures, fixes):
    """every *_respa entry point called on both class sets from the same state, arrays compared after each call
    (UCG/fix_nve_ucgld.cpp:155-173, fix_nve_ucgld_wall_hard.cpp:206-224, fix_ucgstate.cpp:134-136): level 0 of
    initial_integrate_respa integrates positions and lambda with Respa::step[0], every other level is a kick with that
    level's step; final_integrate_respa kicks (and reflects at the wall); fix ucgstate acts at every level"""
    liq = _liq(5)
    liq.ucgvl = np.random.default_rng(5).normal(0, 40.0, liq.n)      # many sites cross the wall
    sims = []
    for cls in (rb.RefSim, rb.HostSim):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"])
        s.command("run_style respa 3 2 3")
        for f in fixes:
            s.command(f)
        s.setup(0)
        sims.append(s)
    ref, gpu = sims
    mc = "mc" in fixes[2]
    calls = [(1, "initial_integrate_respa", 2), (1, "initial_integrate_respa", 1), (1, "initial_integrate_respa", 0),
             (1, "final_integrate_respa", 0), (2, "post_force_respa", 0), (1, "initial_integrate_respa", 0),
             (1, "final_integrate_respa", 1), (2, "post_force_respa", 2), (1, "final_integrate_respa", 2)]
    for ifix, what, lev in calls:
        a0 = ref.get_atoms()
        # same input state on both sides before every call (teacher forcing)
        gpu.set_state(x=a0["x"], v=a0["v"], ucgstate=a0["ucgstate"], ucgl=a0["ucgl"], ucgvl=a0["ucgvl"], ucgp=a0["ucgp"],
                      f=a0["f"], ucgforce=a0["ucgforce"], scores=a0["ucgsoftmaxscores"])
        ref.fix_call_respa(ifix, what, lev)
        gpu.fix_call_respa(ifix, what, lev)
        a, b = ref.get_atoms(), gpu.get_atoms()
        for k in ("x", "v", "ucgl", "ucgvl", "ucgforce", "ucgp"):
            if mc and what == "post_force_respa" and k == "ucgl":
                continue
            assert rel_err(b[k], a[k]) <= 1e-13, (what, lev, k)
        if not (mc and what == "post_force_respa"):       # mc draws from different generators by design
            away = (np.abs(a["ucgp"] - 0.5) > 1e-9) & (np.abs(a["ucgl"] - 0.5) > 1e-12)
            assert np.array_equal(a["ucgstate"][away], b["ucgstate"][away]), (what, lev)
    assert np.abs(ref.get_atoms()["x"] - liq.x).max() > 1e-4        # the sequence did move the sites


def test_respa_langevin_acts_on_the_outermost_level_only(pkg, fixtures):
    """Fix_UCGLD_Langevin::post_force_respa (fix_ucgld_langevin.cpp:217-220): a call at an inner level must not touch
    ucgforce, a call at the outermost level is post_force.  Checked on both class sets through the drag term: with
    gamma2 = 0 impossible (T > 0), so compare ucgforce before/after the inner-level call exactly and the
    outer-level call against drag + bounded noise"""
    liq = _liq(5)
    for cls in (rb.RefSim, rb.HostSim):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"])
        s.command("run_style respa 2 4")
        s.command("fix 1 all nve/ucgld")
        s.command("fix 2 all ucgld/langevin 1.0 1.0 0.5 4711")
        s.command("fix 3 all ucgstate ld")
        s.setup(0)
        before = s.get_atoms()["ucgforce"].copy()
        s.fix_call(1, "post_force_respa_inner")
        assert np.array_equal(s.get_atoms()["ucgforce"], before), cls.__name__
        s.fix_call(1, "post_force_respa_outer")
        after = s.get_atoms()
        g1 = -10.0 / 0.5
        g2 = np.sqrt(10.0) * np.sqrt(24.0 / 0.5 / 0.002)
        noise = (after["ucgforce"] - before - g1 * after["ucgvl"]) / g2
        assert noise.min() >= -0.5 - 1e-9 and noise.max() <= 0.5 + 1e-9, cls.__name__
        assert abs(noise.mean()) < 6 * np.sqrt(1 / 12 / liq.n)


def test_min_post_force(pkg, fixtures):
    """FixUCGState::min_post_force (fix_ucgstate.cpp:138-140) through the force evaluation of a minimiser: displace
    the sites between two evaluations so that the second one sees new scores"""
    liq = _liq(6)
    out = []
    for cls in (rb.RefSim, rb.HostSim):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"])
        s.command("fix 0 all ttarget/stub 1.0")
        s.command("fix 2 all ucgstate")
        s.setup(1)
        x = liq.x + 0.02 * np.sin(np.arange(liq.n * 3).reshape(liq.n, 3))      # a line-search move
        s.set_state(x=x)
        s.min_energy_force(1)
        out.append((s.get_atoms(), s.eng_vdwl()))
    (a, ea), (b, eb) = out
    assert rel_err(b["f"], a["f"]) <= 1e-6
    assert rel_err(b["ucgsoftmaxscores"], a["ucgsoftmaxscores"]) <= 1e-6
    assert rel_err(b["ucgp"], a["ucgp"]) <= 1e-10
    assert rel_err(b["ucgl"], a["ucgl"]) <= 1e-10          # deterministic mode copies ucgp into ucgl
    away = np.abs(a["ucgp"] - 0.5) > 1e-9
    assert np.array_equal(a["ucgstate"][away], b["ucgstate"][away])
    assert abs(eb - ea) <= 1e-8 * abs(ea)


# ------------------------------------------------------------------ offload mode: tracked transfers
@pytest.mark.parametrize("fixes", [
    ["fix 1 all nve/ucgld", "fix 2 all ucgld/langevin 1.5 1.5 0.5 4711", "fix 3 all ucgstate ld"],
    ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld/wall/hard bias_potential 0.2", "fix 2 all ucgstate"],
])
def test_offload_tracked_transfers_change_nothing(pkg, fixtures, monkeypatch, fixes):
    """stock Verlet over an all-UCG deck: with per-field tracking (a field goes up only when the host copy is newer,
    comes down only where stock LAMMPS reads it: x every step, everything before an exchange, on output steps and at
    the end of the run) every array after the run equals the eager mode's bit for bit, over several rebuilds"""
    liq = _liq(8, T=2.0)
    res = []
    for tracked in ("1", "0"):
        monkeypatch.setenv("UCGB200_OFFLOAD_TRACKED", tracked)
        s = rb.HostSim.single_type(liq, fixtures["table4096"], fixtures["state"])
        for f in fixes:
            s.command(f)
        s.setup(1)
        s.run(45, 15)
        res.append((s.get_atoms(), s.eng_vdwl(), s.nbuilds()))
    (a, ea, na), (b, eb, nb) = res
    assert na == nb and na >= 2
    assert ea == eb
    for k in a:
        assert np.array_equal(a[k], b[k]), k


# ------------------------------------------------------------------ per-atom energy / virial (eflag_atom, vflag_atom)
@pytest.mark.parametrize("tabstyle,tablength", [("linear", 4096), ("spline", 1500)])
def test_per_atom_energy_and_virial(pkg, fixtures, tabstyle, tablength):
    """compute pe/atom / stress/atom ask the pair style for per-atom tallies through Pair::ev_setup (ENERGY_ATOM = 2,
    VIRIAL_ATOM = 4); the reference delivers them through [stock] Pair::ev_tally (pair_table_ucgld.cpp:531-533): half
    of every pair's energy and virial to each partner, ghost tallies reverse-communicated to the owners.  The drop-in
    class fills Pair::eatom / Pair::vatom from the device (ucgb200_pair_peratom); they must sum to the global values"""
    liq = _liq(7)
    out = []
    for cls in (rb.RefSim, rb.HostSim):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"], tabstyle=tabstyle, tablength=tablength)
        s.command("fix 0 all ttarget/stub 1.0")
        s.compute_once(3)
        e, v = s.pair_peratom()
        out.append((e, v, s.eng_vdwl(), s.virial()))
    (ea, va, Ea, Wa), (eb, vb, Eb, Wb) = out
    assert rel_err(eb, ea) <= 1e-8
    assert rel_err(vb, va) <= 1e-8
    worst = np.abs(eb - ea).max() / np.abs(ea).max()
    assert abs(eb.sum() - Eb) <= 1e-10 * abs(Eb) and abs(ea.sum() - Ea) <= 1e-10 * abs(Ea)
    assert rel_err(vb.sum(0), Wb[0]) <= 1e-10          # the drop-in's global virial is the per-pair tally (Q3)
    assert rel_err(va.sum(0), Wa[1]) <= 1e-10
    assert abs(Eb - Ea) <= 1e-8 * abs(Ea)
    # the same through pair_style table_ucg_bethe
    out = []
    for cls in (rb.RefSim, rb.HostSim):
        s = cls.single_type(liq, fixtures["table4096"], fixtures["state"], pair="table_ucg_bethe",
                            extra="method bethe pseudo yes prior ucgl")
        s.command("fix 0 all ttarget/stub 1.0")
        s.compute_once(3)
        out.append(s.pair_peratom() + (s.eng_vdwl(),))
    (ea, va, Ea), (eb, vb, Eb) = out
    assert rel_err(eb, ea) <= 1e-8 and rel_err(vb, va) <= 1e-8
    assert abs(eb.sum() - Eb) <= 1e-10 * abs(Eb) and abs(Eb - Ea) <= 1e-8 * abs(Ea)


@pytest.mark.parametrize("style", ["rleucg", "rleucg_mixed_types", "bethe_density"])
def test_per_atom_tallies_of_the_density_styles(pkg, fixtures, tmp_path, style):
    """per-atom energy / virial of table_rleucg_interface and table_ucg_bethe_density: the reference delivers them
    through [stock] ev_tally with newton off (pair_table_rleucg_interface.cpp:439, 488;
    pair_table_ucg_bethe_density.cpp:407, 510, 647, 729): half of every visit to the centre site, half to a LOCAL
    partner, nothing to a ghost.  The device keeps each site's share while it accumulates the global sums, the second
    sweep (CV back-force) adds its virial part"""
    t = fixtures["table4096"]
    ntypes = 2
    if style == "rleucg":
        liq = _liq(6)
        sf = tmp_path / "rle.conf"
        sf.write_text("1 2\n2 density use_entropy\n12.0 1.5\n0.3\n")
        lines = ["newton off", "neighbor 0.3 bin", "timestep 0.002", f"pair_style table_rleucg_interface linear 4096 {sf}",
                 f"pair_coeff 1 1 {t} UCG_00 2.5", f"pair_coeff 1 2 {t} UCG_01 2.5", f"pair_coeff 2 2 {t} UCG_11 2.5",
                 "fix 0 all ttarget/stub 1.0"]
    elif style == "rleucg_mixed_types":
        import test_gpu_cluster_switch as T                  # the config-4 system: two-state sites among plain CG ones
        liq, _ = T._system(6)
        sf = tmp_path / "rle.conf"
        sf.write_text(T.RLE_STATE)
        ntypes = 4
        lines = ["newton off", "neighbor 0.3 bin", "timestep 0.002", f"pair_style table_rleucg_interface linear 4096 {sf}"]
        lines += [f"pair_coeff {i} {j} {t} {kw} 2.5" for i, j, kw in T.PAIRS]
        lines += ["fix 0 all ttarget/stub 1.0"]
    else:
        liq = _liq(6)
        sf = tmp_path / "bd.conf"
        sf.write_text("1 2 2\n1 2\n1 2 density entropy \n12.0 1.5\n0.0 0.5\n")
        lines = ["newton off", "neighbor 0.3 bin", "timestep 0.002", f"pair_style table_ucg_bethe_density linear 4096 {sf}",
                 f"pair_coeff 1 1 2 2 {t} UCG_00 2.5 {t} UCG_01 2.5 {t} UCG_01 2.5 {t} UCG_11 2.5",
                 "fix 0 all ttarget/stub 1.0"]
    out = []
    for cls in (rb.RefSim, rb.HostSim):
        s = cls()
        s.box(liq.box_lo, liq.box_hi, ntypes)
        s.atoms(liq)
        for c in lines:
            s.command(c)
        s.compute_once(3)
        e, v = s.pair_peratom()
        out.append((e, v, s.eng_vdwl(), s.virial()))
    (ea, va, Ea, Wa), (eb, vb, Eb, Wb) = out
    assert np.abs(ea).max() > 0 and np.abs(va).max() > 0
    assert rel_err(eb, ea) <= 1e-8, rel_err(eb, ea)
    assert rel_err(vb, va) <= 1e-8, rel_err(vb, va)
    assert abs(ea.sum() - Ea) <= 1e-10 * abs(Ea) and abs(eb.sum() - Eb) <= 1e-10 * abs(Eb)   # the shares add up to the totals
    assert rel_err(vb.sum(0), Wb[0]) <= 1e-10
    assert abs(Eb - Ea) <= 1e-8 * abs(Ea)