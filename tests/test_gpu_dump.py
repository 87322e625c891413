"""GPU parity of the dump / read_dump taps through the C-ABI:
  * `dump custom` files byte-for-byte equal to the files the reference's own dump_custom.cpp wrote (tests/golden/io),
    with the rows formatted on the device and, again, packed on the device + formatted by the host fallback;
  * selection / ordering / packing and the device "%g" text against the oracle (oracle/ucg_io_oracle.py) at 32 k sites;
  * `read_dump` results bit-for-bit equal to the reference's read_dump.cpp on the same files;
  * at the full 1 M-site size: dump -> read_dump -> dump round trip reproduces the file.
The device order is scrambled by a neighbor build before every check (sites are cell-sorted on the device)."""
import os
import sys

import numpy as np
import pytest

import decks
import io_cases as IC
from test_io_oracle import oracle_dump, read_words

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ucg_io_oracle as IO  # noqa: E402

pytestmark = pytest.mark.gpu


def make_ctx(pkg, fixtures, liq, dyn, build=True):
    from lammps_ucg_dev_b200 import engine
    ctx = pkg.Context(0)
    ctx.set_units(1.0, 1.0, 1.0)
    ctx.set_box(liq.box_lo, liq.box_hi)
    ctx.set_timestep(0.005)
    M = decks.MIXED
    sm = engine.StateMap.create(M["n_actual"], M["n_formal"], M["n_states"], M["formal"], M["mu"])
    for ilo, ihi, jlo, jhi, nsi, nsj, keys in decks.MIXED_COEFF:
        idx = [engine.HostTable.from_file(fixtures["table1024"], k, 2.0, 1, 1024).upload(ctx) for k in keys]
        sm.coeff(ilo, ihi, jlo, jhi, nsi, nsj, idx, [2.0] * len(idx))
    sm.init()
    sm.apply(ctx, [0.0, IC.MASS[1], IC.MASS[2], IC.MASS[2]])
    ctx.set_kT(1.0)
    ctx.neigh_configure(0.3)
    ctx.atoms_upload(liq.n, x=liq.x, v=liq.v, type=liq.type, mask=liq.mask, tag=liq.tag, molecule=liq.molecule,
                     ucgstate=liq.ucgstate, ucgl=liq.ucgl, ucgvl=liq.ucgvl, ucgml=liq.ucgml, ucgp=dyn["ucgp"],
                     f=dyn["f"], ucgforce=dyn["ucgforce"], ucgsoftmaxscores=dyn["scores"])
    if build:
        ctx.neigh_build()   # cell-sorts the sites: the device order now differs from the host order
    return ctx


def product_dump(pkg, ctx, path, group, cols, modify, step, dt=0.005):
    from lammps_ucg_dev_b200 import dumpio
    d = dumpio.DumpCustom(ctx, "dump d %s custom 1 %s %s" % (group, path, cols), groupbit=IC.GROUP_BIT if group == "half" else 1)
    for cid, (grp, bit, names) in IC.COMPUTES.items():
        d.bind_compute("compute %s %s property/atom %s" % (cid, grp, " ".join(names)), groupbit=bit)
    for m in modify:
        d.modify(m)
    d.write(step, time=step * dt)
    st = d.stats()
    d.close()
    return st


@pytest.mark.parametrize("device_format", [1, 0])
@pytest.mark.parametrize("name", sorted(IC.DUMP_CASES))
def test_dump_files_equal_the_reference(pkg, fixtures, tmp_path, monkeypatch, name, device_format):
    from lammps_ucg_dev_b200 import synth
    monkeypatch.setenv("UCGB200_DUMP_DEVICE_FORMAT", str(device_format))
    liq, dyn = IC.make_state(synth)
    ctx = make_ctx(pkg, fixtures, liq, dyn)
    group, cols, modify, step = IC.DUMP_CASES[name]
    p = str(tmp_path / (name + ".dump"))
    product_dump(pkg, ctx, p, group, cols, modify, step)
    want = open(os.path.join(IC.GOLDEN, name + ".dump"), "rb").read()
    assert open(p, "rb").read() == want


def test_pack_count_and_text_against_the_oracle_32k(pkg, fixtures):
    from lammps_ucg_dev_b200 import dumpio, synth
    liq, dyn = IC.make_state(synth, ncell=20)
    ctx = make_ctx(pkg, fixtures, liq, dyn)
    a = IC.atoms_dict(liq, dyn)
    compute = {"p": (IC.GROUP_BIT, ["ucgforce", "ucgvl", "ucgml", "ucgstate", "ucgl", "ucgp"])}
    names = ["id", "mol", "type", "mass", "x", "y", "z", "xs", "ys", "zs", "vx", "vy", "vz", "fx", "fy", "fz", "ucgstate", "ucgl",
             "ucgp", "c_p[1]", "c_p[2]", "c_p[3]", "c_p[4]", "c_p[5]", "c_p[6]"]
    codes = names[:19] + ["p_ucgforce", "p_ucgvl", "p_ucgml", "p_ucgstate", "p_ucgl", "p_ucgp"]
    bits = [-1] * 19 + [IC.GROUP_BIT] * 6
    for groupbit, thresh, sort_id in ((1, (), False), (1, (), True), (IC.GROUP_BIT, (("ucgl", ">=", 0.5), ("fx", "<", 0.0)), True),
                                      (1, (("ucgstate", "==", 0.0), ("zs", ">", 0.25), ("mol", "!=", 17.0)), False)):
        want = IO.pack(a, liq.box_lo, liq.box_hi, IC.MASS, names, groupbit=groupbit, thresh=thresh, sort_id=sort_id, compute=compute)
        kw = dict(groupbit=groupbit, col_groupbit=bits, thresh=thresh, order=dumpio.ORDER_ID if sort_id else dumpio.ORDER_INDEX)
        assert dumpio.dump_count(ctx, codes, **kw) == len(want)
        got = dumpio.dump_pack(ctx, codes, **kw)
        assert got.shape == want.shape and np.array_equal(got, want)          # bit-exact, incl. xs ys zs
        assert dumpio.dump_text(ctx, codes, **kw) == IO.lines(want, names)    # %d / %g on the device


def test_device_g_format_on_extreme_values(pkg, fixtures):
    """every finite double class through the device formatter: random bit patterns, subnormals, exact ties,
    boundary neighbours — the text must be what the C library prints"""
    from lammps_ucg_dev_b200 import dumpio, synth
    liq, dyn = IC.make_state(synth, ncell=16)
    n = liq.n
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 2 ** 63, size=(n, 3), dtype=np.uint64) | (rng.integers(0, 2, size=(n, 3), dtype=np.uint64) << np.uint64(63))
    f = bits.view(np.float64).copy()
    f[~np.isfinite(f)] = 0.0
    k = rng.integers(100000, 1000000, n).astype(np.float64)
    v = np.stack([(k + 0.5) * 10.0 ** rng.integers(-12, 12, n), np.ldexp(2 * rng.integers(0, 4000000, n) + 1.0, rng.integers(-21, 19, n)),
                  np.nextafter((k + 0.5) * 1e-3, rng.choice([-np.inf, np.inf], n))], axis=1)
    v[:8, 0] = [0.0, -0.0, 5e-324, 1.7976931348623157e308, np.inf, -np.inf, 999999.5, 1000005.0]
    dyn["f"], liq.v = f, v
    ctx = make_ctx(pkg, fixtures, liq, dyn)
    cols = ["fx", "fy", "fz", "vx", "vy", "vz"]
    buf = np.concatenate([f, v], axis=1)
    assert dumpio.dump_text(ctx, cols) == IO.lines(buf, cols)


@pytest.mark.parametrize("device_parse", [1, 0])
@pytest.mark.parametrize("name", sorted(IC.READ_CASES))
def test_read_dump_equals_the_reference(pkg, fixtures, monkeypatch, name, device_parse):
    from lammps_ucg_dev_b200 import dumpio, synth
    monkeypatch.setenv("UCGB200_READ_DUMP_DEVICE_PARSE", str(device_parse))
    liq2, dyn2 = IC.second_state(synth)
    ctx = make_ctx(pkg, fixtures, liq2, dyn2)
    gold = np.load(os.path.join(IC.GOLDEN, "read_dump_results.npz"))
    words = IC.READ_CASES[name][3]
    stats = dumpio.read_dump(ctx, "read_dump %s %s" % (os.path.join(IC.GOLDEN, "read_" + name + ".dump"), words))
    a = ctx.atoms_download(["x", "v", "f", "tag", "type", "ucgstate", "ucgl", "ucgvl", "ucgp", "ucgforce"])
    for k, v in a.items():
        assert np.array_equal(v, gold[name + "/" + k]), (name, k)
    lo, hi, per = np.zeros(3), np.zeros(3), np.zeros(3, np.int32)
    ctx._l.ucgb200_get_box(ctx._h, lo.ctypes.data_as(dumpio._dp), hi.ctypes.data_as(dumpio._dp), per.ctypes.data_as(dumpio._ip))
    assert np.array_equal(np.stack([lo, hi]), gold[name + "/box"])
    assert stats["after"] == len(gold[name + "/tag"]) and stats["before"] == liq2.n
    # the oracle's counts are the reference's log lines
    kw = read_words(words)
    _, _, _, ostats = IO.read_dump(IC.atoms_dict(liq2, dyn2), liq2.box_lo, liq2.box_hi,
                                   os.path.join(IC.GOLDEN, "read_" + name + ".dump"), **kw)
    assert stats == ostats
    # the context is usable afterwards: the list is rebuilt for the new positions / box
    ctx.neigh_build()
    assert ctx.natoms()[0] == stats["after"]


def test_device_snapshot_parse_equals_strtod(pkg, fixtures, tmp_path, monkeypatch):
    """every number spelling through read_dump's device converter: "%g", "%.17g" (longer than 2^53: converted by the
    host for that row), fixed, exponents beyond e+-22, signs, leading '+', '.5', '5.', inf — the state must equal
    float(token) bit for bit, and equal the host-only path"""
    from lammps_ucg_dev_b200 import dumpio, synth
    liq, dyn = IC.make_state(synth, ncell=8)
    n = liq.n
    rng = np.random.default_rng(17)
    spell = ["%g", "%.17g", "%.3f", "%.10e", "%+.6g", "%.15g", "%25.16e"]
    vals = rng.normal(size=(n, 4)) * 10.0 ** rng.integers(-30, 31, size=(n, 4))
    vals[:, 3] = rng.uniform(0, 1, n)
    special = ["0", "-0", ".5", "5.", "+1.5e+3", "1E5", "123456789012345678", "0.000000000000000000000001", "1e-320", "inf",
               "9007199254740993", "1e22", "1e23", "8.5e-23", "00012.50", "4.9406564584124654e-324"]
    rows, want = [], np.zeros((n, 4))
    for i in range(n):
        toks = []
        for j in range(4):
            t = (special[(i * 4 + j) % len(special)] if (i + j) % 7 == 0 and j < 3 else spell[(i + j) % len(spell)] % vals[i, j]).strip()
            toks.append(t)
            want[i, j] = float(t)
        rows.append("%d %s junk\n" % (liq.tag[i], " ".join(toks)))
    p = str(tmp_path / "hand.dump")
    with open(p, "w") as f:
        f.write("ITEM: TIMESTEP\n3\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n" % n)
        for d in range(3):
            f.write("%.17g %.17g\n" % (liq.box_lo[d], liq.box_hi[d]))
        f.write("ITEM: ATOMS id vx vy vz ucgl extra\n" + "".join(rows))
    got = []
    for mode in (1, 0):
        monkeypatch.setenv("UCGB200_READ_DUMP_DEVICE_PARSE", str(mode))
        ctx = make_ctx(pkg, fixtures, liq, dyn)
        st = dumpio.read_dump(ctx, "read_dump %s 3 vx vy vz ucgl box no" % p)
        assert st["replaced"] == n
        a = ctx.atoms_download(["v", "ucgl"])
        got.append(np.concatenate([a["v"], a["ucgl"][:, None]], axis=1))
    assert np.array_equal(got[0], got[1], equal_nan=True)
    assert np.array_equal(got[0].view(np.uint64), want.view(np.uint64))


def test_error_texts(pkg, fixtures, tmp_path):
    from lammps_ucg_dev_b200 import dumpio, synth
    liq, dyn = IC.make_state(synth)
    ctx = make_ctx(pkg, fixtures, liq, dyn, build=False)
    with pytest.raises(pkg.UCGError, match="Invalid attribute radius in dump custom command"):
        dumpio.DumpCustom(ctx, "dump d all custom 1 %s id radius" % (tmp_path / "x"))
    with pytest.raises(pkg.UCGError, match="No dump custom arguments specified"):
        dumpio.DumpCustom(ctx, "dump d all custom 1 %s" % (tmp_path / "x"))
    d = dumpio.DumpCustom(ctx, "dump d all custom 1 %s id c_p[2]" % (tmp_path / "x"))
    with pytest.raises(pkg.UCGError, match="Could not find dump custom compute ID: p"):
        d.write(0)
    with pytest.raises(pkg.UCGError, match="Invalid dump_modify thresh operator"):
        d.modify("dump_modify d thresh x ~ 1.0")
    with pytest.raises(pkg.UCGError, match="Invalid keyword radius for atom style in compute property/atom command"):
        d.bind_compute("compute p all property/atom ucgl radius")
    with pytest.raises(pkg.UCGError, match="Dump file does not contain requested snapshot"):
        dumpio.read_dump(ctx, "read_dump %s 7 x y z" % os.path.join(IC.GOLDEN, "read_replace_all.dump"))
    with pytest.raises(pkg.UCGError, match="One of the requested read_dump per-atom fields not found in dump file"):
        dumpio.read_dump(ctx, "read_dump %s 5 x" % os.path.join(IC.GOLDEN, "read_ucg_only_box_no.dump"))
    with pytest.raises(pkg.UCGError, match="Duplicate fields in read_dump command"):
        dumpio.read_dump(ctx, "read_dump %s 0 x x" % os.path.join(IC.GOLDEN, "read_replace_all.dump"))
    # empty system
    e = pkg.Context(0)
    e.set_box(liq.box_lo, liq.box_hi)
    assert dumpio.dump_pack(e, ["id", "x"]).shape == (0, 2) and dumpio.dump_text(e, ["id", "x"]) == b""


def test_full_size_round_trip_1M(pkg, fixtures, tmp_path):
    """dump -> read_dump -> dump at the BASELINE config-2 size: the second file equals the first, the restored
    state equals the dumped one to the 6 digits the file carries, ids and states exactly"""
    from lammps_ucg_dev_b200 import dumpio, synth
    liq = synth.fcc_liquid(63)
    ctx = decks.gpu_single_type(pkg, liq, fixtures)
    ctx.neigh_build()
    cols = "id type x y z vx vy vz ucgstate ucgl ucgp"
    p1, p2 = str(tmp_path / "a.dump"), str(tmp_path / "b.dump")
    d = dumpio.DumpCustom(ctx, "dump d all custom 100 %s %s" % (p1, cols))
    d.modify("dump_modify d sort id")
    d.write(0)
    st = d.stats()
    d.close()
    assert st["rows"] == liq.n
    # a second context holding the same atoms with scrambled dynamic state
    liq2 = synth.fcc_liquid(63, seed=999)
    liq2.ucgl[:] = 0.0
    liq2.ucgstate[:] = 0
    ctx2 = decks.gpu_single_type(pkg, liq2, fixtures)
    ctx2.neigh_build()
    stats = dumpio.read_dump(ctx2, "read_dump %s 0 x y z vx vy vz ucgstate ucgl ucgp" % p1)
    assert stats["replaced"] == liq.n and stats["after"] == liq.n
    d2 = dumpio.DumpCustom(ctx2, "dump d all custom 100 %s %s" % (p2, cols))
    d2.modify("dump_modify d sort id")
    d2.write(0)
    d2.close()
    # the files are equal except where a coordinate printed as the upper box face ("106.327") came back wrapped to
    # the lower one ("0") — the same happens in LAMMPS; every other character is identical
    import pandas as pd
    t1 = pd.read_csv(p1, sep=" ", skiprows=9, header=None).to_numpy()
    t2 = pd.read_csv(p2, sep=" ", skiprows=9, header=None).to_numpy()
    assert open(p1, "rb").read(400).split(b"ITEM: ATOMS")[0] == open(p2, "rb").read(400).split(b"ITEM: ATOMS")[0]
    assert t1.shape == t2.shape == (liq.n, 11)
    L = liq.box_hi - liq.box_lo
    assert np.array_equal(t1[:, [0, 1, 5, 6, 7, 8, 9, 10]], t2[:, [0, 1, 5, 6, 7, 8, 9, 10]])
    dx = np.abs(t1[:, 2:5] - t2[:, 2:5])
    wrapped = dx > 0
    assert wrapped.sum() < 100 and np.all(np.abs(dx[wrapped] - np.broadcast_to(L, dx.shape)[wrapped]) < 1e-3)
    a = ctx2.atoms_download(["x", "ucgl", "ucgstate", "tag"])
    o = np.argsort(a["tag"])
    assert np.array_equal(a["tag"][o], liq.tag) and np.array_equal(a["ucgstate"][o], liq.ucgstate)
    assert np.allclose(a["ucgl"][o], liq.ucgl, rtol=1e-5, atol=1e-6)
    d = np.abs(a["x"][o] - liq.x)
    assert np.all(np.minimum(d, np.abs(d - L)) < 1e-3)


def test_resident_run_with_dumps_is_the_same_run(pkg, fixtures, tmp_path):
    """`run 45` with a dump every 10 and one every 15 steps ([stock] Output scheduling through ucgb200_host_run): the
    pieces between dump steps run with `start/stop` semantics, so the trajectory — including the Langevin temperature
    ramp — equals the uncut ucgb200_run bit for bit; each file holds the snapshots of its own steps, the last one
    formats the final state"""
    from lammps_ucg_dev_b200 import dumpio, synth
    liq = synth.fcc_liquid(8)
    deck = dict(pair_style=0, nve=1, langevin=1, t_start=1.0, t_stop=0.4, t_period=0.5, langevin_seed=4711, ucgstate=2)
    finals = []
    for cut in (False, True):
        ctx = decks.gpu_single_type(pkg, liq, fixtures)
        ctx.deck_configure(**deck)
        ctx.setup()
        if cut:
            d10 = dumpio.DumpCustom(ctx, "dump a all custom 10 %s id x y z ucgl ucgstate" % (tmp_path / "a.dump"))
            d15 = dumpio.DumpCustom(ctx, "dump b all custom 15 %s id ucgp" % (tmp_path / "b.*.dump"))
            d10.modify("dump_modify a sort id")
            d15.modify("dump_modify b pad 4")
            dumpio.run(ctx, 45, [d10, d15], dt=0.002)
            dumpio.run(ctx, 5, [d10, d15], dt=0.002)      # a second `run`: step 45 is not written twice
            d10.close()
            d15.close()
        else:
            ctx.run(45)
            ctx.run(5)
        finals.append(ctx.atoms_download(["x", "v", "ucgl", "ucgvl", "ucgp", "ucgstate", "tag"]))
    for k in finals[0]:
        assert np.array_equal(finals[0][k], finals[1][k]), k
    text = open(tmp_path / "a.dump").read()
    steps = [int(b.split("\n")[1]) for b in text.split("ITEM: TIMESTEP")[1:]]
    assert steps == [0, 10, 20, 30, 40, 50]
    assert sorted(p.name for p in tmp_path.glob("b.*.dump")) == ["b.0000.dump", "b.0015.dump", "b.0030.dump", "b.0045.dump"]
    a = finals[1]
    o = np.argsort(a["tag"])
    buf = np.stack([a["tag"][o], a["x"][o, 0], a["x"][o, 1], a["x"][o, 2], a["ucgl"][o], a["ucgstate"][o]], axis=1)
    last = text.split("ITEM: ATOMS id x y z ucgl ucgstate\n")[-1]
    assert last.encode() == IO.lines(buf, ["id", "x", "y", "z", "ucgl", "ucgstate"])
