"""Pins oracle/ucg_io_oracle.py (the numpy restatement of the reference's dump / read_dump path) against the
golden files the reference's own dump_custom.cpp / read_dump.cpp / reader_native.cpp produced
(tests/golden/io/, generator tests/golden/make_golden_io.py), and checks the host-side data-file parser.
CPU only."""
import os
import sys

import numpy as np
import pytest

import io_cases as IC

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ucg_io_oracle as IO  # noqa: E402


def parse_modify(lines):
    """dump_modify words -> keyword arguments of the oracle"""
    import shlex
    kw = dict(thresh=[], sort_id=False, fmt={}, time=None, units=None)
    for line in lines:
        w = shlex.split(line)
        i = 0
        while i < len(w):
            if w[i] == "sort": kw["sort_id"] = w[i + 1] == "id"; i += 2
            elif w[i] == "thresh": kw["thresh"].append((w[i + 1], w[i + 2], float(w[i + 3]))); i += 4
            elif w[i] == "time": kw["time"] = w[i + 1] == "yes"; i += 2
            elif w[i] == "units": kw["units"] = w[i + 1] == "yes"; i += 2
            elif w[i] == "format":
                if w[i + 1] == "line": kw["fmt"]["line"] = w[i + 2]
                elif w[i + 1] == "int": kw["fmt"]["fint"] = w[i + 2]
                elif w[i + 1] == "float": kw["fmt"]["ffloat"] = w[i + 2]
                else: kw["fmt"].setdefault("percol", {})[int(w[i + 1]) - 1] = w[i + 2]
                i += 3
            else: raise ValueError(w[i])
    return kw


def oracle_dump(liq, dyn, group, cols, modify, step, dt=0.005):
    kw = parse_modify(modify)
    a = IC.atoms_dict(liq, dyn)
    cols = cols.split()
    compute = {cid: (bit, names) for cid, (_, bit, names) in IC.COMPUTES.items()}
    buf = IO.pack(a, liq.box_lo, liq.box_hi, IC.MASS, cols, groupbit=IC.GROUP_BIT if group == "half" else 1,
                  thresh=kw["thresh"], sort_id=kw["sort_id"], compute=compute)
    head = IO.header(step, len(buf), liq.box_lo, liq.box_hi, cols, time=step * dt if kw["time"] else None,
                     units="lj" if kw["units"] else None)
    return head + IO.lines(buf, cols, **kw["fmt"])


@pytest.mark.parametrize("name", sorted(IC.DUMP_CASES))
def test_oracle_dump_equals_reference_file(pkg, name):
    from lammps_ucg_dev_b200 import synth
    liq, dyn = IC.make_state(synth)
    group, cols, modify, step = IC.DUMP_CASES[name]
    want = open(os.path.join(IC.GOLDEN, name + ".dump"), "rb").read()
    assert oracle_dump(liq, dyn, group, cols, modify, step) == want


def read_words(words):
    w = words.split()
    kw = dict(nstep=int(w[0]), fields=[], box=True, replace=True, trim=False)
    i = 1
    while i < len(w) and w[i] not in ("box", "replace", "trim"):
        kw["fields"].append(w[i]); i += 1
    while i < len(w):
        kw[w[i]] = w[i + 1] == "yes"; i += 2
    return kw


@pytest.mark.parametrize("name", sorted(IC.READ_CASES))
def test_oracle_read_dump_equals_reference(pkg, name):
    from lammps_ucg_dev_b200 import synth
    liq2, dyn2 = IC.second_state(synth)
    gold = np.load(os.path.join(IC.GOLDEN, "read_dump_results.npz"))
    kw = read_words(IC.READ_CASES[name][3])
    a, lo, hi, stats = IO.read_dump(IC.atoms_dict(liq2, dyn2), liq2.box_lo, liq2.box_hi,
                                    os.path.join(IC.GOLDEN, "read_" + name + ".dump"), **kw)
    for k in ("x", "v", "f", "tag", "type", "ucgstate", "ucgl", "ucgvl", "ucgp", "ucgforce"):
        assert np.array_equal(a[k], gold[name + "/" + k]), (name, k)
    assert np.array_equal(np.stack([lo, hi]), gold[name + "/box"])
    assert stats["after"] == len(gold[name + "/tag"])


def test_data_file_parser_applies_data_atom_post(pkg, tmp_path):
    """read_data columns of fields_data_atom / fields_data_vel + AtomVecUCG::data_atom_post (atom_vec_ucg.cpp:85-90,
    145-170), against the oracle restatement and — when oracle/_ref is built — the reference's own data_atom_post"""
    from lammps_ucg_dev_b200 import dumpio, synth
    liq, dyn = IC.make_state(synth)
    rng = np.random.default_rng(3)
    n = liq.n
    raw_l = rng.uniform(-0.3, 1.3, n)
    raw_s = rng.integers(-1, 4, n).astype(np.int32)
    x_out = liq.x + rng.integers(-2, 3, (n, 3)) * (liq.box_hi - liq.box_lo)   # some atoms outside the box
    img = rng.integers(-1, 2, (n, 3))
    q = rng.normal(size=n)
    p = str(tmp_path / "ucg.data")
    IO.write_data_file(p, liq.box_lo, liq.box_hi, 2, IC.MASS, liq.tag, liq.molecule, liq.type, q, x_out, raw_s, raw_l, liq.ucgml,
                       v=liq.v, ucgvl=liq.ucgvl, image=img)
    d = dumpio.DataFile(p)
    assert d.natoms == n and d.ntypes == 2
    assert np.array_equal(d.box_lo, liq.box_lo) and np.array_equal(d.box_hi, liq.box_hi)
    a = d.arrays()
    s, l, pp = IO.data_atom_post(raw_s, raw_l)
    assert np.array_equal(a["ucgstate"], s) and np.array_equal(a["ucgl"], l) and np.array_equal(a["ucgp"], pp)
    for k, want in (("tag", liq.tag), ("molecule", liq.molecule), ("type", liq.type), ("v", liq.v), ("ucgvl", liq.ucgvl),
                    ("ucgml", liq.ucgml), ("q", q)):
        assert np.array_equal(a[k], want), k
    assert np.array_equal(a["mass"][1:], IC.MASS[1:])
    # wrapped into the box, image flags count the periods
    assert np.all(a["x"] >= liq.box_lo) and np.all(a["x"] < liq.box_hi)
    assert np.allclose(a["x"], np.mod(x_out - liq.box_lo, liq.box_hi - liq.box_lo) + liq.box_lo, atol=1e-12)
    ix = (a["image"] & 1023) - 512
    assert np.array_equal(ix, img[:, 0] + np.round((x_out[:, 0] - a["x"][:, 0]) / (liq.box_hi[0] - liq.box_lo[0])).astype(int))
    import ref_binding as rb
    if rb.available():
        liq.ucgstate, liq.ucgl = raw_s, raw_l
        r = rb.RefSim()
        r.box(liq.box_lo, liq.box_hi, 2)
        r.atoms(liq)
        ra = r.get_atoms()
        assert np.array_equal(ra["ucgstate"], a["ucgstate"]) and np.array_equal(ra["ucgl"], a["ucgl"])
        assert np.array_equal(ra["ucgp"], a["ucgp"])


def test_data_file_errors(pkg, tmp_path):
    from lammps_ucg_dev_b200 import dumpio
    p = tmp_path / "bad.data"
    p.write_text("t\n\n1 atoms\n1 atom types\n0 1 xlo xhi\n0 1 ylo yhi\n0 1 zlo zhi\n\nAtoms # ucg\n\n1 1 2 0 0.5 0.5 0.5 0 0.5 1\n")
    with pytest.raises(pkg.UCGError, match="Invalid atom type in Atoms section"):
        dumpio.DataFile(str(p))
    p.write_text("t\n\n1 atoms\n1 atom types\n0 1 xlo xhi\n0 1 ylo yhi\n0 1 zlo zhi\n\nAtoms # ucg\n\n1 1 1 0 0.5 0.5 0.5 0 0.5\n")
    with pytest.raises(pkg.UCGError, match="Incorrect format in Atoms section"):
        dumpio.DataFile(str(p))
    with pytest.raises(pkg.UCGError, match="Cannot open file"):
        dumpio.DataFile(str(tmp_path / "missing.data"))


def test_dump_command_grammar_without_a_device(pkg, tmp_path):
    """`dump` / `dump_modify` / `compute property/atom` words are parsed on the host: the reference's messages
    (dump_custom.cpp:67, 1843, 2005, 2293; [stock] Dump::modify_params), no GPU involved"""
    import ctypes as C
    from lammps_ucg_dev_b200 import dumpio
    l = pkg.lib()

    def create(line):
        w = line.split()
        h, err = C.c_void_p(), C.create_string_buffer(512)
        rc = l.ucgb200_host_dump_create(len(w), dumpio._argv(w), 1, C.byref(h), err, 512)
        return rc, h, err.value.decode()

    def modify(h, line):
        import shlex
        w = shlex.split(line)
        err = C.create_string_buffer(512)
        return l.ucgb200_host_dump_modify(h, len(w), dumpio._argv(w), err, 512), err.value.decode()

    f = tmp_path / "x.dump"
    for line, msg in ((f"d all custom 10 {f}", "No dump custom arguments specified"),
                      (f"d all custom 0 {f} id", "output frequency must be > 0"),
                      (f"d all atom 10 {f} id", "Unrecognized dump style 'atom'"),
                      (f"d all custom 10 {f} id radius", "Invalid attribute radius in dump custom command"),
                      (f"d all custom 10 {f} id c_p[0]", "Invalid attribute c_p[0] in dump custom command"),
                      (f"d all custom 10 {tmp_path}/x.bin id", "binary files are not supported")):
        rc, h, err = create(line)
        assert rc != 0 and msg in err, (line, err)
    rc, h, err = create(f"d all custom 10 {f} id type x ucgstate ucgl ucgp c_p[2] ucgforce")
    assert rc == 0, err
    for line, msg in (("thresh x ~ 1.0", "Invalid dump_modify thresh operator"), ("thresh radius > 1.0", "Invalid dump_modify thresh attribute: radius"),
                      ("thresh x > LAST", "thresh LAST is not supported"), ("format 9 %g", "Unknown dump_modify format ID keyword: 9"),
                      ("format int %f", "Dump_modify int format does not contain d character"), ("sort 3", "sort by column is not supported"),
                      ("colour red", "Unknown dump_modify keyword: colour"), ("append", "missing argument"), ("every 0", "Illegal dump_modify command")):
        rc, err = modify(h, line)
        assert rc != 0 and msg in err, (line, err)
    for line in ("sort id thresh ucgl >= 0.5 thresh ucgstate == 1", 'format line "%d %d %g %d %g %g %g %g"', "format float %12.6e format 3 %8.3f",
                 "format none thresh none", "append yes header no time yes units yes pad 6 flush no buffer yes every 50"):
        rc, err = modify(h, line)
        assert rc == 0, (line, err)
    names = ["ucgl", "radius"]
    err = C.create_string_buffer(512)
    assert l.ucgb200_host_dump_bind_compute(h, b"p", 1, 2, dumpio._argv(names), err, 512) != 0
    assert "Invalid keyword radius for atom style in compute property/atom command" in err.value.decode()
    assert l.ucgb200_host_dump_bind_compute(h, b"p", 1, 1, dumpio._argv(["ucgl"]), err, 512) != 0
    assert "does not calculate per-atom array" in err.value.decode()
    assert l.ucgb200_host_dump_bind_compute(h, b"p", 1, 3, dumpio._argv(["ucgforce", "ucgvl", "ucgml"]), err, 512) == 0
    n = C.c_int(0)
    l.ucgb200_host_dump_stats(h, None, None, C.byref(n))
    assert n.value == 50
    l.ucgb200_host_dump_free(h)
