"""CPU restatement (numpy) of the reference's dump / read_dump path for atom_style ucg.

TEST INFRASTRUCTURE ONLY — imported by tests/ (and nothing else); never by the product.
Pinned: tests/test_io_oracle.py checks every function here, byte for byte / bit for bit, against files and
arrays produced by the reference's own dump_custom.cpp / read_dump.cpp / reader_native.cpp compiled verbatim
into oracle/_ref (golden copies under tests/golden/io/, generator tests/golden/make_golden_io.py).

Each function cites the reference lines it follows (paths relative to the reference tree).
"""
from __future__ import annotations

import numpy as np

INT_COLUMNS = ("id", "mol", "type", "proc", "ucgstate")          # vtype Dump::INT, dump_custom.cpp:1485-1500, 1676
PROPERTY = ("ucgstate", "ucgl", "ucgforce", "ucgvl", "ucgp", "ucgml")   # UCG/atom_vec_ucg.cpp:172-181


def column(atoms: dict, box_lo, box_hi, mass, name: str, compute=None) -> np.ndarray:
    """one DumpCustom::pack_*() column over ALL owned atoms, as doubles.
    pack_id :2451, pack_molecule :2463, pack_type :2495, pack_mass :2507, pack_x/y/z :2528-2562,
    pack_xs/ys/zs :2606-2649 ((x - boxlo) * (1/prd)), pack_vx.. :3036, pack_fx.. :3114,
    pack_ucgstate/ucgl/ucgp :3552-3577; c_ID[k]: pack_compute :2354 over ComputePropertyAtom, whose values are
    AtomVecUCG::pack_property_atom (atom_vec_ucg.cpp:184-231): zero outside the compute's group."""
    a = atoms
    n = len(a["tag"])
    if name.startswith("c_"):
        cid, k = name[2:], 0
        if "[" in cid:
            cid, k = cid[:cid.index("[")], int(cid[cid.index("[") + 1:-1])
        groupbit, names = compute[cid]
        prop = names[0 if k == 0 else k - 1]
        src = dict(ucgstate=a["ucgstate"], ucgl=a["ucgl"], ucgforce=a["ucgforce"], ucgvl=a["ucgvl"], ucgp=a["ucgp"],
                   ucgml=a["ucgml"])[prop]
        return np.where((a["mask"] & groupbit) != 0, src.astype(np.float64), 0.0)
    if name == "id": return a["tag"].astype(np.float64)
    if name == "mol": return a["molecule"].astype(np.float64)
    if name == "type": return a["type"].astype(np.float64)
    if name == "mass": return np.asarray(mass, np.float64)[a["type"]]
    if name in ("x", "y", "z"): return a["x"][:, "xyz".index(name)].copy()
    if name in ("xs", "ys", "zs"):
        d = "xyz".index(name[0])
        inv = 1.0 / (box_hi[d] - box_lo[d])
        return (a["x"][:, d] - box_lo[d]) * inv
    if name in ("vx", "vy", "vz"): return a["v"][:, "xyz".index(name[1])].copy()
    if name in ("fx", "fy", "fz"): return a["f"][:, "xyz".index(name[1])].copy()
    if name == "ucgstate": return a["ucgstate"].astype(np.float64)
    if name == "ucgl": return a["ucgl"].copy()
    if name == "ucgp": return a["ucgp"].copy()
    if name == "proc": return np.zeros(n)
    if name == "q": return np.zeros(n)
    raise ValueError(name)


def count(atoms, box_lo, box_hi, mass, groupbit=1, thresh=()) -> np.ndarray:
    """DumpCustom::count(), dump_custom.cpp:721-1368: choose = group, then each threshold un-selects
    (`if (choose[i] && *ptr >= value) choose[i] = 0` for LT, ... :1289-1345); returns clist (local indices, ascending)."""
    choose = (atoms["mask"] & groupbit) != 0
    for attr, op, value in thresh:
        v = column(atoms, box_lo, box_hi, mass, attr)
        if op == "<": choose &= ~(v >= value)
        elif op == "<=": choose &= ~(v > value)
        elif op == ">": choose &= ~(v <= value)
        elif op == ">=": choose &= ~(v < value)
        elif op == "==": choose &= ~(v != value)
        elif op == "!=": choose &= ~(v == value)
        elif op == "|^": choose &= ~(((v == 0.0) & (value == 0.0)) | ((v != 0.0) & (value != 0.0)))
        else: raise ValueError(op)
    return np.nonzero(choose)[0]


def pack(atoms, box_lo, box_hi, mass, cols, groupbit=1, thresh=(), sort_id=False, compute=None) -> np.ndarray:
    """DumpCustom::pack(), :1372-1384, then [stock] Dump::sort by id when `dump_modify sort id`"""
    clist = count(atoms, box_lo, box_hi, mass, groupbit, thresh)
    if sort_id:
        clist = clist[np.argsort(atoms["tag"][clist], kind="stable")]
    buf = np.zeros((len(clist), len(cols)))
    for c, name in enumerate(cols):
        buf[:, c] = column(atoms, box_lo, box_hi, mass, name, compute)[clist]
    return buf


def vformats(cols, line=None, fint=None, ffloat=None, percol=None):
    """DumpCustom::init_style, :261-292: column format > int/float format > line format > default"""
    words = line.split() if line else ["%d" if c in INT_COLUMNS else "%g" for c in cols]
    out = []
    for i, c in enumerate(cols):
        if percol and percol.get(i): f = percol[i]
        elif c in INT_COLUMNS and fint: f = fint
        elif c not in INT_COLUMNS and ffloat: f = ffloat
        else: f = words[i]
        out.append(f + (" " if i + 1 < len(cols) else ""))
    return out


def lines(buf, cols, **fmt) -> bytes:
    """DumpCustom::convert_string / write_lines, :1388-1468 (Python's % operator calls the same C conversions)"""
    vf = vformats(cols, **fmt)
    isint = [c in INT_COLUMNS for c in cols]
    out = []
    for row in buf:
        out.append("".join(vf[j] % (int(row[j]) if isint[j] else row[j]) for j in range(len(cols))) + "\n")
    return "".join(out).encode()


def header(ntimestep, natoms, box_lo, box_hi, cols, periodic=(1, 1, 1), time=None, units=None) -> bytes:
    """DumpCustom::header_item, :651-672"""
    s = ""
    if units is not None: s += "ITEM: UNITS\n%s\n" % units
    if time is not None: s += "ITEM: TIME\n%.16g\n" % time
    s += "ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\n" % (ntimestep, natoms)
    bound = " ".join(("pp" if p else "ff") for p in periodic)
    s += "ITEM: BOX BOUNDS %s\n" % bound
    for d in range(3):
        s += "%1.16e %1.16e\n" % (box_lo[d], box_hi[d])
    s += "ITEM: ATOMS %s\n" % " ".join(cols)
    return s.encode()


# ------------------------------------------------------------------------------------------- read_dump
def read_snapshot(path, nstep):
    """ReaderNative::read_time / skip / read_header (text), reader_native.cpp:54-149, 186-290: returns
    (box_lo, box_hi, labels, rows as lists of words) of the snapshot whose timestep is nstep"""
    with open(path) as f:
        while True:
            line = f.readline()
            if not line: raise ValueError("Dump file does not contain requested snapshot")
            if line.strip() == "ITEM: UNITS": f.readline(); line = f.readline()
            if line.strip() == "ITEM: TIME": f.readline(); line = f.readline()
            assert line.strip() == "ITEM: TIMESTEP", line
            step = int(f.readline())
            f.readline()
            natoms = int(f.readline())
            f.readline()
            lo, hi = np.zeros(3), np.zeros(3)
            for d in range(3):
                w = f.readline().split()
                lo[d], hi[d] = float(w[0]), float(w[1])
            labels = f.readline().split()[2:]
            rows = [f.readline().split() for _ in range(natoms)]
            if step == nstep:
                return lo, hi, labels, rows
            if step > nstep: raise ValueError("Dump file does not contain requested snapshot")


def read_dump(atoms: dict, box_lo, box_hi, path, nstep, fields, box=True, replace=True, trim=False, periodic=(1, 1, 1)):
    """ReadDump::command -> header -> atoms -> process_atoms -> migrate_atoms_by_coords on one process
    (read_dump.cpp:80-152, 443-567, 573-666, 797-935, 1150-1163).  Returns (atoms, box_lo, box_hi, stats)."""
    a = {k: v.copy() for k, v in atoms.items()}
    slo, shi, labels, rows = read_snapshot(path, nstep)
    # ReaderNative::read_header :330-420: x -> x | xs | xu | xsu, first present column wins
    index, scaled = {}, None
    for fld in ("id",) + tuple(fields):
        if fld in ("x", "y", "z"):
            if fld in labels: index[fld], flag = labels.index(fld), "nw"
            else:
                cand = [(labels.index(fld + s), s) for s in ("s", "u", "su") if fld + s in labels]
                if not cand: raise ValueError("One of the requested read_dump per-atom fields not found in dump file")
                index[fld], flag = min(cand)
            sc = flag in ("s", "su")
            if scaled is not None and sc != scaled: raise ValueError("Read_dump xyz fields do not have consistent scaling/wrapping")
            scaled = sc
        else:
            if fld not in labels: raise ValueError("One of the requested read_dump per-atom fields not found in dump file")
            index[fld] = labels.index(fld)
    scaled = bool(scaled)
    sprd = shi - slo
    idmap = {int(t): i for i, t in reversed(list(enumerate(a["tag"])))}   # Atom::map: lowest local index wins
    n = len(a["tag"])
    updated = np.zeros(n, bool)
    nreplace = 0
    for w in rows:
        m = idmap.get(int(float(w[index["id"]])), -1)    # static_cast<tagint>(fields[i][0]) :853
        if m < 0: continue
        updated[m] = True
        if not replace: continue
        nreplace += 1
        for fld in fields:
            val = float(w[index[fld]])
            if fld in ("x", "y", "z"):
                d = "xyz".index(fld)
                a["x"][m, d] = val * sprd[d] + slo[d] if scaled else val      # xfield :1359-1380
            elif fld in ("vx", "vy", "vz"): a["v"][m, "xyz".index(fld[1])] = val
            elif fld in ("fx", "fy", "fz"): a["f"][m, "xyz".index(fld[1])] = val
            elif fld == "ucgstate": a["ucgstate"][m] = int(val)
            elif fld == "ucgl": a["ucgl"][m] = val
            elif fld == "ucgp": a["ucgp"][m] = val
    ntrim = 0
    if trim:   # :919-935: avec->copy(nlocal-1, i) into every hole
        src, flag, nlocal, i = list(range(n)), updated.tolist(), n, 0
        while i < nlocal:
            if not flag[i]:
                src[i], flag[i] = src[nlocal - 1], flag[nlocal - 1]
                nlocal -= 1
                ntrim += 1
            else:
                i += 1
        sel = np.asarray(src[:nlocal], int)
        a = {k: v[sel] for k, v in a.items()}
    lo, hi = (slo, shi) if box else (np.asarray(box_lo, float), np.asarray(box_hi, float))
    prd = hi - lo
    x = a["x"]
    for d in range(3):   # Domain::remap, one coordinate at a time
        if not periodic[d]: continue
        for i in range(len(x)):
            c = x[i, d]
            while c < lo[d]: c += prd[d]
            while c >= hi[d]: c -= prd[d]
            x[i, d] = max(c, lo[d])
    stats = dict(before=n, snapshot=len(rows), purged=0, replaced=nreplace, trimmed=ntrim, added=0, after=len(a["tag"]))
    return a, lo, hi, stats


# ------------------------------------------------------------------------------------------- read_data
def data_atom_post(ucgstate, ucgl):
    """AtomVecUCG::data_atom_post, UCG/atom_vec_ucg.cpp:145-170: returns (ucgstate, ucgl, ucgp)"""
    l = np.where(ucgl < 0, 0.0, np.where(ucgl > 1, 1.0, ucgl))
    s = np.where(ucgstate < 0, 0, np.where(ucgstate > 1, 1, ucgstate)).astype(np.int32)
    return s, l, np.full(len(l), -1.0)


def write_data_file(path, box_lo, box_hi, ntypes, mass, tag, mol, type_, q, x, state, ucgl, ucgml, v=None, ucgvl=None, image=None):
    """a data file in the column order of fields_data_atom / fields_data_vel (atom_vec_ucg.cpp:85-90)"""
    with open(path, "w") as f:
        f.write("LAMMPS data file, atom_style ucg (test fixture)\n\n%d atoms\n%d atom types\n\n" % (len(tag), ntypes))
        for d, nm in enumerate("xyz"):
            f.write("%.17g %.17g %slo %shi\n" % (box_lo[d], box_hi[d], nm, nm))
        f.write("\nMasses\n\n")
        for t in range(1, ntypes + 1):
            f.write("%d %.17g\n" % (t, mass[t]))
        f.write("\nAtoms # ucg\n\n")
        for i in range(len(tag)):
            img = "" if image is None else " %d %d %d" % tuple(image[i])
            f.write("%d %d %d %.17g %.17g %.17g %.17g %d %.17g %.17g%s\n" % (tag[i], mol[i], type_[i], q[i], x[i, 0], x[i, 1], x[i, 2],
                                                                       state[i], ucgl[i], ucgml[i], img))
        if v is not None:
            f.write("\nVelocities\n\n")
            for i in range(len(tag)):
                f.write("%d %.17g %.17g %.17g %.17g\n" % (tag[i], v[i, 0], v[i, 1], v[i, 2], ucgvl[i]))
