"""Generate tests/golden/io/* from oracle/_ref: the reference's own patched dump_custom.cpp, read_dump.cpp,
reader.cpp and reader_native.cpp (compiled verbatim against the shim, oracle/Makefile.ref) write the dump files
and apply the read_dump commands of tests/io_cases.py.  Run in the build container:
    python tests/golden/make_golden_io.py
Inputs are regenerated deterministically by the tests (io_cases.make_state), only outputs are stored."""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.load_package()
from lammps_ucg_dev_b200 import synth  # noqa: E402
import io_cases as IC  # noqa: E402
import ref_binding as rb  # noqa: E402


def ref_session(liq, dyn):
    s = rb.RefSim()
    s.box(liq.box_lo, liq.box_hi, 2)
    s.command("group half")
    s.atoms(liq, ucgp=dyn["ucgp"])
    s.set_state(f=dyn["f"], ucgforce=dyn["ucgforce"], scores=dyn["scores"])
    s.command("mass 1 %r" % float(IC.MASS[1]))
    s.command("mass 2 %r" % float(IC.MASS[2]))
    for cid, (grp, _, names) in IC.COMPUTES.items():
        s.command("compute %s %s property/atom %s" % (cid, grp, " ".join(names)))
    return s


def ref_dump(s, path, group, cols, modify, steps):
    s.command("dump d %s custom 1 %s %s" % (group, path, cols))
    for m in modify:
        s.command("dump_modify d " + m)
    for st in steps:
        s.l.ref_set_ntimestep(s.h, int(st))
        s.command("dump_write d")
    s.command("undump d")


def main():
    os.makedirs(IC.GOLDEN, exist_ok=True)
    td = tempfile.mkdtemp()
    liq, dyn = IC.make_state(synth)
    for name, (group, cols, modify, step) in IC.DUMP_CASES.items():
        s = ref_session(liq, dyn)
        p = os.path.join(td, name + ".dump")
        ref_dump(s, p, group, cols, modify, [step])
        shutil.copy(p, os.path.join(IC.GOLDEN, name + ".dump"))
        print(name, os.path.getsize(p), "bytes")
    out = {}
    liq2, dyn2 = IC.second_state(synth)
    for name, (cols, modify, steps, words) in IC.READ_CASES.items():
        s = ref_session(liq, dyn)
        src = os.path.join(IC.GOLDEN, "read_" + name + ".dump")
        tmp = os.path.join(td, "read_" + name + ".dump")
        ref_dump(s, tmp, "all", cols, modify, steps)
        shutil.copy(tmp, src)
        t = ref_session(liq2, dyn2)
        t.command("read_dump %s %s" % (src, words))
        a = t.get_atoms()
        n = t.nlocal()
        lo, hi = np.zeros(3), np.zeros(3)
        t.l.ref_get_box(t.h, lo.ctypes.data_as(rb._dp), hi.ctypes.data_as(rb._dp))
        for k in ("x", "v", "f", "tag", "type", "ucgstate", "ucgl", "ucgvl", "ucgp", "ucgforce"):
            out[name + "/" + k] = a[k]
        out[name + "/box"] = np.stack([lo, hi])
        out[name + "/ntimestep"] = np.array(t.ntimestep())
        print(name, "nlocal", n, "box", lo, hi)
    np.savez_compressed(os.path.join(IC.GOLDEN, "read_dump_results.npz"), **out)
    shutil.rmtree(td)


if __name__ == "__main__":
    main()
