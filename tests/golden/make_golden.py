"""Generate tests/golden/*.npz from oracle/_ref (the reference's own UCG/*.cpp compiled verbatim
against the LAMMPS-API shim).  Run in the build container (needs /root/reference to have built
oracle/_ref):   python tests/golden/make_golden.py
Inputs are regenerated deterministically by lammps_ucg_dev_b200.synth, so only outputs are stored.
"""
import os
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
import ref_binding as rb  # noqa: E402

NCELL = 5          # 500 sites
TABLEN = 1024
NSTEPS = 25


def fixtures(td):
    return (synth.write_table_file(os.path.join(td, "t.table"), npts=TABLEN),
            synth.write_state_file(os.path.join(td, "s.conf")))


def pack(a, keys):
    return {k: a[k] for k in keys}


def main():
    td = tempfile.mkdtemp()
    tf, sf = fixtures(td)
    out = {}
    keys1 = ("f", "ucgforce", "ucgsoftmaxscores", "num_ucgstates")
    keys2 = ("x", "v", "f", "ucgl", "ucgvl", "ucgp", "ucgstate", "ucgforce", "ucgsoftmaxscores")

    # 1. single evaluation of each pair style on the pristine liquid
    for name, pair, extra in (("ucgld", "table_ucgld", ""),
                              ("bethe", "table_ucg_bethe", "method bethe pseudo yes prior ucgl"),
                              ("bethe_sce", "table_ucg_bethe", "method bethe pseudo no prior ucgl"),
                              ("bethe_mf", "table_ucg_bethe", "method mf pseudo yes prior chemical_potential")):
        liq = synth.fcc_liquid(NCELL)
        s = rb.RefSim.single_type(liq, tf, sf, pair=pair, tablength=TABLEN, extra=extra)
        s.command("fix 0 all ttarget/stub 1.0")
        s.compute_once(1)
        a = s.get_atoms()
        for k in keys1:
            out[f"{name}_once_{k}"] = a[k]
        out[f"{name}_once_E"] = np.array(s.eng_vdwl())
        shipped, tally = s.virial()
        out[f"{name}_once_virial_shipped"] = shipped
        out[f"{name}_once_virial_tally"] = tally
        hi, hj = s.neigh_pairs()
        out[f"{name}_npairs"] = np.array(hi.size)

    # 2. deterministic trajectories (C1-like decks)
    decks = {
        "traj_c1": ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld", "fix 2 all ucgstate"],
        "traj_wall": ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld/wall/hard bias_potential 0.1", "fix 2 all ucgstate ld"],
        # sequential RanMars streams: reproducible bit-for-bit by a serial restatement only
        "traj_langevin": ["fix 1 all nve/ucgld", "fix 2 all ucgld/langevin 1.0 1.0 0.5 4711", "fix 3 all ucgstate ld"],
        "traj_mc": ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld", "fix 2 all ucgstate mc 991 0.3"],
    }
    for name, fixes in decks.items():
        liq = synth.fcc_liquid(NCELL)
        s = rb.RefSim.single_type(liq, tf, sf, tablength=TABLEN)
        for f in fixes:
            s.command(f)
        s.setup(1)
        s.run(NSTEPS, NSTEPS)
        a = s.get_atoms()
        for k in keys2:
            out[f"{name}_{k}"] = a[k]
        out[f"{name}_E"] = np.array(s.eng_vdwl())
        out[f"{name}_nbuilds"] = np.array(s.nbuilds())
        if name == "traj_langevin":
            out[f"{name}_lambda_temp"] = np.array(s.fix_scalar(1))

    # 3. table known answers through Pair::single (all four table styles)
    rsq = np.linspace(0.3, 6.2, 60)
    for style, n in (("lookup", 900), ("linear", 1000), ("spline", 800), ("bitmap", 10)):
        liq = synth.fcc_liquid(3)
        s = rb.RefSim.single_type(liq, tf, sf, tabstyle=style, tablength=n, cut=2.5)
        s.command("fix 0 all ttarget/stub 1.0")
        s.init()
        vals = np.array([[s.pair_single(i, j, r2) for r2 in rsq] for (i, j) in ((1, 1), (1, 2), (2, 2))])
        out[f"single_{style}"] = vals
    out["single_rsq"] = rsq
    np.savez_compressed(os.path.join(HERE, "ucg_ref_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "ucg_ref_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
