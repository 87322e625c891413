"""Golden vector of BASELINE.json configs[0] at its full length: the reference's own UCG/*.cpp (oracle/_ref) run the
32 000-site liquid through 1000 steps of `pair_style table_ucgld linear 4096` + `fix nve/ucgld` + `fix ucgstate`
(SURVEY §8d deck C1: kT = 1 from a t_target provider, neighbor 0.3 bin, dt 0.002).  Stored: the number of neighbor
rebuilds, the pair energy, and the final state of every 16th site by tag (x, v, lambda, ucgp) plus the states of all
sites.  Inputs are regenerated deterministically by lammps_ucg_dev_b200.synth.  ~2 minutes on one core:

    python tests/golden/make_golden_config0.py
"""
import os, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402
g.load_package()
from lammps_ucg_dev_b200 import synth  # noqa: E402
import ref_binding as rb  # noqa: E402

NCELL, TABLEN, NSTEPS, EVERY = 20, 4096, 1000, 16


def main():
    td = tempfile.mkdtemp()
    tf = synth.write_table_file(os.path.join(td, "t.table"), npts=TABLEN)
    sf = synth.write_state_file(os.path.join(td, "s.conf"))
    liq = synth.fcc_liquid(NCELL)
    s = rb.RefSim.single_type(liq, tf, sf, tablength=TABLEN, dt=0.002, skin=0.3)
    s.command("fix 0 all ttarget/stub 1.0")
    s.command("fix 1 all nve/ucgld")
    s.command("fix 2 all ucgstate")
    s.setup(1)
    s.run(NSTEPS, NSTEPS)
    a = s.get_atoms()
    order = np.argsort(a["tag"])
    sub = order[::EVERY]
    np.savez_compressed(os.path.join(HERE, "ucg_ref_config0_1000steps.npz"),
                        ncell=NCELL, tablength=TABLEN, nsteps=NSTEPS, every=EVERY, rebuilds=s.nbuilds(), eng_vdwl=s.eng_vdwl(),
                        tag=a["tag"][sub], x=a["x"][sub], v=a["v"][sub], ucgl=a["ucgl"][sub], ucgp=a["ucgp"][sub],
                        ucgstate_bits=np.packbits(a["ucgstate"][order].astype(np.uint8)),
                        box_lo=liq.box_lo, box_hi=liq.box_hi)
    print("rebuilds", s.nbuilds(), "E", s.eng_vdwl())


if __name__ == "__main__":
    main()
