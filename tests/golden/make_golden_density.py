"""Generate tests/golden/ucg_ref_golden_density.npz from oracle/_ref: pair_style table_rleucg_interface,
table_ucg_bethe_density (the reference compiled with the documented repair, oracle/repair_bethe_density.py)
and fix cluster_switch.  Run in the build container:   python tests/golden/make_golden_density.py
Inputs are regenerated deterministically by the tests (synth + the deck builders in
tests/test_gpu_rleucg.py, test_gpu_bethe_density.py, test_gpu_cluster_switch.py), so only outputs are stored."""
import os
import pathlib
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
import test_gpu_cluster_switch as TC  # noqa: E402

NCELL = 5


def deck(liq, lines, ntypes=2):
    s = rb.RefSim()
    s.box(liq.box_lo, liq.box_hi, ntypes)
    s.atoms(liq)
    for c in lines:
        s.command(c)
    return s


def main():
    td = pathlib.Path(tempfile.mkdtemp())
    t = synth.write_table_file(str(td / "t.table"), npts=4096)
    out = {}
    # --- rleucg, single evaluation (eflag on: Q16)
    (td / "rle.conf").write_text("1 2\n2 density use_entropy\n12.0 1.5\n0.3\n")
    liq = synth.fcc_liquid(NCELL)
    s = deck(liq, ["newton off", "neighbor 0.3 bin", "timestep 0.002", f"pair_style table_rleucg_interface linear 4096 {td}/rle.conf",
                   f"pair_coeff 1 1 {t} UCG_00 2.5", f"pair_coeff 1 2 {t} UCG_01 2.5", f"pair_coeff 2 2 {t} UCG_11 2.5",
                   "fix 0 all ttarget/stub 1.0"])
    s.compute_once(1)
    out["rleucg_f"] = s.get_atoms()["f"]
    out["rleucg_E"] = np.array(s.eng_vdwl())
    out["rleucg_virial"] = s.virial()[0]
    # --- bethe_density, single evaluation
    (td / "bd.conf").write_text("1 2 2\n1 2\n1 2 density entropy \n12.0 1.5\n0.0 0.5\n")
    liq = synth.fcc_liquid(NCELL)
    s = deck(liq, ["newton off", "neighbor 0.3 bin", "timestep 0.002", f"pair_style table_ucg_bethe_density linear 4096 {td}/bd.conf",
                   f"pair_coeff 1 1 2 2 {t} UCG_00 2.5 {t} UCG_01 2.5 {t} UCG_01 2.5 {t} UCG_11 2.5", "fix 0 all ttarget/stub 1.0"])
    s.compute_once(1)
    a = s.get_atoms()
    out["bethe_density_f"] = a["f"]
    out["bethe_density_ucgp"] = a["ucgp"]
    out["bethe_density_E"] = np.array(s.eng_vdwl())
    out["bethe_density_virial"] = s.virial()[0]
    # --- cluster_switch trajectory (deck of tests/test_gpu_cluster_switch.py)
    liq, half = TC._system(NCELL + 1)
    s = TC._ref(liq, half, td, {"table4096": t}, 1.08, 5, 15123, 0.3)
    cwd = os.getcwd()
    os.chdir(td)
    try:
        s.setup(1)
        s.run(13, 1)
    finally:
        os.chdir(cwd)
    a = s.get_atoms()
    out["cluster_type"] = a["type"]
    out["cluster_x"] = a["x"]
    out["cluster_stats"] = np.array([s.fix_vector(2, k) for k in range(7)])
    np.savez_compressed(os.path.join(HERE, "ucg_ref_golden_density.npz"), **out)
    print("wrote ucg_ref_golden_density.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
