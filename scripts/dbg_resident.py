import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import synth
import ref_binding as rb
td = tempfile.mkdtemp()
fx = dict(t=synth.write_table_file(td + "/t.table", npts=4096), s=synth.write_state_file(td + "/s.conf"))
liq = synth.fcc_liquid(6)
fixes = ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld", "fix 2 all ucgstate"]
for plan, th in (([13], 0), ([25], 0), ([40], 0)):
    out = []
    sims = []
    for resident in (False, True):
        s = rb.HostSim.single_type(liq, fx["t"], fx["s"])
        for f in fixes: s.command(f)
        if resident: s.command("run_style ucg/b200")
        s.setup(1)
        snaps = [s.get_atoms()]
        for n in plan:
            s.run(n, th)
            snaps.append(s.get_atoms())
        out.append(snaps)
        sims.append(s)
    for k, (a, b) in enumerate(zip(*out)):
        if k: print("   nbuilds", [o_.nbuilds() for o_ in sims], "ndiff v", int((a["v"] != b["v"]).any(axis=1).sum()), "of", len(a["v"]))
        print(plan, th, "checkpoint", k, {f: float(np.abs(a[f] - b[f]).max()) for f in ("x", "v", "f", "ucgl", "ucgvl", "ucgp", "ucgforce")})
