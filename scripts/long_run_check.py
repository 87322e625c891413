"""How long does the bench deck run before a pair comes closer than the tables' inner cutoff (the reference's fatal
'Pair distance < table inner cutoff')?  Prints the kinetic temperature every 100 steps."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import engine, synth
import bench
td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
liq = synth.fcc_liquid(int(os.environ.get("NCELL", "63")))
ctx = pkg.Context(0)
engine.setup_single_type(ctx, tf, sf, tablength=bench.TABLENGTH, cut=bench.CUT, skin=bench.SKIN, dt=bench.DT, kT=1.0, box=(liq.box_lo, liq.box_hi))
engine.upload_liquid(ctx, liq)
L = bench.LANGEVIN
ctx.deck_configure(pair_style=0, nve=1, langevin=1, t_start=L["t_start"], t_stop=L["t_stop"], t_period=L["t_period"],
                   langevin_seed=L["seed"], ucgstate=2, thermo_every=100)
ctx.setup()
for blk in range(int(os.environ.get("BLOCKS", "40"))):
    try:
        ctx.run(100)
    except Exception as e:
        print("after", 100 * blk, "steps:", e, ctx.status() if hasattr(ctx, "status") else "")
        break
    t = ctx.thermo()
    print(100 * (blk + 1), "E", round(t[0], 3), "KE", round(t[7], 3), "lambdaKE", round(t[8], 3), "rebuilds", int(t[11]), flush=True)
    if (blk + 1) % 5 == 0:
        a = ctx.atoms_download(["ucgl", "ucgstate"])
        print("    lambda min/mean/max", float(a["ucgl"].min()), float(a["ucgl"].mean()), float(a["ucgl"].max()),
              "outside [0,1]:", int(((a["ucgl"] < 0) | (a["ucgl"] > 1)).sum()), "state 1:", int(a["ucgstate"].sum()), flush=True)
