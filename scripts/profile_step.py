"""Short single-GPU run for ncu: set up the 1M-site UCG-LD deck, run a few steps and force
one neighbor rebuild so that every kernel of the step appears in the launch list."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import engine, synth
import bench

ncell = int(os.environ.get("NCELL", "63"))
steps = int(os.environ.get("STEPS", "4"))
td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
liq = synth.fcc_liquid(ncell)
ctx = pkg.Context(0)
engine.setup_single_type(ctx, tf, sf, tablength=bench.TABLENGTH, box=(liq.box_lo, liq.box_hi))
engine.upload_liquid(ctx, liq)
L = bench.LANGEVIN
ctx.deck_configure(pair_style=0, nve=1, langevin=1, t_start=L["t_start"], t_stop=L["t_stop"], t_period=L["t_period"],
                   langevin_seed=L["seed"], ucgstate=2, thermo_every=0)
ctx.setup()
ctx.run(steps)
ctx.neigh_build()
ctx.run(steps)
ctx.sync()
print("profile run ok", ctx.thermo()[:3], ctx.launch_count())
