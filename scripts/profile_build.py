"""Neighbor rebuild alone on the 1M-site liquid (for ncu launch lists / timing)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import engine, synth
import bench
td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
liq = synth.fcc_liquid(int(os.environ.get("NCELL", "63")))
ctx = pkg.Context(0)
engine.setup_single_type(ctx, tf, sf, tablength=bench.TABLENGTH, box=(liq.box_lo, liq.box_hi))
engine.upload_liquid(ctx, liq)
ctx.deck_configure(pair_style=0, nve=1, langevin=1, t_start=1.0, t_stop=1.0, t_period=1.0, langevin_seed=7, ucgstate=2)
ctx.setup()
ctx.run(10)
for rep in range(int(os.environ.get("REPS", "4"))):
    ctx.sync(); t0 = time.perf_counter()
    ctx.neigh_build()
    ctx.sync(); print("rebuild wall ms", 1e3 * (time.perf_counter() - t0), flush=True)
