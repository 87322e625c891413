"""Neighbor rebuild time at 1M sites (wall clock around ucgb200_neigh_build with the context synchronised), A/B over
UCGB200_BUILD_DEFER_KEYS."""
import os, sys, tempfile, time
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
engine.setup_single_type(ctx, tf, sf, tablength=4096, box=(liq.box_lo, liq.box_hi))
engine.upload_liquid(ctx, liq)
for mode in ("1", "0", "1", "0"):
    os.environ["UCGB200_BUILD_DEFER_KEYS"] = mode
    ts = []
    for _ in range(6):
        ctx.sync(); t0 = time.perf_counter(); ctx.neigh_build(); ctx.sync(); ts.append(time.perf_counter() - t0)
    print("defer_keys", mode, "rebuild ms", round(1e3 * float(np.median(ts[1:])), 3))
