"""Time the pair kernel on the 1M-site liquid for every schedule variant (env knobs)."""
import os, sys, tempfile, itertools, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import engine, synth
import bench

ncell = int(os.environ.get("NCELL", "63"))
td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
liq = synth.fcc_liquid(ncell)
ctx = pkg.Context(0)
engine.setup_single_type(ctx, tf, sf, tablength=bench.TABLENGTH, box=(liq.box_lo, liq.box_hi))
engine.upload_liquid(ctx, liq)
L = bench.LANGEVIN
ctx.deck_configure(pair_style=0, nve=1, langevin=1, t_start=1.0, t_stop=1.0, t_period=1.0, langevin_seed=L["seed"], ucgstate=2)
ctx.setup()
ctx.run(20)   # a few steps so the liquid is not the pristine lattice
ctx.timers(2)
res = []
variants = json.loads(os.environ.get("VARIANTS", "null")) or [
    dict(LPA=l, BS=b, PF=p, SMEM_TABLE=s) for l in (4, 8, 16) for b in (512,) for p in (0, 1) for s in (1, 0)]
for v in variants:
    for k, val in v.items():
        os.environ["UCGB200_" + k] = str(val)
    ts = []
    for _ in range(6):
        ctx.pair_ucgld(0, 0)
        ts.append(ctx.last_pair_ms())
    res.append((float(np.median(ts[1:])), v))
    print("%8.4f ms  %s" % res[-1], flush=True)
res.sort(key=lambda r: r[0])
print("BEST", res[0])
