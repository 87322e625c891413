"""Kernel time of the other pair styles and of fix cluster_switch on the 1M-site liquid
(SURVEY §8 rows a11-a13, a18): ms per evaluation, algorithmic GB/s (B_pair of §8d)."""
import os, sys, tempfile, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import engine, synth
import bench
ncell = int(os.environ.get("NCELL", "63"))
td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
out = {}

def timed(ctx, fn, reps=5):
    ctx.timers(2)
    ts = []
    for _ in range(reps):
        fn()
        ts.append(ctx.last_pair_ms())
    ctx.timers(0)
    return float(np.median(ts[1:]))

liq = synth.fcc_liquid(ncell, mol_size=4)
n = liq.n
# bethe / bethe_density share the ucgld type maps
ctx = pkg.Context(0)
engine.setup_single_type(ctx, tf, sf, tablength=4096, box=(liq.box_lo, liq.box_hi))
engine.upload_liquid(ctx, liq)
ctx.neigh_build()
tot, _, _ = ctx.neigh_stats()
mf = tot / n
out["sites"] = n; out["full_neighbors_per_site"] = mf
t = timed(ctx, lambda: ctx.pair_ucgld(0, 0)); out["ucgld_fast_ms"] = t
os.environ["UCGB200_FORCE_GENERAL"] = "1"
t = timed(ctx, lambda: ctx.pair_ucgld(0, 0)); out["ucgld_general_ms"] = t
os.environ["UCGB200_FORCE_GENERAL"] = "0"
t = timed(ctx, lambda: ctx.pair_bethe(0, 0, 1, 0, 2)); out["bethe_ms"] = t
ctx.pair_bethe_density_configure([0, 1], [0, 1], [0.0, 12.0], [0.0, 1.5])
t = timed(ctx, lambda: ctx.pair_bethe_density(1, 1)); out["bethe_density_ms"] = t
del ctx
# rleucg (2 state types) + cluster_switch
import test_gpu_cluster_switch as T
ctx = pkg.Context(0)
ctx.set_units(1.0, 1.0, 1.0); ctx.set_box(liq.box_lo, liq.box_hi); ctx.set_timestep(0.002)
idx = {kw: engine.HostTable.from_file(tf, kw, 2.5, 1, 4096).upload(ctx) for kw in ("UCG_00", "UCG_01", "UCG_11")}
tabindex = np.zeros((5, 5), np.int32)
for i, j, kw in T.PAIRS:
    tabindex[i, j] = tabindex[j, i] = idx[kw]
cutsq = np.zeros((5, 5)); cutsq[1:, 1:] = 2.5 ** 2
ctx.pair_rleucg_configure(4, [0, 1, 1, 2, 3], 3, [0, 2, 1, 1], [0, 1, 0, 0], [0.0, 12.0, 0.0, 0.0], [0.0, 1.5, 0.0, 0.0],
                          [0.0, 0.3, 0.0, 0.0, 0.0], tabindex, cutsq, [0.0, 1.0, 1.0, 1.0, 1.0], 1.0)
ctx.neigh_configure(0.3)
nmol = n // 4; half = nmol // 2
liq.type[:] = 4
sw = liq.molecule > half
liq.type[sw] = np.where((liq.molecule[sw] % 3) == 0, 3, 1)
engine.upload_liquid(ctx, liq)
ctx.neigh_build()
t = timed(ctx, lambda: ctx.pair_rleucg(1, 1)); out["rleucg_ms"] = t
ctx.cluster_configure(half + 1, half, 1.08, 15123, 0.3, [1], [3], T.CONTACTS, 4)
ctx.sync(); t0 = time.perf_counter(); ncl = ctx.cluster_check(); ctx.sync(); out["cluster_check_ms"] = 1e3 * (time.perf_counter() - t0)
t0 = time.perf_counter(); att = ctx.cluster_switch(); ctx.sync(); out["cluster_switch_ms"] = 1e3 * (time.perf_counter() - t0)
out["cluster"] = dict(n_cluster=ncl, attempts=att[0], success=att[1], rounds=float(ctx.cluster_stats()[7]), molecules=nmol)
print(json.dumps(out))
