import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import engine, synth
import bench
td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
liq = synth.fcc_liquid((5, 6, 7))
def build(env):
    for k in ("UCGB200_BUILD_F32", "UCGB200_TILE_CAP", "UCGB200_BUILD_TILED"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = pkg.Context(0)
    engine.setup_single_type(ctx, tf, sf, tablength=4096, box=(liq.box_lo, liq.box_hi))
    engine.upload_liquid(ctx, liq)
    ctx.neigh_build()
    return ctx.neigh_download()
A = build({}); B = build({"UCGB200_BUILD_F32": "0"}); C = build({"UCGB200_BUILD_F32": "0", "UCGB200_TILE_CAP": "64"}); D = build({"UCGB200_TILE_CAP": "64"})
pos = {t: x for t, x in zip(liq.tag, liq.x)}
box = liq.box_hi - liq.box_lo
def dist(ti, tj):
    d = pos[ti] - pos[tj]; d -= box * np.round(d / box); return float(np.sqrt((d * d).sum()))
for name, X in (("F32 vs F64", B), ("F32 vs F64 chunked", C), ("F32 vs F32 chunked", D)):
    same = np.array_equal(A["neigh_tags"], X["neigh_tags"])
    print(name, "equal:", same)
    if not same:
        bad = np.nonzero(A["neigh_tags"] != X["neigh_tags"])[0]
        row = np.searchsorted(A["offsets"], bad[0], side="right") - 1
        o, nn = A["offsets"][row], A["numneigh"][row]
        ti = A["tag_i"][row]
        print(" first differing row", row, "tag", ti, "n", nn, "first diff at entry", bad[0] - o, "rows differing", len(np.unique(np.searchsorted(A["offsets"], bad, side="right") - 1)), "of", len(A["tag_i"]))
        print("  A:", [(int(t), round(dist(ti, t), 4)) for t in A["neigh_tags"][o:o + nn]][-30:])
        print("  X:", [(int(t), round(dist(ti, t), 4)) for t in X["neigh_tags"][o:o + nn]][-30:])
print("B vs C equal:", np.array_equal(B["neigh_tags"], C["neigh_tags"]))
