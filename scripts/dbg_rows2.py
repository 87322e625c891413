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
for ncell in ((5, 6, 7), 8, 12, (9,5,5)):
    liq = synth.fcc_liquid(ncell)
    pos = np.zeros((liq.n + 1, 3)); pos[liq.tag] = liq.x
    box = liq.box_hi - liq.box_lo
    for f32 in ("1", "0"):
        os.environ["UCGB200_BUILD_F32"] = f32
        ctx = pkg.Context(0)
        engine.setup_single_type(ctx, tf, sf, tablength=4096, box=(liq.box_lo, liq.box_hi))
        engine.upload_liquid(ctx, liq)
        ctx.neigh_build()
        nl = ctx.neigh_download()
        ti = np.repeat(nl["tag_i"], nl["numneigh"]); tj = nl["neigh_tags"]
        d = pos[ti] - pos[tj]; d -= box * np.round(d / box); r = np.sqrt((d * d).sum(1))
        nsorted = 0
        for row in range(len(nl["tag_i"])):
            o, nn = nl["offsets"][row], nl["numneigh"][row]
            rr = r[o:o + nn]; sk = rr[rr >= 2.5]
            nsorted += bool(np.all(np.diff(sk) >= 0))
        print(ncell, "F32", f32, "rows with ascending skin part:", nsorted, "of", len(nl["tag_i"]), "stats", ctx.neigh_stats())
