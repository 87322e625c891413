"""experiment: cvf / forces / per-atom virial of table_ucg_bethe_density under the three UCGB200_BD_LOGSUM modes, in the
fast path (shared-memory tables), the general path and the per-atom (PA) instantiation, against mode 0"""
import os, sys, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import engine, synth
import bench
td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
liq = synth.fcc_liquid(int(os.environ.get("NCELL", "6")))
res = {}
for path in ("fast", "general", "pa"):
    for mode in ("0", "1", "2"):
        os.environ["UCGB200_BD_LOGSUM"] = mode
        os.environ["UCGB200_FORCE_GENERAL"] = "1" if path == "general" else "0"
        ctx = pkg.Context(0)
        engine.setup_single_type(ctx, tf, sf, tablength=4096, box=(liq.box_lo, liq.box_hi))
        ctx.pair_bethe_density_configure([0, 1], [0, 1], [0.0, 12.0], [0.0, 1.5])
        engine.upload_liquid(ctx, liq)
        ctx.neigh_build()
        if path == "pa":
            ctx.pair_bethe_density(3, 7)
            e, v = ctx.pair_peratom()
        else:
            ctx.pair_bethe_density(1, 1)
            v = None
        p0, cvf = ctx.pair_bethe_density_priors()
        f = ctx.atoms_download(["f"])["f"]
        E, W = ctx.pair_energy_virial()
        res[(path, mode)] = dict(cvf=cvf, f=f, v=v, E=E, W=W)
        del ctx
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
out = {}
for path in ("fast", "general", "pa"):
    b = res[(path, "0")]
    for mode in ("1", "2"):
        a = res[(path, mode)]
        out[f"{path}_mode{mode}"] = dict(cvf=rel(a["cvf"], b["cvf"]), f=rel(a["f"], b["f"]), W=rel(a["W"], b["W"]), E=abs(a["E"] - b["E"]) / abs(b["E"]),
                                         vatom=(rel(a["v"], b["v"]) if a["v"] is not None else None))
out["general_vs_fast_mode0"] = dict(cvf=rel(res[("general", "0")]["cvf"], res[("fast", "0")]["cvf"]))
out["pa_vs_general_mode0"] = dict(cvf=rel(res[("pa", "0")]["cvf"], res[("general", "0")]["cvf"]))
out["pa_vs_general_mode1"] = dict(cvf=rel(res[("pa", "1")]["cvf"], res[("general", "1")]["cvf"]))
print(json.dumps(out, indent=1))
