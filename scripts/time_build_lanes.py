"""Neighbor rebuild time at 1M sites, lattice and running liquid, A/B over UCGB200_BUILD_LANES (lane-per-site against
warp-per-site row build); wall clock around ucgb200_neigh_build with the context synchronised, plus the stage timers
of a 200-step run (neigh ms per step)."""
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
for lanes in ("1", "0", "1", "0"):
    os.environ["UCGB200_BUILD_LANES"] = lanes
    ctx = pkg.Context(0)
    engine.setup_single_type(ctx, tf, sf, tablength=4096, box=(liq.box_lo, liq.box_hi))
    engine.upload_liquid(ctx, liq)
    L = bench.LANGEVIN
    ctx.deck_configure(pair_style=0, nve=1, langevin=1, t_start=L["t_start"], t_stop=L["t_stop"], t_period=L["t_period"],
                       langevin_seed=L["seed"], ucgstate=2, thermo_every=0)
    ts = []
    for _ in range(6):
        ctx.sync(); t0 = time.perf_counter(); ctx.neigh_build(); ctx.sync(); ts.append(time.perf_counter() - t0)
    lattice = round(1e3 * float(np.median(ts[1:])), 3)
    ctx.setup()
    ctx.run(100)                         # a running liquid: sites off the lattice, rows of uneven length
    ts = []
    for _ in range(6):
        ctx.sync(); t0 = time.perf_counter(); ctx.neigh_build(); ctx.sync(); ts.append(time.perf_counter() - t0)
    liquid = round(1e3 * float(np.median(ts[1:])), 3)
    ctx.timers(3)
    ctx.sync(); t0 = time.perf_counter(); ctx.run(200); ctx.sync(); dt = time.perf_counter() - t0
    tms, _ = ctx.timers(0)
    print("lanes", lanes, "rebuild ms: lattice", lattice, "liquid", liquid, "| 200 steps:", round(5 * dt, 4), "ms/step, stages",
          {k: round(v / 200, 4) for k, v in tms.items()}, flush=True)
    del ctx
