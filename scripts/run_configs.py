"""BASELINE.json configs 3 and 4 at their full sizes on one B200 (resident run): throughput,
stage times and sanity properties.  Config 3: 4M sites, pair table_ucg_bethe_density + fix
nve/ucgld + fix ucgstate.  Config 4: 16M sites, pair table_rleucg_interface (4 state types) + fix
cluster_switch (molecules of 4 sites, rateFreq 10) + fix nve/ucgld/wall/hard."""
import os, sys, tempfile, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import engine, synth
import bench
which = os.environ.get("CONFIGS", "3,4").split(",")
steps = int(os.environ.get("STEPS", "20"))
td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
res = {}

def timed_run(ctx, n, nsteps):
    ctx.setup()
    ctx.run(3)
    ctx.sync()
    t0 = time.perf_counter()
    ctx.run(nsteps)
    ctx.sync()
    dt = time.perf_counter() - t0
    ctx.timers(2)
    ctx.run(4)
    tms, _ = ctx.timers(0)
    th = ctx.thermo()
    return dict(sites=n, steps=nsteps, ms_per_step=1e3 * dt / nsteps, matom_steps_per_s=n * nsteps / dt / 1e6,
                stage_ms_per_step={k: v / 4 for k, v in tms.items()}, rebuilds=int(th[11]), nghost=int(th[13]),
                status=ctx.status()[0])

if "3" in which:
    liq = synth.fcc_liquid(int(os.environ.get("N3", "100")))
    ctx = pkg.Context(0)
    engine.setup_single_type(ctx, tf, sf, tablength=4096, box=(liq.box_lo, liq.box_hi))
    ctx.pair_bethe_density_configure([0, 1], [0, 1], [0.0, 12.0], [0.0, 1.5])
    engine.upload_liquid(ctx, liq)
    ctx.deck_configure(pair_style=3, nve=1, ucgstate=1, thermo_every=0)
    r = timed_run(ctx, liq.n, steps)
    a = ctx.atoms_download(["f", "ucgp", "x"])
    p0, cvf = ctx.pair_bethe_density_priors()
    r.update(finite=bool(np.isfinite(a["f"]).all() and np.isfinite(a["x"]).all()), prior_mean=float(p0.mean()), prior_std=float(p0.std()),
             ucgp_mean=float(a["ucgp"].mean()))
    res["config3_bethe_density_4M"] = r
    del ctx, liq, a
    torch.cuda.empty_cache()

if "4" in which:
    import test_gpu_cluster_switch as T
    liq = synth.fcc_liquid(int(os.environ.get("N4", "159")), mol_size=4)
    n = liq.n
    nmol = n // 4; half = nmol // 2
    liq.type[:] = 4
    sw = liq.molecule > half
    liq.type[sw] = np.where((liq.molecule[sw] % 3) == 0, 3, 1)
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
    engine.upload_liquid(ctx, liq)
    ctx.cluster_configure(half + 1, half, 1.08, 15123, 0.3, [1], [3], T.CONTACTS, 4)
    ctx.deck_configure(pair_style=2, nve=2, thermo_every=0, cluster_freq=10)
    r = timed_run(ctx, n, steps)
    a = ctx.atoms_download(["f", "x", "type"])
    st = ctx.cluster_stats()
    r.update(finite=bool(np.isfinite(a["f"]).all() and np.isfinite(a["x"]).all()), types=np.bincount(a["type"], minlength=5).tolist(),
             cluster_stats=st.tolist(), molecules=nmol)
    res["config4_rleucg_cluster_16M"] = r
print(json.dumps(res))
