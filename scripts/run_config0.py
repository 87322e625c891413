"""BASELINE.json configs[0]: the reference's own CPU-runnable case — 32 000-site 2-substate UCG liquid, pair_style
table_ucgld + fix nve/ucgld + fix ucgstate (deterministic; kT from a t_target provider), 1000 steps (SURVEY §8d deck C1).

    python scripts/run_config0.py cpu   # the reference's UCG/*.cpp (oracle/_ref), one host core; runs anywhere
    python scripts/run_config0.py gpu   # the same deck, resident on one B200 through the C-ABI

Both print one JSON object: throughput, LAMMPS-style stage breakdown, and the observables of the final state (the
trajectories themselves diverge chaotically long before step 1000; the first 25-60 steps are compared exactly in
tests/test_gpu_parity.py and tests/test_gpu_host_classes.py)."""
import os, sys, tempfile, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
import bench
mode = sys.argv[1] if len(sys.argv) > 1 else "cpu"
steps = int(os.environ.get("STEPS", "1000"))
ncell = int(os.environ.get("NCELL", "20"))
td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
pkg = g.load_package()
from lammps_ucg_dev_b200 import engine, synth
liq = synth.fcc_liquid(ncell)


def observables(a, e):
    return {"pe_per_site": e / liq.n, "state1_fraction": float(np.mean(a["ucgstate"])), "lambda_mean": float(np.mean(a["ucgl"])),
            "lambda_std": float(np.std(a["ucgl"])), "ucgp_mean": float(np.mean(a["ucgp"])),
            "ke_per_site": float(0.5 * np.sum(a["v"] ** 2) / liq.n), "finite": bool(np.isfinite(a["x"]).all() and np.isfinite(a["f"]).all())}


if mode == "cpu":
    import ref_binding as rb
    assert rb.available(), "oracle/_ref not built (python -c 'import __graft_entry__ as g; g.build()' where /root/reference exists)"
    s = rb.RefSim.single_type(liq, tf, sf, tablength=bench.TABLENGTH, dt=bench.DT, skin=bench.SKIN)
    s.command("fix 0 all ttarget/stub 1.0")
    s.command("fix 1 all nve/ucgld")
    s.command("fix 2 all ucgstate")
    s.setup(1)
    t0 = time.perf_counter()
    s.run(steps, steps)
    dt = time.perf_counter() - t0
    out = {"impl": "reference (oracle/_ref: UCG/*.cpp verbatim), one host core", "sites": liq.n, "steps": steps, "seconds": dt,
           "matom_steps_per_s": liq.n * steps / dt / 1e6, "ms_per_step": 1e3 * dt / steps, "breakdown_s": s.timers(),
           "rebuilds": s.nbuilds(), "final": observables(s.get_atoms(), s.eng_vdwl())}
else:
    import torch
    ctx = pkg.Context(0)
    engine.setup_single_type(ctx, tf, sf, tablength=bench.TABLENGTH, cut=bench.CUT, skin=bench.SKIN, dt=bench.DT, kT=1.0,
                             box=(liq.box_lo, liq.box_hi))
    engine.upload_liquid(ctx, liq)
    ctx.deck_configure(pair_style=0, nve=1, ucgstate=1, thermo_every=steps)
    ctx.setup()
    ctx.sync()
    t0 = time.perf_counter()
    ctx.run(steps)
    ctx.sync()
    dt = time.perf_counter() - t0
    th = ctx.thermo()
    a = ctx.atoms_download(["x", "v", "f", "ucgl", "ucgp", "ucgstate"])
    out = {"impl": "libucgb200 resident run (ucgb200_setup + ucgb200_run), one B200", "sites": liq.n, "steps": steps, "seconds": dt,
           "matom_steps_per_s": liq.n * steps / dt / 1e6, "ms_per_step": 1e3 * dt / steps, "rebuilds": int(th[11]),
           "final": observables(a, th[0]), "status": ctx.status()[0]}
print(json.dumps(out))
