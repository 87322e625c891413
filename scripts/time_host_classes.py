"""Deck-level throughput of the LAMMPS-facing classes on the bench deck (1M-site UCG-LD liquid, table_ucgld LINEAR 4096,
fix nve/ucgld + ucgld/langevin + ucgstate ld), driven by the serial LAMMPS-like driver (oracle/_hostdrv):
  offload   stock Verlet order; eager: every style call moves the arrays it touches over PCIe; tracked (default for
            all-UCG decks): per-field tracking, x comes down once per step
  resident  the same deck with `run_style ucg/b200` (VerletUCGB200 -> ucgb200_run_between)
and the reference's own classes (oracle/_ref) on one host core at 32 k sites."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import synth
import bench
import ref_binding as rb

td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
out = {}


def timed(cls, ncell, resident, nsteps, warm):
    liq = synth.fcc_liquid(ncell)
    s = cls.ucgld_langevin(liq, tf, sf)
    if resident:
        s.command("run_style ucg/b200")
    s.setup(0)
    s.run(warm, 0)
    s.setup(0)              # every LAMMPS `run` starts with Integrate::setup (outside the loop time)
    t0 = time.perf_counter()
    s.run(nsteps, 0)
    dt = time.perf_counter() - t0
    # the driver's own LAMMPS-style breakdown of the timed run (seconds): pair = force_clear + Pair::compute, neigh =
    # Neighbor::decide / build on the HOST, comm = host forward/reverse comm + pbc/borders, modify = the fixes
    out.setdefault("breakdown_ms_per_step", {})[("resident" if resident else "offload_" + os.environ.get("UCGB200_OFFLOAD_TRACKED", "1")) if cls is rb.HostSim else "reference"] = \
        {k: 1e3 * v / nsteps for k, v in s.timers().items()}
    return liq.n * nsteps / dt / 1e6, dt / nsteps * 1e3


ncell = int(os.environ.get("NCELL", "63"))
os.environ["UCGB200_OFFLOAD_TRACKED"] = "0"
v, ms = timed(rb.HostSim, ncell, False, 10, 3)
out["offload_eager_Matom_steps_per_s"], out["offload_eager_ms_per_step"] = v, ms
os.environ["UCGB200_OFFLOAD_TRACKED"] = "1"
v, ms = timed(rb.HostSim, ncell, False, 60, 5)
out["offload_Matom_steps_per_s"], out["offload_ms_per_step"] = v, ms
v, ms = timed(rb.HostSim, ncell, True, 300, 20)
out["resident_Matom_steps_per_s"], out["resident_ms_per_step"] = v, ms
if rb.available():
    v, ms = timed(rb.RefSim, 20, False, 60, 5)
    out["reference_1core_32k_Matom_steps_per_s"] = v
out["sites"] = 4 * ncell ** 3
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_host_classes_timing.json"), "w"), indent=1)
