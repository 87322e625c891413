"""Deck-level throughput of the LAMMPS-facing classes on the bench deck (1M-site UCG-LD liquid, table_ucgld LINEAR 4096,
fix nve/ucgld + ucgld/langevin + ucgstate ld), driven by the serial LAMMPS-like driver (oracle/_hostdrv):
  offload   stock Verlet order, every style call moves the arrays it touches over PCIe
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
    t0 = time.perf_counter()
    s.run(nsteps, 0)
    dt = time.perf_counter() - t0
    return liq.n * nsteps / dt / 1e6, dt / nsteps * 1e3


ncell = int(os.environ.get("NCELL", "63"))
v, ms = timed(rb.HostSim, ncell, False, 10, 3)
out["offload_Matom_steps_per_s"], out["offload_ms_per_step"] = v, ms
v, ms = timed(rb.HostSim, ncell, True, 300, 20)
out["resident_Matom_steps_per_s"], out["resident_ms_per_step"] = v, ms
if rb.available():
    v, ms = timed(rb.RefSim, 20, False, 60, 5)
    out["reference_1core_32k_Matom_steps_per_s"] = v
out["sites"] = 4 * ncell ** 3
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r01_host_classes_timing.json"), "w"), indent=1)
