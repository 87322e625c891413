"""Short single-GPU program for ncu: one device-formatted dump and one read_dump of it at 1M sites (dump taps)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import dumpio, engine, synth
import bench
td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
liq = synth.fcc_liquid(int(os.environ.get("NCELL", "63")))
ctx = pkg.Context(0)
engine.setup_single_type(ctx, tf, sf, tablength=4096, box=(liq.box_lo, liq.box_hi))
engine.upload_liquid(ctx, liq)
ctx.neigh_build()
p = os.path.join(td, "p.dump")
d = dumpio.DumpCustom(ctx, "dump d all custom 100 %s id type x y z vx vy vz ucgstate ucgl ucgp" % p)
d.modify("dump_modify d sort id")
d.write(0)
d.close()
print(dumpio.read_dump(ctx, "read_dump %s 0 x y z vx vy vz ucgstate ucgl ucgp" % p))
