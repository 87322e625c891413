"""Dump / read_dump tap timing at the 1M-site bench size: a `dump custom` of
id type x y z vx vy vz ucgstate ucgl ucgp (11 columns) written (a) with the rows formatted on the device,
(b) packed on the device and formatted by the host fallback (snprintf, what DumpCustom::convert_string does),
(c) by the reference's own dump_custom.cpp in oracle/_ref on one host core; then read_dump of that file.
Device-side stages are timed with the context synchronised on both sides."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import dumpio, engine, synth
import bench

ncell = int(os.environ.get("NCELL", "63"))
td = tempfile.mkdtemp()
tf, sf = bench.make_fixtures(td)
liq = synth.fcc_liquid(ncell)
n = liq.n
ctx = pkg.Context(0)
engine.setup_single_type(ctx, tf, sf, tablength=4096, box=(liq.box_lo, liq.box_hi))
engine.upload_liquid(ctx, liq)
ctx.neigh_build()
ctx.pair_ucgld(0, 0)
ctx.sync()
cols = "id type x y z vx vy vz ucgstate ucgl ucgp"
names = cols.split()
out = dict(sites=n, columns=names)


def wall(fn, reps=3):
    ts = []
    for _ in range(reps):
        ctx.sync(); t0 = time.perf_counter(); r = fn(); ctx.sync(); ts.append(time.perf_counter() - t0)
    return float(np.min(ts)) * 1e3, r


# device stages alone (no PCIe): selection + radix sort + pack; + format + row scan + emit
sp, keep = dumpio._spec(names, order=dumpio.ORDER_ID)
import ctypes as C
nr, nb = C.c_longlong(0), C.c_longlong(0)
ms, _ = wall(lambda: ctx._ck(ctx._l.ucgb200_dump_count(ctx._h, C.byref(sp), C.byref(nr))))
out["device_select_sort_ms"] = ms
ms, _ = wall(lambda: ctx._ck(ctx._l.ucgb200_dump_text(ctx._h, C.byref(sp), None, C.c_longlong(0), C.byref(nr), C.byref(nb))))
out["device_select_sort_pack_format_ms"] = ms
out["text_bytes"] = nb.value
out["packed_bytes"] = n * len(names) * 8
ms, buf = wall(lambda: dumpio.dump_pack(ctx, names, order=dumpio.ORDER_ID))
out["pack_to_host_ms"] = ms
ms, txt = wall(lambda: dumpio.dump_text(ctx, names, order=dumpio.ORDER_ID))
out["text_to_host_ms"] = ms

for mode, key in ((1, "dump_device_format_ms"), (0, "dump_host_format_ms")):
    os.environ["UCGB200_DUMP_DEVICE_FORMAT"] = str(mode)
    p = os.path.join(td, "d%d.dump" % mode)
    d = dumpio.DumpCustom(ctx, "dump d all custom 100 %s %s" % (p, cols))
    d.modify("dump_modify d sort id")
    d.modify("dump_modify d header no")
    ms, _ = wall(lambda: d.write(0), reps=2)   # two body-only snapshots (warm-up + timed) ...
    d.modify("dump_modify d header yes")
    d.close()
    out[key] = ms
    d = dumpio.DumpCustom(ctx, "dump d all custom 100 %s %s" % (p, cols))   # ... then the file that is compared
    d.modify("dump_modify d sort id")
    d.write(0)
    d.close()
same = open(os.path.join(td, "d1.dump"), "rb").read() == open(os.path.join(td, "d0.dump"), "rb").read()
out["device_and_host_formatted_files_identical"] = bool(same)
out["file_bytes"] = os.path.getsize(os.path.join(td, "d1.dump"))

# read_dump of that file into the same context
ms, st = wall(lambda: dumpio.read_dump(ctx, "read_dump %s 0 x y z vx vy vz ucgstate ucgl ucgp" % os.path.join(td, "d1.dump")), reps=2)
out["read_dump_ms"] = ms
out["read_dump_stats"] = st
fields = np.zeros((n, 10)); fields[:, 0] = liq.tag; fields[:, 1:4] = liq.x
ms, _ = wall(lambda: dumpio.update_by_tag(ctx, ["id", "x", "y", "z", "vx", "vy", "vz", "ucgstate", "ucgl", "ucgp"], fields))
out["update_by_tag_ms_incl_h2d"] = ms

# the reference's DumpCustom on one host core (same columns, sorted)
import ref_binding as rb
if rb.available() and not os.environ.get("NO_REF"):
    s = rb.RefSim(); s.box(liq.box_lo, liq.box_hi, 2); s.atoms(liq)
    p = os.path.join(td, "ref.dump")
    s.command("dump d all custom 100 %s %s" % (p, cols)); s.command("dump_modify d sort id")
    t0 = time.perf_counter(); s.command("dump_write d"); out["reference_dump_ms_1core"] = (time.perf_counter() - t0) * 1e3
    s.command("undump d")
    ref = open(p, "rb").read(); mine = open(os.path.join(td, "d1.dump"), "rb").read()
    # ucgp differs (reference: -1 from data_atom_post, device: after one pair evaluation it is still -1) -> compare all
    out["reference_file_identical"] = bool(ref == mine)
    t0 = time.perf_counter(); s.command("read_dump %s 0 x y z vx vy vz ucgstate ucgl ucgp" % p); out["reference_read_dump_ms_1core"] = (time.perf_counter() - t0) * 1e3
out["step_ms_for_scale"] = 0.82
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r01_dump_timing.json"), "w"), indent=1)
