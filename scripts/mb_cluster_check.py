"""2+ GPU check (torchrun) of fix cluster_switch across bricks (resident NCCL driver): labels, states and
accept decisions are all-reduced / replicated, so the switched atom types and the MC statistics must equal
the one-brick run exactly.  Pair style: table_ucgld with three one-state types (1 = ON, 2 = OFF, 3 = inert
partner molecules), whose forces do not depend on the decomposition."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as g
pkg = g.load_package()
from lammps_ucg_dev_b200 import engine, synth, multigpu
import bench as B

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
device = torch.device("cuda", local)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
grid = multigpu.procgrid_for(world)
per = int(os.environ.get("NCELL", "10"))
ncell = (per * grid[0], per * grid[1], per * grid[2])
nsteps = int(os.environ.get("STEPS", "13"))
td = tempfile.mkdtemp()
tf, sf = B.make_fixtures(td)
liq = synth.fcc_liquid_brick(ncell, grid, rank)
ntot = 4 * ncell[0] * ncell[1] * ncell[2]
nmol = ntot // 4; half = nmol // 2
liq.molecule[:] = (liq.tag - 1) // 4 + 1
liq.type[:] = 3
sw = liq.molecule > half
liq.type[sw] = np.where((liq.molecule[sw] % 3) == 0, 2, 1)
# GROUPBIT=2: fix cluster_switch acts on a group that is not `all` (every third molecule stays outside): the group bits
# of ghosts owned by other bricks travel with the border records
GROUPBIT = int(os.environ.get("GROUPBIT", "1"))
if GROUPBIT != 1:
    liq.mask[:] = np.where(liq.molecule % 3 == 1, 1, 1 | GROUPBIT)
parts = [None] * world
dist.all_gather_object(parts, dict(x=liq.x, v=liq.v, type=liq.type, mask=liq.mask, tag=liq.tag, molecule=liq.molecule,
                                   ucgstate=liq.ucgstate, ucgl=liq.ucgl, ucgvl=liq.ucgvl, ucgml=liq.ucgml))
cat = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
CONTACTS = [(1, 1), (1, 2), (2, 1), (2, 2)]

def configure(ctx, brick):
    ctx.set_units(1.0, 1.0, 1.0); ctx.set_box(liq.box_lo, liq.box_hi); ctx.set_timestep(B.DT)
    idx = engine.HostTable.from_file(tf, "UCG_00", 2.5, 1, 4096).upload(ctx)
    ctx.set_types(3, 3, [0, 1, 1, 1], [[0, 0], [1, 0], [2, 0], [3, 0]], [0.0, 0.0, 0.0, 0.0], [0.0, 1.0, 1.0, 1.0])
    tabindex = np.full((4, 4), idx, np.int32)
    cutsq = np.zeros((4, 4)); cutsq[1:, 1:] = 2.5 ** 2
    ctx.set_pair_maps(tabindex, cutsq)
    ctx.set_kT(1.0)
    ctx.neigh_configure(B.SKIN)
    if brick:
        ctx.halo_configure(rank, world, grid)
        engine.upload_liquid(ctx, liq)
    else:
        ctx.atoms_upload(len(cat["tag"]), ucgp=np.full(len(cat["tag"]), -1.0), **cat)

b = pkg.Context(local, stream=stream.cuda_stream)
configure(b, True)
ids = [pkg.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
b.comm_init(ids[0])
b.cluster_configure(half + 1, half, 1.08, 15123, 0.3, [1], [2], CONTACTS, 3, groupbit=GROUPBIT)
b.deck_configure(pair_style=0, nve=1, thermo_every=0, cluster_freq=5)
b.setup()
b.run(nsteps)
gb = b.atoms_download(["type", "tag", "x"])
stb = b.cluster_stats()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
res = [None]
if rank == 0:
    t = pkg.Context(local, stream=stream.cuda_stream)
    configure(t, False)
    t.cluster_configure(half + 1, half, 1.08, 15123, 0.3, [1], [2], CONTACTS, 3, groupbit=GROUPBIT)
    t.deck_configure(pair_style=0, nve=1, thermo_every=0, cluster_freq=5)
    t.setup(); t.run(nsteps)
    gt = t.atoms_download(["type", "tag", "x"])
    o = np.argsort(gt["tag"])
    res = [dict(tag=gt["tag"][o], type=gt["type"][o], x=gt["x"][o], stats=t.cluster_stats(), init=cat["type"][np.argsort(cat["tag"])])]
dist.broadcast_object_list(res, src=0)
ref = res[0]
idx = np.searchsorted(ref["tag"], gb["tag"])
same = bool(np.array_equal(gb["type"], ref["type"][idx]))
box = liq.box_hi - liq.box_lo
dx = gb["x"] - ref["x"][idx]; dx -= box * np.round(dx / box)
r = torch.tensor([0.0 if same else 1.0, float(np.abs(dx).max()), 0.0 if np.array_equal(stb[:7], ref["stats"][:7]) else 1.0], dtype=torch.float64, device=device)
dist.all_reduce(r, op=dist.ReduceOp.MAX)
if rank == 0:
    changed = int((ref["type"] != ref["init"]).sum())
    print(f"mb_cluster_check: groupbit={GROUPBIT} ranks={world} molecules={nmol} types equal={r[0].item() == 0.0} switched sites={changed} "
          f"stats equal={r[2].item() == 0.0} stats={ref['stats'][:7].tolist()} max|dx|={r[1].item():.2e}", flush=True)
    print("mb_cluster_check OK" if (r[0].item() == 0.0 and r[2].item() == 0.0 and changed > 0 and r[1].item() < 1e-9) else "mb_cluster_check FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
