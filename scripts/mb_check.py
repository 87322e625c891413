"""2+ GPU check (torchrun): the resident multi-brick run (NCCL exchanges issued inside libucgb200,
csrc/comm.cu) against the Python-orchestrated BrickCluster on the same bricks — positions,
velocities, forces, lambda and states after N steps with rebuilds and migrations."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
grid = multigpu.procgrid_for(world)
per = int(os.environ.get("NCELL", "12"))
ncell = (per * grid[0], per * grid[1], per * grid[2])
nsteps = int(os.environ.get("STEPS", "60"))
td = tempfile.mkdtemp()
tf, sf = B.make_fixtures(td)
liq = synth.fcc_liquid_brick(ncell, grid, rank, T=2.0)
L = B.LANGEVIN

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)     # torch's collectives and the contexts share one stream, as in bench.py

def make():
    ctx = multigpu.make_gpu_brick(pkg, local, stream=stream.cuda_stream)
    engine.setup_single_type(ctx, tf, sf, tablength=B.TABLENGTH, cut=B.CUT, skin=B.SKIN, dt=B.DT, kT=1.0, box=(liq.box_lo, liq.box_hi))
    ctx.halo_configure(rank, world, grid)
    engine.upload_liquid(ctx, liq)
    return ctx

# A: Python-orchestrated
a = make()
cl = multigpu.BrickCluster({rank: a}, multigpu.DistTransport(dist, device), multigpu.torch_alloc(device), pkg.Context.halo_record_bytes())
ml = float(liq.ucgml[0])
g1 = np.array([0.0, -ml / L["t_period"], -ml / L["t_period"]])
g2 = np.array([0.0, 1.0, 1.0]) * np.sqrt(ml) * np.sqrt(24.0 / L["t_period"] / B.DT)
cl.setup(dict(dt=B.DT, langevin=1, gfactor1=g1, gfactor2=g2, t_target=1.0, langevin_seed=L["seed"], ucgstate=1))
cl.run(nsteps)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
# B: resident (no collective of the other communicator may be in flight while ncclCommInitRank runs)
b = make()
ids = [pkg.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
b.comm_init(ids[0])
b.deck_configure(pair_style=0, nve=1, langevin=1, t_start=1.0, t_stop=1.0, t_period=L["t_period"], langevin_seed=L["seed"], ucgstate=2)
b.setup()
b.run(nsteps)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
# truth: the union of the bricks as ONE brick on rank 0 (resident, single GPU)
parts = [None] * world
dist.all_gather_object(parts, dict(x=liq.x, v=liq.v, type=liq.type, mask=liq.mask, tag=liq.tag, molecule=liq.molecule,
                                   ucgstate=liq.ucgstate, ucgl=liq.ucgl, ucgvl=liq.ucgvl, ucgml=liq.ucgml))
truth = [None]
if rank == 0:
    cat = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    t = pkg.Context(local, stream=stream.cuda_stream)
    engine.setup_single_type(t, tf, sf, tablength=B.TABLENGTH, cut=B.CUT, skin=B.SKIN, dt=B.DT, kT=1.0, box=(liq.box_lo, liq.box_hi))
    t.atoms_upload(len(cat["tag"]), ucgp=np.full(len(cat["tag"]), -1.0), **cat)
    t.deck_configure(pair_style=0, nve=1, langevin=1, t_start=1.0, t_stop=1.0, t_period=L["t_period"], langevin_seed=L["seed"], ucgstate=2)
    t.setup(); t.run(nsteps)
    gt = t.atoms_download(["x", "v", "ucgl", "tag"])
    o = np.argsort(gt["tag"])
    truth = [{k: v[o] for k, v in gt.items()}]
dist.broadcast_object_list(truth, src=0)
truth = truth[0]
def vs_truth(g):
    idx = np.searchsorted(truth["tag"], g["tag"])
    box = liq.box_hi - liq.box_lo
    dx = g["x"] - truth["x"][idx]; dx -= box * np.round(dx / box)
    return max(np.abs(dx).max(), np.abs(g["v"] - truth["v"][idx]).max(), np.abs(g["ucgl"] - truth["ucgl"][idx]).max())
fields = ["x", "v", "f", "ucgl", "ucgvl", "ucgstate", "ucgp", "tag"]
ga, gb = a.atoms_download(fields), b.atoms_download(fields)
oa, ob = np.argsort(ga["tag"]), np.argsort(gb["tag"])
ok = np.array_equal(ga["tag"][oa], gb["tag"][ob])
worst = 0.0
if ok:
    for k in fields[:-1]:
        d = np.abs(ga[k][oa].astype(np.float64) - gb[k][ob].astype(np.float64)).max()
        worst = max(worst, d)
st = b.comm_stats()
ta, tb = vs_truth(ga), vs_truth(gb)
res = torch.tensor([0.0 if ok else 1.0, worst, ta, tb], dtype=torch.float64, device=device)
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"mb_check: ranks={world} sites/rank~{liq.n} steps={nsteps} python rebuilds={cl.nrebuilds} resident rebuilds={st['rebuilds']} "
          f"owner sets equal={res[0].item() == 0.0} max|diff|={res[1].item():.3e} python-vs-1brick={res[2].item():.3e} resident-vs-1brick={res[3].item():.3e}", flush=True)
    assert res[0].item() == 0.0 and res[1].item() <= 1e-10
    assert cl.nrebuilds == st["rebuilds"] and (nsteps < 40 or st["rebuilds"] >= 3)
    print("mb_check OK", flush=True)
dist.barrier()
dist.destroy_process_group()
