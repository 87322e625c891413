"""Multi-brick rebuild trace (torchrun): 1 M sites per GPU, resident driver, UCGB200_COMM_TRACE=2 / UCGB200_BUILD_TRACE=1."""
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
grid = multigpu.procgrid_for(world)
per = int(os.environ.get("NCELL", "63"))
ncell = (per * grid[0], per * grid[1], per * grid[2])
td = tempfile.mkdtemp()
tf, sf = B.make_fixtures(td)
liq = synth.fcc_liquid_brick(ncell, grid, rank)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = multigpu.make_gpu_brick(pkg, local, stream=stream.cuda_stream)
engine.setup_single_type(ctx, tf, sf, tablength=B.TABLENGTH, cut=B.CUT, skin=B.SKIN, dt=B.DT, kT=1.0, box=(liq.box_lo, liq.box_hi))
ctx.halo_configure(rank, world, grid)
engine.upload_liquid(ctx, liq)
ids = [pkg.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
ctx.comm_init(ids[0])
L = B.LANGEVIN
ctx.deck_configure(pair_style=0, nve=1, langevin=1, t_start=1.0, t_stop=1.0, t_period=L["t_period"], langevin_seed=L["seed"], ucgstate=2)
ctx.setup()
ctx.run(int(os.environ.get("STEPS", "60")))
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    print("trace run done", ctx.comm_stats(), flush=True)
dist.destroy_process_group()
