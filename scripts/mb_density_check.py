"""2+ GPU check (torchrun) of the density pair styles across bricks (resident NCCL driver): the extra
forward exchanges of the one-point probabilities / CV forces (csrc/comm.cu, ucg_mb_forward_scalars).
The reference's results for these styles depend on the decomposition (reactions onto ghost neighbors are
dropped, the probability force is tallied for ghost neighbors only: SURVEY Q17), so forces
cannot be compared with a one-brick run; the pair ENERGY is decomposition-invariant and depends on the
ghost probabilities, so it must equal the one-brick energy."""
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
td = tempfile.mkdtemp()
tf, sf = B.make_fixtures(td)
liq = synth.fcc_liquid_brick(ncell, grid, rank)
parts = [None] * world
dist.all_gather_object(parts, dict(x=liq.x, v=liq.v, type=liq.type, mask=liq.mask, tag=liq.tag, molecule=liq.molecule,
                                   ucgstate=liq.ucgstate, ucgl=liq.ucgl, ucgvl=liq.ucgvl, ucgml=liq.ucgml))
cat = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}

def configure(ctx, style, brick):
    if style == "bethe_density":
        engine.setup_single_type(ctx, tf, sf, tablength=B.TABLENGTH, cut=B.CUT, skin=B.SKIN, dt=B.DT, kT=1.0, box=(liq.box_lo, liq.box_hi))
        ctx.pair_bethe_density_configure([0, 1], [0, 1], [0.0, 12.0], [0.0, 1.5])
        ps = 3
    else:
        ctx.set_units(1.0, 1.0, 1.0); ctx.set_box(liq.box_lo, liq.box_hi); ctx.set_timestep(B.DT)
        idx = [engine.HostTable.from_file(tf, k, 2.5, 1, 4096).upload(ctx) for k in ("UCG_00", "UCG_01", "UCG_11")]
        tabindex = np.zeros((3, 3), np.int32)
        tabindex[1, 1], tabindex[1, 2], tabindex[2, 1], tabindex[2, 2] = idx[0], idx[1], idx[1], idx[2]
        cutsq = np.zeros((3, 3)); cutsq[1:, 1:] = 2.5 ** 2
        ctx.pair_rleucg_configure(2, [0, 1, 1], 1, [0, 2], [0, 1], [0.0, 12.0], [0.0, 1.5], [0.0, 0.3, 0.0], tabindex, cutsq, [0.0, 1.0, 1.0], 1.0)
        ctx.neigh_configure(B.SKIN)
        ps = 2
    if brick:
        ctx.halo_configure(rank, world, grid)
        engine.upload_liquid(ctx, liq)
    else:
        ctx.atoms_upload(len(cat["tag"]), ucgp=np.full(len(cat["tag"]), -1.0), **cat)
    ctx.deck_configure(pair_style=ps, nve=1, thermo_every=1)

ok = True
for style in ("bethe_density", "rleucg"):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    b = pkg.Context(local, stream=stream.cuda_stream)
    configure(b, style, True)
    ids = [pkg.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    b.comm_init(ids[0])
    b.setup()
    e0, _ = b.pair_energy_virial()
    b.run(12)
    e1, _ = b.pair_energy_virial()
    fin = bool(np.isfinite(b.atoms_download(["f"])["f"]).all())
    et = torch.tensor([e0, e1, 0.0 if fin else 1.0, float(b.status()[0])], dtype=torch.float64, device=device)
    dist.all_reduce(et)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    if rank == 0:
        t = pkg.Context(local, stream=stream.cuda_stream)
        configure(t, style, False)
        t.setup()
        r0, _ = t.pair_energy_virial()
        rel0 = abs(et[0].item() - r0) / abs(r0)
        print(f"mb_density_check {style}: ranks={world} E(bricks)={et[0].item():.10f} E(one brick)={r0:.10f} rel={rel0:.2e} "
              f"after 12 steps E={et[1].item():.6f} finite={et[2].item() == 0.0} status={et[3].item()}", flush=True)
        ok = ok and rel0 <= 1e-10 and et[2].item() == 0.0 and et[3].item() == 0.0
if rank == 0:
    print("mb_density_check OK" if ok else "mb_density_check FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
