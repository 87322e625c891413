#!/usr/bin/env python
"""bench.py — UCG-LD throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host cores

A "step" is one MD timestep of the hot path over the whole synthetic liquid:
fix nve/ucgld initial_integrate -> skin check (+ neighbor rebuild when it fires) -> ghost
refresh -> pair table_ucgld -> fix ucgld/langevin -> fix ucgstate ld -> final_integrate.
Workload at N=1: BASELINE.json configs[1] — the 1 000 188-site UCG-LD liquid (fcc n=63,
rho*=0.8442, LINEAR tables of 4096 entries, cut 2.5, skin 0.3, dt 0.002).
value = N_sites * K / t (Matom-steps/s) with everything resident in HBM; e2e = the same step
driven through the C-ABI with HOST (pinned) buffers going in and out every step.
Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "UCG-LD Matom-steps/s"
UNIT = "Matom-steps/s"
NCELL_1GPU = 63          # 4*63^3 = 1 000 188 sites ("1M")
TABLENGTH = 4096
DT = 0.002
SKIN = 0.3
CUT = 2.5
LANGEVIN = dict(t_start=1.0, t_stop=1.0, t_period=1.0, seed=48291)


def workload_config(ncell, nranks):
    sites = int(4 * np.prod(ncell)) if not np.isscalar(ncell) else int(4 * ncell ** 3)
    return {"workload": f"{sites / 1e6:.0f}M-site UCG-LD liquid, table_ucgld LINEAR 4096, fix nve/ucgld + ucgld/langevin + ucgstate ld",
            "sites": sites,
            "ncell": list(ncell) if not np.isscalar(ncell) else [ncell] * 3,
            "tablength": TABLENGTH, "cut": CUT, "skin": SKIN, "dt": DT,
            "decomposition": "1 brick" if nranks == 1 else f"{nranks} bricks, halo over NCCL",
            "l2": "inputs larger than L2 (neighbor rows ~0.4 GB + 0.1 GB site records per step; no flush needed)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def make_fixtures(td):
    import __graft_entry__ as g
    g.load_package()
    from lammps_ucg_dev_b200 import synth
    tf = synth.write_table_file(os.path.join(td, "ucg.table"), npts=TABLENGTH)
    sf = synth.write_state_file(os.path.join(td, "ucg.conf"))
    return tf, sf


# --------------------------------------------------------------------- CPU arm
def cpu_reference(ncell, steps, warmup, td):
    """The reference algorithm on the host: oracle/_ref (the reference's own UCG/*.cpp against
    the LAMMPS-API shim) when it was built, else the C restatement (oracle/).  Serial: the
    UCG package is single-threaded per MPI rank and no MPI exists on this box."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    g.load_package()
    from lammps_ucg_dev_b200 import synth
    tf, sf = make_fixtures(td)
    liq = synth.fcc_liquid(ncell)
    kind = "port"
    try:
        import ref_binding as rb
        if rb.available():
            kind = "reference"
    except Exception:
        rb = None
    if kind == "reference":
        sim = rb.RefSim.ucgld_langevin(liq, tf, sf, tablength=TABLENGTH, dt=DT, skin=SKIN, **LANGEVIN)
    else:
        import oracle_binding as ob
        sim = ob.Oracle.single_type(liq, tf, tablength=TABLENGTH, dt=DT, skin=SKIN)
        sim.fix_nve()
        sim.fix_langevin(LANGEVIN["t_start"], LANGEVIN["t_stop"], LANGEVIN["t_period"], LANGEVIN["seed"])
        sim.fix_ucgstate(mode=1)
    sim.setup()
    if warmup:
        sim.run(warmup)
    t0 = time.perf_counter()
    sim.run(steps)
    dt = time.perf_counter() - t0
    return dict(value=liq.n * steps / dt / 1e6, seconds=dt, sites=liq.n, steps=steps, kind=kind,
                breakdown=sim.timers())


def _replica(job):
    ncell, steps, warmup = job
    with tempfile.TemporaryDirectory() as td:
        return cpu_reference(ncell, steps, warmup, td)


def run_reference(args):
    """The reference's own CPU code (oracle/_ref) on the host cores, on THIS arm's workload: the 1 000 188-site liquid
    of configs[1].  The UCG package is serial per MPI rank and this box has no MPI, so "all the host threads it can
    use" is realised as one independent serial run of the whole liquid per core, all at once; the aggregate is an
    UPPER bound for an MPI run of one liquid over the same cores (no halo exchange, no load imbalance).  A bounded
    sample: `steps` capped so that the run ends within a few minutes (~1.2 s per step per core)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncell = int(os.environ.get("UCGB200_NCELL_PER_GPU", NCELL_1GPU))
    steps = max(1, min(args.steps, 40))
    warm = min(args.warmup, 2)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    cores = max(1, min(cores, 64))
    if args.serial:
        cores = 1
    import multiprocessing as mp
    t0 = time.perf_counter()
    # the truly serial figure first (one core, nothing else running): cpu_baseline.serial_1M
    if cores == 1:
        serial = _replica((ncell, steps, warm))
        rs = [serial]
    else:
        serial = _replica((ncell, max(5, min(steps, 8)), 1))
        with mp.get_context("spawn").Pool(cores) as pool:
            rs = pool.map(_replica, [(ncell, steps, warm)] * cores)
    wall = time.perf_counter() - t0
    value = sum(r["value"] for r in rs)
    slowest = max(r["seconds"] for r in rs)
    r0 = rs[0]
    cfg = workload_config(ncell, 1)
    cfg["reference_arm"] = (f"{cores} concurrent serial runs of the whole {r0['sites']}-site liquid (one per host core), "
                            f"{r0['steps']} timed steps each after setup and {warm} warm-up steps")
    if args.gpus > 1:
        cfg["reference_arm"] += (f"; launched beside the {args.gpus}-GPU arm ({args.gpus} x {r0['sites']} sites, weak scaling): the host has the "
                                 "same cores whatever N is, so this is the same CPU figure as at N = 1 (one brick's liquid per core)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": r0["steps"], "steps_requested": args.steps, "warmup": warm, "ms_per_step": 1e3 * slowest / r0["steps"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": r0["kind"],
                             "sample": f"{cores} concurrent serial runs of the same {r0['sites']}-site liquid x {r0['steps']} steps "
                                       "(one per core; no MPI on this box, so this is an upper bound for an MPI run)",
                             "per_core": [r["value"] for r in rs][:8], "wall_s": wall,
                             "breakdown_s": r0["breakdown"],
                             "serial_1M": {"value": serial["value"], "unit": UNIT, "cores": 1, "sites": serial["sites"],
                                           "steps": serial["steps"], "seconds": serial["seconds"],
                                           "breakdown_s": serial["breakdown"]}},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import __graft_entry__ as g
    pkg = g.load_package()
    from lammps_ucg_dev_b200 import engine, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the UCG hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1:
        from lammps_ucg_dev_b200 import multigpu
        return multigpu.bench(args, rank, world, local, dist)

    td = tempfile.mkdtemp()
    tf, sf = make_fixtures(td)
    # UCGB200_NCELL_PER_GPU: non-default problem size (e.g. 100 -> 4M sites per GPU, the per-GPU share of
    # BASELINE.json configs[4]); the default is the 1M-site configs[1]
    ncell1 = int(os.environ.get("UCGB200_NCELL_PER_GPU", NCELL_1GPU))
    liq = synth.fcc_liquid(ncell1)
    stream = torch.cuda.Stream()          # a real (non-default) stream: CUDA events on it bracket every kernel
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(local, stream=stream.cuda_stream)
    engine.setup_single_type(ctx, tf, sf, tablength=TABLENGTH, cut=CUT, skin=SKIN, dt=DT, kT=1.0,
                             box=(liq.box_lo, liq.box_hi))
    engine.upload_liquid(ctx, liq)
    ctx.deck_configure(pair_style=0, nve=1, langevin=1, t_start=LANGEVIN["t_start"], t_stop=LANGEVIN["t_stop"],
                       t_period=LANGEVIN["t_period"], langevin_seed=LANGEVIN["seed"], ucgstate=2, thermo_every=0)
    ctx.setup()
    ctx.run(args.warmup)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count()
    torch.cuda.synchronize()
    ev0.record(stream)
    ctx.run(args.steps)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    sampler.stop_flag.set()
    sampler.join()
    value = liq.n * args.steps / (ms * 1e-3) / 1e6
    th = ctx.thermo()

    # ---- roofline of the dominant kernel (pair table_ucgld): live CUDA-event timing
    total_full, maxrow, nbuilds = ctx.neigh_stats()
    m_half = 0.5 * total_full / liq.n
    bytes_per_site = 100.0 + 4.0 * m_half          # B_pair = 100 + 4*M
    # (a) the pair kernel alone: CUDA events around every launch, one step per call
    ctx.timers(2)
    pair_ms = []
    for _ in range(max(5, min(args.steps, 20))):
        ctx.run(1)
        pair_ms.append(ctx.last_pair_ms())
    ctx.timers(0)
    pair_avg = float(np.mean(pair_ms))
    # (b) the stages of the step: the SAME loop as the timed region (speculative launches on), event pairs recorded
    # without synchronisation and resolved afterwards (ucgb200_timers mode 3)
    nstage = min(max(args.steps, 20), 100)
    ctx.timers(3)
    ctx.run(nstage)
    tms, tl = ctx.timers(0)
    achieved = bytes_per_site * liq.n / (pair_avg * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "k_pair_ucgld_fast", "kernel_ms": pair_avg,
                "bytes_per_site": bytes_per_site, "half_neighbors_per_site": m_half,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                "stage_ms_per_step": {k: v / nstage for k, v in tms.items()},
                "stage_ms_note": f"a second run of {nstage} steps of the same loop as the timed region (speculative launches on); event pairs "
                                 "recorded without synchronisation (ucgb200_timers mode 3); neigh = rebuild kernels, comm = ghost refresh, "
                                 "modify = fused fix stages; the rest of ms_per_step is launch gaps and the flag read-back"}
    prof = os.path.join(ROOT, "profiles", "pair_traffic.json")
    if os.path.exists(prof):
        try:
            pj = json.load(open(prof))
            roofline["traffic"] = pj.get("dram_bytes_per_launch")
            # the ceilings that bind before HBM does (ncu, same capture): DESIGN.md §4
            roofline["secondary"] = {"l1_data_pipe_pct_of_peak": pj.get("lsu_data_pipe_pct"), "tex_data_pipe_pct_of_peak": pj.get("tex_data_pipe_pct"), "fp64_pipe_pct_of_peak": pj.get("fp64_pipe_pct"),
                                     "source": pj.get("source")}
        except Exception:
            pass

    # ---- e2e: the same step through the C-ABI with HOST buffers every step: the full dynamic state of
    # every site (x, v, lambda, v_lambda, state) goes host -> device, one step runs, and the new state
    # plus forces and probabilities come back into pinned host arrays
    n = liq.n
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    H = dict(x=pin((n, 3), torch.float64), v=pin((n, 3), torch.float64), f=pin((n, 3), torch.float64),
             ucgl=pin((n,), torch.float64), ucgvl=pin((n,), torch.float64), ucgstate=pin((n,), torch.int32),
             ucgp=pin((n,), torch.float64), ucgforce=pin((n,), torch.float64))
    ctx.atoms_download_into(**H)
    up = ("x", "v", "ucgl", "ucgvl", "ucgstate")
    h2d = sum(H[k].nbytes for k in up)
    d2h = sum(v.nbytes for v in H.values())
    e2e_steps = max(3, min(args.steps, 20))

    pipelined = os.environ.get("UCGB200_E2E_PIPELINE", "1") != "0"
    inp = {k: H[k] for k in up}

    def e2e_step():
        if pipelined:
            ctx.step_host(inp, H)          # ucgb200_step_host: D2H of x under the pair kernel, of f under the fix stages
        else:
            ctx.atoms_upload(n, **inp)
            ctx.run(1)
            ctx.atoms_download_into(**H)

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": n * e2e_steps / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "api": ("ucgb200_step_host (upload, one step, download; device->host copies overlap the kernels), pinned host arrays"
                   if pipelined else "ucgb200_atoms_upload + ucgb200_run(1) + ucgb200_atoms_download, pinned host arrays")}

    # ---- the taps either side of the path, after the timed regions: one `dump custom` snapshot of the
    # resident state written to a file (rows selected, sorted, packed and formatted on the device) and read back
    taps = None
    try:
        from lammps_ucg_dev_b200 import dumpio
        path = os.path.join(td, "bench.dump")
        cols = "id type x y z vx vy vz ucgstate ucgl ucgp"
        d = dumpio.DumpCustom(ctx, f"dump b all custom 1000 {path} {cols}")
        d.modify("dump_modify b sort id")
        ts = []
        for rep in range(2):
            ctx.sync(); t0 = time.perf_counter(); d.write(int(th[10])); ctx.sync(); ts.append(time.perf_counter() - t0)
        st = d.stats()
        d.close()
        d = dumpio.DumpCustom(ctx, f"dump c all custom 1000 {path}.one {cols}")
        d.modify("dump_modify c sort id")
        d.write(int(th[10]))
        d.close()
        tr = []
        for rep in range(2):   # the first call also page-locks its text staging
            ctx.sync(); t0 = time.perf_counter()
            rd = dumpio.read_dump(ctx, f"read_dump {path}.one {int(th[10])} x y z vx vy vz ucgstate ucgl ucgp")
            ctx.sync(); tr.append(time.perf_counter() - t0)
        t_read = min(tr)
        taps = {"dump_custom_ms": 1e3 * min(ts), "dump_rows": st["rows"], "dump_bytes": st["bytes"], "columns": len(cols.split()),
                "read_dump_ms": 1e3 * t_read, "read_dump_replaced": rd["replaced"]}
    except Exception as e:   # never fail the bench line because of the taps
        taps = {"error": str(e)[:200]}

    # ---- CPU baseline beside it (bounded sample, rank 0 only)
    cpu = None
    if not args.no_cpu:
        r = cpu_reference(ncell1, 8, 1, td)
        cpu = {"value": r["value"], "unit": UNIT, "cores": 1, "kind": r["kind"],
               "sample": f"the same {r['sites']}-site liquid and deck, {r['steps']} steps after setup and 1 warm-up step, serial "
                         "(one core; the UCG package is single-threaded per MPI rank and this box has no MPI)",
               "seconds": r["seconds"], "breakdown_s": r["breakdown"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(ncell1, 1),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "thermo": {"lambda_temp": th[9], "rebuilds_total": int(th[11]), "nghost": int(th[13])}, "taps": taps}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--serial", action="store_true", help="--impl reference: one core only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
