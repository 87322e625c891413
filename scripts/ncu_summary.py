"""Summarise an `ncu --set full` report (read here, no GPU needed) into profiles/:
    python scripts/ncu_summary.py gpurun_out/prof_pair.ncu-rep profiles/r01_pair_full_v4 [pair_traffic]
writes <out>.json (selected raw metrics per launch) and, with the third argument, refreshes
profiles/pair_traffic.json (dram bytes per launch, read by bench.py for roofline.traffic)."""
import csv, io, json, os, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def to_bytes(val, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
    return float(val) * f if f else None


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: k for k, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[col["Kernel Name"]]}
        for k in KEYS:
            if k in col:
                v = r[col[k]].replace(",", "")
                try:
                    d[k] = float(v)
                except ValueError:
                    d[k] = v
                d[k + "|unit"] = units[col[k]]
        launches.append(d)
    json.dump({"report": os.path.basename(rep), "launches": launches}, open(out + ".json", "w"), indent=1)
    if len(sys.argv) > 3:
        tr = []
        for d in launches:
            rd = to_bytes(d["dram__bytes_read.sum"], d["dram__bytes_read.sum|unit"])
            wr = to_bytes(d["dram__bytes_write.sum"], d["dram__bytes_write.sum|unit"])
            tr.append(rd + wr)
        json.dump({"kernel": launches[0]["kernel"], "dram_bytes_per_launch": sum(tr) / len(tr),
                   "source": out + ".json", "launches": len(tr)}, open("profiles/pair_traffic.json", "w"), indent=1)
    for d in launches:
        print(d["kernel"][:90], d.get("gpu__time_duration.sum"), d.get("gpu__time_duration.sum|unit"))


if __name__ == "__main__":
    main()
