"""Per-source-line instruction and stall-sample shares of one kernel in an `ncu --set full --import-source on` report
(read with `ncu -i ... --page source --print-source cuda,sass --csv`):
    python scripts/ncu_lines.py gpurun_out/x.ncu-rep <kernel-substring> out.json [top]
Needs -lineinfo at compile time (csrc/Makefile has it)."""
import csv, io, json, subprocess, sys


def main():
    rep, kern, out = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern, "-c", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    if not hi:
        json.dump({"error": "no source page for " + kern}, open(out, "w"))
        return
    hdr = rows[hi[0]]
    ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    name = next((r[1] for r in rows[:hi[0]] if r and r[0] == "Function Name"), kern)
    lines = []
    for r in rows[hi[0] + 1:]:
        if r and r[0] == "File Path":
            break
        if r and r[0].isdigit() and len(r) > max(ci, si) and r[2] == "-":
            lines.append((int(r[0]), r[1].strip()[:140], float(r[ci] or 0), float(r[si] or 0)))
    ti, ts = sum(l[2] for l in lines) or 1.0, sum(l[3] for l in lines) or 1.0
    lines.sort(key=lambda l: -l[2])
    res = {"report": rep.split("/")[-1], "kernel": name, "warp_instructions": ti, "stall_samples": ts,
           "lines": [{"line": l[0], "inst_pct": round(100 * l[2] / ti, 2), "sample_pct": round(100 * l[3] / ts, 2), "source": l[1]}
                     for l in lines[:top]]}
    json.dump(res, open(out, "w"), indent=1)
    for l in res["lines"][:12]:
        print("%5d %5.1f%% inst %5.1f%% samp  %s" % (l["line"], l["inst_pct"], l["sample_pct"], l["source"][:100]))


if __name__ == "__main__":
    main()
