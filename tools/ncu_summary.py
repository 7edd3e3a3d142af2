#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/:
per-launch duration, DRAM bytes, L1/L2 hit rates, issue utilisation, stall mix, and the
instructions that collect the most stall samples (needs -lineinfo / --import-source on).

usage: tools/ncu_summary.py gpurun_out/x.ncu-rep [profiles/x.summary.txt]
"""
import csv
import io
import subprocess
import sys


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    col = {h: i for i, h in enumerate(hdr)}
    want = [
        "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    ]
    print(f"# ncu summary of {rep}  (cold-cache, serialised replays: compare shares, not absolutes)", file=out)
    for k, r in enumerate(rows):
        print(f"\n## launch {k}", file=out)
        for w in want:
            if w in col:
                v = r[col[w]]
                if w == "Kernel Name":
                    v = v[:110]
                print(f"{w:70s} {v} {units[col[w]]}", file=out)
        try:
            rd, wr = float(r[col["dram__bytes_read.sum"]]), float(r[col["dram__bytes_write.sum"]])
            ur, uw = units[col["dram__bytes_read.sum"]], units[col["dram__bytes_write.sum"]]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            print(f"{'dram traffic (read + write), bytes':70s} {rd * scale[ur] + wr * scale[uw]:.0f}", file=out)
        except Exception:
            pass
    det = ncu(["-i", rep, "--page", "details", "--print-units", "base"])
    keep = ("Duration", "Executed Ipc Active", "Issue Slots Busy", "No Eligible", "Active Warps Per Scheduler",
            "Eligible Warps Per Scheduler", "Warp Cycles Per Issued", "cycles being stalled", "This stall type",
            "Achieved Occupancy", "Theoretical Occupancy", "DRAM Throughput", "Mem Busy", "Max Bandwidth", "L2 Hit Rate",
            "L1/TEX Hit Rate", "Registers Per Thread")
    print("\n# details page (selected lines, all launches in order)", file=out)
    for line in det.splitlines():
        if any(k in line for k in keep) or line.strip().startswith("void ") or "count_bucket" in line or "find_" in line:
            print(line.rstrip()[:200], file=out)
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    h = next((i for i, r in enumerate(src) if "Source" in r and "# Samples" in r), None)
    if h is not None:
        hd = src[h]
        si, ci = hd.index("Source"), hd.index("# Samples")
        body = []
        for r in src[h + 1:]:
            if r == hd or len(r) <= ci:
                break
            try:
                body.append((int(r[ci] or 0), r[si].strip()))
            except ValueError:
                break
        tot = sum(c for c, _ in body) or 1
        print(f"\n# first launch: instructions with >= 2% of the {tot} stall samples (index, share, SASS)", file=out)
        for i, (c, s) in enumerate(body):
            if c >= 0.02 * tot:
                print(f"{i:5d} {100 * c / tot:5.1f}%  {s[:100]}", file=out)


if __name__ == "__main__":
    main()
