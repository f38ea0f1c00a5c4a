"""Turns ncu outputs under gpurun_out/ into small tracked summaries under profiles/.
   usage: summarize_ncu.py launches <csv> <out.txt>        (launch list: kernel, grid, block, duration, share)
          summarize_ncu.py full <rep.ncu-rep> <out.txt>    (key counters per captured launch)"""
import csv, io, subprocess, sys

mode, src, dst = sys.argv[1:4]
if mode == "launches":
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    gi, bi = h.index("Grid Size"), h.index("Block Size")
    rec = {}
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            rec.setdefault(r[0], {"k": r[ki], "g": r[gi], "b": r[bi]})[r[mi]] = r[vi]
    ours = [v for v in rec.values() if "at::" not in v["k"] and "cub::" not in v["k"]]
    tot = sum(float(v.get("gpu__time_duration.sum", 0)) for v in ours)
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold cache, serialised: compare SHARES)\n")
        f.write("# source: %s ; our kernels only; total %.1f us\n" % (src, tot / 1e3))
        f.write("%-64s %-16s %-12s %10s %7s %14s\n" % ("kernel", "grid", "block", "time_us", "share", "warp_instr"))
        for v in ours:
            t = float(v.get("gpu__time_duration.sum", 0))
            f.write("%-64s %-16s %-12s %10.1f %6.1f%% %14s\n" % (v["k"][:64], v["g"], v["b"], t / 1e3, 100 * t / max(tot, 1),
                                                               v.get("smsp__inst_executed.sum", "-")))
else:
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
    units = rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on ; source: %s\n" % src)
        for n, r in enumerate(rows[2:]):
            f.write("\n== launch %d\n" % n)
            for w in want:
                if w in h:
                    i = h.index(w)
                    f.write("%-86s %s %s\n" % (w, r[i][:110], units[i]))
print("wrote", dst)
