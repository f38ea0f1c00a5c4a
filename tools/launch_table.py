"""Prints the per-launch table of an `ncu --metrics gpu__time_duration.sum --csv` log.
    python tools/launch_table.py gpurun_out/launches.csv [last_n]
"""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if "gpu__time_duration" in r.get("Metric Name", "")]
last = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows)
for r in rows[-last:]:
    print("%4s  %-60s %-16s %-14s %9.1f us" % (r["ID"], r["Kernel Name"][:60], r["Grid Size"], r["Block Size"],
                                              float(r["Metric Value"].replace(",", "")) / 1e3))
