"""Print an ncu `--metrics gpu__time_duration.sum --csv` launch list: kernel, grid, block, microseconds."""
import csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; ix = {k: i for i, k in enumerate(h)}
tot = 0.0
for r in rows[1:]:
    us = float(r[ix["Metric Value"]]) / 1e3
    tot += us
    print(re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:44].ljust(44), r[ix["Grid Size"]].rjust(14), r[ix["Block Size"]].rjust(12), f"{us:8.1f} us")
print(f"{'total':44s} {'':14s} {'':12s} {tot:8.1f} us")
