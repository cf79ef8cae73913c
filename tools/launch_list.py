"""Summarise an ncu launch list (gpu__time_duration.sum csv): the last 1/reps of the launches, one line.
python tools/launch_list.py file.csv [reps]"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
recs = [(r[ik].split("(")[0].replace("void ", "").replace("rg::", ""), float(r[iv].replace(",", ""))) for r in rows[1:]]
n = len(recs) // reps
last = recs[-n:]
print(" ".join(f"{k[:22]}={v / 1e3:.0f}" for k, v in last))
print("launches", len(last), "total us", round(sum(v for k, v in last) / 1e3, 1))
