"""Print the tail of an ncu gpu__time_duration launch list (CSV) starting at the last k_cell_area launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
names = [r[ki] for r in data]
anchor = sys.argv[2] if len(sys.argv) > 2 else 'k_cell_area'
idx = [i for i, n in enumerate(names) if anchor in n]
start = idx[-1]
tot = 0
for r in data[start:]:
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    tot += v
    print(f"{v:9.1f} us  {r[ki][:80]}")
print('total us', round(tot, 1))
