"""Print the kernels of one wall-GMRES iteration (k_wall_from_1d ... next k_wall_from_1d) from an ncu launch list."""
import csv, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(row[ki][:64], float(row[vi].replace(",", "")) / 1000) for row in r]
idx = [i for i, (k, _) in enumerate(seq) if "k_wall_from_1d" in k]
back = int(sys.argv[2]) if len(sys.argv) > 2 else 10
a, b = idx[-back], idx[-back + 1]
for k, v in seq[a:b]:
    print(f"{k:64s} {v:7.2f}")
print("iteration total us", round(sum(v for _, v in seq[a:b]), 1), "launches", b - a)
