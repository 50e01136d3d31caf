"""Per-instruction execution counts of an ncu SASS source page (ncu -i rep --page source --csv --print-source sass).
usage: python tools/ncu_sass_region.py x.csv [first_opcode_substring] [before] [after]  -- prints the region around the
first instruction containing the substring; without arguments prints opcode totals."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < 8: continue
    data.append((r[1].strip(), float(r[ix["Instructions Executed"]] or 0), float(r[ix["# Samples"]] or 0)))
tot = sum(d[1] for d in data)
print("instructions executed", tot, "samples", sum(d[2] for d in data))
if len(sys.argv) > 2:
    key = sys.argv[2]; before = int(sys.argv[3]) if len(sys.argv) > 3 else 40; after = int(sys.argv[4]) if len(sys.argv) > 4 else 200
    i0 = [i for i, d in enumerate(data) if key in d[0]][0]
    for d in data[max(0, i0 - before): i0 + after]:
        print(f"{d[1] / 1e6:8.2f}M {d[2]:6.0f}  {d[0][:100]}")
else:
    c = collections.Counter()
    for d in data:
        op = d[0].split()[1] if d[0].startswith("@") else d[0].split()[0]
        c[op.split(".")[0]] += d[1]
    for k, v in c.most_common(30):
        print(f"{v / 1e6:9.2f}M {100 * v / tot:5.1f}%  {k}")
