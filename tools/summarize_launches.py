"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/summarize_launches.py gpurun_out/launches.csv "<header comment>" > profiles/rN_launch_list_summary.txt"""
import csv, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik]).strip()
    v = float(r[iv].replace(",", ""))
    ms = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu].startswith("u") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print(f"# total kernel time in capture: {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")
print("kernel | launches | total ms | ms/launch | share")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k} | {n} | {ms:.3f} | {ms / n:.4f} | {100 * ms / tot:.1f}%")
