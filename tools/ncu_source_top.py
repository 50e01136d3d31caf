"""Top source lines of an ncu --import-source report by non-barrier warp-stall samples.
usage: ncu -i rep --page source --csv --print-source cuda,sass > x.csv ; python tools/ncu_source_top.py x.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0]: continue
    tot = f(r, "# Samples"); bar = f(r, "stall_barrier")
    data.append((tot - bar, tot, bar, r))
T = sum(d[0] for d in data)
print("non-barrier samples", T, "total", sum(d[1] for d in data))
for nb, tot, bar, r in sorted(data, key=lambda d: -d[0])[:topn]:
    st = {k: f(r, k) for k in hdr if k.startswith("stall_") and "(Not" not in k and k != "stall_barrier"}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{r[0]:>4} {nb:7.0f} {100*nb/T:5.1f}% inst={f(r,'Instructions Executed'):9.0f} "
          f"{' '.join(f'{k[6:]}={v:.0f}' for k, v in top)} | {r[1].strip()[:80]}")
