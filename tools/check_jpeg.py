"""GPU JPEG encoder vs Pillow on a few images: byte equality of the whole file (and, when that fails,
where the streams diverge)."""
import io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from ocr_system_b200 import ops

def pil_jpeg(a, q, opt):
    b = io.BytesIO(); Image.fromarray(a).save(b, format="JPEG", quality=q, optimize=opt); return b.getvalue()

def segs(d):
    i = 2; out = []
    while i < len(d):
        m = d[i + 1]; L = (d[i + 2] << 8) | d[i + 3]
        out.append((m, d[i + 4:i + 2 + L]))
        if m == 0xDA: out.append(("scan", d[i + 2 + L:-2])); break
        i += 2 + L
    return out

rng = np.random.default_rng(0)
cases = []
for (h, w) in [(16, 16), (48, 64), (50, 70), (93, 127), (678 // 2, 960 // 2)]:
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = np.stack([(xx * 3 + yy) % 256, (yy * 2) % 256, (xx + yy * 5) % 256], -1).astype(np.uint8)
    noise = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    text = np.full((h, w, 3), 255, np.uint8); text[(yy // 3 + xx // 5) % 4 == 0] = 20
    cases += [("smooth", smooth), ("noise", noise), ("text", text)]
bad = 0
for name, a in cases:
    for q in (95, 75, 30):
        for opt in (False, True):
            ref = pil_jpeg(a, q, opt)
            got = ops.jpeg_encode(torch.from_numpy(a[None]).cuda(), q, opt)[0]
            ok = got == ref
            if not ok:
                bad += 1
                rs, gs = segs(ref), segs(got)
                diff = [(hex(r[0]) if r[0] != "scan" else "scan") for r, g in zip(rs, gs) if r != g]
                sr, sg = rs[-1][1], gs[-1][1]
                first = next((i for i, (x, y) in enumerate(zip(sr, sg)) if x != y), min(len(sr), len(sg)))
                print(f"MISMATCH {name} {a.shape} q={q} opt={opt}: sizes {len(ref)} vs {len(got)}; differing segments {diff}; scan diverges at byte {first}/{len(sr)}")
print("cases", len(cases) * 6, "mismatches", bad)
