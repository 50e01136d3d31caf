"""Randomised sweep of the PaddleOCR-style device ops against their CPU restatements (oracle/db_post.py = upstream
DBPostProcess on cv2 + restated Clipper; NumPy for CTC; oracle/reading_order.py for the line merge) over parameters
and map shapes the parity tests do not enumerate.

    python tools/sweep_paddle_vs_oracle.py --seeds 0 60
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import cv2
import numpy as np
import torch

from oracle import db_post as D
from oracle import reading_order as R
from ocr_system_b200 import ops
from ocr_system_b200.paddle_ops import CTCLabelDecode, DBPostProcess


def blob_map(rng, h, w):
    """Irregular components: smoothed noise (concave blobs, nested holes, specks) instead of text boxes."""
    k = int(rng.choice([5, 9, 15, 25]))
    a = cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (k, k), 0)
    a = (a - a.min()) / max(float(a.max() - a.min()), 1e-6)
    return np.clip((a - 0.5) * float(rng.uniform(2.0, 6.0)) + 0.5, 0, 1).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs=2, default=[0, 40])
    ap.add_argument("--tiny", action="store_true", help="maps from 1 x 1 to 48 x 48 and thin strips")
    a = ap.parse_args()
    bad, checked, t0 = [], 0, time.time()
    for seed in range(*a.seeds):
        rng = np.random.default_rng(seed)
        # ---- DBPostProcess
        h, w = int(rng.choice([96, 320, 640, 736, 960])), int(rng.choice([128, 352, 800, 960, 1120]))
        if a.tiny:
            h, w = (int(rng.integers(1, 49)), int(rng.integers(1, 49))) if seed % 2 else (int(rng.integers(1, 5)), int(rng.integers(50, 900)))
            if seed % 4 == 2:
                h, w = w, h
            pred = blob_map(rng, max(h, 32), max(w, 32))[:h, :w].copy() if seed % 3 else (rng.random((h, w)) > 0.4).astype(np.float32) * 0.9
        else:
            pred = blob_map(rng, h, w) if seed % 3 == 0 else D.synth_prob_map(h, w, seed, n_boxes=int(rng.integers(5, 400)))
        kw = dict(thresh=float(rng.choice([0.2, 0.3, 0.5])), box_thresh=float(rng.choice([0.5, 0.6, 0.7])),
                  unclip_ratio=float(rng.choice([1.5, 1.6, 2.0])), max_candidates=int(rng.choice([50, 1000])),
                  use_dilation=bool(rng.integers(0, 2)), score_mode=str(rng.choice(["fast", "slow"])))
        dst = (max(1, int(h * rng.choice([1.0, 1.5, 0.75]))), max(1, int(w * rng.choice([1.0, 1.333, 2.0]))))
        sl = [(dst[0], dst[1], h / dst[0], w / dst[1])]
        ref = D.DBPostProcess(**kw)({"maps": pred[None, None]}, sl, with_scores=True)[0]
        got = DBPostProcess(**kw)({"maps": pred[None, None]}, sl, with_scores=True)[0]
        checked += 1
        ok = len(got["points"]) == len(ref["points"]) and (len(ref["points"]) == 0 or (
            np.array_equal(np.asarray(got["points"]), np.asarray(ref["points"]))
            and np.abs(np.asarray(got["scores"], np.float64) - np.asarray(ref["scores"], np.float64)).max() <= 1e-4))
        if not ok:
            nd = -1
            if len(got["points"]) == len(ref["points"]):
                nd = int((np.asarray(got["points"]) != np.asarray(ref["points"])).any(axis=(1, 2)).sum())
            bad.append((seed, "db", h, w, kw, len(ref["points"]), len(got["points"]), nd))
            print("MISMATCH db", seed, h, w, kw, "boxes", len(ref["points"]), len(got["points"]), "differing", nd, flush=True)
        # ---- CTC
        n, t, c = int(rng.integers(1, 70)), int(rng.integers(1, 60)), int(rng.choice([2, 37, 97, 6625]))
        p = rng.random((n, t, c)).astype(np.float32)
        if seed % 2:
            p = np.round(p * 8) / 8           # many exact ties -> first index wins
        idx, pos, ln, conf = [x.cpu().numpy() for x in ops.ctc_greedy(torch.from_numpy(p).cuda())]
        am, mx = p.argmax(2), p.max(2)
        checked += 1
        for b in range(n):
            sel = np.ones(t, bool); sel[1:] = am[b, 1:] != am[b, :-1]; sel &= am[b] != 0
            want = am[b][sel]
            wc = float(np.mean(mx[b][sel])) if sel.any() else 0.0
            if ln[b] != len(want) or not np.array_equal(idx[b, :ln[b]], want) or abs(conf[b] - wc) > 1e-4:
                bad.append((seed, "ctc", n, t, c, b))
                print("MISMATCH ctc", seed, n, t, c, b, flush=True)
                break
        # ---- reading order
        from reading_pages import page
        from ocr_system_b200 import ocr_postprocessor as PP

        kind = ["grid", "ties", "float"][seed % 3]
        items = page(seed, int(rng.choice([1, 2, 7, 64, 65, 300, 1000])), kind)
        ratio = float(rng.choice([0.0, 0.3, 0.5, 0.7, 2.0, -0.5]))
        merged = PP.process_ocr_result(items, y_tolerance_ratio=ratio)
        order, line_of, nl, lc, ly = R.reading_order([it[0] for it in items], [it[2] for it in items], ratio)
        checked += 1
        got_order = [int(b.text[1:]) for m in merged for b in m.blocks]
        if got_order != order or len(merged) != nl or [m.confidence for m in merged] != lc or [m.y_position for m in merged] != ly:
            bad.append((seed, "reading_order", kind, len(items), ratio))
            print("MISMATCH reading_order", seed, kind, len(items), ratio, flush=True)
    print(json.dumps({"seeds": a.seeds, "checked": checked, "mismatches": len(bad), "seconds": round(time.time() - t0, 1)}))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sweep_paddle_vs_oracle.json", "w") as f:
        json.dump({"seeds": a.seeds, "checked": checked, "mismatches": [list(map(str, b)) for b in bad]}, f)


if __name__ == "__main__":
    main()
