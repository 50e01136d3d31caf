import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2, torch
from oracle import db_post as D
from ocr_system_b200.paddle_ops import DBPostProcess
full = D.synth_prob_map(960, 960, 0, n_boxes=300)
for (x0, y0, x1, y1) in ((770, 520, 850, 570), (430, 540, 480, 585)):
    pred = np.full((960, 960), 0.05, np.float32)
    pred[y0:y1, x0:x1] = full[y0:y1, x0:x1]
    sl = [(960, 960, 1.0, 1.0)]
    kw = dict(thresh=0.3, box_thresh=0.1, unclip_ratio=1.5)
    ref = D.DBPostProcess(**kw)({"maps": pred[None, None]}, sl, with_scores=True)[0]
    got = DBPostProcess(**kw)({"maps": pred[None, None]}, sl, with_scores=True)[0]
    print("ref", ref["points"].tolist(), ref["scores"]); print("got", got["points"].tolist(), got["scores"])
    # same box embedded in a small map
    sub = np.ascontiguousarray(pred[y0-20:y1+20, x0-20:x1+20])
    sl2 = [(sub.shape[0], sub.shape[1], 1.0, 1.0)]
    ref = D.DBPostProcess(**kw)({"maps": sub[None, None]}, sl2, with_scores=True)[0]
    got = DBPostProcess(**kw)({"maps": sub[None, None]}, sl2, with_scores=True)[0]
    print("small ref", ref["scores"], "got", got["scores"])
