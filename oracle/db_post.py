"""CPU ORACLE for DBPostProcess -- test / baseline infrastructure only.

[upstream PaddleOCR, NOT in the reference tree] ppocr/postprocess/db_postprocess.py
``DBPostProcess`` restated on the libraries upstream uses where they exist in this
image (cv2.findContours / minAreaRect / boxPoints / fillPoly / mean) plus:
  * shapely Polygon.area / .length  -> shoelace / perimeter in float64
  * pyclipper (Clipper 6.4.2) PyclipperOffset(JT_ROUND, ET_CLOSEDPOLYGON).Execute
    -> ``clipper_offset_round`` below, a restatement of clipper.cpp ClipperOffset
    (AddPath / FixOrientations / DoOffset / OffsetPoint / DoRound).  The trailing
    Clipper union only removes duplicate / collinear vertices of an outward offset of a
    convex quad, which min-area-rect does not see.
PARITY UNPINNED: neither paddleocr nor pyclipper/shapely can be installed here and the
reference ships no vectors for this op.  The only in-repo contract is the
consumer format in backend/utils/ocr_postprocessor.py:19-24,70-93 (quad as 4 x [x, y]).
"""
from __future__ import annotations

import math

import cv2
import numpy as np


def _round_clipper(v: float) -> int:
    return int(v - 0.5) if v < 0 else int(v + 0.5)


def clipper_offset_round(path, delta: float, arc_tolerance: float = 0.25):
    """ClipperOffset.AddPath(path, jtRound, etClosedPolygon); Execute(delta) before the union.
    path: sequence of (x, y), truncated to integers like pyclipper's IntPoint conversion."""
    pts = [(int(p[0]), int(p[1])) for p in path]
    hi = len(pts) - 1
    if hi < 0:
        return []
    while hi > 0 and pts[0] == pts[hi]:
        hi -= 1
    contour = [pts[0]]
    for i in range(1, hi + 1):
        if contour[-1] != pts[i]:
            contour.append(pts[i])
    if len(contour) < 3:
        return []
    # FixOrientations: Orientation == (Area >= 0); reverse when false
    a = 0.0
    n = len(contour)
    j = n - 1
    for i in range(n):
        a += (float(contour[j][0]) + contour[i][0]) * (float(contour[j][1]) - contour[i][1])
        j = i
    if not (-a * 0.5 >= 0):
        contour.reverse()
    if abs(delta) < 1e-20:
        return contour
    y = arc_tolerance if arc_tolerance > 0 else 0.25
    if arc_tolerance > abs(delta) * 0.25:
        y = abs(delta) * 0.25
    steps = math.pi / math.acos(1 - y / abs(delta))
    if steps > abs(delta) * math.pi:
        steps = abs(delta) * math.pi
    m_sin, m_cos = math.sin(2 * math.pi / steps), math.cos(2 * math.pi / steps)
    steps_per_rad = steps / (2 * math.pi)
    if delta < 0:
        m_sin = -m_sin
    normals = []
    for i in range(n):
        p1, p2 = contour[i], contour[(i + 1) % n]
        dx, dy = float(p2[0] - p1[0]), float(p2[1] - p1[1])
        if dx == 0 and dy == 0:
            normals.append((0.0, 0.0))
            continue
        f = 1.0 / math.sqrt(dx * dx + dy * dy)
        dx *= f
        dy *= f
        normals.append((dy, -dx))
    out = []
    k = n - 1
    for j in range(n):
        nk, nj = normals[k], normals[j]
        sx, sy = contour[j]
        sin_a = nk[0] * nj[1] - nj[0] * nk[1]
        done = False
        if abs(sin_a * delta) < 1.0:
            cos_a = nk[0] * nj[0] + nj[1] * nk[1]
            if cos_a > 0:
                out.append((_round_clipper(sx + nk[0] * delta), _round_clipper(sy + nk[1] * delta)))
                done = True
        elif sin_a > 1.0:
            sin_a = 1.0
        elif sin_a < -1.0:
            sin_a = -1.0
        if done:
            continue  # note: clipper returns before `k = j`
        if sin_a * delta < 0:
            out.append((_round_clipper(sx + nk[0] * delta), _round_clipper(sy + nk[1] * delta)))
            out.append((sx, sy))
            out.append((_round_clipper(sx + nj[0] * delta), _round_clipper(sy + nj[1] * delta)))
        else:  # DoRound
            ang = math.atan2(sin_a, nk[0] * nj[0] + nk[1] * nj[1])
            st = max(_round_clipper(steps_per_rad * abs(ang)), 1)
            X, Y = nk
            for _ in range(st):
                out.append((_round_clipper(sx + X * delta), _round_clipper(sy + Y * delta)))
                X2 = X
                X = X * m_cos - m_sin * Y
                Y = X2 * m_sin + Y * m_cos
            out.append((_round_clipper(sx + nj[0] * delta), _round_clipper(sy + nj[1] * delta)))
        k = j
    return out


def get_mini_boxes(contour):
    """upstream get_mini_boxes: minAreaRect -> boxPoints -> order [tl, tr, br, bl]-ish by x then y."""
    bounding_box = cv2.minAreaRect(contour)
    points = sorted(list(cv2.boxPoints(bounding_box)), key=lambda x: x[0])
    if points[1][1] > points[0][1]:
        i1, i4 = 0, 1
    else:
        i1, i4 = 1, 0
    if points[3][1] > points[2][1]:
        i2, i3 = 2, 3
    else:
        i2, i3 = 3, 2
    box = [points[i1], points[i2], points[i3], points[i4]]
    return box, min(bounding_box[1])


def box_score_fast(bitmap, _box):
    h, w = bitmap.shape[:2]
    box = _box.copy()
    xmin = np.clip(np.floor(box[:, 0].min()).astype("int32"), 0, w - 1)
    xmax = np.clip(np.ceil(box[:, 0].max()).astype("int32"), 0, w - 1)
    ymin = np.clip(np.floor(box[:, 1].min()).astype("int32"), 0, h - 1)
    ymax = np.clip(np.ceil(box[:, 1].max()).astype("int32"), 0, h - 1)
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
    box[:, 0] = box[:, 0] - xmin
    box[:, 1] = box[:, 1] - ymin
    cv2.fillPoly(mask, box.reshape(1, -1, 2).astype("int32"), 1)
    return cv2.mean(bitmap[ymin : ymax + 1, xmin : xmax + 1], mask)[0]


def box_score_slow(bitmap, contour):
    """upstream box_score_slow: mean of the probabilities over cv2.fillPoly(contour)."""
    h, w = bitmap.shape[:2]
    contour = np.reshape(contour.copy(), (-1, 2))
    xmin = np.clip(np.min(contour[:, 0]), 0, w - 1)
    xmax = np.clip(np.max(contour[:, 0]), 0, w - 1)
    ymin = np.clip(np.min(contour[:, 1]), 0, h - 1)
    ymax = np.clip(np.max(contour[:, 1]), 0, h - 1)
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
    contour[:, 0] = contour[:, 0] - xmin
    contour[:, 1] = contour[:, 1] - ymin
    cv2.fillPoly(mask, contour.reshape(1, -1, 2).astype("int32"), 1)
    return cv2.mean(bitmap[ymin : ymax + 1, xmin : xmax + 1], mask)[0]


def unclip(box, unclip_ratio):
    """shapely area/length + pyclipper offset, restated.  Returns int array [K,2] or None."""
    b = np.asarray(box, np.float64)
    x, y = b[:, 0], b[:, 1]
    area = 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))
    length = float(np.sum(np.hypot(np.roll(x, -1) - x, np.roll(y, -1) - y)))
    if length == 0:
        return None
    distance = area * unclip_ratio / length
    pts = clipper_offset_round([(float(p[0]), float(p[1])) for p in box], distance)
    if len(pts) < 3:
        return None
    return np.array(pts, np.int32)


def boxes_from_bitmap(pred, bitmap, dest_width, dest_height, box_thresh=0.6, unclip_ratio=1.5, max_candidates=1000,
                      min_size=3, score_mode="fast"):
    height, width = bitmap.shape
    contours, _ = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    num = min(len(contours), max_candidates)
    boxes, scores = [], []
    for index in range(num):
        contour = contours[index]
        points, sside = get_mini_boxes(contour)
        if sside < min_size:
            continue
        points = np.array(points)
        score = box_score_fast(pred, points.reshape(-1, 2)) if score_mode == "fast" else box_score_slow(pred, contour)
        if box_thresh > score:
            continue
        ex = unclip(points, unclip_ratio)
        if ex is None:
            continue
        box, sside = get_mini_boxes(ex.reshape(-1, 1, 2))
        if sside < min_size + 2:
            continue
        box = np.array(box)
        box[:, 0] = np.clip(np.round(box[:, 0] / width * dest_width), 0, dest_width)
        box[:, 1] = np.clip(np.round(box[:, 1] / height * dest_height), 0, dest_height)
        boxes.append(box.astype("int32"))
        scores.append(score)
    return np.array(boxes, dtype="int32").reshape(-1, 4, 2), scores


class DBPostProcess:
    """upstream signature: DBPostProcess(thresh, box_thresh, max_candidates, unclip_ratio,
    use_dilation, score_mode, box_type)(outs_dict, shape_list) -> [{"points": int32[K,4,2]}]."""

    def __init__(self, thresh=0.3, box_thresh=0.7, max_candidates=1000, unclip_ratio=2.0, use_dilation=False,
                 score_mode="fast", box_type="quad", **kwargs):
        assert score_mode in ("fast", "slow") and box_type == "quad"
        self.score_mode = score_mode
        self.thresh, self.box_thresh = thresh, box_thresh
        self.max_candidates, self.unclip_ratio = max_candidates, unclip_ratio
        self.min_size = 3
        self.dilation_kernel = np.array([[1, 1], [1, 1]]) if use_dilation else None

    def __call__(self, outs_dict, shape_list, with_scores: bool = False):
        pred = np.asarray(outs_dict["maps"])[:, 0, :, :]
        segmentation = pred > self.thresh
        out = []
        for b in range(pred.shape[0]):
            # upstream passes shape_list as a float64 ndarray: src dims are np.float64 scalars
            src_h, src_w = np.float64(shape_list[b][0]), np.float64(shape_list[b][1])
            mask = segmentation[b]
            if self.dilation_kernel is not None:
                mask = cv2.dilate(np.array(mask).astype(np.uint8), self.dilation_kernel)
            boxes, scores = boxes_from_bitmap(pred[b], mask, src_w, src_h, self.box_thresh, self.unclip_ratio,
                                              self.max_candidates, self.min_size, self.score_mode)
            d = {"points": boxes}
            if with_scores:
                d["scores"] = np.asarray(scores, np.float64)
            out.append(d)
        return out


def synth_prob_map(h=960, w=960, seed=0, n_boxes=500, hole_frac=0.1):
    """background U(0,0.12); ~n_boxes axis-aligned / slightly rotated text boxes
    (h 10-16, w 18-60) filled U(0.75,0.99) with softened borders, a fraction with interior holes."""
    rng = np.random.default_rng(seed)
    pred = rng.uniform(0, 0.12, (h, w)).astype(np.float32)
    occ = np.zeros((h, w), np.uint8)
    placed = 0
    tries = 0
    while placed < n_boxes and tries < n_boxes * 20:
        tries += 1
        bh, bw = int(rng.integers(10, 17)), int(rng.integers(18, 61))
        cx, cy = float(rng.uniform(40, w - 40)), float(rng.uniform(20, h - 20))
        ang = float(rng.uniform(-8, 8)) if rng.random() < 0.5 else 0.0
        rect = cv2.boxPoints(((cx, cy), (bw, bh), ang)).astype(np.int32)
        x0, y0 = max(int(rect[:, 0].min()) - 3, 0), max(int(rect[:, 1].min()) - 3, 0)
        x1, y1 = min(int(rect[:, 0].max()) + 4, w), min(int(rect[:, 1].max()) + 4, h)
        if occ[y0:y1, x0:x1].any():
            continue
        occ[y0:y1, x0:x1] = 1
        m = np.zeros((y1 - y0, x1 - x0), np.uint8)
        cv2.fillPoly(m, [rect - np.array([x0, y0], np.int32)], 1)
        roi = pred[y0:y1, x0:x1]
        val = rng.uniform(0.75, 0.99, m.shape).astype(np.float32)
        roi[:] = np.where(m > 0, val, roi)
        er = cv2.erode(m, np.ones((3, 3), np.uint8))
        edge = (m > 0) & (er == 0)
        roi[:] = np.where(edge, roi * np.float32(0.55), roi)  # softened border (still > thresh 0.3)
        if rng.random() < hole_frac and bh >= 12 and bw >= 24:
            hx, hy = int(cx), int(cy)
            hw_, hh_ = int(rng.integers(2, 5)), int(rng.integers(2, 4))
            pred[hy - hh_ // 2 : hy - hh_ // 2 + hh_, hx - hw_ // 2 : hx - hw_ // 2 + hw_] = rng.uniform(0, 0.1)
        placed += 1
    return pred
