"""CPU unit tests of ocr-system_b200/csrc/db_geom.h (the geometry the DB kernels run per
candidate), compiled for the host by tests/geom_host.cpp and compared with cv2:
fillPoly coverage, minAreaRect (incl. cv2's tie-breaking on findContours borders and holes),
get_mini_boxes, the Clipper offset restatement, and the whole per-candidate chain."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G():
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    so = os.path.join(HERE, "_build", "libgeom_host.so")
    src = os.path.join(HERE, "geom_host.cpp")
    hdr = os.path.join(HERE, "..", "ocr-system_b200", "csrc", "db_geom.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, src])
    lib = C.CDLL(so)
    lib.geom_mini_box.restype = C.c_float
    lib.geom_unclip_distance.restype = C.c_double
    return lib


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def test_fillpoly_cover_matches_cv2(G):
    rng = np.random.default_rng(1)
    n = 0
    for it in range(4000):
        h, w = int(rng.integers(3, 60)), int(rng.integers(3, 90))
        ww, hh = rng.uniform(0, w * 0.55), rng.uniform(0, h * 0.55)
        if it % 4 == 0:
            hh = rng.uniform(0, 2)
        ang = rng.uniform(-90, 90) if it % 3 else 0.0
        rect = cv2.boxPoints(((rng.uniform(0.3 * w, 0.7 * w), rng.uniform(0.3 * h, 0.7 * h)), (ww, hh), ang))
        quad = np.ascontiguousarray(rect.astype(np.int32))
        if not (quad[:, 0].min() >= 0 and quad[:, 0].max() < w and quad[:, 1].min() >= 0 and quad[:, 1].max() < h):
            continue
        ref = np.zeros((h, w), np.uint8)
        cv2.fillPoly(ref, quad.reshape(1, -1, 2), 1)
        got = np.zeros((h, w), np.uint8)
        G.geom_fill_quad(P(quad), h, w, P(got))
        assert np.array_equal(ref, got), quad.tolist()
        n += 1
    assert n > 2000


def test_fillpoly_cover_matches_cv2_when_the_page_border_cuts_the_quad(G):
    """box_score_fast near the page edge: the mask is the quad's bounding box clipped to the image, so the quad sticks
    out by a few pixels; cv2 then draws the boundary with cv::clipLine'd end points and takes the scan-line slopes
    from them."""
    rng = np.random.default_rng(1)
    n = 0
    for it in range(6000):
        bw, bh, ang = rng.uniform(15, 60), rng.uniform(6, 16), rng.uniform(-12, 12)
        rect = cv2.boxPoints(((40, 20), (bw, bh), ang))
        xmin, xmax = int(np.floor(rect[:, 0].min())), int(np.ceil(rect[:, 0].max()))
        ymin, ymax = int(np.floor(rect[:, 1].min())), int(np.ceil(rect[:, 1].max()))
        cl, cr, ct, cb = [int(v) for v in rng.integers(0, 4, 4)]
        if it % 2:
            cr = cb = 0
        x0, y0 = xmin + cl, ymin + ct
        w, h = xmax - cr - x0 + 1, ymax - cb - y0 + 1
        if w < 3 or h < 3:
            continue
        quad = np.ascontiguousarray((rect - np.array([x0, y0], np.float32)).astype(np.int32))
        ref = np.zeros((h, w), np.uint8)
        cv2.fillPoly(ref, quad.reshape(1, -1, 2), 1)
        got = np.zeros((h, w), np.uint8)
        G.geom_fill_quad(P(quad), h, w, P(got))
        assert np.array_equal(ref, got), (quad.tolist(), h, w)
        n += 1
    assert n > 5000


def _random_mask(rng, h=120, w=160):
    m = np.zeros((h, w), np.uint8)
    for _ in range(int(rng.integers(3, 14))):
        c = (int(rng.integers(10, w - 10)), int(rng.integers(10, h - 10)))
        t = rng.integers(0, 3)
        if t == 0:
            cv2.ellipse(m, c, (int(rng.integers(1, 25)), int(rng.integers(1, 12))), float(rng.uniform(0, 180)), 0, 360, 1, -1)
        elif t == 1:
            r = cv2.boxPoints((c, (float(rng.uniform(1, 50)), float(rng.uniform(1, 16))), float(rng.uniform(-30, 30))))
            cv2.fillPoly(m, [r.astype(np.int32)], 1)
        else:
            cv2.fillPoly(m, [(rng.integers(-15, 15, (3, 2)) + np.array(c)).astype(np.int32)], 1)
    for _ in range(int(rng.integers(0, 6))):
        c = (int(rng.integers(10, w - 10)), int(rng.integers(10, h - 10)))
        cv2.ellipse(m, c, (int(rng.integers(1, 6)), int(rng.integers(1, 4))), float(rng.uniform(0, 180)), 0, 360, 0, -1)
    return m


def test_min_area_rect_matches_cv2_on_contours_and_holes(G):
    """Component pixel set -> row extremes -> hull -> calipers == cv2.minAreaRect(contour), including
    which of several equal-area rectangles cv2 returns (hull start convention)."""
    rng = np.random.default_rng(3)
    seen = {0: 0, 1: 0}
    for _ in range(60):
        m = _random_mask(rng)
        h, w = m.shape
        cs, hier = cv2.findContours(m * 255, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
        if hier is None:
            continue
        _, labf = cv2.connectedComponents(m, connectivity=8)
        _, labb = cv2.connectedComponents(1 - m, connectivity=4)
        border = set(np.unique(np.concatenate([labb[0], labb[-1], labb[:, 0], labb[:, -1]])).tolist())
        for ci, c in enumerate(cs):
            hole = hier[0][ci][3] != -1
            x0, y0 = c[0, 0]
            if not hole:
                ys, xs = np.nonzero(labf == labf[y0, x0])
                mode, sx, sy = 1, 0, 0
            else:
                cand = None
                for (px, py) in c[:, 0, :]:
                    nb = {int(labb[py + dy, px + dx]) for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1))
                          if 0 <= py + dy < h and 0 <= px + dx < w and m[py + dy, px + dx] == 0}
                    cand = nb if cand is None else cand & nb
                cand -= border
                if len(cand) != 1:
                    continue
                hm = labb == cand.pop()
                adj = np.zeros_like(hm)
                adj[1:] |= hm[:-1]; adj[:-1] |= hm[1:]; adj[:, 1:] |= hm[:, :-1]; adj[:, :-1] |= hm[:, 1:]
                ys, xs = np.nonzero(adj & (m > 0))
                hy, hx = np.argwhere(hm)[0]
                mode, sx, sy = 2, int(hx) - 1, int(hy)
            pts = np.ascontiguousarray(np.stack([xs, ys], 1).astype(np.int32))
            ref = cv2.minAreaRect(c)
            out = np.zeros(5, np.float32)
            G.geom_min_area_rect_ex(P(pts), len(pts), mode, sx, sy, P(out))
            refv = np.array([ref[0][0], ref[0][1], ref[1][0], ref[1][1], ref[2]], np.float32)
            assert np.array_equal(out.view(np.uint32), refv.view(np.uint32)), (ref, out.tolist())   # to the bit
            seen[int(hole)] += 1
    assert seen[0] > 200 and seen[1] > 10


def test_min_area_rect_and_box_points_bit_equal_to_cv2_on_random_hulls(G):
    """dbg_min_area_rect on cv::convexHull's vertex order == cv2.minAreaRect to the last float bit (centre, size and
    angle: cv2 4.13 picks the calipers side by exact cross products and reports the rectangle in the frame whose
    angle lies in [-90, 0)), and dbg_box_points == cv2.boxPoints."""
    rng = np.random.default_rng(1)
    n = 0
    for _ in range(30000):
        k = int(rng.integers(3, 14))
        pts = rng.integers(0, int(rng.integers(4, 300)), (k, 2)).astype(np.int32)
        hull = cv2.convexHull(pts.reshape(-1, 1, 2))
        if len(hull) < 3:
            continue
        ref = cv2.minAreaRect(pts.reshape(-1, 1, 2))
        refv = np.array([ref[0][0], ref[0][1], ref[1][0], ref[1][1], ref[2]], np.float32)
        hp = np.ascontiguousarray(hull[:, 0, :].astype(np.int32))
        out = np.zeros(5, np.float32)
        G.geom_min_area_rect_hull(P(hp), len(hp), P(out))
        assert np.array_equal(out.view(np.uint32), refv.view(np.uint32)), (hp.tolist(), ref, out.tolist())
        bp = np.zeros(8, np.float32)
        G.geom_box_points(P(refv), P(bp))
        assert np.array_equal(bp.reshape(4, 2).view(np.uint32), cv2.boxPoints(ref).view(np.uint32)), (ref, bp.tolist())
        n += 1
    assert n > 25000


def test_clipper_offset_matches_oracle_restatement(G):
    from oracle import db_post as D

    rng = np.random.default_rng(2)
    for _ in range(500):
        rect = cv2.boxPoints(((rng.uniform(50, 900), rng.uniform(50, 900)), (rng.uniform(3, 300), rng.uniform(3, 60)),
                              rng.uniform(-90, 90)))
        box, _ = D.get_mini_boxes(rect.reshape(-1, 1, 2))
        box = np.ascontiguousarray(np.array(box, np.float32))
        ratio = float(rng.choice([1.5, 1.6, 2.0]))
        dist = G.geom_unclip_distance(P(box), C.c_double(ratio))
        ref = D.clipper_offset_round([(float(p[0]), float(p[1])) for p in box], dist)
        o = np.zeros((512, 2), np.int32)
        m = G.geom_clipper_offset(P(box), C.c_double(dist), P(o), 512)
        assert m == len(ref) and np.array_equal(o[:m], np.array(ref, np.int32).reshape(-1, 2))
        ex = D.unclip(box, ratio)
        assert ex is not None and len(ex) == m


def test_candidate_chain_matches_oracle(G):
    """pixel set -> final int32 box + score, vs the cv2-based upstream restatement, per contour."""
    from oracle import db_post as D

    for seed, nb in ((7, 150), (8, 200), (9, 120)):
        _candidate_chain(G, D, D.synth_prob_map(480, 640, seed, n_boxes=nb))


def _candidate_chain(G, D, pred):
    h, w = pred.shape
    mask = (pred > np.float32(0.3)).astype(np.uint8)
    cs, hier = cv2.findContours(mask * 255, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
    _, labf = cv2.connectedComponents(mask, connectivity=8)
    tot = 0
    for ci, c in enumerate(cs):
        if hier[0][ci][3] != -1:
            continue
        x0, y0 = c[0, 0]
        ys, xs = np.nonzero(labf == labf[y0, x0])
        pts = np.ascontiguousarray(np.stack([xs, ys], 1).astype(np.int32))
        # upstream steps for THIS contour (boxes_from_bitmap's loop body)
        boxes, scores = [], []
        points, sside = D.get_mini_boxes(c)
        if sside >= 3:
            points = np.array(points)
            score = D.box_score_fast(pred, points.reshape(-1, 2))
            if not (0.6 > score):
                ex = D.unclip(points, 1.5)
                if ex is not None:
                    box, ss = D.get_mini_boxes(ex.reshape(-1, 1, 2))
                    if ss >= 5:
                        box = np.array(box)
                        box[:, 0] = np.clip(np.round(box[:, 0] / w * np.float64(w)), 0, w)
                        box[:, 1] = np.clip(np.round(box[:, 1] / h * np.float64(h)), 0, h)
                        boxes.append(box.astype("int32")); scores.append(score)
        out8 = np.zeros(8, np.int32); sc = C.c_double(); b8 = np.zeros(8, np.float32); why = C.c_int()
        ok = G.geom_candidate(P(pts), len(pts), 0, 0, 0, P(np.ascontiguousarray(pred)), h, w, C.c_double(0.6),
                              C.c_double(1.5), 3, C.c_double(w), C.c_double(h), P(out8), C.byref(sc), P(b8), C.byref(why))
        tot += 1
        assert ok == (len(boxes) == 1), (ci, ok, len(boxes), why.value)
        if ok:
            assert abs(sc.value - scores[0]) <= 1e-6
            # same four vertices in the same order (get_mini_boxes' tl, tr, br, bl)
            assert np.array_equal(boxes[0].reshape(-1, 2), out8.reshape(-1, 2)), (ci, boxes[0].tolist(), out8.tolist())
    assert tot > 100


def test_fillpoly_of_a_contour_is_the_enclosed_region():
    """What score_mode='slow' rests on (k_dbpost.cu db_cand_score_slow_kernel): cv2.fillPoly of a findContours border
    covers, for an outer border, everything that cannot leave the component's outline over non-component pixels, and
    for a hole border, the border pixels plus the 4-connected non-component region that holds the hole."""
    import cv2
    from scipy import ndimage as ndi

    rng = np.random.default_rng(1)
    s8, s4 = np.ones((3, 3), int), np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    total = holes = 0
    for _ in range(120):
        h, w = rng.integers(20, 80, 2)
        mask = (ndi.gaussian_filter(rng.random((h, w)), rng.uniform(0.6, 2.5)) > rng.uniform(0.45, 0.55)).astype(np.uint8)
        cs, hier = cv2.findContours(mask, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
        if hier is None:
            continue
        fg, _ = ndi.label(mask, s8)
        bg, _ = ndi.label(1 - mask, s4)
        for c, hh in zip(cs, hier[0]):
            pts = c.reshape(-1, 2)
            fill = np.zeros((h, w), np.uint8)
            cv2.fillPoly(fill, [pts.astype(np.int32)], 1)
            comp = fg[pts[0, 1], pts[0, 0]]
            not_c = fg != comp
            if hh[3] < 0:                                   # outer border
                lab, _ = ndi.label(np.pad(not_c, 1, constant_values=True), s4)
                region = ~(lab == lab[0, 0])[1:-1, 1:-1]
                assert np.array_equal(region, fill > 0)
            else:                                           # hole border: find the hole among the bg neighbours
                lab, _ = ndi.label(not_c, s4)
                ok = False
                x, y = pts[0]
                for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                    xx, yy = x + dx, y + dy
                    if 0 <= xx < w and 0 <= yy < h and not mask[yy, xx]:
                        hm = bg == bg[yy, xx]
                        region = (ndi.binary_dilation(hm, s4) & (mask > 0)) | (lab == lab[yy, xx])
                        ok |= np.array_equal(region, fill > 0)
                assert ok
                holes += 1
            total += 1
    assert total > 2000 and holes > 200


def test_clipper_offset_is_the_round_join_offset_of_the_quad():
    """pyclipper is not in this image, so the Clipper restatement (oracle/db_post.py, which db_geom.h is held to above) is
    "parity unpinned" against the library.  What does not need the library is the geometry it must produce -- the
    Minkowski sum of the quad with a disc of radius d, joins rounded (JT_ROUND, arc tolerance 0.25), coordinates rounded
    to integers: every vertex lies d from the quad (never a mitred or squared corner, never short of d), the arcs are
    there (the point d along each corner's bisector lies on the outline), and the area is Steiner's A + P d + pi d^2 up
    to the integer rounding of the outline."""
    from oracle import db_post as D

    def seg_dist(p, poly):
        a, b = poly, np.roll(poly, -1, axis=0)
        ab = b - a
        t = np.clip(((p - a) * ab).sum(1) / np.maximum((ab * ab).sum(1), 1e-30), 0.0, 1.0)
        return float(np.min(np.hypot(*(p - (a + t[:, None] * ab)).T)))

    def shoelace(q):
        return 0.5 * abs(float(np.dot(q[:, 0], np.roll(q[:, 1], -1)) - np.dot(q[:, 1], np.roll(q[:, 0], -1))))

    rng = np.random.default_rng(4)
    checked = 0
    for _ in range(400):
        rect = cv2.boxPoints(((rng.uniform(100, 800), rng.uniform(100, 800)), (rng.uniform(8, 400), rng.uniform(6, 80)),
                              rng.uniform(-90, 90)))
        box = np.round(rect).astype(np.float64)
        area = shoelace(box)
        length = float(np.sum(np.hypot(*(np.roll(box, -1, axis=0) - box).T)))
        if area < 4:
            continue
        ratio = float(rng.choice([1.5, 1.6, 2.0]))
        d = area * ratio / length
        ex = D.unclip(box, ratio).astype(np.float64)
        dist = np.array([seg_dist(p, box) for p in ex])
        assert d - 1.0 <= dist.min() and dist.max() <= d + 0.75, (d, dist.min(), dist.max())
        centre = box.mean(0)
        if d >= 5:
            for k in range(4):
                e0 = box[k] - box[k - 1]
                e1 = box[k] - box[(k + 1) % 4]
                u = e0 / np.linalg.norm(e0) + e1 / np.linalg.norm(e1)
                u /= np.linalg.norm(u)
                assert np.dot(u, box[k] - centre) > 0
                assert seg_dist(box[k] + d * u, ex) <= 1.0, (d, k)     # a chord instead of the arc would be 0.29 d away
        per = float(np.sum(np.hypot(*(np.roll(ex, -1, axis=0) - ex).T)))
        assert abs(shoelace(ex) - (area + length * d + np.pi * d * d)) <= 0.75 * per
        checked += 1
    assert checked > 300
