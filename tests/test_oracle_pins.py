"""CPU pins for the parts of the oracle that the reference itself cannot pin.

a15 det resize+normalize and a17 CTC greedy decode are upstream PaddleOCR ops that the reference neither vendors
nor calls: "parity unpinned" by the reference.  What CAN be pinned without it is pinned here:
the C restatement against the libraries upstream builds on (cv2.resize + NumPy; NumPy argmax/max/mean), and the
drop-in goldens produced by the unmodified reference against the oracle's composition of its own stages.
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from dropin_images import MAX_DIM, image_in_mode  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "dropin_golden.json")))


@pytest.mark.parametrize("h,w", [(960, 678), (2000, 1413), (501, 333), (64, 1999), (31, 17)])
def test_det_resize_normalize_equals_cv2_resize_plus_numpy(oracle, h, w):
    """upstream DetResizeForTest('max', 960) + NormalizeImage + ToCHWImage, exact."""
    import cv2

    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got, shape = oracle.det_resize_normalize(img, 960)
    ratio = 960.0 / max(h, w) if max(h, w) > 960 else 1.0
    rh, rw = int(h * ratio), int(w * ratio)
    rh, rw = max(int(round(rh / 32) * 32), 32), max(int(round(rw / 32) * 32), 32)
    r = cv2.resize(img, (rw, rh))
    mean = np.array([0.485, 0.456, 0.406], np.float32).reshape(1, 1, 3)
    std = np.array([0.229, 0.224, 0.225], np.float32).reshape(1, 1, 3)
    want = ((r.astype("float32") * np.float32(1.0 / 255.0) - mean) / std).transpose(2, 0, 1)
    assert got.shape == want.shape == (3, rh, rw)
    assert np.array_equal(got, want)                      # bit-equal float32
    assert tuple(shape) == (h, w, rh / h, rw / w)


@pytest.mark.parametrize("limit_type,limit,h,w", [("min", 736, 300, 200), ("min", 736, 501, 1333), ("min", 960, 64, 1999),
                                                 ("resize_long", 960, 333, 501), ("resize_long", 640, 2000, 1413),
                                                 ("resize_long", 1280, 31, 17), ("min", 64, 7, 5)])
def test_det_resize_normalize_other_limit_types_equal_cv2_resize_plus_numpy(oracle, limit_type, limit, h, w):
    """upstream's "min" / "resize_long" limit types enlarge images: the same cv2.resize(INTER_LINEAR) + NumPy float
    stage, exact, with the target size from a literal restatement of resize_image_type0."""
    import cv2

    rng = np.random.default_rng(h * 11 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got, shape = oracle.det_resize_normalize(img, limit, limit_type)
    rh, rw = oracle.det_target_size_upstream(h, w, limit, limit_type)
    assert rh % 32 == 0 and rw % 32 == 0 and (limit_type != "min" or min(rh, rw) >= min(h, w))
    r = cv2.resize(img, (int(rw), int(rh)))
    mean = np.array([0.485, 0.456, 0.406], np.float32).reshape(1, 1, 3)
    std = np.array([0.229, 0.224, 0.225], np.float32).reshape(1, 1, 3)
    want = ((r.astype("float32") * np.float32(1.0 / 255.0) - mean) / std).transpose(2, 0, 1)
    assert got.shape == want.shape == (3, rh, rw)
    assert np.array_equal(got, want)
    assert tuple(shape) == (h, w, rh / h, rw / w)


def test_ctc_greedy_equals_numpy(oracle):
    """upstream CTCLabelDecode: argmax (first max wins), max, drop repeats / blank, mean."""
    rng = np.random.default_rng(0)
    n, t, c = 37, 40, 211
    logits = rng.normal(size=(n, t, c)).astype(np.float32) * 3
    # planted repeats, blanks and exact ties
    logits[:, 5] = logits[:, 4]
    logits[:, 9, 0] = 50
    logits[:, 10, 0] = 50
    p = np.exp(logits - logits.max(-1, keepdims=True))
    p = (p / p.sum(-1, keepdims=True)).astype(np.float32)
    p[:, 20, 7] = p[:, 20, 3] = p[:, 20].max(-1) + np.float32(0.25)      # tie: index 3 wins
    idx, pos, ln, conf = oracle.ctc_greedy(p)
    am, mx = p.argmax(2), p.max(2)
    for b in range(n):
        keep = [k for k in range(t) if am[b, k] != 0 and (k == 0 or am[b, k] != am[b, k - 1])]
        assert ln[b] == len(keep)
        assert idx[b, :ln[b]].tolist() == am[b, keep].tolist()
        assert pos[b, :ln[b]].tolist() == keep
        want = float(np.mean(mx[b, keep])) if keep else 0.0
        assert abs(float(conf[b]) - want) <= 1e-6


def test_resize_nearest_equals_pillow_and_reference_golden(oracle):
    from PIL import Image

    rng = np.random.default_rng(2)
    for (h, w, oh, ow) in [(877, 620, 600, 424), (3508, 2480, 960, 678), (100, 333, 47, 200), (17, 13, 5, 4)]:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        im = Image.frombytes("P", (w, h), a.tobytes())
        want = np.frombuffer(im.resize((ow, oh), Image.Resampling.LANCZOS).tobytes(), np.uint8).reshape(oh, ow)
        assert np.array_equal(oracle.resize_nearest(a, ow, oh), want)
    # the reference's own resize_if_needed on a mode-P page (golden from the unmodified module)
    import hashlib

    case = [c for c in GOLD["cases"] if c["kind"] == "mode" and c["mode"] == "P" and c["method"] == "resize_if_needed"][0]
    p = image_in_mode(oracle, "P", 0)
    idx = np.frombuffer(p.tobytes(), np.uint8).reshape(p.size[1], p.size[0])
    tw, th = oracle.target_size(p.size[0], p.size[1], MAX_DIM)
    assert [tw, th] == case["want"]["size"]
    assert hashlib.sha256(oracle.resize_nearest(idx, tw, th).tobytes()).hexdigest() == case["want"]["sha"]


@pytest.mark.parametrize("case", [c for c in GOLD["cases"] if c["kind"] == "optimize_for_ocr"],
                         ids=lambda c: f"{c['mode']}-{c['seed']}-{'+'.join(c['kwargs'])}")
def test_oracle_composition_reproduces_reference_optimize_for_ocr_flags(oracle, case):
    """reference optimize_for_ocr :191-242 with every flag, restated as a composition of oracle stages."""
    import hashlib

    kw = dict(apply_contrast=True, apply_sharpness=True, apply_denoise=False, grayscale=False)
    kw.update(case["kwargs"])
    img = np.asarray(image_in_mode(oracle, case["mode"], case["seed"]))
    tw, th = oracle.target_size(img.shape[1], img.shape[0], MAX_DIM)
    x = oracle.resize_lanczos(img, tw, th)
    if kw["grayscale"] and x.ndim == 3:
        x = oracle.gray_pil(x)
    if kw["apply_denoise"]:
        x = oracle.median3(x)
    if kw["apply_contrast"]:
        x = oracle.contrast(x, 1.2)
    if kw["apply_sharpness"]:
        x = oracle.sharpness(x, 1.1)
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() == case["want"]["sha"]


def test_otsu_restatement_equals_cv2(oracle):
    import cv2

    rng = np.random.default_rng(4)
    imgs = [oracle.gray_pil(oracle.synth_page(300, 220, s)) for s in range(3)]
    imgs += [rng.integers(0, 256, (64, 80), dtype=np.uint8), np.full((20, 20), 77, np.uint8),
             (rng.random((90, 70)) < 0.3).astype(np.uint8) * 200 + 20]
    for g in imgs:
        t, mask = cv2.threshold(g, 0, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
        assert oracle.otsu_threshold(g) == int(t)


ANGLE_GOLD = json.load(open(os.path.join(HERE, "golden", "angle_golden.json")))


def _numpy_simd_tag():
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as feats
        return bool(feats.get("AVX512_SKX"))   # numpy's bundled SVML loops are built for AVX512_SKX
    except Exception:  # noqa: BLE001
        return False


def test_deskew_angle_is_numpys_arctan2_not_glibcs(oracle):
    """The reference's per-segment angle is np.degrees(np.arctan2(dy, dx)) (image_preprocessing.py:421).  Goldens from
    the unmodified reference (tests/golden/make_angle_golden.py) include pages whose median segment is one where
    numpy's SIMD arctan2 and glibc's atan2 differ in the last place: the oracle must give numpy's value.  The golden
    values are those of the numpy build / CPU dispatch recorded in the file; on a host whose numpy dispatches
    differently the reference itself would give other last digits, so there the test only checks self-consistency."""
    import numpy as np

    same_dispatch = np.__version__ == ANGLE_GOLD["numpy"] and _numpy_simd_tag() == ANGLE_GOLD["numpy_avx512_skx"]
    assert ANGLE_GOLD["differing"], "the golden set must contain pages where numpy and glibc disagree"
    for c in ANGLE_GOLD["cases"]:
        small = oracle.resize_lanczos(oracle.synth_page(c["h"], c["w"], c["seed"]), *_target(c["w"], c["h"], c["max_dim"]))
        _, angle, lines = oracle.deskew(small)
        assert len(lines) == c["n_lines"]
        if same_dispatch:
            assert float(angle).hex() == c["angle_hex"], c["seed"]
        # numpy's vectorised arctan2 (the product's host layer) == its scalar one (the reference's loop), always
        from ocr_system_b200 import ops

        la = ops.line_angles(lines)
        ref = []
        for x1, y1, x2, y2 in lines:
            a = np.degrees(np.arctan2(y2 - y1, x2 - x1))
            ref.append(a + 90 if a < -45 else (a - 90 if a > 45 else a))
        assert np.array_equal(la, np.array(ref))
        got, _, _ = ops.deskew_decide(lines[None], np.array([len(lines)], np.int32), small.shape[0], small.shape[1])
        assert got[0] == (angle if abs(angle) <= 45 else 0.0)


def _target(w, h, md):
    if max(w, h) <= md:
        return w, h
    return (md, int(h * md / w)) if w > h else (int(w * md / h), md)


def test_rotation_matrix_equals_cv2_getRotationMatrix2D(oracle):
    """image_preprocessing.py:442: the matrix the reference hands to warpAffine (OpenCV evaluates cos / sin with the C
    library, unlike numpy's arctan2 above) -- oracle restatement and the library's host entry, 5000 angles, bit-equal."""
    import cv2
    import numpy as np
    from ocr_system_b200 import ops

    rng = np.random.default_rng(5)
    for a in np.concatenate([rng.uniform(-45, 45, 5000), [0.5, -0.5, 45.0, -45.0, 1e-9]]):
        for cx, cy in ((339, 480), (141, 200)):
            m = cv2.getRotationMatrix2D((cx, cy), float(a), 1.0)
            assert np.array_equal(m, oracle.rotation_matrix(cx, cy, float(a)))
            assert np.array_equal(m, ops.rotation_matrix(cx, cy, float(a)))


def test_adaptive_threshold_oracle_under_both_opencv_dispatch_modes(oracle):
    """The float Gaussian of cv2.adaptiveThreshold is the one dispatch-dependent OpenCV step of the path.  Goldens from
    the unmodified reference in both modes (tests/golden/make_adaptive_golden.py); additionally the live cv2 of this
    process in whichever mode it is in."""
    import hashlib

    import cv2
    import numpy as np
    from adaptive_inputs import plane
    from conftest import cv2_dispatch

    with open(os.path.join(HERE, "golden", "adaptive_dispatch_golden.json")) as f:
        gold = json.load(f)
    for c in gold["cases"]:
        g = plane(c["seed"])
        for mode, key in (("avx2", "sha_default"), ("plain", "sha_plain")):
            got = oracle.adaptive_gauss11(g, 2, cv_dispatch=mode)
            assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == c[key], (c["seed"], mode)
        live = cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)
        assert np.array_equal(live, oracle.adaptive_gauss11(g, 2, cv_dispatch=cv2_dispatch())), c["seed"]


TALL = [(4, 2212, 3, 2000), (8, 800, 6, 723), (8, 801, 6, 724), (8, 801, 6, 801), (8, 801, 6, 900), (8, 801, 12, 700),
        (20, 2001, 15, 1000), (20, 2000, 15, 1000), (1, 101, 1, 50), (2, 201, 1, 100), (3000, 8, 2000, 6), (5, 530, 2, 300)]


@pytest.mark.parametrize("w,h,tw,th", TALL)
def test_resize_pass_order_of_very_tall_images_equals_pillow(oracle, w, h, tw, th):
    """Pillow (12.2.0) runs the vertical pass FIRST when both passes are needed, h > 100 * w and the image shrinks
    vertically; the uint8 intermediate makes that visible (found by the degenerate-shape sweep of the drop-in against
    the real reference: 4 x 2212 -> 3 x 2000).  Live Pillow is the reference's own resampler (image_preprocessing.py:110)."""
    import numpy as np
    from PIL import Image

    rng = np.random.default_rng(w * 7919 + h)
    for c in (1, 3):
        a = rng.integers(0, 256, (h, w, 3) if c == 3 else (h, w), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(a).resize((tw, th), Image.Resampling.LANCZOS))
        assert np.array_equal(oracle.resize_lanczos(a, tw, th), ref), (w, h, tw, th, c)



def test_sauvola_restatement_equals_the_per_pixel_definition(oracle):
    """No library in this image implements Sauvola and the reference has no such operator ("parity unpinned"): what can
    be pinned is that the integral-image restatement (which the GPU kernel is compared with) equals the operator's
    definition evaluated window by window -- mean and population standard deviation of the window clipped to the page,
    T = m (1 + k (s / R - 1)), pixel > T -> 255 -- including windows larger than the page and 1-pixel pages."""
    rng = np.random.default_rng(5)
    for (h, w, window, k, r) in [(23, 31, 5, 0.2, 128.0), (17, 9, 25, 0.34, 128.0), (40, 40, 15, 0.5, 64.0), (1, 1, 25, 0.2, 128.0),
                                 (3, 50, 7, 0.1, 128.0)]:
        g = rng.integers(0, 256, (h, w), dtype=np.uint8)
        g[: h // 2] = (g[: h // 2] // 64) * 64          # flat-ish areas: ties of pixel and threshold are likely
        got = oracle.sauvola(g, window, k, r)
        rad = window // 2
        want = np.zeros_like(g)
        near_tie = np.zeros(g.shape, bool)
        for y in range(h):
            for x in range(w):
                win = g[max(0, y - rad): y + rad + 1, max(0, x - rad): x + rad + 1].astype(np.float64)
                m, s = win.mean(), win.std()            # numpy's two-pass population std: an independent evaluation
                t = m * (1.0 + k * (s / r - 1.0))
                want[y, x] = 255 if float(g[y, x]) > t else 0
                near_tie[y, x] = abs(float(g[y, x]) - t) < 1e-9
        assert np.array_equal(got[~near_tie], want[~near_tie]), (h, w, window)
        assert near_tie.mean() < 0.05
