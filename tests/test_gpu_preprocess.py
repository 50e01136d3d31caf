"""GPU parity: every preprocessing kernel (through the C-ABI) vs the CPU oracle.

Bit-exact for all integer / byte work; the adaptive threshold is compared
bit-exactly too (the oracle restates OpenCV's plain, non-SIMD summation order).
Reference call sites: backend/utils/image_preprocessing.py (line numbers in the
kernels' headers)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _t(a, dev):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _page(O, h, w, seed):
    return O.synth_page(h, w, seed)


def _rand(h, w, c, seed):
    rng = np.random.default_rng(seed)
    shape = (h, w, c) if c > 1 else (h, w)
    return rng.integers(0, 256, size=shape, dtype=np.uint8)


def test_synth_device_equals_host(oracle, cuda):
    from ocr_system_b200 import ops

    dev = ops.synth_pages(3, 877, 620, seed0=5, device=cuda).cpu().numpy()
    for i in range(3):
        assert np.array_equal(dev[i], oracle.synth_page(877, 620, 5 + i))


@pytest.mark.parametrize("h,w,md", [(3508, 2480, 960), (3508, 2480, 2000), (2480, 3508, 960), (1200, 850, 600),
                                    (501, 333, 200), (2000, 1413, 1999), (700, 5000, 960)])
def test_resize_lanczos_rgb(oracle, cuda, h, w, md):
    from ocr_system_b200 import ops

    img = _page(oracle, h, w, 1) if h * w > 10**6 else _rand(h, w, 3, 7)
    tw, th = oracle.target_size(w, h, md)
    assert ops.target_size(w, h, md) == (tw, th)
    got = ops.resize_lanczos(_t(img[None], cuda), tw, th).cpu().numpy()[0]
    assert np.array_equal(got, oracle.resize_lanczos(img, tw, th))


@pytest.mark.parametrize("n,h,w,tw,th", [(1, 1024, 1600, 600, 384), (1, 777, 1264, 333, 205), (1, 900, 640, 470, 661),
                                         (3, 1200, 848, 300, 424), (1, 2150, 1600, 372, 500), (2, 2300, 1600, 347, 499),
                                         (2, 640, 1008, 401, 255)])
def test_resize_lanczos_tensor_core_shapes(oracle, cuda, n, h, w, tw, th):
    """Widths that are multiples of 16 take the bulk-staged tensor-core kernel (k_resize.cu): scales 1.36-4.6 (one and
    two K steps; the largest falls back to the staged kernel because a vertical tile's taps exceed 64 rows), output
    sizes that are not multiples of 8 / 64, batches (the last rows of the batch are cut at the end of the buffer)."""
    from ocr_system_b200 import ops

    imgs = np.stack([_rand(h, w, 3, 11 + i) for i in range(n)])
    got = ops.resize_lanczos(_t(imgs, cuda), tw, th).cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], oracle.resize_lanczos(imgs[i], tw, th)), i


@pytest.mark.parametrize("env", ["", "LUMINA_RESIZE_NO_TMA", "LUMINA_RESIZE_STAGED", "LUMINA_RESIZE_DP4A"])
def test_resize_lanczos_kernel_variants_and_batch_slices(oracle, cuda, env, monkeypatch):
    """The shipped kernel stages chunks with one TMA tensor copy; the per-row bulk-copy form (spans wider than a TMA box),
    the plane-staged tensor-core kernel and the dp4a kernel stay as fallbacks -- all four give the reference's bytes, also
    on a batch that starts inside a larger allocation (the tensor map is built on the slice) and on the last page of the
    allocation (copies / boxes that reach past the end)."""
    from ocr_system_b200 import ops

    if env:
        monkeypatch.setenv(env, "1")
    imgs = np.stack([_rand(700, 1008, 3, 31 + i) for i in range(3)])
    x = _t(imgs, cuda)
    for sl, (tw, th) in ((slice(1, 3), (275, 191)), (slice(0, 3), (575, 399)), (slice(2, 3), (230, 160))):
        got = ops.resize_lanczos(x[sl], tw, th).cpu().numpy()
        for k, i in enumerate(range(*sl.indices(3))):
            assert np.array_equal(got[k], oracle.resize_lanczos(imgs[i], tw, th)), (env, sl, i)


def test_resize_lanczos_gray_batch_and_one_axis(oracle, cuda):
    from ocr_system_b200 import ops

    imgs = np.stack([_rand(640, 480, 1, s) for s in range(3)])
    got = ops.resize_lanczos(_t(imgs, cuda), 201, 268).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], oracle.resize_lanczos(imgs[i], 201, 268))
    rgb = _rand(300, 400, 3, 3)
    for (ow, oh) in ((400, 123), (177, 300), (40, 30)):  # one-axis and 10x reductions (generic path)
        got = ops.resize_lanczos(_t(rgb[None], cuda), ow, oh).cpu().numpy()[0]
        assert np.array_equal(got, oracle.resize_lanczos(rgb, ow, oh)), (ow, oh)


@pytest.mark.parametrize("h,w", [(960, 678), (101, 77), (33, 1000)])
def test_gray_and_binarize(oracle, cuda, h, w):
    from ocr_system_b200 import ops

    img = np.stack([_rand(h, w, 3, s) for s in range(2)])
    x = _t(img, cuda)
    gp, gc, bz = ops.gray_pil(x).cpu().numpy(), ops.gray_cv(x).cpu().numpy(), ops.binarize(x, 128).cpu().numpy()
    for i in range(2):
        assert np.array_equal(gp[i], oracle.gray_pil(img[i]))
        assert np.array_equal(gc[i], oracle.gray_cv(img[i]))
        assert np.array_equal(bz[i], oracle.threshold(oracle.gray_pil(img[i]), 128))
    g = _t(oracle.gray_pil(img[0])[None], cuda)
    assert np.array_equal(ops.binarize(g, 100).cpu().numpy()[0], oracle.threshold(oracle.gray_pil(img[0]), 100))


@pytest.mark.parametrize("h,w,c", [(960, 678, 3), (101, 77, 3), (101, 77, 1), (64, 3, 3), (3, 50, 1), (2000, 1413, 3)])
def test_contrast_sharpness_median(oracle, cuda, h, w, c):
    from ocr_system_b200 import ops

    n = 2 if h < 1000 else 1
    img = np.stack([(_page(oracle, h, w, s) if (c == 3 and h >= 900) else _rand(h, w, c, s)) for s in range(n)])
    x = _t(img, cuda)
    mean = ops.contrast_mean(x).cpu().numpy()
    for f in (1.2, 1.3, 0.6):
        got = ops.enhance_contrast(x, f).cpu().numpy()
        for i in range(n):
            assert mean[i] == oracle.contrast_mean(img[i])
            assert np.array_equal(got[i], oracle.contrast(img[i], f)), ("contrast", f)
    for f in (1.1, 1.2, 0.4, 2.5):
        got = ops.enhance_sharpness(x, f).cpu().numpy()
        for i in range(n):
            assert np.array_equal(got[i], oracle.sharpness(img[i], f)), ("sharp", f)
    got = ops.contrast_sharpness(x, 1.2, 1.1).cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], oracle.sharpness(oracle.contrast(img[i], 1.2), 1.1))
    got = ops.median3(x).cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], oracle.median3(img[i]))


@pytest.mark.parametrize("dispatch", ["plain", "avx2"])
@pytest.mark.parametrize("h,w,c", [(960, 678, 3), (101, 77, 1), (130, 67, 3), (11, 9, 1), (64, 70, 1), (40, 132, 3)])
def test_adaptive_binarize(oracle, cuda, h, w, c, dispatch):
    from ocr_system_b200 import ops

    img = np.stack([(_page(oracle, h, w, s) if h >= 900 else _rand(h, w, c, s)) for s in range(2)])
    got = ops.adaptive_binarize(_t(img, cuda), 2, cv_dispatch=dispatch).cpu().numpy()
    for i in range(2):
        g = oracle.gray_pil(img[i]) if c == 3 else img[i]
        assert np.array_equal(got[i], oracle.adaptive_gauss11(g, 2, cv_dispatch=dispatch))
    assert ops.default_cv_dispatch() == "avx2"


def test_adaptive_binarize_equals_the_reference_under_both_opencv_dispatch_modes(oracle, cuda):
    """tests/golden/adaptive_dispatch_golden.json: the unmodified reference's adaptive_binarize under OpenCV's default
    dispatch (AVX2 + FMA3: fused multiply-add in the filter's vector loops) and under the plain path, on planes where
    the two differ, every width residue mod 8 (the scalar remainders are not fused)."""
    import hashlib
    import json
    import os
    import sys

    from ocr_system_b200 import ops

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    from adaptive_inputs import plane

    with open(os.path.join(here, "golden", "adaptive_dispatch_golden.json")) as f:
        gold = json.load(f)
    assert sum(1 for c in gold["cases"] if c["differing_px"]) >= 10
    for c in gold["cases"]:
        g = plane(c["seed"])
        assert g.shape == (c["h"], c["w"])
        for mode, key in (("avx2", "sha_default"), ("plain", "sha_plain")):
            got = ops.adaptive_binarize(_t(g[None], cuda), 2, cv_dispatch=mode).cpu().numpy()[0]
            assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == c[key], (c["seed"], mode)


@pytest.mark.parametrize("orientation", [1, 2, 3, 4, 5, 6, 7, 8])
def test_exif_transpose(oracle, cuda, orientation):
    from ocr_system_b200 import ops

    img = _rand(37, 53, 3, orientation)
    got = ops.exif_transpose(_t(img[None], cuda), orientation).cpu().numpy()[0]
    assert np.array_equal(got, oracle.exif_transpose(img, orientation))


def _resized_page(O, seed, md):
    pg = O.synth_page(3508, 2480, seed)
    tw, th = O.target_size(2480, 3508, md)
    return O.resize_lanczos(pg, tw, th)


@pytest.mark.parametrize("md", [960, 2000])
def test_canny_ppht_deskew(oracle, cuda, md):
    from ocr_system_b200 import ops

    seeds = [0, 1, 2] if md == 960 else [0]
    imgs = np.stack([_resized_page(oracle, s, md) for s in seeds])
    x = _t(imgs, cuda)
    edges = ops.canny(x, 50, 150)
    e = edges.cpu().numpy()
    ref_edges = [oracle.canny(oracle.gray_cv(im), 50, 150) for im in imgs]
    for i in range(len(seeds)):
        assert np.array_equal(e[i], ref_edges[i]), "canny edge map"
    lines, nlines = ops.hough_lines_p(edges)
    lines, nlines = lines.cpu().numpy(), nlines.cpu().numpy()
    for i in range(len(seeds)):
        ref = oracle.ppht(ref_edges[i])
        assert nlines[i] == len(ref)
        assert np.array_equal(lines[i, : nlines[i]], ref), "HoughLinesP line list / order"
    out, angles = ops.deskew(x)
    out = out.cpu().numpy()
    for i in range(len(seeds)):
        ref_img, ref_angle, _ = oracle.deskew(imgs[i])
        assert angles[i] == ref_angle
        assert np.array_equal(out[i], ref_img)


@pytest.mark.parametrize("h,w,c", [(960, 678, 3), (211, 97, 1), (50, 300, 3)])
def test_warp_affine_teacher_forced(oracle, cuda, h, w, c):
    from ocr_system_b200 import ops

    img = _page(oracle, h, w, 3) if c == 3 and h > 900 else _rand(h, w, c, 4)
    angs = [0.5, -0.7, 2.9, -13.0, 44.9]
    mats = np.stack([oracle.rotation_matrix(w // 2, h // 2, a).reshape(6) for a in angs])
    for a, m in zip(angs, mats):
        assert np.array_equal(ops.rotation_matrix(w // 2, h // 2, a).reshape(6), m)
    x = _t(np.stack([img] * len(angs)), cuda)
    apply = np.array([1, 1, 0, 1, 1], np.uint8)
    got = ops.warp_affine_cubic(x, mats, apply).cpu().numpy()
    for i, a in enumerate(angs):
        ref = oracle.warp_affine_cubic(img, mats[i]) if apply[i] else img
        assert np.array_equal(got[i], ref), a


@pytest.mark.parametrize("h,w", [(960, 678), (3508, 2480), (300, 500), (20, 31)])
def test_det_resize_normalize(oracle, cuda, h, w):
    from ocr_system_b200 import ops

    img = _rand(h, w, 3, 9)
    got, shape_list = ops.det_resize_normalize(_t(img[None], cuda), 960)
    ref, sl = oracle.det_resize_normalize(img, 960)
    assert got.shape[1:] == ref.shape
    assert tuple(shape_list[0]) == tuple(sl)
    # float stage tolerance from north_star: <= 1e-4 abs (observed: exact)
    assert np.max(np.abs(got.cpu().numpy()[0] - ref)) <= 1e-4


@pytest.mark.parametrize("n,t,c", [(8, 40, 6625), (3, 7, 13), (2, 200, 97), (5, 1, 2)])
def test_ctc_greedy(oracle, cuda, n, t, c):
    from ocr_system_b200 import ops

    rng = np.random.default_rng(n * 1000 + t)
    p = rng.random((n, t, c)).astype(np.float32)
    p[:, :, 0] += (rng.random((n, t)) < 0.3) * 2.0          # blanks
    if t > 2:
        p[:, 1::3] = p[:, 0:-1:3][:, : p[:, 1::3].shape[1]]   # repeats
    if c > 8:
        p[0, 0, 5] = p[0, 0, 3] = 9.0                         # exact tie -> first index
    idx, pos, ln, conf = [a.cpu().numpy() for a in ops.ctc_greedy(_t(p, cuda))]
    ridx, rpos, rln, rconf = oracle.ctc_greedy(p)
    assert np.array_equal(ln, rln) and np.array_equal(idx, ridx) and np.array_equal(pos, rpos)
    assert np.max(np.abs(conf - rconf)) <= 1e-4


@pytest.mark.parametrize("variant", ["l2", "cluster", "nopipe"])
def test_ppht_fallback_variants_are_exact_too(oracle, cuda, variant, monkeypatch):
    """The default HoughLinesP kernel keeps accumulator + edge bitmask in (distributed) shared memory;
    larger pages fall back to a cluster kernel with the mask in L2, then to L2 atomics.  All three
    must give cv2's exact line list."""
    from ocr_system_b200 import ops

    monkeypatch.setenv("LUMINA_PPHT", variant)
    imgs = np.stack([_resized_page(oracle, s, 960) for s in (0, 3)])
    edges_ref = [oracle.canny(oracle.gray_cv(im), 50, 150) for im in imgs]
    lines, nlines = ops.hough_lines_p(_t(np.stack(edges_ref), cuda))
    lines, nlines = lines.cpu().numpy(), nlines.cpu().numpy()
    for i in range(2):
        ref = oracle.ppht(edges_ref[i])
        assert nlines[i] == len(ref) and np.array_equal(lines[i, : nlines[i]], ref)


def test_ppht_degenerate_inputs(oracle, cuda):
    from ocr_system_b200 import ops

    h, w = 300, 420
    empty = np.zeros((h, w), np.uint8)
    one_line = empty.copy(); one_line[150, 30:400] = 255
    cross = empty.copy(); cross[40:260, 210] = 255; cross[150, 20:400] = 255
    full = np.full((64, 64), 255, np.uint8)
    for e in (empty, one_line, cross):
        lines, nl = ops.hough_lines_p(_t(e[None], cuda))
        ref = oracle.ppht(e)
        assert int(nl[0]) == len(ref) and np.array_equal(lines[0, : int(nl[0])].cpu().numpy(), ref)
    lines, nl = ops.hough_lines_p(_t(full[None], cuda), threshold=20, min_line_length=10, max_line_gap=2)
    ref = oracle.ppht(full, threshold=20, min_len=10, max_gap=2)
    assert int(nl[0]) == len(ref) and np.array_equal(lines[0, : int(nl[0])].cpu().numpy(), ref)


def test_pipelined_stream_equals_sequential_steps(oracle, cuda):
    """run_device_stream (steps overlapped on CUDA streams, HoughLinesP split into prepare + lines) must give
    exactly what run_device gives batch by batch."""
    import torch
    from ocr_system_b200 import ops
    from ocr_system_b200.pipeline import PagePipeline

    batches = [ops.synth_pages(6, 877, 620, seed0=s) for s in (0, 6, 12, 18)]
    pipe = PagePipeline(max_dimension=400)
    want = [pipe.run_device(b) for b in batches]
    got = list(pipe.run_device_stream(batches))
    torch.cuda.synchronize()
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert np.array_equal(g.angles, w.angles)
        for name in ("pages", "gray", "binary", "det_input"):
            assert torch.equal(getattr(g, name), getattr(w, name)), name
    # the two halves of the Hough call on one stream are the whole call
    edges = ops.canny(ops.resize_if_needed(batches[0], 400))
    l1, n1 = ops.hough_lines_p(edges)
    l2, n2 = ops.HoughJob(edges).prepare().lines()
    assert torch.equal(n1, n2)
    for i in range(edges.shape[0]):
        assert torch.equal(l1[i, : int(n1[i])], l2[i, : int(n2[i])])


def test_deskew_angle_equals_the_reference_on_pages_where_numpy_and_glibc_disagree(oracle, cuda):
    """tests/golden/angle_golden.json: angles of the unmodified reference on pages whose median Hough segment is one
    where numpy's arctan2 (SIMD math on AVX-512 builds) and glibc's atan2 differ in the last place.  The device
    chain + numpy host decision must give the reference's float64, through the per-page op and through the pipeline."""
    import json

    import torch
    from ocr_system_b200 import ops
    from ocr_system_b200.pipeline import PagePipeline
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "angle_golden.json")) as f:
        ANGLE_GOLD = json.load(f)
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as feats
        skx = bool(feats.get("AVX512_SKX"))
    except Exception:  # noqa: BLE001
        skx = False
    same_dispatch = np.__version__ == ANGLE_GOLD["numpy"] and skx == ANGLE_GOLD["numpy_avx512_skx"]
    cases = ANGLE_GOLD["cases"]
    pages = np.stack([oracle.synth_page(c["h"], c["w"], c["seed"]) for c in cases])
    res = PagePipeline(max_dimension=cases[0]["max_dim"]).run_device(_t(pages, cuda))
    small = ops.resize_if_needed(_t(pages, cuda), cases[0]["max_dim"])
    _, angles = ops.deskew(small)
    for i, c in enumerate(cases):
        _, want, _ = oracle.deskew(small[i].cpu().numpy())
        assert angles[i] == want and res.angles[i] == want, c["seed"]
        if same_dispatch:
            assert float(angles[i]).hex() == c["angle_hex"], c["seed"]


def test_degenerate_shapes_resize_pass_order_median_and_zero_size(oracle, cuda):
    """What the degenerate-shape sweep against the real reference found (tools/sweep_dropin_vs_reference.py --tiny):
    Pillow's vertical-first pass order on images taller than 100 x their width, the median on planes narrower than
    4 bytes, and the ValueError for a target side that rounds to 0."""
    from test_oracle_pins import TALL
    from ocr_system_b200 import ops

    for (w, h, tw, th) in TALL:
        rng = np.random.default_rng(w * 7919 + h)
        for c in (1, 3):
            a = rng.integers(0, 256, (2, h, w, 3) if c == 3 else (2, h, w), dtype=np.uint8)
            got = ops.resize_lanczos(_t(a, cuda), tw, th).cpu().numpy()
            for i in range(2):
                assert np.array_equal(got[i], oracle.resize_lanczos(a[i], tw, th)), (w, h, tw, th, c)
    for (h, w) in [(10, 1), (687, 1), (5, 2), (1, 1), (1, 3), (7, 3), (1, 2000), (2000, 1)]:
        for c in (1, 3):
            a = np.random.default_rng(h * 31 + w).integers(0, 256, (3, h, w, 3) if c == 3 else (3, h, w), dtype=np.uint8)
            got = ops.median3(_t(a, cuda)).cpu().numpy()
            sharp = ops.contrast_sharpness(_t(a, cuda), 1.2, 1.1).cpu().numpy()
            for i in range(3):
                assert np.array_equal(got[i], oracle.median3(a[i])), (h, w, c)
                assert np.array_equal(sharp[i], oracle.sharpness(oracle.contrast(a[i], 1.2), 1.1)), (h, w, c)
    with pytest.raises(ValueError, match="height and width must be > 0"):
        ops.resize_if_needed(_t(np.zeros((1, 1034, 2, 3), np.uint8), cuda), 16)


def test_ingest_zero_copy_and_fallback_agree(cuda):
    """PIL pages reach the device the same whether Pillow's storage can be viewed through Arrow (one allocator
    block) or has to go through np.asarray (multi-block image, odd L width)."""
    import torch
    from PIL import Image
    from ocr_system_b200 import ops
    from ocr_system_b200.image_preprocessing import ImagePreprocessor

    rng = np.random.default_rng(3)
    x4 = torch.from_numpy(rng.integers(0, 256, (2, 37, 53, 4), dtype=np.uint8)).to(cuda)   # 1961 px: not a multiple of 16
    assert torch.equal(ops.rgbx_to_rgb(x4), x4[..., :3].contiguous())
    ip = ImagePreprocessor(max_dimension=4000)
    big = rng.integers(0, 256, (1200, 900, 3), dtype=np.uint8)
    old = Image.core.get_block_size()
    try:
        Image.core.set_block_size(1 << 20)                 # 1 MB blocks: the 4.3 MB image below is multi-block
        multi = Image.fromarray(big).copy()
    finally:
        Image.core.set_block_size(old)
    single = Image.fromarray(big).copy()
    assert ip._zero_copy_view(multi) is None and ip._zero_copy_view(single) is not None
    for imgs in ([multi, multi], [single, single, single], [single.convert("L")], [single.convert("L").crop((0, 0, 899, 1200))]):
        got = ip._upload(imgs).cpu().numpy()
        want = np.stack([np.asarray(im).reshape(im.size[1], im.size[0], -1) for im in imgs])
        assert np.array_equal(got, want)


def test_otsu_equals_cv2_and_sauvola_equals_the_float64_restatement(oracle, cuda):
    """north_star extras (not reference call sites): Otsu is pinned by cv2 (threshold and mask bit-equal); Sauvola by
    the NumPy float64 restatement with exact integer window sums (windows clipped to the page)."""
    import cv2
    import torch

    from ocr_system_b200 import ops

    rng = np.random.default_rng(5)
    pages = np.stack([oracle.gray_pil(oracle.synth_page(700, 500, s)) for s in range(3)] +
                     [rng.integers(0, 256, (700, 500), dtype=np.uint8)])
    x = torch.from_numpy(pages).to(cuda)
    mask, thr = ops.otsu_binarize(x)
    for i in range(len(pages)):
        t, ref = cv2.threshold(pages[i], 0, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
        assert int(thr[i]) == int(t) == oracle.otsu_threshold(pages[i])
        assert np.array_equal(mask[i].cpu().numpy(), ref)
    for window, k in ((25, 0.2), (15, 0.34), (49, 0.1), (3, 0.5)):
        got = ops.sauvola_binarize(x, window, k, 128.0).cpu().numpy()
        for i in range(len(pages)):
            assert np.array_equal(got[i], oracle.sauvola(pages[i], window, k, 128.0)), (window, k, i)
    odd = torch.from_numpy(np.ascontiguousarray(pages[0][:133, :77])).to(cuda)[None]     # ragged tiles, odd pitch
    assert np.array_equal(ops.sauvola_binarize(odd, 25, 0.2).cpu().numpy()[0], oracle.sauvola(pages[0][:133, :77], 25, 0.2))
    m2, t2 = ops.otsu_binarize(odd)
    assert int(t2[0]) == oracle.otsu_threshold(pages[0][:133, :77])
