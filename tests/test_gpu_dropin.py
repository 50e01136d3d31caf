"""GPU parity of the drop-in surface and of the batched host-buffer API.

* ``ImagePreprocessor`` public methods on PIL objects of every mode, and the flag combinations of
  ``optimize_for_ocr`` / ``preprocess_for_azure`` / the pdf_to_images resize loop, against goldens produced by the
  UNMODIFIED reference module (tests/golden/make_dropin_golden.py -> dropin_golden.json).
* ``PagePipeline.run_host`` / ``run_host_stream`` (the API the headline e2e number is measured through): every
  yielded batch equals the oracle for THAT batch, also while the consumer still holds the previous result.
* thread-safety of the singleton-style objects, device guards, unaligned slices.
"""
import json
import os
import sys
import threading

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from dropin_images import MAX_DIM, digest, image_in_mode, mixed_pdf_pages, run  # noqa: E402

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(HERE, "golden", "dropin_golden.json")))
MODE_CASES = [c for c in GOLD["cases"] if c["kind"] == "mode"]
OCR_CASES = [c for c in GOLD["cases"] if c["kind"] == "optimize_for_ocr"]
AZ_CASES = [c for c in GOLD["cases"] if c["kind"] == "preprocess_for_azure"]

# Documented deviations from the reference (DESIGN.md "PIL modes"): the GPU JPEG encoder writes 3-component YCbCr
# 4:2:0 files only; Pillow also writes CMYK (4 components), raw-YCbCr and grayscale ("1") files.
DEVIATIONS = {("CMYK", "compress_for_azure"): "OSError", ("YCbCr", "compress_for_azure"): "OSError",
              ("1", "compress_for_azure"): "OSError"}

METHODS = {
    "resize_if_needed": lambda ip, im: ip.resize_if_needed(im),
    "enhance_contrast": lambda ip, im: ip.enhance_contrast(im, 1.2),
    "enhance_sharpness": lambda ip, im: ip.enhance_sharpness(im, 1.1),
    "denoise": lambda ip, im: ip.denoise(im),
    "convert_to_grayscale": lambda ip, im: ip.convert_to_grayscale(im),
    "binarize": lambda ip, im: ip.binarize(im, 120),
    "adaptive_binarize": lambda ip, im: ip.adaptive_binarize(im),
    "deskew": lambda ip, im: ip.deskew(im),
    "optimize_for_ocr": lambda ip, im: ip.optimize_for_ocr(im),
    "preprocess_for_azure": lambda ip, im: ip.preprocess_for_azure(im),
    "compress_for_azure": lambda ip, im: ip.compress_for_azure(im, target_size_mb=0.05),
}


@pytest.fixture(scope="module")
def ip(cuda):
    import cv2

    cv2.setUseOptimized(False)
    from ocr_system_b200.image_preprocessing import ImagePreprocessor

    return ImagePreprocessor(max_dimension=MAX_DIM, cv_dispatch="plain")   # the goldens were made with setUseOptimized(False)


@pytest.mark.parametrize("case", MODE_CASES, ids=[f"{c['mode']}-{c['method']}" for c in MODE_CASES])
def test_every_pil_mode_matches_the_reference(oracle, ip, case):
    im = image_in_mode(oracle, case["mode"], 0)
    got = run(lambda: METHODS[case["method"]](ip, im))
    dev = DEVIATIONS.get((case["mode"], case["method"]))
    if dev is not None:
        assert got == {"raises": dev}
        return
    assert got == case["want"]


@pytest.mark.parametrize("case", OCR_CASES, ids=[f"{c['mode']}-{c['seed']}-{'+'.join(c['kwargs'])}" for c in OCR_CASES])
def test_optimize_for_ocr_flags_match_the_reference(oracle, ip, case):
    im = image_in_mode(oracle, case["mode"], case["seed"])
    assert digest(ip.optimize_for_ocr(im, **case["kwargs"])) == case["want"]


@pytest.mark.parametrize("case", AZ_CASES, ids=[f"{c['mode']}-{'+'.join(c['kwargs'])}" for c in AZ_CASES])
def test_preprocess_for_azure_flags_match_the_reference(oracle, ip, case):
    im = image_in_mode(oracle, case["mode"], case["seed"])
    assert digest(ip.preprocess_for_azure(im, **case["kwargs"])) == case["want"]


def test_resize_pages_groups_by_shape_and_matches_per_page_reference(oracle, ip):
    want = [c for c in GOLD["cases"] if c["kind"] == "resize_pages"][0]["want"]
    pages = mixed_pdf_pages(oracle)
    small = pages[3]
    got = ip.resize_pages(pages)
    assert [digest(p) for p in got] == want
    assert got[3] is small                       # a page that needs no resize is passed through (reference :289-292)


# ------------------------------------------------------------------ PagePipeline host-buffer API
def _oracle_chain(O, page, md, enhance):
    tw, th = O.target_size(page.shape[1], page.shape[0], md)
    small = O.resize_lanczos(page, tw, th) if max(page.shape[:2]) > md else page
    img, angle, _ = O.deskew(small)
    if enhance:
        img = O.sharpness(O.contrast(img, 1.2), 1.1)
    g = O.gray_pil(img)
    return img, angle, g, O.adaptive_gauss11(g, 2, cv_dispatch="avx2")   # PagePipeline's default: OpenCV's default dispatch


def _host_batches(O, n_batches, per, h, w, seed0=0):
    import torch

    out = []
    for b in range(n_batches):
        a = np.stack([O.synth_page(h, w, seed0 + b * per + i) for i in range(per)])
        out.append(torch.from_numpy(a).pin_memory())
    return out


@pytest.mark.parametrize("enhance", [False, True])
def test_run_host_stream_every_batch_equals_the_oracle(oracle, cuda, enhance):
    """4 different pinned batches through the double-buffered stream; the consumer keeps batch i-1's host result
    while batch i is produced (keep=2) and checks BOTH after each step."""
    import torch
    from ocr_system_b200.pipeline import PagePipeline

    h, w, md, per = 1754, 1240, 960, 3          # A4 @ 150 dpi -> 678x960 (the bench geometry after the resize)
    batches = _host_batches(oracle, 4, per, h, w)
    want = [[_oracle_chain(oracle, b[i].numpy(), md, enhance) for i in range(per)] for b in batches]
    pipe = PagePipeline(max_dimension=md, enhance=enhance)

    def check(bi, out_host):
        for i in range(per):
            img, angle, g, binary = want[bi][i]
            assert out_host["angles"][i] == angle, (bi, i)
            assert np.array_equal(out_host["pages"][i].numpy(), img), (bi, i, "pages")
            assert np.array_equal(out_host["binary"][i].numpy(), binary), (bi, i, "binary")

    prev = None
    seen = 0
    for bi, (out_host, res, h2d, d2h) in enumerate(pipe.run_host_stream(batches)):
        assert h2d == batches[bi].numel()
        check(bi, out_host)
        if prev is not None:
            check(bi - 1, prev)                  # still intact although the next batch has been produced
            assert prev["pages"].data_ptr() != out_host["pages"].data_ptr()
        # device-side results of the same step
        for i in range(per):
            assert np.array_equal(res.gray[i].cpu().numpy(), want[bi][i][2])
        prev = out_host
        seen += 1
    assert seen == 4
    # run_host (fresh buffers) and run_device agree with the stream for an arbitrary batch
    out_host, res, _, _ = pipe.run_host(batches[2])
    check(2, out_host)
    rd = pipe.run_device(batches[1].to(cuda))
    torch.cuda.synchronize()
    for i in range(per):
        assert rd.angles[i] == want[1][i][1]
        assert np.array_equal(rd.pages[i].cpu().numpy(), want[1][i][0])
    # a second stream on the same pipeline object reuses its buffers: results must not leak between runs
    for bi, (out_host, _res, _a, _b) in enumerate(pipe.run_host_stream(list(reversed(batches)))):
        check(3 - bi, out_host)


def test_run_host_stream_without_resize_or_rotation_does_not_alias_the_upload_slot(oracle, cuda):
    import torch
    from ocr_system_b200.pipeline import PagePipeline

    rng = np.random.default_rng(3)
    batches = [torch.from_numpy(rng.integers(0, 256, (2, 300, 200, 3), dtype=np.uint8)).pin_memory() for _ in range(3)]
    pipe = PagePipeline(max_dimension=960, deskew=False)
    held = []
    for out_host, res, _, _ in pipe.run_host_stream(batches, keep=3):
        held.append((out_host, res.pages))
    torch.cuda.synchronize()
    for b, (oh, dev_pages) in zip(batches, held):
        assert np.array_equal(oh["pages"].numpy(), b.numpy())
        assert np.array_equal(dev_pages.cpu().numpy(), b.numpy())


# ------------------------------------------------------------------ threads, devices, alignment
def test_four_threads_share_one_preprocessor(oracle, cuda):
    """The reference singleton is called from asyncio.to_thread workers (ocr_service.py:674-676)."""
    from PIL import Image
    from oracle import reference_port as RP
    from ocr_system_b200.image_preprocessing import ImagePreprocessor

    ipx = ImagePreprocessor(max_dimension=600)
    pages = [oracle.synth_page(877, 620, 20 + i) for i in range(4)]
    want = [RP.preprocess_for_azure(p, 600) for p in pages]
    got = [[None] * 3 for _ in range(4)]
    errs = []

    def work(t):
        try:
            for rep in range(3):
                got[t][rep] = ipx.preprocess_for_azure(Image.fromarray(pages[t]))
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for t in range(4):
        for rep in range(3):
            assert got[t][rep] == want[t], (t, rep)


def test_two_threads_share_one_pipeline(oracle, cuda):
    from ocr_system_b200.pipeline import PagePipeline

    h, w, md = 877, 620, 600
    pipe = PagePipeline(max_dimension=md)
    sets = [_host_batches(oracle, 3, 2, h, w, seed0=100 * t) for t in range(2)]
    want = [[[_oracle_chain(oracle, b[i].numpy(), md, False) for i in range(2)] for b in s] for s in sets]
    errs = []

    def work(t):
        try:
            for bi, (out_host, _res, _a, _b) in enumerate(pipe.run_host_stream(sets[t])):
                for i in range(2):
                    assert out_host["angles"][i] == want[t][bi][i][1]
                    assert np.array_equal(out_host["pages"][i].numpy(), want[t][bi][i][0])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_ops_follow_the_tensor_device_not_the_current_device(oracle):
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from ocr_system_b200 import ops
    from ocr_system_b200.pipeline import PagePipeline

    torch.cuda.set_device(0)
    page = oracle.synth_page(877, 620, 9)
    x = torch.from_numpy(page[None]).to("cuda:1")
    tw, th = oracle.target_size(620, 877, 600)
    small = ops.resize_lanczos(x, tw, th)
    assert small.device.index == 1
    ref = oracle.resize_lanczos(page, tw, th)
    assert np.array_equal(small.cpu().numpy()[0], ref)
    out, angles = ops.deskew(small)
    img, angle, _ = oracle.deskew(ref)
    assert angles[0] == angle and np.array_equal(out.cpu().numpy()[0], img)
    res = PagePipeline(max_dimension=600, device="cuda:1").run_device(x)
    assert np.array_equal(res.pages.cpu().numpy()[0], img)
    assert torch.cuda.current_device() == 0


def test_sliced_unaligned_batches(oracle, cuda):
    import torch
    from ocr_system_b200 import ops

    rng = np.random.default_rng(5)
    imgs = rng.integers(0, 256, (3, 33, 31, 3), dtype=np.uint8)     # 3069 bytes per page: pages[1:] is unaligned
    x = torch.from_numpy(imgs).to(cuda)
    g = ops.gray_pil(x[1:]).cpu().numpy()
    b = ops.binarize(x[1:], 100).cpu().numpy()
    for i in range(2):
        assert np.array_equal(g[i], oracle.gray_pil(imgs[1 + i]))
        assert np.array_equal(b[i], oracle.threshold(oracle.gray_pil(imgs[1 + i]), 100))
    planes = torch.from_numpy(imgs[..., 0].copy()).to(cuda)
    assert np.array_equal(ops.binarize(planes[1:], 77).cpu().numpy()[1], oracle.threshold(imgs[2, ..., 0], 77))


def test_resize_nearest_kernel_equals_pillow(oracle, cuda):
    import torch
    from PIL import Image
    from ocr_system_b200 import ops

    rng = np.random.default_rng(11)
    for (h, w, oh, ow) in [(877, 620, 600, 424), (100, 333, 47, 200), (17, 13, 5, 4), (1000, 700, 999, 699)]:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        got = ops.resize_nearest(torch.from_numpy(a[None]).to(cuda), ow, oh).cpu().numpy()[0]
        im = Image.frombytes("P", (w, h), a.tobytes())
        assert np.array_equal(got, np.frombuffer(im.resize((ow, oh), Image.Resampling.LANCZOS).tobytes(), np.uint8).reshape(oh, ow))
        assert np.array_equal(got, oracle.resize_nearest(a, ow, oh))


# ---------------------------------------------------------------------------------------------- OCRService adapter (f4)
class _StubOCRService:
    """The reference's page loop restated for the test (backend/services/ocr_service.py:604-627 calling :398-417):
    pdf_to_images, then per page preprocess_for_azure(image, apply_deskew=, apply_binarize=, target_size_mb=) and the
    (here: fake) Azure call.  `module` plays the role of `services.ocr_service` (it owns `image_preprocessor`)."""

    def __init__(self, module, apply_deskew=True, apply_binarize=False, target_size_mb=2.0):
        self.m, self._apply_deskew, self._apply_binarize, self._target_size_mb = module, apply_deskew, apply_binarize, target_size_mb
        self.sent = []

    def _analyze_with_azure(self, data: bytes):
        self.sent.append(data)
        return {"content": f"page of {len(data)} bytes"}

    def _process_single_image_sync(self, image, page_number=1):
        b = self.m.image_preprocessor.preprocess_for_azure(image, apply_deskew=self._apply_deskew,
                                                           apply_binarize=self._apply_binarize,
                                                           target_size_mb=self._target_size_mb)
        return {"page_number": page_number, "processed_image_bytes": b, "result": self._analyze_with_azure(b)}

    def process_pdf_as_images_sync(self, pdf_path):
        images = self.m.image_preprocessor.pdf_to_images(pdf_path)
        return [self._process_single_image_sync(im, i) for i, im in enumerate(images, start=1)]


@pytest.mark.parametrize("binarize", [False, True])
def test_ocr_service_page_loop_becomes_one_batched_submission(oracle, cuda, binarize):
    import types

    from PIL import Image

    from ocr_system_b200.image_preprocessing import ImagePreprocessor
    from ocr_system_b200.ocr_service_adapter import install

    inner = ImagePreprocessor(max_dimension=600)
    pages = [Image.fromarray(oracle.synth_page(1000, 720, 40 + i)) for i in range(5)]
    inner.pdf_to_images = lambda pdf_path, dpi=None: pages          # poppler is not in the image: the rasteriser is stubbed
    module = types.SimpleNamespace(image_preprocessor=None)
    proxy = install(module, inner)
    svc = _StubOCRService(module, apply_binarize=binarize, target_size_mb=0.2)
    out = svc.process_pdf_as_images_sync("document.pdf")
    assert proxy.batched_calls == 1 and len(out) == 5
    for i, o in enumerate(out):
        want = inner.preprocess_for_azure(pages[i], apply_deskew=True, apply_binarize=binarize, target_size_mb=0.2)
        assert o["processed_image_bytes"] == want and svc.sent[i] == want, i
    # an image the adapter never saw takes the per-image path; proxy attributes fall through to the drop-in
    other = Image.fromarray(oracle.synth_page(500, 400, 3))
    assert module.image_preprocessor.preprocess_for_azure(other) == inner.preprocess_for_azure(other)
    assert proxy.batched_calls == 1 and proxy.max_dimension == 600


def test_preprocess_pages_for_azure_decodes_jpeg_bytes_on_the_device(oracle, cuda):
    """bytes inputs: baseline JPEG files go through the device decoder, everything else through the host codec;
    either way the result is what the per-page call on Pillow's decode returns."""
    import io

    from PIL import Image

    from ocr_system_b200 import ops
    from ocr_system_b200.image_preprocessing import ImagePreprocessor

    ip_ = ImagePreprocessor(max_dimension=600)

    def enc(a, fmt="JPEG", **kw):
        b = io.BytesIO()
        Image.fromarray(a).save(b, fmt, **kw)
        return b.getvalue()

    p = [oracle.synth_page(1000, 720, 60 + i) for i in range(4)]
    files = [enc(p[0], quality=75), enc(p[1], quality=90, optimize=True), enc(p[2], quality=75, progressive=True),
             enc(p[3], "PNG"), enc(np.asarray(Image.fromarray(p[0]).convert("L")), quality=80), enc(p[1][:800], quality=75)]
    assert [ops.jpeg_probe(f) is not None for f in files] == [True, True, False, False, True, True]
    launches = ops.launch_count()
    got = ip_.preprocess_pages_for_azure(files, target_size_mb=0.2)
    assert ops.launch_count() > launches
    for i, f in enumerate(files):
        want = ip_.preprocess_for_azure(Image.open(io.BytesIO(f)), target_size_mb=0.2)
        assert got[i] == want, i


def test_tiff_and_png_files_take_the_host_codec_and_match_the_reference_sequence(oracle, cuda, tmp_path):
    """load_image (image_preprocessing.py:57-68) on TIFF / PNG paths -- the formats north_star's "PDF/TIFF page rasters"
    arrive in (uncompressed, LZW and Group-4 TIFF; PNG is what pdf_to_images asks poppler for): decoded by the host
    codec exactly as in the reference, then the device chain; result == the reference call sequence on the same file."""
    import io

    from PIL import Image

    from ocr_system_b200.image_preprocessing import ImagePreprocessor
    from oracle import reference_port as RP

    ip_ = ImagePreprocessor(max_dimension=600)
    page = oracle.synth_page(1000, 720, 77)
    cases = [("a.tif", dict()), ("b.tif", dict(compression="tiff_lzw")), ("c.png", dict()),
             ("d.tif", dict(compression="group4"))]
    for name, kw in cases:
        im = Image.fromarray(page)
        if kw.get("compression") == "group4":
            im = im.convert("1")                      # bilevel fax page: load_image converts it to RGB
        path = tmp_path / name
        im.save(path, **kw)
        loaded = ip_.load_image(path)
        assert loaded.mode in ("RGB", "L")
        ref_in = Image.open(path)
        ref_in = ref_in if ref_in.mode in ("RGB", "L") else ref_in.convert("RGB")
        want = RP.preprocess_for_azure(np.asarray(ref_in), max_dim=600, target_size_mb=0.2)
        assert ip_.preprocess_for_azure(path, target_size_mb=0.2) == want, name
        assert ip_.preprocess_pages_for_azure([path, path.read_bytes()], target_size_mb=0.2) == [want, want], name


def test_dropin_equals_the_real_reference_module_on_random_calls(oracle, cuda):
    """The unmodified reference module travels with the snapshot (oracle/_ref, placed by oracle/make_ref.py): run it
    beside the drop-in on random PIL images (modes, sizes, contents) through every public method with random arguments
    and require identical results -- PIL mode / size / bytes, float64 angle, JPEG file bytes, exception type
    (tools/sweep_dropin_vs_reference.py; 5720 calls of the full sweep: profiles/r2_sweep_dropin_vs_reference.json).
    cv2 must be in its default dispatch here, like in the application."""
    import cv2

    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    from sweep_dropin_vs_reference import sweep

    was = cv2.useOptimized()
    cv2.setUseOptimized(True)
    try:
        res = sweep(1000, 1016, verbose=False)
        import sweep_dropin_vs_reference as S

        S.TINY = True          # degenerate sizes (1 x 1 ... ), extreme aspect ratios, target sides that round to 0
        try:
            tiny = sweep(2000, 2040, verbose=False)
        finally:
            S.TINY = False
    finally:
        cv2.setUseOptimized(was)
    if res is None:
        pytest.skip("oracle/_ref is not in this snapshot (no /root/reference when build() ran)")
    checked, bad = res
    assert checked == 16 * 11 and not bad, bad[:3]
    assert tiny[0] == 40 * 11 and not tiny[1], tiny[1][:3]

