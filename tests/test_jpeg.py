"""JPEG encode-to-size (backend/utils/image_preprocessing.py:496-557 compress_for_azure).

The oracle for this row is the codec the reference calls: Pillow / libjpeg-turbo.  Its outputs for seeded
images are pinned in tests/golden/jpeg_golden.json (sha256 + size, generated in the build container by
tests/golden/make_jpeg_golden.py, which also runs the REFERENCE's own compress_for_azure).
CPU : the goldens still match the Pillow that is installed (a library drift would show here first).
GPU : the CUDA encoder through the C-ABI produces the same FILES, byte for byte; compress_for_azure of the
      drop-in returns the reference's bytes (quality ladder and resize fall-back included).
"""
import hashlib
import io
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from jpeg_images import image  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "jpeg_golden.json")))


def _sha(b):
    return hashlib.sha256(b).hexdigest()


def _pil(a, q, opt):
    from PIL import Image

    b = io.BytesIO()
    Image.fromarray(a).save(b, format="JPEG", quality=q, optimize=opt)
    return b.getvalue()


def test_installed_pillow_matches_pinned_goldens():
    import PIL

    if PIL.__version__ != GOLD["pillow"]:
        pytest.skip(f"goldens were made with Pillow {GOLD['pillow']}, found {PIL.__version__}")
    for f in GOLD["files"][::7]:
        d = _pil(image(f["kind"], f["h"], f["w"]), f["quality"], f["optimize"])
        assert (len(d), _sha(d)) == (f["size"], f["sha"])


@pytest.mark.gpu
def test_gpu_files_equal_pinned_pillow_files(cuda):
    import torch
    from ocr_system_b200 import ops

    enc = ops.JpegEncoder()
    bad = []
    for f in GOLD["files"]:
        a = image(f["kind"], f["h"], f["w"])
        files, sizes = enc.encode(torch.from_numpy(a[None]).to(cuda), f["quality"], f["optimize"])
        if (len(files[0]), _sha(files[0])) != (f["size"], f["sha"]) or int(sizes[0]) != f["size"]:
            bad.append((f["kind"], f["h"], f["w"], f["quality"], f["optimize"], len(files[0]), f["size"]))
    assert not bad, bad[:10]


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", [(678, 960), (339, 481), (1000, 707)])
def test_gpu_batch_equals_live_pillow_on_pages(cuda, oracle, h, w):
    import torch
    from ocr_system_b200 import ops

    pages = np.stack([oracle.synth_page(h, w, s) for s in range(3)])
    pages[2] = np.random.default_rng(5).integers(0, 256, (h, w, 3), dtype=np.uint8)   # worst case: noise
    x = torch.from_numpy(pages).to(cuda)
    enc = ops.JpegEncoder()
    for q, opt, reuse in [(95, True, False), (85, True, True), (35, True, True), (90, False, True)]:
        files, sizes = enc.encode(x, q, opt, reuse_dct=reuse)
        for i in range(3):
            ref = _pil(pages[i], q, opt)
            assert files[i] == ref, (i, q, opt, len(files[i]), len(ref))
    # the size cap: a page that does not fit is reported, not copied
    files, sizes = enc.encode(x, 95, True, max_bytes=int(sizes.min()) + 10)
    assert sum(f is None for f in files) >= 1 and all(s > 0 for s in sizes)


@pytest.mark.gpu
def test_gpu_compress_for_azure_equals_reference(cuda):
    from PIL import Image
    from ocr_system_b200.image_preprocessing import ImagePreprocessor

    ip = ImagePreprocessor()
    for c in GOLD["compress_for_azure"]:
        d = ip.compress_for_azure(Image.fromarray(image(c["kind"], c["h"], c["w"])), target_size_mb=c["target_size_mb"])
        assert (len(d), _sha(d)) == (c["size"], c["sha"]), c
    # L input is converted like the reference (:522-525); the decoded image is what Pillow would decode
    g = Image.fromarray(image("photo", 93, 127)).convert("L")
    d = ip.compress_for_azure(g)
    assert d == _pil(np.asarray(g.convert("RGB")), 95, True)
    assert Image.open(io.BytesIO(d)).size == (127, 93)


@pytest.mark.gpu
@pytest.mark.parametrize("binarize", [False, True])
def test_gpu_preprocess_for_azure_files_equal_reference_sequence(cuda, oracle, binarize):
    """The app's whole per-page job (ocr_service.py:412-417): host raster -> JPEG bytes, batched and per page,
    against the reference's call sequence on Pillow/OpenCV (oracle/reference_port.preprocess_for_azure)."""
    from PIL import Image
    from oracle import reference_port as RP
    from ocr_system_b200.image_preprocessing import ImagePreprocessor

    pages = [oracle.synth_page(877, 620, s) for s in (0, 1, 2)] + [oracle.synth_page(620, 877, 3)]
    imgs = [Image.fromarray(p) for p in pages]
    from conftest import cv2_dispatch

    ip = ImagePreprocessor(max_dimension=400, cv_dispatch=cv2_dispatch())   # RP runs the live cv2 of this process
    got = ip.preprocess_pages_for_azure(imgs, apply_binarize=binarize)
    for p, im, g in zip(pages, imgs, got):
        want = RP.preprocess_for_azure(p, 400, apply_binarize=binarize)
        assert g == want
        assert ip.preprocess_for_azure(im, apply_binarize=binarize) == want


@pytest.mark.gpu
def test_gpu_jpeg_degenerate_sizes_equal_pillow(cuda):
    """Sizes below one block / one MCU and strongly non-square pages (dummy blocks on both edges)."""
    import torch
    from ocr_system_b200 import ops

    rng = np.random.default_rng(11)
    for (h, w) in [(1, 1), (1, 17), (9, 7), (8, 8), (15, 31), (17, 16), (33, 2), (2, 257)]:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for q, opt in [(95, True), (60, False), (1, True)]:
            got = ops.jpeg_encode(torch.from_numpy(a[None]).to(cuda), q, opt)[0]
            assert got == _pil(a, q, opt), (h, w, q, opt)


@pytest.mark.gpu
def test_gpu_batched_azure_path_mixed_sizes_and_modes(cuda, oracle):
    """preprocess_pages_for_azure groups pages by (size, mode); results come back in input order and equal the
    per-page call."""
    from PIL import Image
    from ocr_system_b200.image_preprocessing import ImagePreprocessor

    a = Image.fromarray(oracle.synth_page(877, 620, 0))
    b = Image.fromarray(oracle.synth_page(620, 877, 1))
    c = Image.fromarray(oracle.synth_page(877, 620, 2)).convert("L")
    d = Image.fromarray(oracle.synth_page(300, 200, 3))          # below max_dimension: no resize
    imgs = [a, b, c, a, d, b]
    ip = ImagePreprocessor(max_dimension=400)
    got = ip.preprocess_pages_for_azure(imgs, target_size_mb=0.05)
    assert len(got) == len(imgs) and got[0] == got[3] and got[1] == got[5]
    for im, g in zip(imgs, got):
        assert g == ip.preprocess_for_azure(im, target_size_mb=0.05)
        assert Image.open(io.BytesIO(g)).mode == "RGB" and len(g) <= int(0.05 * 1024 * 1024)
