"""CPU-side checks: the C-ABI library loads and exports every symbol include/lumina_b200.h
declares (no compute calls), host-only entry points agree with the oracle / reference rules,
the product path never touches oracle/, and the drop-in keeps the reference's error behaviour."""
import collections
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def abi():
    import __graft_entry__ as g

    g._load_build_module().build()
    from ocr_system_b200 import _abi

    return _abi


def test_library_exports_every_declared_symbol(abi):
    hdr = open(os.path.join(ROOT, "include", "lumina_b200.h")).read()
    declared = set(re.findall(r"\b(lumina_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = C.CDLL(abi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lumina_b200.h but not exported"
    assert declared == set(abi.PROTOTYPES), declared ^ set(abi.PROTOTYPES)
    assert abi.lib().lumina_abi_version() == 4


def test_target_size_matches_reference_rule(abi, oracle):
    from ocr_system_b200 import ops

    for (w, h, md) in [(2480, 3508, 2000), (2480, 3508, 960), (3508, 2480, 960), (1000, 1000, 2000), (4000, 3000, 2000),
                       (2001, 17, 2000), (333, 5000, 1234)]:
        assert ops.target_size(w, h, md) == oracle.target_size(w, h, md)
    assert ops.det_target_size(960, 678) == oracle.det_target_size(960, 678) == (960, 672)
    assert ops.det_target_size(3508, 2480) == oracle.det_target_size(3508, 2480)


def test_host_angle_and_rotation_match_oracle(abi, oracle):
    from ocr_system_b200 import ops

    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 200):
        lines = rng.integers(0, 1500, (n, 4)).astype(np.int32)
        assert ops.median_angle(lines) == oracle.median_angle(lines)
    assert ops.median_angle(np.zeros((0, 4), np.int32)) == 0.0
    for a in (0.5, -1.3, 44.0):
        assert np.array_equal(ops.rotation_matrix(706, 1000, a), oracle.rotation_matrix(706, 1000, a))


def test_errors_are_codes_not_crashes(abi):
    lib = abi.lib()
    rc = lib.lumina_rgb2gray_pil_u8(None, None, 16, None)  # null pointers: rejected before any launch
    assert rc == -1
    assert b"null" in lib.lumina_last_error_string()
    with pytest.raises(abi.LuminaError):
        abi.check(rc)
    assert lib.lumina_ppht_workspace_bytes(0, 10, 10, 1.0, 0.01) == 0


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ocr-system_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"
                assert "liblumina_oracle" not in src and "reference_port" not in src, f


def test_no_cpu_fallback(abi):
    import torch

    from ocr_system_b200 import ops

    with pytest.raises(TypeError):
        ops.gray_pil(torch.zeros((1, 4, 4, 3), dtype=torch.uint8))  # CPU tensor: refused, never computed on host


def test_dropin_surface_and_error_behaviour(abi, tmp_path):
    from ocr_system_b200.image_preprocessing import ImagePreprocessor, image_preprocessor

    ip = ImagePreprocessor()
    assert ip.max_dimension == 2000 and ip.target_dpi == 300      # config.py:69, image_preprocessing.py:45-51
    for name in ("load_image", "load_image_bytes", "resize_if_needed", "get_optimal_size", "enhance_contrast",
                 "enhance_sharpness", "denoise", "convert_to_grayscale", "auto_orient", "binarize", "optimize_for_ocr",
                 "pdf_to_images", "get_pdf_page_count", "save_image", "image_to_bytes", "get_image_info", "deskew",
                 "adaptive_binarize", "compress_for_azure", "preprocess_for_azure"):
        assert callable(getattr(ip, name)) and callable(getattr(image_preprocessor, name))
    with pytest.raises(FileNotFoundError):
        ip.load_image(tmp_path / "missing.png")
    assert ip.get_optimal_size(2480, 3508) == (1413, 2000)
    assert ip.get_optimal_size(100, 100) == (100, 100)
    from PIL import Image

    small = Image.new("RGB", (64, 48), (10, 20, 30))
    assert ip.resize_if_needed(small) is small                     # no resize needed -> same object, no GPU
    png = ip.image_to_bytes(small, "PNG")
    assert ip.load_image_bytes(png).size == (64, 48)
    p = ip.save_image(small, tmp_path / "a" / "b.jpg")
    assert p.exists() and ip.load_image(p).mode == "RGB"
    info = ip.get_image_info(small)
    assert info["needs_resize"] is False and info["size_optimal"] == (64, 48)
    try:
        import pdf2image  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            ip.pdf_to_images(tmp_path / "x.pdf")


def test_ctc_dictionary_handling(abi, tmp_path):
    from ocr_system_b200.paddle_ops import CTCLabelDecode

    d = tmp_path / "dict.txt"
    d.write_text("a\nb\nक\n", encoding="utf-8")
    dec = CTCLabelDecode(str(d), use_space_char=True)
    assert dec.character == ["blank", "a", "b", "क", " "]
    assert len(CTCLabelDecode().character) == 37


def test_gray_source_rule_holds_for_every_pil_mode():
    """``ImagePreprocessor._gray_source`` (the input of convert_to_grayscale / binarize / adaptive_binarize / deskew for
    PIL objects that are neither RGB nor L) rests on two Pillow facts: ``convert('L')`` of a YCbCr image is its Y band,
    and for every other mode ``convert('L') == convert('RGB').convert('L')``.  Pin both against the Pillow in this image
    so that a Pillow upgrade that breaks the rule fails here, not as a silent pixel difference on the GPU box."""
    from PIL import Image

    rng = np.random.default_rng(7)
    rgb = Image.fromarray(rng.integers(0, 256, (37, 53, 3), dtype=np.uint8))
    rgba = Image.fromarray(rng.integers(0, 256, (37, 53, 4), dtype=np.uint8), "RGBA")
    pal_t = rgb.convert("P")
    pal_t.info["transparency"] = 3
    images = [rgba, rgba.convert("LA"), rgb.convert("RGBX"), rgb.convert("CMYK"), rgb.convert("HSV"), rgb.convert("P"), pal_t,
              rgb.convert("1"), rgba.convert("RGBa"),
              Image.fromarray(rng.integers(-500, 1000, (37, 53)).astype(np.int32), "I"),
              Image.fromarray((rng.random((37, 53)) * 400 - 50).astype(np.float32), "F"),
              Image.fromarray(rng.integers(0, 65536, (37, 53)).astype(np.uint16))]
    for im in images:
        assert np.array_equal(np.asarray(im.convert("L")), np.asarray(im.convert("RGB").convert("L"))), im.mode
    ycc = rgb.convert("YCbCr")
    assert np.array_equal(np.asarray(ycc.convert("L")), np.asarray(ycc.getchannel(0)))
    la_pre = rgba.convert("La")            # Pillow refuses La -> L and La -> RGB alike: the drop-in raises the same error type
    for target in ("L", "RGB"):
        with pytest.raises(ValueError):
            la_pre.convert(target)


def test_ctc_right_to_left_and_label_decode(abi, tmp_path):
    """Host half of CTCLabelDecode: upstream's pred_reverse for right-to-left dictionaries (Latin / digit runs keep
    their order) and ``decode(label)`` for the ground-truth strings; checked against a literal restatement of upstream's
    regex loop (ppocr/postprocess/rec_postprocess.py, BaseRecLabelDecode.pred_reverse)."""
    import re

    from ocr_system_b200.paddle_ops import CTCLabelDecode

    def upstream_pred_reverse(pred):
        pred_re, cur = [], ""
        for c in pred:
            if not bool(re.search("[a-zA-Z0-9 :*./%+-]", c)):
                if cur != "":
                    pred_re.append(cur)
                pred_re.append(c)
                cur = ""
            else:
                cur += c
        if cur != "":
            pred_re.append(cur)
        return "".join(pred_re[::-1])

    d = tmp_path / "arabic_dict.txt"
    alphabet = list("ابتثجحخ") + list("abXY019") + list(" :*./%+-") + ["،", "ـ"]
    d.write_text("\n".join(alphabet) + "\n", encoding="utf-8")
    dec = CTCLabelDecode(str(d))
    assert dec.reverse and not CTCLabelDecode(character=alphabet).reverse
    rng = np.random.default_rng(3)
    for _ in range(300):
        symbols = [alphabet[k] for k in rng.integers(0, len(alphabet), int(rng.integers(0, 24)))]
        assert dec.pred_reverse(symbols) == upstream_pred_reverse(symbols)
    assert dec.pred_reverse(list("اب12 ab.ت")) == "ت12 ab.با"
    label = np.array([[1, 2, 0, 8, 9, 0], [0, 0, 0, 0, 0, 0]])
    assert dec.decode_label(label) == [(upstream_pred_reverse([dec.character[k] for k in (1, 2, 8, 9)]), 1.0), ("", 0.0)]
    plain = CTCLabelDecode(character=list("abc"))
    assert plain.decode_label([[1, 1, 0, 3]]) == [("aac", 1.0)]


def test_resize_plan_cache_is_a_bounded_lru_that_never_evicts_a_plan_in_use(abi):
    """ops._ResizePlans: device coefficient tables per geometry are kept in a bounded LRU (a service sees arbitrary scan
    sizes); a plan that a caller is launching with is never destroyed.  Host logic only: create / destroy are injected."""
    import threading

    from ocr_system_b200 import ops

    created, destroyed = [], []
    cache = ops._ResizePlans(capacity=3, create=lambda *k: created.append(k) or ("plan", k),
                             destroy=lambda dev, h: destroyed.append(h[1]))
    geo = [(0, 100 + i, 200, 50, 100) for i in range(6)]
    for g in geo[:3]:
        with cache.use(*g) as h:
            assert h == ("plan", g)
    with cache.use(*geo[0]):                       # refresh geo[0]: geo[1] is now the oldest
        pass
    assert created == geo[:3] and destroyed == [] and len(cache) == 3
    with cache.use(*geo[3]):
        assert destroyed == [geo[1]] and len(cache) == 3
    with cache.use(*geo[0]) as h0:                 # geo[0] is held while two more geometries arrive
        with cache.use(*geo[4]):
            with cache.use(*geo[5]):
                assert geo[0] not in destroyed and h0 == ("plan", geo[0])
    assert set(destroyed) == {geo[1], geo[2], geo[3]} and len(cache) == 3
    assert created.count(geo[0]) == 1
    # every plan in use: the cache may exceed its bound rather than free tables under a running launch
    tight = ops._ResizePlans(capacity=1, create=lambda *k: k, destroy=lambda dev, h: destroyed.append(("tight", h)))
    with tight.use(*geo[0]):
        with tight.use(*geo[1]):
            assert len(tight) == 2 and not any(d[0] == "tight" for d in destroyed if isinstance(d[0], str))
    with tight.use(*geo[2]):
        pass
    assert len(tight) == 1
    # threads hammering a small cache: a handle is never destroyed between __enter__ and __exit__ of its user
    live, errors = collections.Counter(), []
    lock = threading.Lock()

    def destroy(dev, h):
        with lock:
            if live[h] > 0:
                errors.append(h)

    shared = ops._ResizePlans(capacity=2, create=lambda *k: k, destroy=destroy)

    def worker(seed):
        rng = np.random.default_rng(seed)
        for _ in range(400):
            g = geo[int(rng.integers(0, 6))]
            with shared.use(*g) as h:
                with lock:
                    live[h] += 1
                with lock:
                    live[h] -= 1

    ts = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors and len(shared) <= 2 + 4


def test_det_target_size_limit_types_match_upstream_rule(abi, oracle):
    """lumina_det_target_size_ex (host entry) against the literal python restatement of upstream's resize_image_type0 for
    all three limit types; the old entry stays the "max" rule; bad arguments are status codes / ValueError."""
    from ocr_system_b200 import ops
    from ocr_system_b200.paddle_ops import DetResizeNormalize

    rng = np.random.default_rng(11)
    for _ in range(6000):
        h, w = int(rng.integers(1, 5000)), int(rng.integers(1, 5000))
        limit = int(rng.choice([32, 64, 320, 640, 736, 960, 1280, 1536, 2000, int(rng.integers(1, 3000))]))
        for lt in ("max", "min", "resize_long"):
            assert ops.det_target_size(h, w, limit, lt) == oracle.det_target_size_upstream(h, w, limit, lt), (h, w, limit, lt)
        assert ops.det_target_size(h, w, limit) == oracle.det_target_size(h, w, limit)
    for (h, w, limit) in [(48, 48, 96), (80, 80, 100), (1008, 1008, 960), (3508, 2480, 960)]:   # round-half-even ties and the headline
        for lt in ("max", "min", "resize_long"):
            assert ops.det_target_size(h, w, limit, lt) == oracle.det_target_size_upstream(h, w, limit, lt)
    with pytest.raises(ValueError):
        ops.det_target_size(100, 100, 960, "area")
    with pytest.raises(ValueError):
        DetResizeNormalize(limit_type="area")
    oh, ow = C.c_int(), C.c_int()
    L = abi.lib()
    assert L.lumina_det_target_size_ex(100, 100, 960, 3, C.byref(oh), C.byref(ow)) < 0
    assert L.lumina_det_target_size_ex(0, 100, 960, 0, C.byref(oh), C.byref(ow)) < 0
    assert L.lumina_det_target_size_ex(100, 100, 960, 1, C.byref(oh), C.byref(ow)) == 0 and (oh.value, ow.value) == (960, 960)


def test_dropin_accepts_every_call_the_reference_accepts(abi):
    """Public surface of the reference's two hot-path modules (tests/golden/signature_golden.json, recorded from the
    unmodified reference by tests/golden/make_signature_golden.py): every public method / function exists in the drop-in
    with the same parameter names, kinds and defaults in the same positions (extra trailing parameters must have
    defaults), and the boundary dataclasses carry the same fields in the same order."""
    import dataclasses
    import inspect
    import json
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_signature_golden import params

    from ocr_system_b200 import image_preprocessing as ip
    from ocr_system_b200 import ocr_postprocessor as pp

    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "signature_golden.json")))
    assert len(gold["ImagePreprocessor"]) >= 20 and len(gold["ocr_postprocessor"]) >= 5

    def check(want, fn, label):
        got = params(fn)
        assert got[: len(want)] == want, (label, want, got)
        for name, kind, default in got[len(want):]:
            assert default is not None or kind in ("VAR_POSITIONAL", "VAR_KEYWORD"), (label, name)

    for name, want in gold["ImagePreprocessor"].items():
        check(want, getattr(ip.ImagePreprocessor, name), f"ImagePreprocessor.{name}")
    for name, want in gold["ocr_postprocessor"].items():
        check(want, getattr(pp, name), f"ocr_postprocessor.{name}")
    for name, fields in gold["dataclasses"].items():
        assert [f.name for f in dataclasses.fields(getattr(pp, name))] == fields, name
    assert gold["module_singleton"] == "ImagePreprocessor" and isinstance(ip.image_preprocessor, ip.ImagePreprocessor)
    for name in gold["ImagePreprocessor"]:
        if name != "__init__":
            assert callable(getattr(ip.image_preprocessor, name)), name      # the lazy singleton forwards every method


def test_batching_preprocessor_host_logic_with_a_stub_device(abi):
    """ocr_service_adapter.BatchingPreprocessor (the reference's page loop, ocr_service.py:604-660, as one batched
    submission): look-ahead, max_batch split, argument change mid-document, unknown images, per-thread state -- against a
    stub inner preprocessor, so the host logic is covered without a GPU (the GPU test checks the bytes)."""
    import threading

    from PIL import Image

    from ocr_system_b200.ocr_service_adapter import BatchingPreprocessor, install

    class Inner:
        max_dimension = 2000

        def __init__(self):
            self.batches, self.singles = [], []

        def pdf_to_images(self, path, dpi=None):
            return [Image.new("RGB", (8, 8), (i, 0, 0)) for i in range(int(path))]

        def preprocess_pages_for_azure(self, pages, *flags):
            self.batches.append((len(pages), flags))
            return [b"B%d|%r" % (p.getpixel((0, 0))[0], flags) for p in pages]

        def preprocess_for_azure(self, image, *flags):
            self.singles.append(flags)
            return b"S%d|%r" % (image.getpixel((0, 0))[0], flags)

    inner = Inner()
    bp = BatchingPreprocessor(inner, max_batch=4)
    assert bp.max_dimension == 2000                                  # everything else is forwarded
    flags = (True, False, True, True, 2.0)
    pages = bp.pdf_to_images("10")
    got = [bp.preprocess_for_azure(p, apply_deskew=True, apply_binarize=False) for p in pages]
    assert got == [b"B%d|%r" % (i, flags) for i in range(10)]
    assert [n for n, _ in inner.batches] == [4, 4, 2] and bp.batched_calls == 3 and not inner.singles
    # an image the adapter has not seen takes the per-image path with the caller's arguments
    other = Image.new("RGB", (8, 8), (99, 0, 0))
    assert bp.preprocess_for_azure(other, False, True) == b"S99|%r" % ((False, True, True, True, 2.0),)
    # arguments change in the middle of a document: the pages from there on are resubmitted with the new ones
    inner.batches.clear()
    pages = bp.pdf_to_images("3")
    a = bp.preprocess_for_azure(pages[0])
    b = bp.preprocess_for_azure(pages[1], apply_binarize=True)
    c = bp.preprocess_for_azure(pages[2], apply_binarize=True)
    f2 = (True, True, True, True, 2.0)
    assert (a, b, c) == (b"B0|%r" % (flags,), b"B1|%r" % (f2,), b"B2|%r" % (f2,))
    assert [n for n, _ in inner.batches] == [3, 2]
    # a page asked for twice is preprocessed again (per image), never answered from a stale entry
    assert bp.preprocess_for_azure(pages[0]).startswith(b"S0|")
    # two threads, two documents: the look-ahead state is per thread
    results = {}

    def doc(n):
        ps = bp.pdf_to_images(str(n))
        results[n] = [bp.preprocess_for_azure(p) for p in ps]

    ts = [threading.Thread(target=doc, args=(n,)) for n in (5, 7)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert [len(results[n]) for n in (5, 7)] == [5, 7] and all(r.startswith(b"B") for n in (5, 7) for r in results[n])

    class Svc:
        image_preprocessor = None

    proxy = install(Svc, preprocessor=inner, max_batch=2)
    assert Svc.image_preprocessor is proxy and proxy._max_batch == 2


def test_preprocess_pages_for_azure_grouping_and_slicing_host_logic(abi):
    """Host half of the batched Azure path with the device calls stubbed out: pages are grouped by (size, mode), a group
    larger than the cap is submitted in slices, results come back in input order."""
    from PIL import Image

    from ocr_system_b200.image_preprocessing import ImagePreprocessor

    calls = []

    class Stub(ImagePreprocessor):
        _MAX_BATCH_PAGES = 3

        def _upload(self, images):
            calls.append((images[0].size, images[0].mode, len(images)))
            return [im.getpixel((0, 0)) for im in images]

        def preprocess_device(self, x, *flags):
            return x, None

        def compress_pages_for_azure(self, x, target_size_mb=2.0, **kw):
            return [repr(v).encode() for v in x]

    ip = Stub(max_dimension=400)
    imgs = []
    for i in range(11):
        size, mode = [((8, 8), "RGB"), ((8, 6), "RGB"), ((8, 8), "L")][i % 3 if i < 9 else 0]
        imgs.append(Image.new(mode, size, i if mode == "L" else (i, 0, 0)))
    got = ip.preprocess_pages_for_azure(imgs)
    assert got == [repr(im.getpixel((0, 0))).encode() for im in imgs]
    assert sorted(calls) == sorted([((8, 8), "RGB", 3), ((8, 8), "RGB", 2), ((8, 6), "RGB", 3), ((8, 8), "L", 3)])
