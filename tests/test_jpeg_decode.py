"""Page ingest: baseline JPEG decode (backend/utils/image_preprocessing.py:57-75 load_image /
load_image_bytes -> Image.open(...)).

The oracle for this row is the decoder the reference calls: Pillow / libjpeg-turbo, which is installed wherever the
tests run, so every case compares with a LIVE Pillow decode of the same file (files are made by Pillow's encoder:
all sampling modes, grayscale, optimised and standard tables, restart intervals, odd and tiny sizes).
CPU : (1) oracle/jpeg_decode.c (sequential restatement) == Pillow; (2) the host build of jpegd_core.h -- the same
      parser / sub-sequence decoder / IDCT / upsampling the kernels run, driven by a sequential simulation of the
      kernels' schedule -- == Pillow, and every sub-sequence's exit state is re-verified in the write pass.
GPU : the CUDA decoder through the C-ABI == Pillow, byte for byte, incl. whole A4 pages and mixed batches.
"""
import ctypes as C
import io
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _img(rng, h, w, kind):
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "smooth":
        y, x = np.mgrid[0:h, 0:w]
        return np.stack([x * 255 // max(w - 1, 1), y * 255 // max(h - 1, 1), (x + y) * 255 // max(w + h - 2, 1)], -1).astype(np.uint8)
    a = np.full((h, w, 3), 240, np.uint8) + rng.integers(0, 12, (h, w, 3), dtype=np.uint8)
    for i in range(0, h - 6, 14):
        a[i:i + 5, 3:max(4, w - 3):2] = rng.integers(0, 70, dtype=np.uint8)
    return a


def _save(a, gray=False, **kw):
    from PIL import Image

    im = Image.fromarray(a)
    if gray:
        im = im.convert("L")
        kw.pop("subsampling", None)
    b = io.BytesIO()
    try:
        im.save(b, "JPEG", **kw)
    except OSError:      # libjpeg refuses a few tiny-image / optimize / restart combinations
        return None
    return b.getvalue()


def _pil(data):
    from PIL import Image

    return np.asarray(Image.open(io.BytesIO(data)))


def _cases(sizes, qualities=(30, 75, 95, 100), seed=0):
    rng = np.random.default_rng(seed)
    for (h, w) in sizes:
        for kind in ("noise", "smooth", "page"):
            a = _img(rng, h, w, kind)
            for q in qualities:
                for ss in (0, 1, 2):
                    for opt in (False, True):
                        for gray in (False, True):
                            if gray and ss:
                                continue
                            for rst in (0, 1, 3):
                                kw = dict(quality=q, optimize=opt, subsampling=ss)
                                if rst:
                                    kw["restart_marker_blocks"] = rst
                                d = _save(a, gray, **kw)
                                if d is not None:
                                    yield (h, w, kind, q, ss, opt, gray, rst), d


SMALL = [(1, 1), (2, 2), (3, 5), (8, 8), (16, 16), (17, 33), (9, 4), (31, 47), (64, 48), (100, 150)]


def test_oracle_decoder_equals_pillow(oracle):
    n = 0
    for key, d in _cases(SMALL + [(233, 177)]):
        got = oracle.jpeg_decode(d)
        ref = _pil(d)
        assert got is not None and got.shape == ref.shape and np.array_equal(got, ref), key
        n += 1
    assert n > 2500


def test_oracle_decoder_reports_files_outside_the_subset(oracle):
    from PIL import Image

    a = _img(np.random.default_rng(1), 40, 56, "page")
    b = io.BytesIO()
    Image.fromarray(a).save(b, "JPEG", progressive=True)
    assert oracle.jpeg_decode(b.getvalue()) is None
    b = io.BytesIO()
    Image.fromarray(a).convert("CMYK").save(b, "JPEG")
    assert oracle.jpeg_decode(b.getvalue()) is None
    with pytest.raises(ValueError):
        oracle.jpeg_decode(b"\xff\xd8\xff\xd9")


def _build_host(name, *flags):
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    so = os.path.join(HERE, "_build", name)
    src = os.path.join(HERE, "jpegd_host.cpp")
    hdr = os.path.join(HERE, "..", "ocr-system_b200", "csrc", "jpegd_core.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", *flags, "-o", so, src])
    return C.CDLL(so)


@pytest.fixture(scope="module")
def J():
    return _build_host("libjpegd_host.so")


def _host_decode(J, data, sub_bits):
    buf = np.frombuffer(data, np.uint8)
    whc, st = (C.c_int * 3)(), (C.c_longlong * 4)()
    rc = J.jdh_decode(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), sub_bits, None, whc, st)
    assert rc == 0, rc
    out = np.empty((whc[1], whc[0], whc[2]), np.uint8)
    rc = J.jdh_decode(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), sub_bits, out.ctypes.data_as(C.c_void_p), whc, st)
    assert rc == 0, rc
    return (out[:, :, 0] if whc[2] == 1 else out), list(st)


def test_parallel_schedule_on_the_host_equals_pillow(J):
    """jpegd_core.h as the kernels use it: guessed entry states, exit-state propagation until nothing changes,
    block-index scan with restart bases, write pass, DC prefix sums."""
    n = 0
    for key, d in _cases(SMALL + [(233, 177)], qualities=(30, 95)):
        ref = _pil(d)
        for sub_bits in (128, 1024):
            got, st = _host_decode(J, d, sub_bits)
            assert got.shape == ref.shape and np.array_equal(got, ref), (key, sub_bits)
            assert st[3] == st[1], (key, sub_bits, st)     # every exit state confirmed by the write pass
            n += 1
    assert n > 2000


def test_long_codes_without_second_level_tables_walk_the_ladder():
    """JD_MAX_SUB = 1: only one 10-bit prefix gets a second-level table, every other long code takes the maxcode
    ladder (the fallback for unusual optimised tables)."""
    J1 = _build_host("libjpegd_host_sub1.so", "-DJD_MAX_SUB=1")
    rng = np.random.default_rng(5)
    for q, opt in ((100, False), (97, True), (60, False)):
        d = _save(_img(rng, 120, 168, "noise"), quality=q, optimize=opt, subsampling=2)
        got, st = _host_decode(J1, d, 1024)
        assert np.array_equal(got, _pil(d)) and st[3] == st[1]


def test_parallel_schedule_on_a_full_a4_page(J, oracle):
    d = _save(oracle.synth_page(3508, 2480, 3), quality=75)
    got, st = _host_decode(J, d, 1024)
    assert np.array_equal(got, _pil(d))
    assert st[0] < 64 and st[3] == st[1], st


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_decoder_equals_pillow_on_every_mode(cuda):
    from ocr_system_b200 import ops

    n = 0
    for key, d in _cases(SMALL + [(233, 177), (480, 640)], qualities=(30, 95)):
        ref = _pil(d)
        got = ops.jpeg_decode([d]).cpu().numpy()[0]
        if ref.ndim == 2:
            got = got[:, :, 0]
        assert got.shape == ref.shape and np.array_equal(got, ref), key
        n += 1
    assert n > 1000


@pytest.mark.gpu
def test_gpu_decoder_batches_of_a4_pages(cuda, oracle):
    """Whole pages, several per batch, different content / file sizes / Huffman tables per page."""
    import torch
    from ocr_system_b200 import ops

    pages = [oracle.synth_page(3508, 2480, s) for s in range(3)]
    files = [_save(pages[0], quality=75), _save(pages[1], quality=90, optimize=True), _save(pages[2], quality=50),
             _save(np.full((3508, 2480, 3), 255, np.uint8), quality=75)]          # a blank page: periodic bit stream
    dec = ops.JpegDecoder()
    blob, offs = dec.pack(files)
    out, status = dec.decode(blob, offs)
    torch.cuda.synchronize()
    assert status.cpu().tolist() == [0, 0, 0, 0]
    out = out.cpu().numpy()
    for i, f in enumerate(files):
        assert np.array_equal(out[i], _pil(f)), i
    # same decoder object, second batch (workspace / staging reuse), grayscale pages with restart intervals
    gfiles = [_save(p, gray=True, quality=80, restart_marker_rows=1) for p in pages[:2]]
    blob, offs = dec.pack(gfiles)
    out, status = dec.decode(blob, offs)
    torch.cuda.synchronize()
    assert status.cpu().tolist() == [0, 0]
    for i, f in enumerate(gfiles):
        assert np.array_equal(out[i, :, :, 0].cpu().numpy(), _pil(f)), i


@pytest.mark.gpu
def test_gpu_decoder_flags_truncated_files_and_rejects_other_formats(cuda, oracle):
    import torch
    from PIL import Image
    from ocr_system_b200 import ops, _abi

    a = oracle.synth_page(600, 400, 1)
    good = _save(a, quality=75)
    cut = good[: len(good) // 2] + b"\xff\xd9"
    dec = ops.JpegDecoder()
    blob, offs = dec.pack([good, cut])
    out, status = dec.decode(blob, offs)
    torch.cuda.synchronize()
    assert status.cpu().tolist() == [0, 1]
    assert np.array_equal(out[0].cpu().numpy(), _pil(good))
    b = io.BytesIO()
    Image.fromarray(a).save(b, "JPEG", progressive=True)
    assert ops.jpeg_probe(b.getvalue()) is None
    assert ops.jpeg_probe(b"not a jpeg at all") is None
    with pytest.raises(_abi.LuminaError):
        ops.jpeg_decode([b.getvalue()])
    with pytest.raises(_abi.LuminaError):          # mixed geometries in one batch
        ops.jpeg_decode([good, _save(oracle.synth_page(300, 400, 1), quality=75)])


@pytest.mark.gpu
def test_run_encoded_stream_every_batch_equals_the_chain_on_pillow_rasters(cuda, oracle):
    """The e2e path for files: what comes back for batch i is the chain applied to Pillow's decode of batch i's
    files (and a consumer may hold batch i-1 while batch i is produced)."""
    import torch
    from ocr_system_b200 import ops
    from ocr_system_b200.pipeline import PagePipeline

    h, w, nb, bs = 1000, 720, 4, 3
    batches, refs = [], []
    packer = ops.JpegDecoder()
    for b in range(nb):
        files = [_save(oracle.synth_page(h, w, 10 * b + i), quality=75) for i in range(bs)]
        blob, offs = packer.pack(files)
        batches.append((blob.clone().pin_memory(), offs))
        refs.append(np.stack([_pil(f) for f in files]))
    pipe = PagePipeline(max_dimension=480, device=cuda)
    held = None
    for i, (out, res, h2d, d2h) in enumerate(pipe.run_encoded_stream(batches)):
        want = pipe.run_device(torch.from_numpy(refs[i]).to(cuda))
        torch.cuda.synchronize()
        assert h2d == int(batches[i][1][-1])
        assert np.array_equal(out["pages"].numpy(), want.pages.cpu().numpy()), i
        assert np.array_equal(out["binary"].numpy(), want.binary.cpu().numpy()), i
        assert np.array_equal(out["angles"], want.angles), i
        # and against the CPU oracle for the first page of the batch
        tw, th = oracle.target_size(w, h, 480)
        ref_img, ref_angle, _ = oracle.deskew(oracle.resize_lanczos(refs[i][0], tw, th))
        assert out["angles"][0] == ref_angle and np.array_equal(out["pages"].numpy()[0], ref_img)
        if held is not None:   # keep=2: the previous batch's host results are still intact
            assert np.array_equal(held[0]["pages"].numpy(), held[1])
        held = (out, out["pages"].numpy().copy())


@pytest.mark.gpu
def test_gpu_decoder_survives_corrupt_scans(cuda, oracle):
    """Bit flips, truncations and garbage in the entropy-coded data: the kernels must stay inside their buffers
    (compute-sanitizer is closed on this GPU pool, so the guard is structural: every store index is derived from a
    bounds-checked block number) and report every page as decoded (0) or corrupt (1);
    pages whose bytes were not touched decode exactly."""
    import torch
    from ocr_system_b200 import ops

    rng = np.random.default_rng(11)
    good = _save(oracle.synth_page(400, 304, 5), quality=75)
    good_rst = _save(oracle.synth_page(400, 304, 6), quality=75, restart_marker_blocks=2)
    sos = good.index(b"\xff\xda")
    files = [good, good_rst]
    for k in range(14):
        src = bytearray(good if k % 2 == 0 else good_rst)
        start = src.index(b"\xff\xda") + 14
        if k % 3 == 0:
            for _ in range(1 + k):
                src[int(rng.integers(start, len(src) - 2))] ^= 1 << int(rng.integers(0, 8))
        elif k % 3 == 1:
            src = src[: int(rng.integers(start + 4, len(src) - 2))] + b"\xff\xd9"
        else:
            a = int(rng.integers(start, len(src) - 64))
            src[a:a + 48] = rng.integers(0, 256, 48, dtype=np.uint8).tobytes()
        files.append(bytes(src))
    dec = ops.JpegDecoder()
    blob, offs = dec.pack(files)
    out, status = dec.decode(blob, offs)
    torch.cuda.synchronize()
    st = status.cpu().tolist()
    assert st[0] == 0 and st[1] == 0 and set(st) <= {0, 1}
    assert np.array_equal(out[0].cpu().numpy(), _pil(good)) and np.array_equal(out[1].cpu().numpy(), _pil(good_rst))
    assert sos > 0
