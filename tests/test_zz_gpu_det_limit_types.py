"""GPU parity of upstream's other DetResizeForTest limit types ("min", "resize_long": images are enlarged) through
``ops.det_resize_normalize`` / ``paddle_ops.DetResizeNormalize`` against the oracle (which tests/test_oracle_pins.py pins
to cv2.resize + NumPy, exact).  The kernel is the one ``limit_type="max"`` uses; only the host size rule differs.
Added after the last GPU session of round 2 (no GPU minutes were left to run it), hence in a file that sorts last."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("limit_type,limit,h,w", [("min", 736, 300, 200), ("min", 736, 501, 1333), ("resize_long", 960, 333, 501),
                                                 ("resize_long", 640, 2000, 1413), ("min", 64, 7, 5)])
def test_det_resize_normalize_limit_types(oracle, cuda, limit_type, limit, h, w):
    import torch

    from ocr_system_b200.paddle_ops import DetResizeNormalize

    rng = np.random.default_rng(h * 13 + w)
    imgs = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    got, shape_list = DetResizeNormalize(limit_side_len=limit, limit_type=limit_type)(torch.from_numpy(imgs).to(cuda))
    for i in range(2):
        ref, sl = oracle.det_resize_normalize(imgs[i], limit, limit_type)
        assert got.shape[1:] == ref.shape
        assert tuple(shape_list[i]) == tuple(sl)
        # float stage tolerance from north_star: <= 1e-4 abs
        assert np.max(np.abs(got[i].cpu().numpy() - ref)) <= 1e-4


def test_preprocess_pages_for_azure_slices_large_groups(cuda, oracle):
    """A (size, mode) group larger than the per-submission cap (LUMINA_MAX_BATCH_PAGES, 64 by default) goes through the
    device in slices -- raster objects and JPEG files alike -- and the bytes are those of the per-page call."""
    import io

    from PIL import Image

    from ocr_system_b200.image_preprocessing import ImagePreprocessor

    pages = [Image.fromarray(oracle.synth_page(640, 452, s)) for s in range(5)]
    files = []
    for p in pages:
        buf = io.BytesIO()
        p.save(buf, format="JPEG", quality=80)
        files.append(buf.getvalue())
    ip = ImagePreprocessor(max_dimension=400)
    ip._MAX_BATCH_PAGES = 2
    want = [ip.preprocess_for_azure(p, target_size_mb=0.05) for p in pages]
    assert ip.preprocess_pages_for_azure(pages, target_size_mb=0.05) == want
    want_f = [ip.preprocess_for_azure(f, target_size_mb=0.05) for f in files]
    assert ip.preprocess_pages_for_azure(files, target_size_mb=0.05) == want_f
    assert ip.preprocess_pages_for_azure(pages + files, target_size_mb=0.05) == want + want_f
