"""Batched submission for the reference's ``OCRService`` page loop.

The reference processes a PDF one page at a time (``backend/services/ocr_service.py:604-660``): ``pdf_to_images`` and
then, per page under ``Semaphore(1)`` (``:398-475``), ``image_preprocessor.preprocess_for_azure(page, ...)`` followed
by the Azure call.  On a GPU the preprocessing of all pages of a document is ONE batched call
(``ImagePreprocessor.preprocess_pages_for_azure``); this module lets the reference benefit from that without touching
its loop, its dataclasses or its Azure client (all out of scope here):

    import services.ocr_service as svc
    from ocr_system_b200.ocr_service_adapter import install
    install(svc)           # svc.image_preprocessor becomes a BatchingPreprocessor around the B200 drop-in

``BatchingPreprocessor`` is the drop-in preprocessor plus a look-ahead: the page list that ``pdf_to_images`` returned
is remembered; when the service asks for the first of those pages, every remembered page is preprocessed in one
batched device call with the very arguments the service passed, and the remaining per-page calls are dictionary
look-ups.  Images the adapter has not seen (``process_image_sync``) take the ordinary per-image path.  The bytes are
the ones the per-page call returns (tests/test_gpu_dropin.py), so ``OCROutput.processed_image_bytes`` is unchanged.
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional, Sequence, Tuple

from PIL import Image


class BatchingPreprocessor:
    """Proxy around an ``ImagePreprocessor`` (the B200 drop-in): same attributes and methods; ``pdf_to_images`` +
    ``preprocess_for_azure`` cooperate so that the pages of a document are preprocessed as one batch."""

    def __init__(self, inner, max_batch: int = 64):
        self._inner = inner
        self._max_batch = int(max_batch)
        self._lock = threading.Lock()
        # per calling thread (the service runs in asyncio.to_thread workers, ocr_service.py:674-676)
        self._tls = threading.local()
        self.batched_calls = 0      # how many batched device submissions were made (diagnostics / tests)

    def __getattr__(self, name):
        return getattr(self._inner, name)

    # ---- the two methods the page loop uses -------------------------------------------------------------------
    def pdf_to_images(self, pdf_path, dpi: Optional[int] = None) -> List[Image.Image]:
        pages = self._inner.pdf_to_images(pdf_path, dpi)
        self.expect_pages(pages)
        return pages

    def expect_pages(self, pages: Sequence[Image.Image]) -> None:
        """Announce a page list that is about to be preprocessed one by one (what ``pdf_to_images`` does itself)."""
        self._tls.pending = {id(p): p for p in pages}
        self._tls.order = [id(p) for p in pages]
        self._tls.done = {}
        self._tls.key = None

    def preprocess_for_azure(self, image, apply_deskew: bool = True, apply_binarize: bool = False,
                             apply_contrast: bool = True, apply_sharpness: bool = True,
                             target_size_mb: float = 2.0) -> bytes:
        pending: Dict[int, Image.Image] = getattr(self._tls, "pending", None) or {}
        key: Tuple = (apply_deskew, apply_binarize, apply_contrast, apply_sharpness, target_size_mb)
        done: Dict[int, bytes] = getattr(self._tls, "done", {})
        k = id(image)
        if k in done and self._tls.key == key:
            pending.pop(k, None)
            return done.pop(k)
        if k not in pending:
            return self._inner.preprocess_for_azure(image, apply_deskew, apply_binarize, apply_contrast, apply_sharpness,
                                                    target_size_mb)
        # first request for a remembered page: submit it together with the pages that follow it
        order = self._tls.order
        start = order.index(k)
        ids = [i for i in order[start:start + self._max_batch] if i in pending]
        batch = [pending[i] for i in ids]
        results = self._inner.preprocess_pages_for_azure(batch, apply_deskew, apply_binarize, apply_contrast,
                                                         apply_sharpness, target_size_mb)
        with self._lock:
            self.batched_calls += 1
        self._tls.key = key
        done.clear()
        done.update(dict(zip(ids, results)))
        self._tls.done = done
        pending.pop(k, None)
        return done.pop(k)


def install(ocr_service_module, preprocessor=None, max_batch: int = 64) -> BatchingPreprocessor:
    """Point the reference module's ``image_preprocessor`` (``from utils.image_preprocessing import image_preprocessor``,
    ocr_service.py:39) at a ``BatchingPreprocessor`` around the B200 drop-in.  Returns the proxy."""
    if preprocessor is None:
        from .image_preprocessing import image_preprocessor as preprocessor
    proxy = BatchingPreprocessor(preprocessor, max_batch=max_batch)
    ocr_service_module.image_preprocessor = proxy
    return proxy
