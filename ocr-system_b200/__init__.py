"""ocr-system_b200 -- B200-native page-image hot path for Lumina OCR.

Importable as ``ocr_system_b200`` (the directory name carries a hyphen, so
``ocr_system_b200/__init__.py`` at the repo root points its ``__path__`` here).

  _abi                 ctypes binding of include/lumina_b200.h (the C-ABI .so)
  ops                  batched page-tensor operators (torch = allocator + streams)
  image_preprocessing  drop-in ``ImagePreprocessor`` / ``image_preprocessor``
                       (reference: backend/utils/image_preprocessing.py)
  paddle_ops           DBPostProcess / CTCLabelDecode / DetPreprocess (upstream PaddleOCR API)
  pipeline             PagePipeline: batched host->HBM->host page chain, page sharding
"""
__version__ = "0.1.0"
