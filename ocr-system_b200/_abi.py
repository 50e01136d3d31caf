"""ctypes binding of the C-ABI in ``include/lumina_b200.h``.

This is the binding a maintainer of the (Python) reference would add: one
``CDLL`` and typed prototypes, no torch types cross the boundary.  There is no
CPU fallback: a missing library is a hard ``ImportError``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblumina_b200.so")


class LuminaError(RuntimeError):
    """A C-ABI call returned a negative status."""


_P = C.c_void_p
_I = C.c_int
_Z = C.c_size_t
_F = C.c_float
_D = C.c_double

# name -> (restype, argtypes)   (mirrors include/lumina_b200.h, one entry per symbol)
PROTOTYPES = {
    "lumina_abi_version": (_I, []),
    "lumina_last_error_string": (C.c_char_p, []),
    "lumina_launch_count": (C.c_uint64, []),
    "lumina_exif_transpose_u8": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "lumina_target_size": (None, [_I, _I, _I, C.POINTER(_I), C.POINTER(_I)]),
    "lumina_resize_plan_create": (_I, [_I, _I, _I, _I, C.POINTER(_P)]),
    "lumina_resize_plan_destroy": (None, [_P]),
    "lumina_resize_workspace_bytes": (_Z, [_P, _I, _I]),
    "lumina_resize_lanczos_u8": (_I, [_P, _P, _P, _I, _I, _P, _Z, _P]),
    "lumina_nearest_table_host": (None, [_I, _I, _P]),
    "lumina_resize_nearest_u8": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "lumina_rgb2gray_pil_u8": (_I, [_P, _P, _Z, _P]),
    "lumina_rgb2gray_cv_u8": (_I, [_P, _P, _Z, _P]),
    "lumina_contrast_mean_u8": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "lumina_contrast_apply_u8": (_I, [_P, _P, _I, _I, _I, _I, _P, _F, _P]),
    "lumina_sharpness_u8": (_I, [_P, _P, _I, _I, _I, _I, _F, _P]),
    "lumina_contrast_sharpness_u8": (_I, [_P, _P, _I, _I, _I, _I, _P, _F, _F, _P]),
    "lumina_median3_u8": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "lumina_binarize_u8": (_I, [_P, _P, _Z, _I, _I, _P]),
    "lumina_rgbx_to_rgb_u8": (_I, [_P, _P, _Z, _P]),
    "lumina_adaptive_gauss11_u8": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "lumina_adaptive_gauss11_ex_u8": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "lumina_canny_workspace_bytes": (_Z, [_I, _I, _I]),
    "lumina_canny_u8": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "lumina_ppht_workspace_bytes": (_Z, [_I, _I, _I, _D, _D]),
    "lumina_ppht": (_I, [_P, _I, _I, _I, _D, _D, _I, _I, _I, _P, _P, _I, _P, _Z, _P]),
    "lumina_ppht_prepare": (_I, [_P, _I, _I, _I, _D, _D, _P, _Z, _P]),
    "lumina_ppht_lines": (_I, [_P, _I, _I, _I, _D, _D, _I, _I, _I, _P, _P, _I, _P, _Z, _P]),
    "lumina_median_angle_host": (_D, [_P, _I]),
    "lumina_copy_lines_to_host": (_I, [_P, _I, _I, _I, _P, _P]),
    "lumina_deskew_decide_host": (None, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "lumina_deskew_decide_angles_host": (None, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "lumina_rotation_matrix_host": (None, [_D, _D, _D, _D, _P]),
    "lumina_warp_affine_cubic_u8": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "lumina_invert_affine_host": (None, [_P, _P]),
    "lumina_det_target_size": (None, [_I, _I, _I, C.POINTER(_I), C.POINTER(_I)]),
    "lumina_det_target_size_ex": (_I, [_I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I)]),
    "lumina_skew_workspace_bytes_for": (_Z, [_I, _I, _I]),
    "lumina_skew_estimate_fast": (_I, [_P, _I, _I, _I, _P, _P, _Z, _P]),
    "lumina_otsu_u8": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "lumina_sauvola_u8": (_I, [_P, _P, _I, _I, _I, _I, _D, _D, _P]),
    "lumina_det_resize_normalize": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _F, _P]),
    "lumina_ctc_workspace_bytes": (_Z, [_I, _I]),
    "lumina_ctc_greedy": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "lumina_db_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "lumina_db_postprocess": (_I, [_P, _I, _I, _I, _F, _D, _D, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "lumina_db_postprocess_ex": (_I, [_P, _I, _I, _I, _F, _D, _D, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "lumina_db_mask_ccl": (_I, [_P, _I, _I, _I, _F, _P, _P, _P, _Z, _P]),
    "lumina_reading_order": (_I, [_P, _P, _P, _I, _I, _D, _I, _P, _P, _P, _P, _P, _P]),
    "lumina_jpeg_workspace_bytes": (_Z, [_I, _I, _I]),
    "lumina_jpeg_encode_rgb": (_I, [_P, _I, _I, _I, _I, _I, _P, _Z, _P, _P, _Z, _P]),
    "lumina_jpeg_probe": (_I, [_P, _Z, _P]),
    "lumina_jpeg_decode_workspace_bytes": (_Z, [_I, _I, _I, _I, _I, _I, _Z]),
    "lumina_jpeg_decode_stage_bytes": (_Z, [_I]),
    "lumina_jpeg_decode_batch": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _Z, _P]),
    "lumina_synth_pages_u8": (_I, [_P, _I, _I, _I, C.c_uint64, _P]),
    "lumina_synth_prob_maps_f32": (_I, [_P, _I, _I, _I, C.c_uint64, _P]),
    "lumina_synth_ctc_f32": (_I, [_P, _I, _I, _I, C.c_uint64, C.c_uint32, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the typed library.  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension was not built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU fallback for this path."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().lumina_last_error_string().decode("utf-8", "replace")
        raise LuminaError(f"lumina_b200 error {rc}: {msg}")
