"""PaddleOCR-style operators with the upstream call signatures, on the GPU.

The reference never vendors or calls these: they are the local-engine
ops BASELINE.json's north_star names, restated from upstream PaddleOCR
(ppocr/postprocess/{db_postprocess,rec_postprocess}.py, ppocr/data/imaug/operators.py)
-- parity unpinned by the reference; pinned by the restated oracle + cv2.

  DetResizeNormalize  DetResizeForTest(limit_side_len, limit_type) + NormalizeImage + ToCHWImage
  DBPostProcess       (outs_dict{"maps": [N,1,H,W]}, shape_list) -> [{"points": int32[K,4,2]}]
  CTCLabelDecode      (preds[N,T,C]) -> [(text, conf)]
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops


def _to_cuda(x, dtype) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"expected ndarray or tensor, got {type(x)}")
    if not torch.cuda.is_available():
        raise RuntimeError("ocr_system_b200 needs a CUDA device: there is no CPU fallback")
    return x.to(device="cuda", dtype=dtype, non_blocking=True).contiguous()


class DetResizeNormalize:
    def __init__(self, limit_side_len: int = 960, mean=ops.DET_MEAN, std=ops.DET_STD, scale: float = 1.0 / 255.0,
                 limit_type: str = "max"):
        """``limit_type``: upstream's "max" (PaddleOCR's inference default, det_limit_type), "min" or "resize_long"."""
        if limit_type not in ops.DET_LIMIT_TYPES:
            raise ValueError(f"limit_type must be one of {sorted(ops.DET_LIMIT_TYPES)}, got {limit_type!r}")
        self.limit_side_len, self.mean, self.std, self.scale = limit_side_len, mean, std, scale
        self.limit_type = limit_type

    def __call__(self, pages) -> Tuple[torch.Tensor, np.ndarray]:
        """pages: uint8 [N,H,W,3] (ndarray or tensor) -> (CUDA f32 [N,3,oh,ow], shape_list [N,4])."""
        return ops.det_resize_normalize(_to_cuda(pages, torch.uint8), self.limit_side_len, self.mean, self.std,
                                        self.scale, self.limit_type)


class DBPostProcess:
    """upstream signature (ppocr/postprocess/db_postprocess.py):
    DBPostProcess(thresh, box_thresh, max_candidates, unclip_ratio, use_dilation, score_mode, box_type)
    (outs_dict{"maps": [N,1,H,W]}, shape_list[N x (src_h, src_w, ratio_h, ratio_w)]) -> [{"points": int32[K,4,2]}]

    Supported: box_type="quad" (what upstream ships in its det configs) with score_mode "fast" or "slow", with or
    without ``use_dilation`` (RapidOCR's default).  box_type="poly" needs the ordered contour (approxPolyDP) and a
    general polygon offset (Clipper union of a non-convex path) that the label-based device pipeline never builds;
    it raises instead of silently doing something different."""

    def __init__(self, thresh=0.3, box_thresh=0.7, max_candidates=1000, unclip_ratio=2.0, use_dilation=False,
                 score_mode="fast", box_type="quad", **kwargs):
        if score_mode not in ("fast", "slow"):
            raise ValueError("score_mode must be 'fast' or 'slow'")
        if box_type != "quad":
            raise NotImplementedError("box_type='poly' is not implemented (only 'quad')")
        self.score_mode = score_mode
        self.use_dilation = bool(use_dilation)
        self.thresh, self.box_thresh = thresh, box_thresh
        self.max_candidates, self.unclip_ratio = max_candidates, unclip_ratio
        self.min_size = 3

    def __call__(self, outs_dict, shape_list, with_scores: bool = False):
        maps = outs_dict["maps"]
        pred = _to_cuda(maps, torch.float32)
        if pred.dim() != 4:
            raise ValueError("maps must be [N,1,H,W]")
        pred = pred[:, 0, :, :].contiguous()
        sl = np.asarray(shape_list, dtype=np.float64)
        src_hw = sl[:, :2].astype(np.int32)
        boxes, scores, counts = ops.db_postprocess(pred, src_hw, self.thresh, self.box_thresh, self.unclip_ratio,
                                                   self.max_candidates, self.min_size, self.use_dilation, self.score_mode)
        counts_h = counts.cpu().numpy()
        kmax = int(counts_h.max(initial=0))
        boxes_h = boxes[:, : max(kmax, 1)].cpu().numpy()
        scores_h = scores[:, : max(kmax, 1)].cpu().numpy() if with_scores else None
        out = []
        for b in range(pred.shape[0]):
            d = {"points": boxes_h[b, : counts_h[b]].copy()}
            if with_scores:
                d["scores"] = scores_h[b, : counts_h[b]].astype(np.float64)
            out.append(d)
        return out


class CTCLabelDecode:
    """Greedy CTC decode (upstream BaseRecLabelDecode/CTCLabelDecode).

    character_dict_path: text file, one symbol per line; None -> "0123456789abcdefghijklmnopqrstuvwxyz".
    Class 0 is the blank; ``use_space_char`` appends ' '."""

    def __init__(self, character_dict_path: Optional[str] = None, use_space_char: bool = False,
                 character: Optional[Sequence[str]] = None):
        if character is not None:
            chars = list(character)
        elif character_dict_path is None:
            chars = list("0123456789abcdefghijklmnopqrstuvwxyz")
        else:
            chars = []
            with open(character_dict_path, "rb") as f:
                for line in f.readlines():
                    chars.append(line.decode("utf-8").strip("\n").strip("\r\n"))
            if use_space_char:
                chars.append(" ")
        self.reverse = character_dict_path is not None and "arabic" in str(character_dict_path)
        self.character = ["blank"] + chars

    def decode_indices(self, preds):
        """Device part only: (idx[N,T], pos[N,T], len[N], conf[N]) CUDA tensors."""
        return ops.ctc_greedy(_to_cuda(preds, torch.float32))

    _LATIN = frozenset("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 :*./%+-")

    @classmethod
    def pred_reverse(cls, symbols: Sequence[str]) -> str:
        """upstream BaseRecLabelDecode.pred_reverse (right-to-left dictionaries): the order of the decoded symbols is
        reversed, but a run of Latin letters / digits / `` :*./%+-`` keeps its own left-to-right order."""
        parts: List[str] = []
        run = ""
        for c in symbols:
            if any(ch in cls._LATIN for ch in c):
                run += c
            else:
                if run:
                    parts.append(run)
                parts.append(c)
                run = ""
        if run:
            parts.append(run)
        return "".join(reversed(parts))

    def _text(self, symbols: Sequence[str]) -> str:
        return self.pred_reverse(symbols) if self.reverse else "".join(symbols)

    def decode_label(self, label) -> List[Tuple[str, float]]:
        """upstream ``decode(label)`` (no duplicate removal): drop the blank class, map the rest; the confidence of a
        ground-truth string is 1.0 (0.0 for an empty one: upstream's ``conf_list = [0]`` before ``np.mean``)."""
        out = []
        for row in np.asarray(label):
            symbols = [self.character[int(k)] for k in row if int(k) != 0]
            out.append((self._text(symbols), 1.0 if symbols else 0.0))
        return out

    def __call__(self, preds, label=None, *args, **kwargs):
        """-> [(text, conf)]; with ``label`` (int [N,L], 0 = blank / padding) -> ([(text, conf)], [(label_text, 1.0)])
        like upstream."""
        if isinstance(preds, (tuple, list)):
            preds = preds[-1]
        n_cls = preds.shape[2]
        if n_cls > len(self.character):
            raise ValueError(f"preds have {n_cls} classes but the dictionary holds {len(self.character)}")
        idx, _pos, ln, conf = self.decode_indices(preds)
        idx, ln, conf = idx.cpu().numpy(), ln.cpu().numpy(), conf.cpu().numpy()
        chars = self.character
        out = [(self._text([chars[k] for k in idx[b, : ln[b]]]), float(conf[b])) for b in range(idx.shape[0])]
        if label is None:
            return out
        return out, self.decode_label(label)
