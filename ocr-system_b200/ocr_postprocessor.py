"""Reading order and line merging of OCR boxes -- drop-in for ``backend/utils/ocr_postprocessor.py``.

Same names, arguments and results as the reference module (TextBlock / MergedLine dataclasses,
``parse_rapidocr_output``, ``group_into_lines``, ``sort_and_merge_lines``, ``process_ocr_result``,
``format_merged_output``, ``extract_text_ordered``).  Parsing and string joining are host work (strings
never go to the GPU); the geometry -- stable sort by y centre, tolerance grouping against the running
mean of the open line, stable sort by left edge, per-line means -- runs in ``reading_order_kernel``
(``csrc/k_reading.cu``), one CTA per page, and ``process_ocr_results_batch`` does a whole batch of pages
in one launch.  There is no CPU fallback: without the CUDA library these functions raise.

Coordinates and confidences travel as float64 (Python floats), sums follow CPython's ``sum()``.
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops


@dataclass
class TextBlock:
    """A detected text block with its quad (ocr_postprocessor.py:19-39)."""
    text: str
    confidence: float
    box: List[List[float]]  # [[x1,y1], [x2,y2], [x3,y3], [x4,y4]]

    @property
    def y_center(self) -> float:
        return (self.box[0][1] + self.box[2][1]) / 2

    @property
    def x_left(self) -> float:
        return min(p[0] for p in self.box)

    @property
    def height(self) -> float:
        return abs(self.box[2][1] - self.box[0][1])


@dataclass
class MergedLine:
    """A merged line (ocr_postprocessor.py:42-48)."""
    text: str
    confidence: float
    y_position: float
    blocks: List[TextBlock]


def _as_block(item) -> Optional[TextBlock]:
    """One detector item -> TextBlock, or None for a shape the reference silently skips."""
    if all(hasattr(item, k) for k in ("box", "text", "score")):                  # dataclass items of newer RapidOCR
        quad = item.box if isinstance(item.box, list) else item.box.tolist()
        return TextBlock(text=item.text, confidence=item.score, box=quad)
    if isinstance(item, (list, tuple)) and len(item) >= 3:                        # [box, text, confidence]
        quad = item[0].tolist() if hasattr(item[0], "tolist") else item[0]
        return TextBlock(text=str(item[1]), confidence=float(item[2]), box=quad)
    return None


def parse_rapidocr_output(result) -> List[TextBlock]:
    """ocr_postprocessor.py:51-98.  Accepts a RapidOCR result object (``.ocr_result``) or a plain sequence;
    an item that raises while being read is reported on stdout and skipped, exactly like the reference."""
    items = getattr(result, "ocr_result", result) if result is not None else None
    blocks: List[TextBlock] = []
    if items is None:
        return blocks
    for item in items:
        try:
            blk = _as_block(item)
        except Exception as e:  # noqa: BLE001 - reference behaviour: report and continue
            print(f"  ⚠️ Failed to parse item: {e}")
            continue
        if blk is not None:
            blocks.append(blk)
    return blocks


def _device_order(pages: Sequence[Sequence[TextBlock]], y_tolerance_ratio: float, one_line: bool = False):
    """One launch for a batch of pages -> per page (order, line_of, nlines, line_conf, line_y) on the host."""
    counts = [len(p) for p in pages]
    offsets = np.zeros(len(pages) + 1, np.int32)
    np.cumsum(counts, out=offsets[1:])
    total = int(offsets[-1])
    if total == 0:
        return [(np.zeros(0, np.int32), np.zeros(0, np.int32), 0, np.zeros(0), np.zeros(0)) for _ in pages]
    try:
        boxes = np.array([b.box for p in pages for b in p], dtype=np.float64)   # one C-level pass over the quads
    except ValueError as e:
        raise ValueError("a text box must be four (x, y) points") from e
    if boxes.shape != (total, 4, 2):
        raise ValueError("a text box must be four (x, y) points")
    conf = np.array([b.confidence for p in pages for b in p], dtype=np.float64)
    dev = torch.device("cuda", torch.cuda.current_device())
    order, line_of, nlines, line_conf, line_y = ops.reading_order(
        torch.from_numpy(boxes).to(dev), torch.from_numpy(conf).to(dev), torch.from_numpy(offsets), y_tolerance_ratio,
        one_line=one_line)
    order, line_of, nlines = order.cpu().numpy(), line_of.cpu().numpy(), nlines.cpu().numpy()
    line_conf, line_y = line_conf.cpu().numpy(), line_y.cpu().numpy()
    out = []
    for i in range(len(pages)):
        a, b = int(offsets[i]), int(offsets[i + 1])
        out.append((order[a:b], line_of[a:b], int(nlines[i]), line_conf[a:b], line_y[a:b]))
    return out


def _lines_from(blocks: Sequence[TextBlock], order, line_of, nl) -> List[List[TextBlock]]:
    # line_of is non-decreasing along the reading order: lines are consecutive slices
    starts = np.searchsorted(line_of, np.arange(nl + 1), side="left").tolist()
    seq = [blocks[i] for i in order.tolist()]
    return [seq[starts[l]:starts[l + 1]] for l in range(nl)]


def group_into_lines(blocks: List[TextBlock], y_tolerance_ratio: float = 0.5) -> List[List[TextBlock]]:
    """ocr_postprocessor.py:101-143.  Lines top to bottom; inside a line the blocks keep the order of the
    y-sorted list (the reference appends in that order), i.e. ascending y_center, ties by input order."""
    if not blocks:
        return []
    order, line_of, nl, _, _ = _device_order([blocks], y_tolerance_ratio)[0]
    lines = _lines_from(blocks, order, line_of, nl)
    pos = {id(b): i for i, b in enumerate(blocks)}
    return [sorted(ln, key=lambda b: (b.y_center, pos[id(b)])) for ln in lines]


def _merge(blocks, order, line_of, nl, line_conf, line_y) -> List[MergedLine]:
    merged = []
    for l, ln in enumerate(_lines_from(blocks, order, line_of, nl)):
        merged.append(MergedLine(text=" ".join(b.text for b in ln), confidence=float(line_conf[l]),
                                 y_position=float(line_y[l]), blocks=ln))
    return merged


def sort_and_merge_lines(lines: List[List[TextBlock]], space_threshold_ratio: float = 2.0) -> List[MergedLine]:
    """ocr_postprocessor.py:146-182 for lines that are already grouped: every given line is sorted by x_left
    (stable), merged with single spaces, and the merged lines are sorted by mean y (stable).  Each line is
    sent to the device as its own one-line page (``one_line=True``)."""
    lines = [ln for ln in lines]
    if any(len(ln) == 0 for ln in lines):
        raise ZeroDivisionError("division by zero")  # the reference divides by len(sorted_line)
    if not lines:
        return []
    res = _device_order(lines, 0.0, one_line=True)
    merged = []
    for ln, (order, line_of, nl, line_conf, line_y) in zip(lines, res):
        merged.extend(_merge(ln, order, line_of, nl, line_conf, line_y))
    merged.sort(key=lambda m: m.y_position)
    return merged


def process_ocr_results_batch(results, y_tolerance_ratio: float = 0.5) -> List[List[MergedLine]]:
    """``process_ocr_result`` for many pages in one kernel launch."""
    pages = [parse_rapidocr_output(r) for r in results]
    live = [i for i, p in enumerate(pages) if p]
    out: List[List[MergedLine]] = [[] for _ in pages]
    if live:
        for i, r in zip(live, _device_order([pages[i] for i in live], y_tolerance_ratio)):
            out[i] = _merge(pages[i], *r)
    return out


def process_ocr_result(result, y_tolerance_ratio: float = 0.5, merge_lines: bool = True) -> List[MergedLine]:
    """ocr_postprocessor.py:185-213 (``merge_lines`` is accepted and ignored, as in the reference)."""
    return process_ocr_results_batch([result], y_tolerance_ratio)[0]


def format_merged_output(merged_lines: List[MergedLine], show_confidence: bool = False) -> str:
    """ocr_postprocessor.py:216-226: ``NN. text`` per line, with ``[conf]`` between them on request."""
    def row(i: int, ln: MergedLine) -> str:
        conf = f"[{ln.confidence:.2f}] " if show_confidence else ""
        return f"{i:02d}. {conf}{ln.text}"

    return "\n".join(row(i, ln) for i, ln in enumerate(merged_lines, 1))


def extract_text_ordered(result, y_tolerance: float = 0.5) -> str:
    """ocr_postprocessor.py:233-243."""
    return format_merged_output(process_ocr_result(result, y_tolerance_ratio=y_tolerance), show_confidence=False)
