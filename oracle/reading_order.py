"""CPU oracle for the reading-order / line-merge step -- TEST INFRASTRUCTURE ONLY.

Plain-Python restatement of ``backend/utils/ocr_postprocessor.py`` of the reference on index arrays instead
of TextBlock objects, so the CUDA path (``reading_order_kernel``) can be compared element by element.
Pinned against the reference itself: ``tests/golden/make_reading_golden.py`` imports the reference module
in the build container and stores its outputs in ``tests/golden/reading_order_golden.json``
(``tests/test_reading_order.py`` checks this file against them).  Nothing in the product imports this.
"""
from typing import List, Sequence, Tuple


def y_center(box) -> float:      # ocr_postprocessor.py:26-29
    return (box[0][1] + box[2][1]) / 2


def x_left(box) -> float:        # ocr_postprocessor.py:31-34
    return min(p[0] for p in box)


def height(box) -> float:        # ocr_postprocessor.py:36-39
    return abs(box[2][1] - box[0][1])


def group_into_lines(boxes: Sequence, y_tolerance_ratio: float = 0.5) -> List[List[int]]:
    """ocr_postprocessor.py:101-143 on indices: lines of block indices in y-sorted (append) order."""
    n = len(boxes)
    if n == 0:
        return []
    srt = sorted(range(n), key=lambda i: y_center(boxes[i]))                 # :116 (stable)
    avg_height = sum(height(boxes[i]) for i in srt) / n                       # :119
    y_tolerance = avg_height * y_tolerance_ratio                              # :120
    lines: List[List[int]] = []
    cur = [srt[0]]
    cur_y = y_center(boxes[srt[0]])
    for i in srt[1:]:                                                         # :126-137
        if abs(y_center(boxes[i]) - cur_y) <= y_tolerance:
            cur.append(i)
            cur_y = sum(y_center(boxes[j]) for j in cur) / len(cur)
        else:
            lines.append(cur)
            cur = [i]
            cur_y = y_center(boxes[i])
    if cur:
        lines.append(cur)
    return lines


def sort_and_merge_lines(boxes: Sequence, conf: Sequence[float], lines: List[List[int]]):
    """ocr_postprocessor.py:146-182 on indices -> list of (block indices left to right, mean conf, mean y),
    sorted by mean y (stable)."""
    merged = []
    for ln in lines:
        s = sorted(ln, key=lambda i: x_left(boxes[i]))                        # :162
        avg_conf = sum(conf[i] for i in s) / len(s)                           # :170
        avg_y = sum(y_center(boxes[i]) for i in s) / len(s)                   # :171
        merged.append((s, avg_conf, avg_y))
    merged.sort(key=lambda m: m[2])                                           # :181
    return merged


def reading_order(boxes: Sequence, conf: Sequence[float], y_tolerance_ratio: float = 0.5
                  ) -> Tuple[List[int], List[int], int, List[float], List[float]]:
    """The flat form the kernel emits: (order, line_of, nlines, line_conf, line_y)."""
    merged = sort_and_merge_lines(boxes, conf, group_into_lines(boxes, y_tolerance_ratio))
    order, line_of = [], []
    for l, (s, _, _) in enumerate(merged):
        order += s
        line_of += [l] * len(s)
    return order, line_of, len(merged), [m[1] for m in merged], [m[2] for m in merged]
