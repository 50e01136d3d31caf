"""Seeded synthetic OCR results ([box, text, conf] items) shared by the golden generator and the tests."""
import random

CASES = [(1, 1, "grid", 0.5), (2, 2, "ties", 0.5), (3, 37, "grid", 0.5), (4, 120, "grid", 0.3),
         (5, 300, "grid", 0.7), (6, 64, "ties", 0.5), (7, 200, "ties", 0.0), (8, 90, "float", 0.5),
         (9, 500, "float", 0.4), (10, 1000, "grid", 0.5), (11, 33, "float", 2.0),
         # zero tolerance on jittered rows: two lines one ulp apart whose rounded means come out in the other order, so the
         # reference's final sort by mean y (:181) moves a line (found by tools/sweep_paddle_vs_oracle.py)
         (27, 1000, "grid", 0.0)]


def page(seed, n, kind):
    rnd = random.Random(seed)
    items = []
    if kind == "grid":          # text-like layout: rows of words, jittered, shuffled detector order
        y = 40.0
        while len(items) < n:
            h = rnd.choice([18.0, 22.0, 27.5, 31.0])
            x = rnd.uniform(20, 60)
            while x < 900 and len(items) < n:
                w = rnd.uniform(30, 160)
                jy = rnd.choice([0.0, 0.5, -0.5, 1.25, -2.0, 3.0])
                tilt = rnd.choice([0.0, 0.0, 1.0, -1.5])
                box = [[x, y + jy], [x + w, y + jy + tilt], [x + w, y + jy + h + tilt], [x, y + jy + h]]
                items.append(box)
                x += w + rnd.uniform(4, 40)
            y += h * rnd.uniform(0.9, 1.8)
        rnd.shuffle(items)
    elif kind == "ties":        # integer boxes with many equal y centres and equal left edges
        for _ in range(n):
            x, y = float(rnd.randrange(0, 8) * 50), float(rnd.randrange(0, 6) * 30)
            h = float(rnd.choice([0, 10, 20]))
            items.append([[x, y], [x + 40.0, y], [x + 40.0, y + h], [x, y + h]])
    elif kind == "float":       # arbitrary doubles (sums are not exact: exercises the compensated sum)
        for _ in range(n):
            x, y = rnd.uniform(0, 1000), rnd.uniform(0, 1400)
            w, h = rnd.uniform(5, 200), rnd.uniform(1e-3, 60)
            items.append([[x, y], [x + w, y + rnd.uniform(-3, 3)], [x + w * 0.99, y + h], [x - rnd.uniform(0, 5), y + h]])
    res = []
    for i, box in enumerate(items):
        conf = rnd.random() if kind == "float" else float(rnd.randrange(50, 100)) / 100.0
        res.append([box, f"w{i}", conf])
    return res
