"""Recipe for oracle/_ref/ (test / bench infrastructure; git-ignored, travels to the GPU box with the snapshot).

The reference's hot-path module is one pure-Python file, backend/utils/image_preprocessing.py.  This recipe places
an UNMODIFIED copy of it under oracle/_ref/ together with a two-line `config` stand-in (the module's only import from
the rest of the application is `from config import settings`, used for one default: OCR_MAX_IMAGE_DIMENSION, whose
value is read from the reference's own backend/config.py here), so that `bench.py --impl reference` and the CPU
baseline time the reference's own code on the GPU box's host cores (`cpu_baseline.kind == "reference"`) instead of
oracle/reference_port.py's restatement.  Nothing under oracle/_ref/ is committed; without /root/reference the recipe
is a no-op and the port is used.

    python oracle/make_ref.py            (also run by __graft_entry__.build())
"""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/backend"
OUT = os.path.join(HERE, "_ref")


def make(force: bool = False) -> str:
    src = os.path.join(REF, "utils", "image_preprocessing.py")
    if not os.path.exists(src):
        return "reference tree not present: oracle/_ref left as it is"
    os.makedirs(OUT, exist_ok=True)
    dst = os.path.join(OUT, "image_preprocessing.py")
    if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
        shutil.copyfile(src, dst)
    m = re.search(r"OCR_MAX_IMAGE_DIMENSION\s*:\s*int\s*=\s*(\d+)", open(os.path.join(REF, "config.py")).read())
    max_dim = int(m.group(1)) if m else 2000
    with open(os.path.join(OUT, "config.py"), "w") as f:
        f.write("# stand-in for backend/config.py (written by oracle/make_ref.py): the one setting the module reads\n"
                f"class _Settings:\n    OCR_MAX_IMAGE_DIMENSION = {max_dim}\n\n\nsettings = _Settings()\n")
    return f"oracle/_ref: reference image_preprocessing.py ({os.path.getsize(dst)} bytes), OCR_MAX_IMAGE_DIMENSION={max_dim}"


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
