"""Builds ocr-system_b200/liblumina_b200.so (C-ABI, sm_100a) from csrc/*.cu with nvcc.

In-tree build: the .so sits next to this file so it travels with a repo snapshot
(git-ignored).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
SO = os.path.join(HERE, "liblumina_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off"]
# float32 evaluation order is part of the spec for these files: never contract to FMA
NO_FMA = {"k_stencil.cu", "k_det.cu", "k_point.cu", "k_ppht.cu", "k_dbpost.cu", "k_ctc.cu", "k_binarize.cu"}


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the lumina_b200 extension cannot be built")


def _deps_mtime() -> float:
    m = 0.0
    for d in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(d):
            m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _deps_mtime():
        return SO
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [nvcc, *ARCH, *COMMON]
        if src in NO_FMA:
            cmd.append("-fmad=false")
        if os.environ.get("LUMINA_PPHT_PROFILE") and src == "k_ppht.cu":
            cmd.append("-DLUMINA_PPHT_PROFILE=1")  # per-phase clock64 ticks in the stats buffer (tools/ppht_stats.py)
        if src == "k_jpegd.cu":  # experiment knobs (sub-sequence length / CTA size of the entropy kernel)
            for k in ("LUMINA_JD_SWL", "LUMINA_JD_CHUNK"):
                if os.environ.get(k):
                    cmd.append(f"-D{k}={os.environ[k]}")
        cmd += ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, *ARCH, "-shared", "-cudart", "static", "-o", SO, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
