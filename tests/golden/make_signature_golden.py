"""Public surface of the reference modules on the hot path -> tests/golden/signature_golden.json.

Imports the UNMODIFIED reference (``/root/reference/backend/utils/{image_preprocessing,ocr_postprocessor}.py``, build
container only) and records, for every public callable, its parameter names, kinds and defaults (repr), plus the
dataclass fields of the post-processor's boundary types.  tests/test_abi_and_host.py requires the drop-in modules to
accept the same calls (same names / defaults in the same positions; extra trailing keyword parameters are allowed).

    python tests/golden/make_signature_golden.py
"""
import dataclasses
import inspect
import json
import logging
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))


def params(f):
    return [[p.name, p.kind.name, None if p.default is inspect.Parameter.empty else repr(p.default)]
            for p in inspect.signature(f).parameters.values()]


def surface(ip, pp):
    out = {"ImagePreprocessor": {}, "ocr_postprocessor": {}, "dataclasses": {}}
    for name, f in inspect.getmembers(ip.ImagePreprocessor, callable):
        if name == "__init__" or not name.startswith("_"):
            out["ImagePreprocessor"][name] = params(f)
    for name, f in inspect.getmembers(pp, inspect.isfunction):
        if f.__module__ == pp.__name__ and not name.startswith("_"):
            out["ocr_postprocessor"][name] = params(f)
    for name, c in inspect.getmembers(pp, inspect.isclass):
        if c.__module__ == pp.__name__ and dataclasses.is_dataclass(c):
            out["dataclasses"][name] = [f.name for f in dataclasses.fields(c)]
    out["module_singleton"] = type(ip.image_preprocessor).__name__
    return out


if __name__ == "__main__":
    logging.disable(logging.CRITICAL)
    sys.path.insert(0, "/root/reference/backend")
    import utils.image_preprocessing as ip
    import utils.ocr_postprocessor as pp

    with open(os.path.join(HERE, "signature_golden.json"), "w") as fh:
        json.dump(surface(ip, pp), fh, sort_keys=True)
        fh.write("\n")
    print("wrote signature_golden.json")
