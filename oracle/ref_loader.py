"""Loads the REAL reference functions — TEST INFRASTRUCTURE ONLY.

Two sources, in this order:

* the reference tree itself (``/root/reference``, build container only): each hot-path file is imported by path;
* ``oracle/_ref/*.code`` (.pyc format) — CPython bytecode compiled from those same files by ``oracle/build_ref.py`` (outputs only,
  git-ignored, travels to the GPU box): what lets ``bench.py``'s CPU arm and ``tests/test_gpu_vs_reference.py`` run
  the reference's own functions where the tree does not exist.

Used by ``tests/golden/make_golden.py`` (fixtures), ``tests/test_oracle_vs_reference.py`` (pins
``oracle/reference_port.py`` against the reference), ``oracle/reference_live.py`` (the reference-backed step of the
CPU baseline) and the GPU parity tests.  Never imported by the product package.

The reference's package ``__init__``s do not import under current torchvision / without ``webcolors``
(SURVEY.md §4), so each hot-path file is loaded on its own.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("UDAPE_REFERENCE_ROOT", "/root/reference"))

_FILES = {
    "function": "adain/function.py",
    "style_net": "lib/models/Style_net.py",
    "keypoint_detection": "lib/keypoint_detection.py",
    "loss": "lib/models/loss.py",
    "ema": "lib/models/ema.py",
    "utils": "utils.py",
    "dataset_util": "lib/datasets/util.py",
    "adain_net": "adain/net.py",
}


BYTECODE_ROOT = Path(__file__).resolve().parent / "_ref"


def source_available() -> bool:
    return all((REFERENCE_ROOT / f).is_file() for f in _FILES.values())


def bytecode_available() -> bool:
    """oracle/_ref holds bytecode of every hot-path file, compiled by THIS interpreter version."""
    try:
        import json

        manifest = json.loads((BYTECODE_ROOT / "MANIFEST.json").read_text())
    except (OSError, ValueError):
        return False
    return (manifest.get("magic") == importlib.util.MAGIC_NUMBER.hex()
            and all((BYTECODE_ROOT / f"{n}.code").is_file() for n in _FILES))


def available() -> bool:
    return source_available() or bytecode_available()


def kind() -> str | None:
    """Where load() takes the reference from: "source" | "bytecode" | None."""
    return "source" if source_available() else ("bytecode" if bytecode_available() else None)


def load(name: str) -> types.ModuleType:
    """Import one reference hot-path file as module ``_udape_ref_<name>``."""
    mod_name = f"_udape_ref_{name}"
    if mod_name in sys.modules:
        return sys.modules[mod_name]
    if name == "adain_net":
        # adain/net.py does `from function import ...` (it is run with adain/ as the working directory)
        sys.modules.setdefault("function", load("function"))
    if source_available():
        spec = importlib.util.spec_from_file_location(mod_name, REFERENCE_ROOT / _FILES[name])
    elif bytecode_available():
        path = str(BYTECODE_ROOT / f"{name}.code")
        spec = importlib.util.spec_from_loader(mod_name, importlib.machinery.SourcelessFileLoader(mod_name, path), origin=path)
    else:
        raise ImportError(f"reference file {_FILES[name]}: neither {REFERENCE_ROOT} nor {BYTECODE_ROOT} holds it")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[mod_name] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(mod_name, None)
        raise
    return mod
