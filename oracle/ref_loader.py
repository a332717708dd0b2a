"""Loads the REAL reference functions by file path — TEST INFRASTRUCTURE ONLY.

``/root/reference`` exists only in the build container, never on the GPU box, so nothing in
the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may depend on this module at run time.  It is
used (a) by ``tests/golden/make_golden.py`` to generate the committed fixtures and (b) by
``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent) to pin
``oracle/reference_port.py`` against the reference itself.

The reference's package ``__init__``s do not import under current torchvision / without
``webcolors``, so each hot-path file is loaded on its own with
``importlib.util.spec_from_file_location``.
"""
from __future__ import annotations

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


def available() -> bool:
    return all((REFERENCE_ROOT / f).is_file() for f in _FILES.values())


def load(name: str) -> types.ModuleType:
    """Import one reference hot-path file as module ``_udape_ref_<name>``."""
    mod_name = f"_udape_ref_{name}"
    if mod_name in sys.modules:
        return sys.modules[mod_name]
    path = REFERENCE_ROOT / _FILES[name]
    if name == "adain_net":
        # adain/net.py does `from function import ...` (it is run with adain/ as the working directory)
        sys.modules.setdefault("function", load("function"))
    spec = importlib.util.spec_from_file_location(mod_name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[mod_name] = mod
    spec.loader.exec_module(mod)
    return mod
