"""Compiles the reference's hot-path files to CPython bytecode under ``oracle/_ref/`` — TEST INFRASTRUCTURE ONLY.

The reference is pure Python; what a C reference's ``gcc`` recipe is for a compiled one, ``py_compile`` is here: the
eight hot-path files are compiled FROM THE SOURCES WHERE THEY LIE under ``/root/reference`` and only the outputs
(``*.code`` files: marshalled code objects in the .pyc format — the snapshot sent to the GPU box skips ``*.pyc`` — + a manifest) are written, into ``oracle/_ref/`` — git-ignored, so the history stays free of
reference code, but not gpurun-ignored, so the directory travels to the GPU box like the built ``.so``.  There
``oracle/ref_loader.py`` imports the bytecode (same image, same CPython 3.12 magic number), which lets

  * ``bench.py --impl reference`` / ``cpu_baseline`` time the reference's OWN functions on the box's host cores
    (``cpu_baseline.kind = "reference"``) instead of the restated port, and
  * ``tests/test_gpu_vs_reference.py`` compare the CUDA path with the reference itself on the GPU box.

No reference source text is copied anywhere.  ``__graft_entry__.build()`` runs this when ``/root/reference`` exists;
on the GPU box (no reference tree) the prebuilt files are used as they are.

    python oracle/build_ref.py            # compile (only files whose source hash changed)
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import py_compile
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
if str(HERE.parent) not in sys.path:
    sys.path.insert(0, str(HERE.parent))

from oracle import ref_loader  # noqa: E402  (the file table lives there)


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref holds bytecode of every hot-path file for this interpreter."""
    root = ref_loader.REFERENCE_ROOT
    if not all((root / f).is_file() for f in ref_loader._FILES.values()):
        return ref_loader.bytecode_available()
    OUT.mkdir(exist_ok=True)
    manifest_path = OUT / "MANIFEST.json"
    try:
        manifest = json.loads(manifest_path.read_text())
    except (OSError, ValueError):
        manifest = {}
    magic = importlib.util.MAGIC_NUMBER.hex()
    files = manifest.get("files", {}) if manifest.get("magic") == magic else {}
    for name, rel in ref_loader._FILES.items():
        src = root / rel
        digest = hashlib.sha256(src.read_bytes()).hexdigest()
        out = OUT / f"{name}.code"
        if files.get(name, {}).get("sha256") == digest and out.is_file():
            continue
        # dfile: the name tracebacks show (relative to the reference root); unchecked hash: valid without the source
        py_compile.compile(str(src), cfile=str(out), dfile=f"<reference>/{rel}", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        files[name] = {"source": rel, "sha256": digest}
        if verbose:
            print(f"[oracle/_ref] {rel} -> {out.relative_to(HERE.parent)}")
    manifest_path.write_text(json.dumps({"magic": magic, "python": sys.version.split()[0], "files": files}, indent=1))
    return True


if __name__ == "__main__":
    ok = build()
    print("[oracle/_ref]", "ready" if ok else "reference tree absent and no prebuilt bytecode")
