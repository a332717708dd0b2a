"""In-tree build of the C-ABI library (``libudape_b200.so``) with nvcc for sm_100a.

The shared object is written next to the sources (``uda_poseestimation_b200/lib``) so it
travels with the repository snapshot to the GPU box; it is git-ignored.  Nothing here
needs a GPU: nvcc cross-compiles.

    python -m uda_poseestimation_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_DIR = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
OBJ_DIR = REPO_DIR / "build" / "udape_obj"
LIB_NAME = "libudape_b200.so"

SOURCES = ["api.cu", "adain.cu", "clamp.cu", "decode.cu", "loss.cu", "heatmap.cu", "ema.cu", "optim.cu", "dp.cu", "rewarp.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall",
    "-Xptxas", "-v",
    # parity first: no --use_fast_math (IEEE division / sqrt / accurate expf)
]


def find_nvcc() -> str:
    cand = [os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"]
    for c in cand:
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def lib_path() -> Path:
    # UDAPE_LIB=/path/to/variant.so loads an alternative build (kernel A/B experiments on one box)
    override = os.environ.get("UDAPE_LIB")
    return Path(override) if override else LIB_DIR / LIB_NAME


def _newest_dep_mtime() -> float:
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((REPO_DIR / "include").glob("*.h"))
    deps.append(Path(__file__))
    return max(p.stat().st_mtime for p in deps)


def needs_build() -> bool:
    if os.environ.get("UDAPE_LIB"):
        return False
    lp = lib_path()
    return (not lp.exists()) or lp.stat().st_mtime < _newest_dep_mtime()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library. Idempotent.

    One process per GPU means several ranks may get here at once on a fresh checkout: the build runs under an
    exclusive file lock (the others wait, then find the library current), and the library is linked to a
    temporary name and renamed into place, so no rank can ``dlopen`` a half-written file."""
    if not force and not needs_build():
        return lib_path()
    import fcntl

    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    with open(OBJ_DIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():   # another rank built it while this one waited
                return lib_path()
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    nvcc = find_nvcc()
    header_mtime = max(
        [p.stat().st_mtime for p in CSRC.glob("*.cuh")]
        + [p.stat().st_mtime for p in (REPO_DIR / "include").glob("*.h")]
        + [Path(__file__).stat().st_mtime]
    )
    sources = [s for s in SOURCES if (CSRC / s).exists()]

    def compile_one(src: str) -> Path:
        s = CSRC / src
        o = OBJ_DIR / (s.stem + ".o")
        if not force and o.exists() and o.stat().st_mtime >= max(s.stat().st_mtime, header_mtime):
            return o
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(REPO_DIR / "include"), "-c", str(s), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = (r.stdout or "") + (r.stderr or "")
        (OBJ_DIR / (s.stem + ".ptxas.log")).write_text(log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        if verbose:
            sys.stderr.write(log)
        return o

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    # default visibility is hidden; the extern "C" entry points are exported explicitly
    tmp = lib_path().with_name(f".{LIB_NAME}.{os.getpid()}.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-cudart", "static",
           "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    os.replace(tmp, lib_path())
    return lib_path()


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
