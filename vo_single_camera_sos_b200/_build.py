"""In-tree nvcc build of libsosfront.so (sm_100a only).

The shared library is written next to this file so that it travels to the GPU box with the
repository snapshot; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_DIR = PKG_DIR.parent
CSRC_DIR = PKG_DIR / "csrc"
INCLUDE_DIR = REPO_DIR / "include"
BUILD_DIR = REPO_DIR / "build" / "sosfront"
LIB_PATH = PKG_DIR / "libsosfront.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    f"-I{INCLUDE_DIR}", f"-I{CSRC_DIR}",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libsosfront.so cannot be built")
    return exe


def sources() -> list[Path]:
    return sorted(CSRC_DIR.glob("*.cu"))


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link libsosfront.so. Returns the library path."""
    srcs = sources()
    headers = sorted(CSRC_DIR.glob("*.cuh")) + sorted(INCLUDE_DIR.glob("*.h"))
    if not force and not _stale(LIB_PATH, srcs + headers):
        return LIB_PATH
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: Path) -> Path:
        obj = BUILD_DIR / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs)]  # static cudart: independent of torch's
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
