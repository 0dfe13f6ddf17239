"""In-tree nvcc build of libgkn_b200.so (sm_100a only).

The shared library is compiled straight from ``csrc/*.cu`` with nvcc (static cudart, no torch
headers, no libcuda link dependency) and written next to this file so that it travels with the
repo snapshot to the GPU box.  A content hash of the sources and flags is stored beside it; the
build is skipped when the hash matches.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "libgkn_b200.so"
HASH_PATH = PKG_DIR / "libgkn_b200.so.hash"

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-shared",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    files = sorted(CSRC.glob("*")) + sorted(INCLUDE.glob("*.h"))
    for f in files:
        if f.is_file():
            h.update(f.name.encode())
            h.update(f.read_bytes())
    return h.hexdigest()


def find_nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    return None


def is_current() -> bool:
    return LIB_PATH.exists() and HASH_PATH.exists() and HASH_PATH.read_text().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA library if it is missing or stale.  Returns the path of the .so."""
    if not force and is_current():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libgkn_b200.so (set NVCC=/path/to/nvcc)")
    # Several ranks of one torchrun job may get here at once with a stale library: serialise the
    # build with an exclusive file lock, re-check under the lock (the first rank builds, the others
    # find it current) and give every process its own temporary output name.
    import fcntl

    with open(LIB_PATH.with_suffix(".so.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():
                return LIB_PATH
            tmp = LIB_PATH.with_suffix(f".so.{os.getpid()}.tmp")
            cmd = [nvcc, *NVCC_FLAGS, f"-I{INCLUDE}", "-o", str(tmp), *map(str, _sources())]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if proc.returncode != 0:
                tmp.unlink(missing_ok=True)
                raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
            if verbose:
                print(proc.stderr)
            os.replace(tmp, LIB_PATH)
            HASH_PATH.write_text(_digest() + "\n")
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
