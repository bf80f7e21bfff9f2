"""Builds libgemmgan_sm100a.so in-tree with nvcc (sm_100a only, -lineinfo).

The .so lands next to this file so that it travels with the repo snapshot to the GPU box.
Incremental: a translation unit is recompiled only when it or a header changed.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG / "build"
LIB = PKG / "libgemmgan_sm100a.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    "-I", str(PKG.parent / "include"),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libgemmgan_sm100a.so cannot be built (there is no CPU fallback)")


def _headers_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "gemmgan.h"]):
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc: str, src: Path, hdr: str, verbose: bool) -> Path:
    obj = BUILD / (src.stem + ".o")
    stamp = BUILD / (src.stem + ".stamp")
    digest = hashlib.sha256(src.read_bytes() + hdr.encode()).hexdigest()
    if obj.exists() and stamp.exists() and stamp.read_text() == digest:
        return obj
    cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (BUILD / (src.stem + ".ptxas.log")).write_text(r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed on {src.name}")
    if verbose:
        sys.stderr.write(r.stderr)
    stamp.write_text(digest)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    nvcc = _nvcc()
    BUILD.mkdir(exist_ok=True)
    if force:
        for p in BUILD.glob("*.stamp"):
            p.unlink()
    srcs = sorted(CSRC.glob("*.cu"))
    hdr = _headers_digest()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, hdr, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if (not LIB.exists()) or LIB.stat().st_mtime < newest or force:
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
