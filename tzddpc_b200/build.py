"""Build libtzddpc.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m tzddpc_b200.build [--force]

The library is a plain C-ABI shared object (include/tzddpc.h); it links the CUDA runtime
statically so that it does not depend on which libcudart PyTorch happened to load.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT_DIR = HERE / "lib"
LIB = OUT_DIR / "libtzddpc.so"
SOURCES = ["tz_abi.cu", "tz_fused.cu", "tz_fast.cu", "tz_bucket1.cu", "tz_bucket2.cu", "tz_bucket3.cu", "tz_big.cu", "tz_zono.cu", "tz_identify.cu", "tz_rng.cu", "tz_gain.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("TZ_EXTRA_NVCC", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _closure(path: Path, seen=None) -> list:
    """`path` and every header it includes with quotes, transitively (paths relative to the including file)."""
    seen = {} if seen is None else seen
    path = path.resolve()
    if path in seen or not path.exists():
        return list(seen)
    seen[path] = True
    for inc in _INC.findall(path.read_text()):
        _closure(path.parent / inc, seen)
    return list(seen)


def _digest_of(src: Path) -> str:
    h = hashlib.sha256()
    for p in sorted(_closure(src)):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _digest() -> str:
    h = hashlib.sha256()
    for s in SOURCES:
        h.update(_digest_of(CSRC / s).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> Path:
    """Compiles the translation units whose source or (transitively) included headers changed since their object was
    built, then links.  Per-object digests live next to the objects; the library digest in lib/libtzddpc.sha256."""
    OUT_DIR.mkdir(exist_ok=True)
    stamp = OUT_DIR / "libtzddpc.sha256"
    digest = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objdir = OUT_DIR / "obj"
    objdir.mkdir(exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = objdir / (src.replace(".cu", ".o"))
        ostamp = objdir / (src.replace(".cu", ".sha256"))
        d = _digest_of(CSRC / src)
        if not force and obj.exists() and ostamp.exists() and ostamp.read_text().strip() == d:
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print("[tzddpc_b200.build]", " ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        ostamp.write_text(d)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", str(LIB), *map(str, objs)]
    if verbose:
        print("[tzddpc_b200.build]", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    print(build(force=args.force))
