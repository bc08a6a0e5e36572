"""Builds libhmm_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m pytorch_hmm_b200.build [--force] [--verbose]

The .so lands in pytorch_hmm_b200/lib/ so that it travels with the source tree (it is git-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
DEBUG_BUILD = os.environ.get("HMMB200_DEBUG_BUILD") == "1"     # timing traces / forced kernel variants (tools/*_trace.py); never shipped
LIBNAME = "libhmm_b200_dbg.so" if DEBUG_BUILD else "libhmm_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]
LINK_FLAGS = ["--shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
OBJDIR = os.path.join(HERE, "lib", "obj_dbg" if DEBUG_BUILD else "obj")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def lib_path() -> str:
    return os.path.join(LIBDIR, LIBNAME)


def _deps():
    deps = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    deps.append(os.path.join(HERE, "..", "include", "hmm_b200.h"))
    return deps


def source_hash() -> str:
    """sha256 over the CUDA sources, their headers, the public header and the compile flags: what the .so was built from."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS + LINK_FLAGS).encode())
    for d in _deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _hash_file(target: str) -> str:
    return target + ".srchash"


def is_stale(target: str = None) -> bool:
    """True when the library is missing or was built from other sources than the ones in the tree (content hash, so a copy
    of the tree that does not preserve mtimes -- the GPU-box snapshot -- does not trigger a rebuild)."""
    target = target or lib_path()
    if not os.path.exists(target) or not os.path.exists(_hash_file(target)):
        return True
    with open(_hash_file(target)) as f:
        return f.read().strip() != source_hash()


_stale = is_stale


def build_library(force: bool = False, verbose: bool = False) -> str:
    target = lib_path()
    if not force and not _stale(target):
        return target
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libhmm_b200.so (and there is no CPU fallback)")
    os.makedirs(OBJDIR, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    if DEBUG_BUILD:
        extra.append("-DHMMB200_DEBUG_HOOKS")

    headers = [d for d in _deps() if not d.endswith(".cu")]
    flag_tag = os.path.join(OBJDIR, ".flags")
    flags_now = " ".join(NVCC_FLAGS + extra)
    flags_same = os.path.exists(flag_tag) and open(flag_tag).read() == flags_now
    h_time = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):                                   # one translation unit per .cu, compiled side by side
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        if flags_same and not verbose and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), h_time):
            return src, obj, subprocess.CompletedProcess([], 0, "", "")      # object is newer than its source and every header
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    objs = []
    for src, obj, r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    with open(flag_tag, "w") as f:
        f.write(flags_now)
    subprocess.check_call([nvcc] + LINK_FLAGS + ["-o", target] + objs)
    with open(_hash_file(target), "w") as f:
        f.write(source_hash() + "\n")
    return target


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
