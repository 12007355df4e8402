"""Builds libtstereo.so (the C-ABI engine, include/tstereo.h) in-tree with nvcc for sm_100a.

    python -m temporalstereo_b200.build [--force] [--verbose]

The library is plain CUDA C++ (no torch headers): every csrc/*.cu is compiled to an object under
build/ and linked into temporalstereo_b200/libtstereo.so.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(PKG, "libtstereo.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def source_id() -> str:
    """sha1 over every source that goes into libtstereo.so; baked into the library as
    tstereo_build_id() so a stale .so is detected at load time (_lib.load)."""
    import hashlib
    h = hashlib.sha1()
    files = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")))
    files.append(os.path.join(ROOT, "include", "tstereo.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile and link in-tree.  Serialised across processes with a file lock (every rank of a torchrun job calls
    `_lib.load()` at once), and the library is linked to a temporary name and renamed into place, so no process can ever
    dlopen a half-written libtstereo.so."""
    import fcntl
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "tstereo.h")]
    jobs = []
    objs = []
    sid = source_id()
    idfile = os.path.join(OBJ, "build_id.txt")
    if not os.path.exists(idfile) or open(idfile).read() != sid:
        force = True                                # any source changed: rebuild everything (cheap, parallel)
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + [f'-DTSTEREO_BUILD_ID="{sid}"'] + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    print(log, file=sys.stderr)
    if force or jobs or _stale(LIB, objs):
        tmp = LIB + f".tmp{os.getpid()}"
        run([nvcc, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
        os.replace(tmp, LIB)
    with open(idfile, "w") as f:
        f.write(sid)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
