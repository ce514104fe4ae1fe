"""Build the CUDA library in-tree (``symtensor_b200/lib/libsymtensor_b200.so``) with nvcc for sm_100a.

``python -m symtensor_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsymtensor_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "symtensor_b200.h"))
    if not force and not _stale(LIB, deps):
        return LIB
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libsymtensor_b200.so")
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [d for d in deps if not d.endswith(".cu")]
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    jobs = []
    objs = []
    for src in sources():  # one translation unit per file, compiled in parallel; only the stale ones
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc] + compile_flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, proc in jobs:
        out, _ = proc.communicate()
        log.append(out)
        if proc.returncode != 0:
            for _, other in jobs:
                if other.poll() is None:
                    other.kill()
            raise RuntimeError(f"nvcc failed on {src}:\n" + out)
    res = subprocess.run([nvcc, "-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc (link) failed:\n" + res.stdout + res.stderr)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
