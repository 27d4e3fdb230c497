"""In-tree build of libaukit_cuda.so (hand-written CUDA for sm_100a, nvcc only).

    python -m aukit_b200.build [--force] [--verbose]

The shared library is a plain C-ABI object (include/aukit_cuda.h): it links only the CUDA
runtime, not torch.  It is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libaukit_cuda.so")
LUA_LIB = os.path.join(LIBDIR, "aukit_cuda.so")  # the Lua C module (luaopen_aukit_cuda)

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "aukit_cuda.h"))
    return hs


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr, flush=True)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = _headers()
    objs, jobs = [], []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            extra = ["-Xptxas", "-v"] if verbose else []
            jobs.append([NVCC, *ARCH, *CUFLAGS, *extra, "-c", s, "-o", o])
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda c: _run(c, verbose), jobs))
    if force or jobs or _stale(LIB, objs):
        _run([NVCC, *ARCH, "-shared", "-o", LIB, *objs], verbose)
    lua_src = os.path.join(CSRC, "lua_binding.c")
    if os.path.exists(lua_src) and (force or _stale(LUA_LIB, [lua_src] + hdrs)):
        # Lua C module: Lua API symbols stay undefined and resolve in the host interpreter
        _run(["gcc", "-O2", "-fPIC", "-shared", "-o", LUA_LIB, lua_src, "-L" + LIBDIR, "-laukit_cuda",
              "-Wl,-rpath,$ORIGIN"], verbose)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
