"""Build the C-ABI CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every ``csrc/*.cu`` is compiled to an object file under ``csrc/_build/`` (in parallel, one nvcc per source) and the
objects are linked into ``csrc/libcheckerpose_b200.so``.  ``build_library(force=True)`` recompiles everything from
scratch (what ``__graft_entry__.build()`` does); without ``force`` only stale objects are rebuilt."""
from __future__ import annotations

import concurrent.futures
import glob
import hashlib
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
IN_TREE_LIB = os.path.join(CSRC, "libcheckerpose_b200.so")
# CHECKERPOSE_B200_LIB selects another build of the same library (kernel A/B experiments, scripts/kbench.py);
# bench.py and the tests assert that it is NOT set (they measure the in-tree build)
LIB_PATH = os.environ.get("CHECKERPOSE_B200_LIB") or IN_TREE_LIB
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH_FLAGS + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(os.path.dirname(CSRC)), "include", "checkerpose_b200.h")]


def _obj_dir(defines) -> str:
    tag = "default" if not defines else hashlib.sha1(" ".join(sorted(defines)).encode()).hexdigest()[:10]
    d = os.path.join(CSRC, "_build", tag)
    os.makedirs(d, exist_ok=True)
    return d


def _compile_one(nvcc, src, obj, defines, verbose):
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return res.stderr


def build_library(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """Compile csrc/*.cu into csrc/libcheckerpose_b200.so (or ``out``, with extra -D ``defines``); returns its path."""
    nvcc = os.environ.get("NVCC", "nvcc")
    out = out or IN_TREE_LIB
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max([hdr_t] + [os.path.getmtime(s) for s in sources()]):
        return out      # up to date (the objects under _build/ do not travel to the GPU box; the library does)
    odir = _obj_dir(tuple(defines))
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(odir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))
    logs = []
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            logs = list(ex.map(lambda j: _compile_one(nvcc, j[0], j[1], defines, verbose), jobs))
    if jobs or not os.path.exists(out) or any(os.path.getmtime(o) > os.path.getmtime(out) for o in objs):
        cmd = [nvcc] + ARCH_FLAGS + ["-shared", "-o", out] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print("\n".join(logs))
    return out


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
