"""In-tree build of libyagi_b200.so with nvcc for sm_100a (no JIT cache, no torch types)."""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libyagi_b200.so")
# A/B builds (tools/build_variant.sh) live next to it; YG_LIB=<name> loads lib/libyagi_b200_<name>.so instead
if os.environ.get("YG_LIB"):
    SO = os.path.join(LIBDIR, "libyagi_b200_%s.so" % os.environ["YG_LIB"])

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas=-v",
] + os.environ.get("YG_NVCC_EXTRA", "").split()      # e.g. -DYG_LARGE_ACQUIRE_FENCE=0 for an A/B build


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libyagi_b200.so cannot be built (there is no CPU fallback)")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime() -> float:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "yagi_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src: str, force: bool, hdr_mtime: float, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_mtime):
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, os.path.basename(src)[:-3] + ".ptxas.log")
    with open(log, "w") as f:
        f.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout + r.stderr))
    if verbose:
        print("compiled", os.path.basename(src))
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sources()
    hdr_mtime = _deps_mtime()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, hdr_mtime, verbose), srcs))
    if force or not os.path.exists(SO) or any(os.path.getmtime(o) > os.path.getmtime(SO) for o in objs):
        cmd = [_nvcc(), "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        if verbose:
            print("linked", SO)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
