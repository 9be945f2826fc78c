"""In-tree build of the C-ABI library: csrc/*.cu -> _C/libb200sr.so (sm_100a only).

    python -m mobilesuperresolution_b200.build [--force]

nvcc is driven directly (no torch headers: the ABI is plain C).  Translation units compile in
parallel; objects are cached under _C/obj and rebuilt when a source or header is newer.
The CUDA runtime is linked SHARED so the library shares libcudart.so.12 -- and with it the
current device and stream handles -- with the PyTorch process that loads it.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB = os.path.join(OUT_DIR, "libb200sr.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "--expt-relaxed-constexpr", "-ccbin", "/usr/bin/g++"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "b200sr.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, force: bool, hdr_mtime: float, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    s = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(s), hdr_mtime):
        return obj
    cmd = [NVCC, *ARCH, *FLAGS, *os.environ.get("B200SR_NVCC_EXTRA", "").split(), "-c", s, "-o", obj]   # developer -D switches (timing experiments)
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        with open(obj + ".ptxas.log", "w") as fh:
            fh.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    hdr = _headers_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, hdr, verbose), srcs))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-cudart", "shared", "-ccbin", "/usr/bin/g++", "-o", LIB, *objs,
               "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
