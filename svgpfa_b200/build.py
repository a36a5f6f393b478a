"""In-tree build of ``libsvgpfa_b200.so`` (hand-written CUDA for sm_100a, no torch dependency).

    python -m svgpfa_b200.build            # incremental
    python -m svgpfa_b200.build --force

nvcc cross-compiles without a GPU; the built ``.so`` is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libsvgpfa_b200.so")
SOURCES = ["api.cu", "indpoints.cu", "quad.cu", "quad_mma.cu", "quad_mma64.cu", "spike.cu", "panel.cu", "lbfgs.cu"]
# measurement probes and test hooks: a separate library, never loaded by the product path
PROBES_LIB = os.path.join(PKG, "libsvgpfa_b200_probes.so")
PROBES_SOURCES = ["probes.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v",
              "--expt-relaxed-constexpr", "-I", INCLUDE, "-I", CSRC]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "svgpfa_b200.h"))
    headers.append(os.path.join(INCLUDE, "svgpfa_b200_probes.h"))
    headers.append(os.path.abspath(__file__))

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            with open(obj + ".log", "w") as f:
                f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
            if verbose:
                print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES) + len(PROBES_SOURCES)) as ex:
        all_objs = list(ex.map(compile_one, SOURCES + PROBES_SOURCES))
    for lib, objs in ((LIB, all_objs[:len(SOURCES)]), (PROBES_LIB, all_objs[len(SOURCES):])):
        if force or _stale(lib, objs):
            cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                          "-cudart", "static", "-Xcompiler", "-fPIC"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
