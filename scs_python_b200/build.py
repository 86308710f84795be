"""Build libscsb200.so in-tree with nvcc for sm_100a (no JIT cache, travels with gpurun).

    python -m scs_python_b200.build [--force]

Every .cu under csrc/ is compiled with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
and linked into scs_python_b200/libscsb200.so (the C ABI declared in include/scs_b200.h).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libscsb200.so")
SOURCES = ["sparse.cu", "tiled.cu", "linsys.cu", "cones.cu", "aa.cu", "dist.cu", "scs_solver.cu", "batch.cu", "rw.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=default", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 backend cannot be built (no CPU fallback exists)")


def _newer(src: str, dst: str) -> bool:
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "scs_b200.h")]
    return any(_newer(d, LIB) for d in deps)


def build_alt(cfg: int = 0) -> str:
    """Development build of the alternative tile geometry of the tiled SpMV engine (csrc/tiled.cuh,
    -DB200_TILED_CFG) into libscsb200_cfg<k>.so; loaded with SCS_B200_LIBPATH by tools/spmv_variants.py."""
    nvcc = _nvcc()
    out = os.path.join(HERE, "libscsb200_cfg%d.so" % cfg)
    objdir = os.path.join(HERE, "build", "cfg%d" % cfg)
    os.makedirs(objdir, exist_ok=True)

    def one(src):
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        r = subprocess.run([nvcc, *NVCC_FLAGS, "-DB200_TILED_CFG=%d" % cfg, "-c", os.path.join(CSRC, src), "-o", o],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return o
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker",
                        "-Bsymbolic", "-lcudart", "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return out


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "scs_b200.h"))

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if not force and not _newer(s, o) and not any(_newer(h, o) for h in headers):
            return o, ""
        cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return o, r.stderr

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in res]
    for _, err in res:
        if verbose and err.strip():
            sys.stderr.write(err)
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xlinker", "-Bsymbolic", "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    if "--alt" in sys.argv:
        i = sys.argv.index("--alt")
        print(build_alt(int(sys.argv[i + 1]) if i + 1 < len(sys.argv) and sys.argv[i + 1].isdigit() else 0))
    else:
        build(force="--force" in sys.argv)
