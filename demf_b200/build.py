"""Builds demf_b200/libdemf_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m demf_b200.build [--force]

No torch involved: the library has plain C entry points (include/demf_b200.h) and links the
CUDA runtime statically, so it works with whatever allocator/stream the host process uses.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(PKG, "libdemf_b200.so")
SOURCES = ["capi.cu", "msda.cu", "fps.cu", "ball_query.cu", "point_ops.cu", "rows.cu", "sa_fused.cu", "ball_grid.cu", "glue.cu", "postprocess.cu", "bn_rows.cu", "gemm_tf32.cu", "loss.cu", "sa_pipe.cu", "mha.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "umma.cuh"), os.path.join(CSRC, "ball_grid.cuh"), os.path.join(PKG, "..", "include", "demf_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libdemf_b200.so. Returns its path."""
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    flags = list(NVCC_FLAGS) + os.environ.get("DEMF_NVCC_EXTRA", "").split()   # e.g. -DDEMF_SAP_PROF (diagnostic builds)
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        if force or _stale(o, [s] + HEADERS):
            jobs.append([nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        for res in ex.map(lambda c: subprocess.run(c, env=env, capture_output=True, text=True), jobs):
            if verbose or res.returncode:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode:
                raise RuntimeError("nvcc failed: " + " ".join(res.args))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs
        res = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if res.returncode:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed: " + " ".join(cmd))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
