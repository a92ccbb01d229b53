"""Builds libnas3d_b200.so (hand-written sm_100a kernels + the C-ABI of include/nas3d_b200.h).

In-tree build with plain nvcc so the .so travels with the repo snapshot; no torch headers are
involved (the library has no torch types in its signatures).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libnas3d_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


VARIANTS = {
    # A/B builds of the same sources (selected at run time with NAS3D_LIB=<path>)
    "noffma2": ["-DNAS3D_NO_FFMA2"],      # scalar FFMA instead of packed FFMA2 (common.cuh)
    "pwminb4": ["-DNAS3D_PW_MINB=4"],     # pointwise kernels: register cap for 4 / 6 CTAs per SM
    "pwminb6": ["-DNAS3D_PW_MINB=6"],
    # fused 1x1 backward: ring depth / CTAs per SM
    "pbs6m3": ["-DNAS3D_PB_S=6", "-DNAS3D_PB_MINB=3"],
    "pbs8m3": ["-DNAS3D_PB_S=8", "-DNAS3D_PB_MINB=3"],
}


def build_library(force=False, verbose=False, variant=None):
    extra = []
    lib_path, obj_dir = LIB_PATH, OBJ_DIR
    if variant:
        extra = VARIANTS[variant]
        lib_path = os.path.join(LIB_DIR, "libnas3d_b200_%s.so" % variant)
        obj_dir = os.path.join(HERE, "build", variant)
    return _build(force, verbose, extra, lib_path, obj_dir)


def _build(force, verbose, extra, LIB_PATH, OBJ_DIR):
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "nas3d_b200.h"))
    jobs = []
    objs = []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for %s" % cmd[-3])
    if jobs or force or _stale(LIB_PATH, objs):
        cmd = [NVCC, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var))
