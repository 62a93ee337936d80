"""Builds the in-tree native artefacts of bammmotif2_b200 with nvcc / g++ (no torch, no JIT cache):

    libbamm_b200.so   the C-ABI library (include/bamm_b200.h): csrc/capi.cu + csrc/kernels.cuh, sm_100a only

Run as `python -m bammmotif2_b200.build` or through __graft_entry__.build().
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libbamm_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # keep the reference's operation order (no FMA contraction) in the model update
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-misleading-indentation",
    "-shared",
]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build_lib(force=False, verbose=False):
    srcs = [os.path.join(HERE, "csrc", "capi.cu")]
    deps = srcs + [os.path.join(HERE, "csrc", "kernels.cuh"), os.path.join(ROOT, "include", "bamm_b200.h"), __file__]
    if not force and _newer(LIB, deps):
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    env = dict(os.environ)
    env.pop("CXX", None); env.pop("CC", None)          # the image exports a gcc wrapper without libgomp specs
    subprocess.check_call(cmd + ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else cmd, env=env)
    return LIB


def build_all(force=False, verbose=False):
    return [build_lib(force, verbose)]


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(p)
