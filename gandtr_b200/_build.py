"""Build recipe of libgandtr_b200.so (sm_100a only, in-tree).

`python -m gandtr_b200._build` or `__graft_entry__.build()`. nvcc cross-compiles without a GPU. The
built library is git-ignored but travels to the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
BUILD_DIR = os.path.join(PKG_DIR, "_build")
LIB_PATH = os.path.join(PKG_DIR, "libgandtr_b200.so")

SOURCES = ["lib.cu", "clahe_sm100.cu", "gem_whiten_sm100.cu", "score_exact_sm100.cu", "score_topk_sm100.cu",
           "map_eval_sm100.cu", "resize_sm100.cu", "mining_sm100.cu", "whiten_learn_sm100.cu", "jpeg_nvjpeg.cu"]
HEADERS = ["common.cuh", "select.cuh", "clahe_math.cuh", "tc_common.cuh", "resize_math.h"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
              "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libgandtr_b200.so")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu of gandtr_b200/csrc for sm_100a and link libgandtr_b200.so. Returns its path."""
    nvcc = _nvcc()
    os.makedirs(BUILD_DIR, exist_ok=True)
    common_deps = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(INCLUDE, "gandtr_b200.h"), __file__]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + common_deps):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed.append(src)
    if failed:
        raise RuntimeError("nvcc failed for: %s" % ", ".join(failed))
    if force or procs or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + ["-ldl"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("linking libgandtr_b200.so failed")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
