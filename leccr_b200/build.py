"""In-tree build of the C-ABI library (nvcc, sm_100a only).

`python -m leccr_b200.build` compiles leccr_b200/csrc/api.cu into leccr_b200/_lib/libleccr_b200.so.
nvcc cross-compiles without a GPU, so this also runs on the CPU-only build box.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "api.cu")
LIB_DIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIB_DIR, "libleccr_b200.so")
DEPS = [
    os.path.join(PKG, "csrc", f)
    for f in ("api.cu", "ptx.cuh", "gemm_sm100.cuh", "epilogues.cuh", "kernels.cuh")
] + [os.path.join(ROOT, "include", "leccr_b200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; leccr_b200 needs the CUDA toolkit to build its sm_100a library")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources. Returns the .so path."""
    if not force and not is_stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [
        nvcc_path(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-lineinfo", "-std=c++17",
        "-shared", "-Xcompiler", "-fPIC",
        "-o", LIB, SRC, "-ldl",
    ]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
