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


HASH = LIB + ".srchash"


def source_hash():
    """Content hash of everything the library is compiled from (mtimes do not survive a snapshot copy)."""
    import hashlib

    h = hashlib.sha256()
    for d in DEPS:
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale():
    if not os.path.exists(LIB) or not os.path.exists(HASH):
        return True
    with open(HASH) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    """Compile the library if it is missing or was built from other sources. Returns the .so path.
    Safe under torchrun: one process builds (file lock), into a temporary file that is renamed into place."""
    if not force and not is_stale():
        return LIB
    import fcntl

    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():  # another rank built it while we waited
                return LIB
            tmp = f"{LIB}.{os.getpid()}.tmp"
            cmd = [
                nvcc_path(),
                "-gencode", "arch=compute_100a,code=sm_100a",
                "-O3", "-lineinfo", "-std=c++17",
                "-shared", "-Xcompiler", "-fPIC",
                "-o", tmp, SRC, "-ldl",
            ]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
            digest = source_hash()
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            if verbose:
                sys.stderr.write(res.stderr)
            os.replace(tmp, LIB)
            with open(HASH + ".tmp", "w") as f:
                f.write(digest + "\n")
            os.replace(HASH + ".tmp", HASH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
