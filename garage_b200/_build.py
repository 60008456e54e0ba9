"""In-tree build of libgarage_ec.so (sm_100a only; nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, "libgarage_ec.so")
SOURCES = ["garage_ec.cu"]
DEPS = ["garage_ec.cu", "rs_kernels.cuh", "gf256.h", "gf256_tables.inc"]


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def nvcc_cmd(out=SO, extra=()):
    return [
        nvcc_path(), "-O3", "-std=c++17",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo", "-Xcompiler", "-fPIC,-O3", "-shared",
        "-I", os.path.join(ROOT, "include"),
        *extra,
        "-o", out,
        *[os.path.join(CSRC, s) for s in SOURCES],
    ]


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.join(ROOT, "include", "garage_ec.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile the CUDA extension if sources are newer than the .so (or force)."""
    if not force and not stale():
        return SO
    if not all(os.path.exists(os.path.join(CSRC, d)) for d in DEPS):
        raise RuntimeError("garage_b200/csrc sources missing")
    cmd = nvcc_cmd(extra=("-Xptxas", "-v") if verbose else ())
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return SO


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
