"""Build every native artefact in-tree (no JIT cache): the CUDA library (sm_100a), the host mesh
helpers, and the test-infrastructure oracle. `python -m gproshan_b200.build [--force]`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    # update_step uses explicitly rounded intrinsics; keep IEEE div/sqrt and denormals everywhere else too
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, cwd=None):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, cwd=cwd)


def _cc():
    return "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def build_cuda(force=False, verbose=False):
    out = os.path.join(HERE, "libptp_b200.so")
    srcs = [os.path.join(CSRC, "ptp_api.cu"), os.path.join(CSRC, "ptp_device.cuh"), os.path.join(ROOT, "include", "ptp_b200.h")]
    if force or _newer(out, srcs):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, srcs[0]]
        if os.path.exists("/usr/bin/g++"):
            cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
        _run(cmd)
    return out


def build_meshgen(force=False):
    out = os.path.join(HERE, "libptp_meshgen.so")
    src = os.path.join(CSRC, "meshgen.c")
    if force or _newer(out, [src]):
        _run([_cc(), "-O3", "-fopenmp", "-fPIC", "-shared", "-o", out, src, "-lm"])
    return out


def build_oracle(force=False):
    """Test infrastructure: the C restatement always, the reference CPU build when /root/reference exists."""
    odir = os.path.join(ROOT, "oracle")
    _run(["make", "-C", odir] + (["-B"] if force else []))


def build_shim(force=False):
    """The C++ drop-in (reference signatures) linked against the reference's own CPU sources, for the tests.
    Only possible where /root/reference exists; the built .so travels to the GPU box."""
    sdir = os.path.join(HERE, "shim")
    _run(["make", "-C", sdir] + (["-B"] if force else []))


def build_all(force=False, verbose=False):
    build_meshgen(force)
    build_cuda(force, verbose)
    build_oracle(force)
    build_shim(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
