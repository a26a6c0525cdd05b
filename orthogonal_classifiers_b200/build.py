"""In-tree build of the C-ABI CUDA library (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "liboc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
    "-shared", "-Xcompiler", "-fPIC",
]


def _sources():
    out = [os.path.join(ROOT, "include", "oc_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


FLAGS_FILE = os.path.join(PKG, "liboc_b200.flags")   # the flags the library next to it was built with (travels with the .so)


def _flags_key() -> str:
    return " ".join(NVCC_FLAGS + os.environ.get("OC_NVCC_EXTRA", "").split())


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    # an experimental build (OC_NVCC_EXTRA=-D...) must not stay in place as the shipped library
    try:
        with open(FLAGS_FILE) as f:
            if f.read().strip() != _flags_key():
                return True
    except OSError:
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/oc_b200.cu -> liboc_b200.so next to the package (in-tree, so
    it travels to the GPU box).  nvcc cross-compiles without a GPU."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + [
        "-I", os.path.join(ROOT, "include"), "-I", CSRC,
        "-o", LIB, os.path.join(CSRC, "oc_b200.cu"),
    ]
    cmd += os.environ.get("OC_NVCC_EXTRA", "").split()  # experiments only (e.g. -DOC_HW_MINB=5)
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(FLAGS_FILE, "w") as f:
        f.write(_flags_key() + "\n")
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
