"""Build recipe for libb200denoise.so (nvcc, sm_100a only, in-tree output).

The shared library has no torch / Python dependency: plain CUDA runtime + the C-ABI of
``include/b200denoise.h``.  ``python -m audio_denoising_b200._build`` (or ``__graft_entry__.build()``)
rebuilds it when any source is newer than the binary.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libb200denoise.so")
SOURCES = ["api.cu", "dsp.cu", "griffinlim.cu", "gl_fast.cu", "gl_fast_n512.cu", "gl_fast_n2048.cu", "gl_warp.cu", "gl_reg.cu", "model.cu", "cell.cu", "invmel_tc.cu", "unet_mma.cu", "unet_tc.cu", "stream.cu", "ingest.cu"]


def _newest_source_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(PKG), "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def needs_build() -> bool:
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest_source_mtime()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    for s in SOURCES:
        obj = os.path.join(PKG, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
               "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3", "-c", os.path.join(CSRC, s), "-o", obj]
        cmd[1:1] = os.environ.get("B2D_EXTRA_NVCC_FLAGS", "").split()  # experiment builds only (tools/), empty for the product
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[b2d build] {s} FAILED\n{out}\n")
        elif verbose or "warning" in out:
            sys.stderr.write(f"[b2d build] {s}\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed building libb200denoise.so")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static", "-o", LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
