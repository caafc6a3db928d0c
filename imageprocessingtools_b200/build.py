"""Builds the native parts IN-TREE (the .so files travel to the GPU box with the snapshot):

    libppmx_gpu.so   CUDA kernels + the C ABI of include/ppmx_gpu.h   (nvcc, sm_100a only)
    libppmx_host.so  the C host layer of include/ppmx_host.h          (gcc, links the above)
    ppmx-b200        the command line                                  (gcc)

No GPU is needed to build: nvcc cross-compiles for sm_100a.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
GPU_SO = os.path.join(PKG, "libppmx_gpu.so")
TUNING_SO = os.path.join(PKG, "libppmx_gpu_tuning.so")
HOST_SO = os.path.join(PKG, "libppmx_host.so")
CLI = os.path.join(PKG, "ppmx-b200")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false",  # the FP64 operators must round every product and sum separately
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]
# no -march=native / -mfma: contribution tables must match the reference's glibc doubles
GCC_FLAGS = ["-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-std=gnu99"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build_all(force: bool = False, verbose: bool = False) -> None:
    hdrs = [os.path.join(PKG, "..", "include", "ppmx_gpu.h"), os.path.join(PKG, "..", "include", "ppmx_host.h"),
            os.path.join(CSRC, "ppmx_kernels.h"), os.path.join(CSRC, "ppmx_common.cuh"), os.path.join(CSRC, "ppmx_ctx.h"), os.path.join(CSRC, "ppmx_conv.cuh")]
    cu = [os.path.join(CSRC, f) for f in ("ppmx_color.cu", "ppmx_geometry.cu", "ppmx_bicubic.cu", "ppmx_conv.cu", "ppmx_conv_sep.cu", "ppmx_conv_ua.cu", "ppmx_conv_vw.cu",
                                          "ppmx_fused.cu", "ppmx_gpu.cu", "ppmx_chain.cu")]
    cu = [f for f in cu if os.path.exists(f)]
    # the release library, and the same sources with -DPPMX_TUNING (alternative kernel variants behind
    # ppmx_gpu_set_tuning("variant", n)) for tools/sweep.py and the variant tests
    todo = [(so, flags) for so, flags in ((GPU_SO, []), (TUNING_SO, ["-DPPMX_TUNING"]))
            if force or _newer(so, cu + hdrs)]
    if todo:
        # one nvcc per translation unit and flavour, in parallel, then one link step each
        objdir = os.path.join(PKG, "build")
        os.makedirs(objdir, exist_ok=True)
        compile_flags = [f for f in NVCC_FLAGS if f not in ("-shared", "-cudart", "static")]
        procs = []
        for so, extra in todo:
            tag = "tuning." if extra else ""
            for src in cu:
                obj = os.path.join(objdir, tag + os.path.basename(src) + ".o")
                if not force and not _newer(obj, [src] + hdrs):
                    procs.append((None, obj, so))  # this object is still current
                    continue
                cmd = [_nvcc()] + compile_flags + extra + ["-c", "-o", obj, src]
                if verbose:
                    print(" ".join(cmd), flush=True)
                procs.append((subprocess.Popen(cmd), obj, so))
        objs = {}
        failed = []
        for proc, obj, so in procs:
            if proc is not None and proc.wait() != 0:
                failed.append(obj)
                if os.path.exists(obj):
                    os.remove(obj)
            objs.setdefault(so, []).append(obj)
        if failed:
            raise RuntimeError("nvcc failed for " + ", ".join(failed))
        for so, _ in todo:
            _run([_nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "-Bsymbolic", "-o", so] + objs[so],
                 verbose)
    host_c = os.path.join(CSRC, "ppmx_host.c")
    if force or _newer(HOST_SO, [host_c, GPU_SO] + hdrs):
        _run(["gcc"] + GCC_FLAGS + ["-shared", "-o", HOST_SO, host_c, "-L" + PKG, "-lppmx_gpu", "-lm", "-lpthread",
                                    "-Wl,-rpath,$ORIGIN"], verbose)
    cli_c = os.path.join(CSRC, "ppmx_cli.c")
    if force or _newer(CLI, [cli_c, HOST_SO]):
        _run(["gcc"] + GCC_FLAGS + ["-o", CLI, cli_c, "-L" + PKG, "-lppmx_host", "-lppmx_gpu", "-lm", "-lpthread",
                                    "-Wl,-rpath,$ORIGIN"], verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
