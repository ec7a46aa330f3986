"""Build libaudiort_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m audio_raytracer_b200.build [--force] [--verbose]

Flags that matter for correctness:
  -fmad=false              no FMA contraction (the reference is managed C#: every op rounds once);
                           the exact paths use explicit __f*_rn intrinsics as well
  -prec-div/-prec-sqrt     IEEE division and square root (nvcc defaults, stated explicitly)
  -ftz=false               subnormals kept (Unity's f32tof16 relies on a subnormal multiply)
  -Xcompiler -ffp-contract=off  the host-side ProcessAudioDataJob finalisation is not contracted either
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libaudiort_cuda.so")
SOURCES = ["audiort_api.cu", "k0_pack.cu", "k1_trace.cu", "k1_trace_grid.cu", "k1_bounce.cu", "k1_query_fan.cu", "k2_permeation.cu", "k2_permeation_grid.cu", "k2_permeation_binned.cu", "k3_reduce.cu", "k4_fan_build.cu", "k5_grid_build.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-O2", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libaudiort_cuda cannot be built (there is no CPU fallback)")


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "audiort.h"))
    deps.append(os.path.abspath(__file__))
    return deps


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """Compile and link. ``defines``/``out`` build an experimental variant next to the product library."""
    if out is None and os.environ.get("AUDIORT_LIB"):
        return os.environ["AUDIORT_LIB"]          # an explicitly selected prebuilt variant
    if out is None and not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build" if out is None else "build_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    log = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + ARCH + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, cmd, p in procs:
        text, _ = p.communicate()
        log.append("$ " + " ".join(cmd) + "\n" + text)
        if p.returncode != 0:
            failed = True
    with open(os.path.join(objdir, "nvcc.log"), "w") as f:
        f.write("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed, see audio-raytracer_b200/build/nvcc.log")
    target = out or LIB
    link = [nvcc] + ARCH + ["-shared", "-o", target] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(link)
    if verbose:
        sys.stdout.write("\n".join(log))
    return target


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
