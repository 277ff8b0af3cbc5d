"""Build recipes (explicit compiler invocations, no build system).

  build_native()  nvcc, sm_100a only -> vbz_compression_b200/libvbz_b200.so   (the product)
  build_cpusim()  g++, -DVBZ_CPUSIM  -> tests/cpusim/libvbz_b200_cpusim.so    (TEST-ONLY emulator build:
                  the same kernel and host sources run on a fiber-based CUDA emulator so the
                  CPU test-suite can exercise them in the GPU-less build container; never shipped,
                  never loaded by the package)
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvbz_b200.so")
SIM_DIR = os.path.join(ROOT, "tests", "cpusim")
SIM_LIB = os.path.join(SIM_DIR, "libvbz_b200_cpusim.so")

CU_SOURCES = ["vbzk_launch.cu"]
HOST_SOURCES = ["host/batch.cpp", "host/c_api.cpp", "host/plugin.cpp"]
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _sources():
    deps = []
    for d, _, fs in os.walk(CSRC):
        deps += [os.path.join(d, f) for f in fs if f.endswith((".cu", ".cuh", ".cpp", ".h"))]
    deps += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    return deps


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libvbz_b200.so cannot be built (there is no CPU fallback)")
    return p


def build_native(force=False, verbose=False):
    deps = _sources() + [os.path.abspath(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    objs = []
    for src in CU_SOURCES + HOST_SOURCES:
        obj = os.path.join(objdir, src.replace("/", "_") + ".o")
        cmd = [nvcc] + NVCC_ARCH + common + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        if src.endswith(".cpp"):
            cmd.insert(1, "-x")
            cmd.insert(2, "cu")
        subprocess.run(cmd, check=True)
        objs.append(obj)
    subprocess.run([nvcc] + NVCC_ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"], check=True)
    return LIB


def build_cpusim(force=False):
    deps = _sources() + [os.path.join(SIM_DIR, f) for f in ("cpusim.h", "cpusim.cpp", "cuda_runtime_sim.h")] + [os.path.abspath(__file__)]
    if not force and not _stale(SIM_LIB, deps):
        return SIM_LIB
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES + HOST_SOURCES] + [os.path.join(SIM_DIR, "cpusim.cpp")]
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-DVBZ_CPUSIM", "-fvisibility=hidden", "-Wall", "-Wno-unused-function",
           "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", SIM_DIR, "-o", SIM_LIB]
    for s in srcs:
        cmd += ["-x", "c++", s]
    subprocess.run(cmd, check=True)
    return SIM_LIB


if __name__ == "__main__":
    what = sys.argv[1:] or ["native"]
    if "native" in what:
        print(build_native(force="--force" in what, verbose="-v" in what))
    if "cpusim" in what:
        print(build_cpusim(force="--force" in what))
