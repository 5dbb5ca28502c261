"""In-tree native build of the product library (nvcc, sm_100a only) and the scene generator.

    python -m stratusgfx_b200.build [--force]

Outputs (git-ignored, but they travel to the GPU box with the repo snapshot):
    stratusgfx_b200/csrc/libstratusgi.so     CUDA kernels + C-ABI of include/stratusgi.h
    stratusgfx_b200/scene/libsgi_scene.so    host-only procedural input generator
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(CSRC, "libstratusgi.so")
CU_SOURCES = ["sgi_abi.cu", "sgi_band.cu", "k_prepass.cu", "k_direct.cu", "k_ssao.cu", "k_taa.cu", "k_cubemaps.cu", "k_vpl.cu", "k_tiles.cu", "k_gather.cu", "k_denoise.cu", "k_atrous_tma.cu"]
HEADERS = ["sgi_internal.h", "sgi_math.cuh", "sgi_merge.cuh", "sgi_denoise.cuh", os.path.join("..", "..", "include", "stratusgi.h")]

# -fmad=false: the integer outputs of this path hinge on float comparisons, so no mul+add may be
# contracted behind the source's back (see csrc/sgi_math.cuh).  IEEE sqrt/div are nvcc defaults.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("SGI_EXTRA_NVCC_FLAGS", "").split()


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build_native(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(deps):
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objs, todo = [], []
    for s in srcs:
        o = s[:-3] + ".o"
        if force or not os.path.exists(o) or os.path.getmtime(o) < _newest([s] + deps[len(srcs):]):
            todo.append((s, o))
        objs.append(o)

    def compile_one(job):
        s, o = job
        return s, subprocess.run([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)

    # one nvcc per translation unit, side by side (the twelve files take 80 s one after the other)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4) or 1) as pool:
        for s, res in pool.map(compile_one, todo):
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed for " + s)
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


HOST_DIR = os.path.join(_HERE, "host")
HOST_BIN = os.path.join(HOST_DIR, "host_frame")


def build_host(force=False):
    """The C++ host side above the C-ABI (host/stratus_gi_backend.h) and its driver program (host/host_frame.cpp): g++ against
    the two in-tree libraries and the CUDA runtime (device memory only)."""
    src = os.path.join(HOST_DIR, "host_frame.cpp")
    deps = [src, os.path.join(HOST_DIR, "stratus_gi_backend.h"), os.path.join(_HERE, "..", "include", "stratusgi.h"), LIB]
    if not force and os.path.exists(HOST_BIN) and os.path.getmtime(HOST_BIN) >= _newest(deps):
        return HOST_BIN
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(cuda, "include"), src, "-o", HOST_BIN,
                           "-L" + CSRC, "-L" + os.path.join(_HERE, "scene"), "-L" + os.path.join(cuda, "lib64"),
                           "-l:libstratusgi.so", "-l:libsgi_scene.so", "-lcudart",
                           "-Wl,-rpath,$ORIGIN/../csrc", "-Wl,-rpath,$ORIGIN/../scene", "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    return HOST_BIN


def build_all(force=False, verbose=False):
    from .scene import build as build_scene
    build_scene(force)
    lib = build_native(force, verbose)
    build_host(force)
    return lib


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print(LIB)
