"""In-tree build of the native libraries (no JIT cache, nothing installed outside the repo).

    python vit-fpga_b200/build.py [--force]

Produces
    vit-fpga_b200/lib/libnetcuda.so        CUDA kernels (sm_100a SASS) + host runtime + the extern "C" ABI
    vit-fpga_b200/lib/libnetcuda_host.so   cuda::net_cuda (the net::net_abstract implementation): what an application links
    vit-fpga_b200/lib/libnetcuda_hostdrv.so  C entry points that drive net_cuda through net::net_abstract* (tests/ and bench.py only)

nvcc cross-compiles for sm_100a without a GPU; the .so files travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
# NETCUDA_BUILD_TAG=<tag>: a second build beside the product one (lib_<tag>/, build_<tag>/), selected at run time with
# NETCUDA_LIB_DIR -- for A/B runs of two kernel versions inside one GPU call and for NETCUDA_DEBUG_TIMELINE builds
_TAG = os.environ.get("NETCUDA_BUILD_TAG", "")
LIB = os.path.join(HERE, "lib" + ("_" + _TAG if _TAG else ""))
OBJ = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
INCLUDE = os.path.join(ROOT, "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("NETCUDA_CXX", "/usr/bin/g++")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-ccbin", CXX]
NVCC_FLAGS += os.environ.get("NETCUDA_EXTRA_NVCC_FLAGS", "").split()  # (A/B builds, together with NETCUDA_BUILD_TAG)
if os.environ.get("NETCUDA_DEBUG_TIMELINE"):  # clock64 timeline hooks for tools/*_timeline.py (never in the shipped build)
    NVCC_FLAGS += ["-DNETCUDA_DEBUG_TIMELINE"]
CU_SOURCES = ["gemm.cu", "elementwise.cu", "attention.cu", "mlp_stream.cu", "mlp_umma_stream.cu", "runtime.cu", "weights_io.cu", "frame_ring.cu", "staging.cpp"]
CU_HEADERS = ["ptx.cuh", "gemm_tcgen05.cuh", "kernels.h"]
HOST_SOURCES = ["net_cuda.cpp"]
HOST_DRIVER_SOURCES = ["host_capi.cpp"]  # ctypes driver for tests/ and bench.py: NOT part of the shipped host library


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + cmd[0])


def build(force: bool = False, verbose: bool = False) -> dict:
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in CU_HEADERS] + [os.path.join(INCLUDE, "netcuda.h"), os.path.abspath(__file__)]
    jobs = []
    objs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + headers):
            jobs.append([NVCC] + NVCC_FLAGS + ["-I", INCLUDE, "-c", s, "-o", o])
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(_run, jobs))
    lib_cuda = os.path.join(LIB, "libnetcuda.so")
    if force or jobs or _newer(lib_cuda, objs):
        _run([NVCC, "-shared", "-o", lib_cuda] + objs + ["-ccbin", CXX, "-cudart", "static", "-Xlinker", "-z,defs", "-lpthread",
                                                        "-ldl", "-lrt"])
    lib_host = os.path.join(LIB, "libnetcuda_host.so")
    host_srcs = [os.path.join(HOST, s) for s in HOST_SOURCES]
    host_deps = host_srcs + [os.path.join(INCLUDE, h) for h in ("netCUDA.h", "netAbstract.h", "defines.h", "netcuda.h")] + [lib_cuda]
    if force or _newer(lib_host, host_deps):
        # gnu++14: the only language level the reference states (.vscode/c_cpp_properties.json:13)
        _run([CXX, "-std=gnu++14", "-O2", "-fPIC", "-Wall", "-shared", "-I", INCLUDE, "-o", lib_host] + host_srcs +
             ["-L", LIB, "-lnetcuda", "-pthread", "-Wl,-rpath,$ORIGIN", "-Wl,-z,defs"])
    lib_drv = os.path.join(LIB, "libnetcuda_hostdrv.so")
    drv_srcs = [os.path.join(HOST, s) for s in HOST_DRIVER_SOURCES]
    if force or _newer(lib_drv, drv_srcs + [os.path.join(INCLUDE, h) for h in ("netCUDA.h", "netAbstract.h", "defines.h")] + [lib_host]):
        _run([CXX, "-std=gnu++14", "-O2", "-fPIC", "-Wall", "-shared", "-I", INCLUDE, "-o", lib_drv] + drv_srcs +
             ["-L", LIB, "-lnetcuda_host", "-lnetcuda", "-pthread", "-Wl,-rpath,$ORIGIN", "-Wl,-z,defs"])
    if verbose:
        print("built", lib_cuda, ",", lib_host, "and", lib_drv)
    return {"libnetcuda": lib_cuda, "libnetcuda_host": lib_host, "libnetcuda_hostdrv": lib_drv}


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
