"""Builds libobvhs_cuda.so (sm_100a) in-tree with nvcc. No torch, no JIT cache: the .so travels with the repo snapshot.

    python -m obvhs_b200.build [--force] [--verbose]

Flags that matter for parity with the reference's CPU arithmetic: -fmad=false (rustc never contracts
a*b+c), IEEE division/sqrt, no flush-to-zero (all nvcc defaults, stated explicitly).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libobvhs_cuda.so")
SOURCES = ["api.cu", "ploc.cu", "sort.cu", "bvh2.cu", "collapse.cu", "splits.cu", "reinsertion.cu", "cwbvh_build.cu", "traverse.cu", "query.cu", "comm.cu"]
HEADERS = ["common.cuh", "compact.cuh", "sort_tile.cuh", "cwbvh_exponent.h", os.path.join("..", "..", "include", "obvhs_cuda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name: str, defines: list[str]) -> str:
    """Tuning only: the library compiled with extra -D flags into obvhs_b200/lib_variants/<name>/libobvhs_cuda.so (git-ignored;
    travels to the GPU box). Select it with OBVHS_LIB_PATH (read by api.load_library)."""
    out_dir = os.path.join(_HERE, "lib_variants", name)
    obj_dir = os.path.join(out_dir, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append(subprocess.Popen([nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]))
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError("nvcc failed")
    lib = os.path.join(out_dir, "libobvhs_cuda.so")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", lib, *objs, "-Xlinker",
                           "--exclude-libs,ALL", "-ldl"])
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(_HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", LIB_PATH, *objs,
            "-Xlinker", "--exclude-libs,ALL", "-ldl"]
    subprocess.check_call(link)
    # a static archive of the same objects: what a Rust build.rs would link (north_star)
    ar = os.path.join(LIB_DIR, "libobvhs_cuda.a")
    if os.path.exists(ar):
        os.remove(ar)
    subprocess.check_call(["ar", "rcs", ar, *objs])
    return LIB_PATH


HOST_TEST_SRC = os.path.join(_HERE, "..", "tests", "cpp", "host_api.cpp")
HOST_TEST_BIN = os.path.join(LIB_DIR, "host_api_test.bin")
HOST_TEST_BIN_STATIC = os.path.join(LIB_DIR, "host_api_test_static.bin")  # linked against libobvhs_cuda.a, as a Rust build.rs links it


def build_host_test(force: bool = False) -> str:
    """The C++ host side (include/obvhs.hpp) compiled with plain g++ against the shared library: tests/cpp/host_api.cpp ->
    obvhs_b200/lib/host_api_test.bin (run by tests/test_gpu_cpp_host.py on the GPU box; it travels with the snapshot)."""
    build()
    deps = [HOST_TEST_SRC, os.path.join(_HERE, "..", "include", "obvhs.hpp"), os.path.join(_HERE, "..", "include", "obvhs_cuda.h"), LIB_PATH]
    if not force and os.path.exists(HOST_TEST_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_TEST_BIN) for d in deps):
        return HOST_TEST_BIN
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", HOST_TEST_SRC, "-o", HOST_TEST_BIN, "-L", LIB_DIR, "-lobvhs_cuda",
                           "-Wl,-rpath,$ORIGIN"])
    # the same program linked STATICALLY against libobvhs_cuda.a + libcudart_static, exactly the link line of
    # rust/obvhs-cuda-sys/build.rs (static=obvhs_cuda, static=cudart_static, stdc++, dl, rt, pthread)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "lib64")
    subprocess.check_call(["g++", "-std=c++17", "-O2", HOST_TEST_SRC, "-o", HOST_TEST_BIN_STATIC, os.path.join(LIB_DIR, "libobvhs_cuda.a"),
                           "-L", cuda_lib, "-lcudart_static", "-ldl", "-lrt", "-lpthread"])
    return HOST_TEST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_host_test(force="--force" in sys.argv))
