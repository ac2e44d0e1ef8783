"""The C++ host side above the C ABI (include/obvhs.hpp, the compiled-language mirror of the reference's interface): CPU tests
check that it compiles against the shared library and fails loudly without a GPU; the GPU test runs tests/cpp/host_api.cpp, which
checks builds / traversals / queries / rebuilds against a brute-force intersection and against each other."""
import os
import subprocess

import pytest

from obvhs_b200 import build as b


def test_cpp_host_side_compiles_and_links():
    path = b.build_host_test()
    assert os.path.exists(path) and os.access(path, os.X_OK)
    # and statically, against libobvhs_cuda.a + libcudart_static: the link line of rust/obvhs-cuda-sys/build.rs
    assert os.path.exists(b.HOST_TEST_BIN_STATIC) and os.access(b.HOST_TEST_BIN_STATIC, os.X_OK)
    ldd = subprocess.run(["ldd", b.HOST_TEST_BIN_STATIC], capture_output=True, text=True).stdout
    assert "libobvhs_cuda" not in ldd and "libcudart" not in ldd


def test_cpp_host_side_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    p = subprocess.run([b.build_host_test()], capture_output=True, text=True, timeout=120)
    assert p.returncode == 3, (p.returncode, p.stdout, p.stderr)
    assert "obvhs::Error" in p.stdout and "no CPU fallback" in p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("linkage", ["shared", "static"])
def test_cpp_host_program_on_gpu(linkage):
    # the binaries built by __graft_entry__.build() travel with the snapshot; only missing ones are built here.
    # static: the same program linked against libobvhs_cuda.a + libcudart_static, as a Rust build.rs links the library
    want = b.HOST_TEST_BIN if linkage == "shared" else b.HOST_TEST_BIN_STATIC
    if not os.path.exists(want):
        b.build_host_test(force=True)
    p = subprocess.run([want], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.returncode, p.stdout[-3000:], p.stderr[-2000:])
    assert "-> ok" in p.stdout
