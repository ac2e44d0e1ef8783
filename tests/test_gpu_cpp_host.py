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


def test_cpp_host_side_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    p = subprocess.run([b.build_host_test()], capture_output=True, text=True, timeout=120)
    assert p.returncode == 3, (p.returncode, p.stdout, p.stderr)
    assert "obvhs::Error" in p.stdout and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_cpp_host_program_on_gpu():
    # the binary built by __graft_entry__.build() travels with the snapshot; only a missing one is built here
    path = b.HOST_TEST_BIN if os.path.exists(b.HOST_TEST_BIN) else b.build_host_test()
    p = subprocess.run([path], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.returncode, p.stdout[-3000:], p.stderr[-2000:])
    assert "-> ok" in p.stdout
