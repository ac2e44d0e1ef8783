"""The Rust shim (rust/obvhs-cuda-sys/src/lib.rs) is written but cannot be compiled in this image (no rustc/cargo). This keeps it
honest: every function include/obvhs_cuda.h declares must be declared in the `extern "C"` block with the same number of parameters,
and the POD sizes asserted there must be the ones the library asserts. (lib.rs's extern block is generated from the header: `python scripts/gen_rust_sys.py`.)"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_decls():
    h = open(os.path.join(ROOT, "include", "obvhs_cuda.h")).read()
    body = re.sub(r"/\*.*?\*/", "", h[h.index("typedef struct ObvhsContext ObvhsContext;"):], flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(obvhs_cuda_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", body, flags=re.S):
        args = " ".join(args.split())
        out[name] = 0 if args in ("", "void") else len(args.split(","))
    return out


def _rust_decls():
    src = open(os.path.join(ROOT, "rust", "obvhs-cuda-sys", "src", "lib.rs")).read()
    ext = src[src.index('extern "C" {'):]
    out = {}
    for name, args in re.findall(r"pub fn (obvhs_cuda_[a-z0-9_]+)\s*\(([^;]*?)\)\s*(?:->[^;]*)?;", ext, flags=re.S):
        args = " ".join(args.split()).rstrip(", ")
        out[name] = 0 if args == "" else len([a for a in args.split(",") if a.strip()])
    return out, src


def test_every_c_symbol_is_bound_with_the_same_arity():
    c = _c_decls()
    r, _ = _rust_decls()
    assert len(c) >= 81
    assert set(c) == set(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    assert {k: v for k, v in c.items() if r[k] != v} == {}


def test_pod_sizes_and_files_present():
    _, src = _rust_decls()
    for ty, size in (("Aabb", 32), ("Triangle", 48), ("Bvh2Node", 48), ("CwBvhNode", 80), ("Ray", 64), ("RayNew", 32), ("RayOd", 24), ("RayHit", 16)):
        assert f"size_of::<{ty}>() == {size}" in src
    for f in ("rust/obvhs-cuda-sys/Cargo.toml", "rust/obvhs-cuda-sys/build.rs", "rust/obvhs-cuda/Cargo.toml", "rust/obvhs-cuda/src/lib.rs"):
        assert os.path.exists(os.path.join(ROOT, f)), f
    build_rs = open(os.path.join(ROOT, "rust", "obvhs-cuda-sys", "build.rs")).read()
    from obvhs_b200 import build as b

    for s in b.SOURCES:  # build.rs compiles the same translation units as obvhs_b200/build.py
        assert f'"{s}"' in build_rs, s
    assert "compute_100a,code=sm_100a" in build_rs and "-fmad=false" in build_rs
