"""The C-ABI library loads and exports every symbol include/obvhs_cuda.h declares (no compute: no GPU needed)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from obvhs_b200 import build

    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "obvhs_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(obvhs_cuda_\w+)\s*\(", text)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/obvhs_cuda.h but not exported"


def test_python_bindings_cover_header(lib_path):
    from obvhs_b200 import api

    assert sorted(api.SIGNATURES) == declared_symbols()
    api.load_library()


def declared_arity():
    """name -> number of parameters, from the header's prototypes."""
    text = open(os.path.join(ROOT, "include", "obvhs_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(obvhs_cuda_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = " ".join(args.split())
        out[name] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_python_bindings_have_the_header_arity():
    # a ctypes signature with a missing or extra argument corrupts the call silently: compare every binding with its prototype
    from obvhs_b200 import api

    arity = declared_arity()
    assert set(arity) == set(api.SIGNATURES)
    wrong = {n: (len(sig[1]), arity[n]) for n, sig in api.SIGNATURES.items() if len(sig[1]) != arity[n]}
    assert wrong == {}


def test_presets_match_reference_table(lib_path):
    # src/lib.rs:233-305
    from obvhs_b200.api import BvhBuildParams, PlocSearchDistance, SortPrecision

    f = BvhBuildParams.fast_build()
    assert (f.pre_split, f.ploc_search_distance, f.search_depth_threshold, f.sort_precision, f.max_prims_per_leaf) == (
        False, PlocSearchDistance.Low, 2, SortPrecision.U64, 8)
    assert abs(f.reinsertion_batch_ratio - 0.02) < 1e-9
    m = BvhBuildParams.medium_build()
    assert (m.ploc_search_distance, m.search_depth_threshold) == (PlocSearchDistance.Medium, 3)
    assert BvhBuildParams.fastest_build().reinsertion_batch_ratio == 0.0
    assert BvhBuildParams.slow_build().pre_split and BvhBuildParams.very_slow_build().sort_precision == SortPrecision.U128


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "obvhs_cuda.h"\n_Static_assert(sizeof(ObvhsCwBvhNode)==80,"");_Static_assert(sizeof(ObvhsRay)==64,"");\n'
                   '_Static_assert(sizeof(ObvhsBvh2Node)==48,"");_Static_assert(sizeof(ObvhsAabb)==32,"");\n'
                   '_Static_assert(sizeof(ObvhsTriangle)==48,"");_Static_assert(sizeof(ObvhsRayHit)==16,"");\n'
                   '_Static_assert(sizeof(ObvhsRayNew)==32,"");_Static_assert(sizeof(ObvhsRayOd)==24,"");_Static_assert(sizeof(ObvhsRayHit8)==8,"");\n'
                   'int main(void){return 0;}\n')
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_no_cpu_fallback_without_device(lib_path):
    """On a box without a GPU the product path must fail loudly, never compute on the CPU."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from obvhs_b200 import api\n"
            "try:\n    api.Context(0)\nexcept api.ObvhsError as e:\n    print('RAISED', e.code)\nelse:\n    print('CREATED')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300).stdout
    import torch

    if torch.cuda.is_available():
        assert "CREATED" in out
    else:
        assert "RAISED -2" in out, out
