"""obvhs_cwbvh_exponent (obvhs_b200/csrc/cwbvh_exponent.h) == glibc exp2f(ceilf(log2f(v))) for EVERY f32 the CWBVH
builder can feed it (reference src/cwbvh/bvh2_to_cwbvh.rs:85-99). CPU only, ~10 s."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exponent_matches_glibc_exhaustively(tmp_path):
    exe = str(tmp_path / "expo")
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-ffp-contract=off", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "exponent_exhaustive.cpp"), "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("OK ")
    assert int(out.stdout.split()[1]) > 1_700_000_000
