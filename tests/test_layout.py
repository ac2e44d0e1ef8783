"""Repository invariants the judge checks: the product never touches the oracle, and nothing GPU-side reads /root/reference."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _files(sub, exts):
    for d, _, fs in os.walk(os.path.join(ROOT, sub)):
        if "__pycache__" in d or os.sep + "build" in d or os.sep + "lib" in d:
            continue
        for f in fs:
            if f.endswith(exts):
                yield os.path.join(d, f)


def test_product_never_imports_or_links_the_oracle():
    for path in _files("obvhs_b200", (".py", ".cu", ".cuh", ".h")):
        text = open(path, errors="replace").read()
        assert not re.search(r"oracle_bind|obvhs_oracle|libobvhs_oracle|import oracle|from oracle", text), path


def test_nothing_that_runs_on_the_gpu_box_reads_the_reference_checkout():
    for path in list(_files("obvhs_b200", (".py",))) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        assert "/root/reference" not in open(path).read(), path
    for path in _files("tests", (".py",)):
        if os.path.basename(path) in ("make_kitchen_fixture.py", "test_layout.py"):
            continue  # the fixture generator runs in the build container only
        assert "/root/reference" not in open(path).read(), path
