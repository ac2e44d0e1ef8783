import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def kitchen_tris():
    from obvhs_b200 import test_util as tu

    return tu.kitchen()


@pytest.fixture(scope="session")
def scenes(kitchen_tris):
    """Small/medium deterministic scenes used by both the oracle and the GPU parity tests."""
    from obvhs_b200 import test_util as tu

    return {
        "cornell": tu.cornell_box(),
        "ico_plane": np.concatenate([tu.icosphere(1), tu.plane()], axis=0),
        "flat4": tu.flat_plane(4),
        "terrain32": tu.demoscene(32, 0),
        "soup4k": tu.triangle_soup(4096, 3),
        "kitchen": kitchen_tris,
        "splitty": tu.soup_with_large_triangles(3000, 40, 5),
    }
