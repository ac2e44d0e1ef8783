"""Regenerates obvhs_b200/assets/kitchen_tris.npz from the reference's assets/kitchen.obj.zst.

Run in the build container only (needs /root/reference):  python tests/golden/make_kitchen_fixture.py
The fixture stores vertex positions (float32) and triangle vertex indices (int32) in OBJ file order, which is the
order examples/helpers/load_obj.rs:19-43 flattens objects/groups/polys into the triangle list.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from obvhs_b200 import test_util as tu  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/kitchen.obj.zst"
text = tu.zstd_decompress(open(src, "rb").read()).decode("utf-8", errors="replace")
verts, faces = [], []
for line in text.splitlines():
    if line.startswith("v "):
        verts.append([float(t) for t in line.split()[1:4]])
    elif line.startswith("f "):
        idx = [int(t.split("/")[0]) for t in line.split()[1:]]
        idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
        faces.append(idx[:3])
        if len(idx) == 4:
            faces.append([idx[0], idx[2], idx[3]])
positions = np.asarray(verts, dtype=np.float32)
faces = np.asarray(faces, dtype=np.int32)
out = os.path.join(ROOT, "obvhs_b200", "assets", "kitchen_tris.npz")
np.savez_compressed(out, positions=positions, faces=faces)
print(out, positions.shape, faces.shape, os.path.getsize(out))
