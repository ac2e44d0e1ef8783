"""python scripts/gen_rust_sys.py -- regenerates the `extern "C"` block of rust/obvhs-cuda-sys/src/lib.rs from include/obvhs_cuda.h
(everything above the block -- constants, POD structs, size assertions -- is kept as it is)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
h = open(os.path.join(ROOT, "include", "obvhs_cuda.h")).read()
body = re.sub(r"/\*.*?\*/", "", h[h.index("typedef struct ObvhsContext ObvhsContext;"):], flags=re.S)
decls = re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(obvhs_cuda_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", body, flags=re.S)
TY = {"int": "c_int", "void": "c_void", "size_t": "usize", "uint32_t": "u32", "uint64_t": "u64", "uint8_t": "u8", "float": "f32", "double": "f64",
      "char": "c_char", "ObvhsAabb": "Aabb", "ObvhsTriangle": "Triangle", "ObvhsBvh2Node": "Bvh2Node", "ObvhsCwBvhNode": "CwBvhNode", "ObvhsRay": "Ray",
      "ObvhsRayNew": "RayNew", "ObvhsRayOd": "RayOd", "ObvhsRayHit": "RayHit", "ObvhsRayHit8": "RayHit8", "ObvhsBuildParams": "BuildParams", "ObvhsContext": "Context", "ObvhsBvh2": "Bvh2",
      "ObvhsCwBvh": "CwBvh"}


def conv(t):
    t = t.strip()
    const = "const" in t.split()
    t = t.replace("const", "").strip()
    stars = t.count("*")
    r = TY[t.replace("*", "").strip()]
    for i in range(stars):
        r = ("*const " if (const and i == 0) else "*mut ") + r
    return r


out = []
for ret, name, args in decls:
    params = []
    args = " ".join(args.split())
    if args not in ("void", ""):
        for a in args.split(","):
            m = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)(\[[A-Z_a-z0-9]*\])?$", a.strip())
            ty, nm, arr = m.group(1), m.group(2), m.group(3)
            if arr:
                ty += "*"
            if nm in ("type", "ref", "box", "in"):
                nm += "_"
            params.append(f"{nm}: {conv(ty)}")
    r = "" if ret.strip() == "void" else f" -> {conv(ret)}"
    line = f"    pub fn {name}({', '.join(params)}){r};"
    if len(line) > 130:
        line = f"    pub fn {name}(\n        " + ",\n        ".join(params) + f",\n    ){r};"
    out.append(line)
path = os.path.join(ROOT, "rust", "obvhs-cuda-sys", "src", "lib.rs")
src = open(path).read()
head = src[: src.index('extern "C" {')]
open(path, "w").write(head + 'extern "C" {\n' + "\n".join(out) + "\n}\n")
print(len(decls), "declarations")
