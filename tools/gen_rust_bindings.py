#!/usr/bin/env python
"""Prints `unsafe fn` lines for integration/impact_voxel_cuda/src/lib.rs from the prototypes of include/impact_voxel_cuda.h
(for the entry points lib.rs does not bind yet, or all of them with --all). The output goes inside `define_lib! { ... }`."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hdr = open(os.path.join(ROOT, "include", "impact_voxel_cuda.h")).read()
rs = open(os.path.join(ROOT, "integration", "impact_voxel_cuda", "src", "lib.rs")).read()
hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
SCALARS = {"uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "int": "i32", "float": "f32", "double": "f64", "size_t": "usize",
           "uint8_t": "u8", "uint16_t": "u16", "unsigned char": "u8", "char": "c_char", "void": "c_void"}
STRUCTS = {"ivx_voxel": "Voxel", "ivx_chunk_submesh": "ChunkSubmesh", "ivx_index_materials": "VoxelMeshIndexMaterials",
           "ivx_chunk_desc": "IvxChunkDesc", "ivx_sdf_node": "IvxSdfNode"}


def camel(name):
    return "".join(p.capitalize() for p in name.split("_"))


def rust_type(c):
    c = c.strip()
    const = c.startswith("const ")
    if const:
        c = c[6:].strip()
    stars = c.count("*")
    base = c.replace("*", "").replace("const", "").strip()
    if base in SCALARS:
        t = SCALARS[base]
    elif base in STRUCTS:
        t = STRUCTS[base]
    elif base.startswith("ivx_"):
        t = camel(base)
    else:
        raise ValueError(c)
    for _ in range(stars):
        t = ("*const " if const else "*mut ") + t
        const = False if stars > 1 else const
    return t


out = []
for m in re.finditer(r"^(int|void|uint32_t|uint64_t|const char\*)\s+(ivx_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.M | re.S):
    ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
    if "--all" not in sys.argv and re.search(r"\bfn %s\b" % name, rs):
        continue
    params = []
    for a in [x.strip() for x in args.split(",") if x.strip() and x.strip() != "void"]:
        arr = re.search(r"\[(\d*)\]$", a)
        if arr:
            a = a[: arr.start()].strip()
        mm = re.match(r"(.*?)([A-Za-z_][A-Za-z_0-9]*)$", a)
        ctype, pname = mm.group(1).strip(), mm.group(2)
        if arr:
            ctype += "*"
        if ctype.endswith("* const*"):
            ctype = ctype.replace("* const*", "**")
        params.append(f"{pname}: {rust_type(ctype)}")
    r = {"int": "i32", "void": "()", "uint32_t": "u32", "uint64_t": "u64", "const char*": "*const c_char"}[ret]
    line = f"    unsafe fn {name}({', '.join(params)}) -> {r};"
    out.append(line)
print("\n".join(out))
