"""CPU-only checks of the product's host side: the C ABI loads and exports every declared symbol,
refuses to run without a device, and its graph compiler (program.cpp) equals the oracle's, bit for bit."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as H
from impact_b200 import _lib as L
from impact_b200.graph import PROG_NODE_DTYPE, SDFGraph
from impact_b200.voxel import compile_program_host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_declared_in_the_header():
    lib = L.lib()
    header = open(os.path.join(ROOT, "include", "impact_voxel_cuda.h")).read()
    declared = set(re.findall(r"^(?:int|void|const char\*|uint32_t|uint64_t)\s+(ivx_[a-z_]+)\s*\(", header, re.M))
    assert declared, "no declarations parsed"
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert lib.ivx_abi_version() == 1


def test_the_rust_shim_binds_every_entry_point_with_the_header_s_arity():
    # integration/impact_voxel_cuda/src/lib.rs is the reference-side binding a maintainer adds; it is not
    # compiled here (no Rust toolchain), so at least keep it in step with the header: every function, same parameter count
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "impact_voxel_cuda.h")).read(), flags=re.S)
    shim = open(os.path.join(ROOT, "integration", "impact_voxel_cuda", "src", "lib.rs")).read()
    protos = re.findall(r"^(?:int|void|const char\*|uint32_t|uint64_t)\s+(ivx_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", header, re.M | re.S)
    assert len(protos) >= 75
    for name, args in protos:
        m = re.search(r"unsafe fn %s\(([^;]*?)\)\s*->" % name, shim, re.S)
        assert m, f"{name} is not bound in lib.rs"
        n_c = 0 if args.strip() in ("", "void") else len(args.split(","))
        n_rs = 0 if not m.group(1).strip() else len(m.group(1).split(","))
        assert n_c == n_rs, (name, n_c, n_rs)


def test_struct_layouts_match_the_header():
    assert C.sizeof(L.Config) == 24
    assert C.sizeof(L.ProgramInfo) == 32
    assert C.sizeof(L.ObjectInfo) == 104
    assert C.sizeof(L.MeshInfo) == 64
    assert C.sizeof(L.AbsorbStats) == 20
    assert PROG_NODE_DTYPE.itemsize == 144


def test_header_is_plain_c_and_record_sizes_match_the_bindings(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (and as C++), and the record types the bindings mirror
    with numpy dtypes must have the sizes the C compiler gives them."""
    import shutil
    import subprocess

    header = os.path.join(ROOT, "include", "impact_voxel_cuda.h")
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", header])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", header])
    src = tmp_path / "sizes.c"
    src.write_text(
        '#include <stdio.h>\n#include "impact_voxel_cuda.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ivx_voxel), sizeof(ivx_chunk_desc), '
        "sizeof(ivx_chunk_submesh), sizeof(ivx_index_materials), sizeof(ivx_sdf_node), sizeof(ivx_node), "
        "sizeof(ivx_inertial_moments), sizeof(ivx_isometry), sizeof(ivx_surface_voxel), sizeof(ivx_voxel_contact)); return 0; }\n")
    exe = tmp_path / "sizes"
    subprocess.check_call([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    from impact_b200.graph import SDF_NODE_DTYPE as NODE_DTYPE

    assert sizes == [L.VOXEL_DTYPE.itemsize, L.CHUNK_DTYPE.itemsize, L.SUBMESH_DTYPE.itemsize,
                     L.INDEX_MATERIALS_DTYPE.itemsize, NODE_DTYPE.itemsize, PROG_NODE_DTYPE.itemsize, 40, 28,
                     L.SURFACE_VOXEL_DTYPE.itemsize, L.CONTACT_DTYPE.itemsize], sizes


def test_create_fails_loudly_without_a_device():
    import torch

    lib = L.lib()
    cfg = L.Config(lib.ivx_abi_version(), 0, None, 0)
    h = C.c_void_p()
    rc = lib.ivx_create(C.byref(cfg), C.byref(h))
    if torch.cuda.is_available():
        assert rc == 0
        lib.ivx_destroy(h)
    else:
        assert rc == 7 and not h.value  # IVX_ERR_NO_DEVICE: no CPU fallback
    bad = L.Config(999, 0, None, 0)
    assert lib.ivx_create(C.byref(bad), C.byref(h)) == 1


GRAPHS = {
    "sphere": H.sphere_graph, "box": H.box_graph, "union": H.sphere_union_graph, "complex": H.complex_graph,
    "noisy_sphere": H.noisy_sphere_graph, "noisy_box": H.noisy_box_graph, "zoo": H.csg_zoo_graph,
    "asteroid_like": H.asteroid_like_graph,
}


@pytest.mark.parametrize("name", sorted(GRAPHS))
def test_host_graph_compile_equals_oracle_bit_for_bit(oracle, name):
    g = GRAPHS[name]()
    nodes, depth, lo, hi = compile_program_host(g)
    ref = oracle.Generator(g.nodes(), g.root_node_id)
    rn = ref.nodes()
    assert len(nodes) == len(rn) and depth == ref.stack_size
    for f in PROG_NODE_DTYPE.names:
        if f == "_pad":
            continue
        a, b = nodes[f], rn[f]
        if a.dtype.kind == "f":
            assert H.f32_bits_equal(a, b).all(), f
        else:
            assert np.array_equal(a, b), f
    rlo, rhi = ref.domain()
    assert H.f32_bits_equal(lo, rlo).all() and H.f32_bits_equal(hi, rhi).all()


def test_compile_unrolls_shared_subgraphs_child_one_first(oracle):
    g = SDFGraph()
    s = g.sphere(3.0)
    a = g.translation(s, [5.0, 0, 0])
    b = g.translation(s, [-5.0, 0, 0])
    g.subtraction(a, b, 0.5)
    nodes, depth, _, _ = compile_program_host(g)
    assert [int(k) for k in nodes["kind"]] == [0, 3, 0, 3, 8]  # sphere duplicated per parent
    assert depth == 2
    assert nodes["transform"][0][12] == -5.0 and nodes["transform"][2][12] == 5.0
    # root margin = MAX_F32; children of a smooth combine get margin + 2.5 * 0.25 * k * log2(leaves)
    assert nodes["margin"][4] == np.float32(0.02) * np.float32(127)
    assert nodes["margin"][0] == nodes["margin"][4] + np.float32(2.5) * (np.float32(0.25) * np.float32(0.5) * np.float32(1.0))


def test_compile_errors_mirror_the_reference():
    g = SDFGraph()
    g.add_node((7, (0, 0), 0, 0, [0.0]))  # union of itself
    with pytest.raises(ValueError, match="cycle"):
        compile_program_host(g)
    g = SDFGraph()
    g.add_node((3, (5, 0), 0, 0, [0, 0, 0]))  # translation of a missing node
    with pytest.raises(ValueError, match="Missing SDF node 5"):
        compile_program_host(g)
