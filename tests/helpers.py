"""Shared test inputs (the reference's own benchmark graphs) and comparison helpers."""
from __future__ import annotations

import numpy as np

from impact_b200.graph import SDFGraph, VoxelTypeGenerator

# ---- graphs: engine/src/benchmark/benchmarks/generation.rs:24-124 and BASELINE.md §4 ----------


def sphere_graph(radius=31.0):
    """BASELINE config 1: `Sphere(r = 31)` → 64³ grid."""
    g = SDFGraph()
    g.sphere(radius)
    return g


def box_graph(extent=80.0):
    g = SDFGraph()  # generate_box (generation.rs:24-38)
    g.box([extent] * 3)
    return g


def sphere_union_graph(scale=1.0):
    g = SDFGraph()  # generate_sphere_union (generation.rs:40-59)
    s1 = g.sphere(60.0 * scale)
    s2 = g.sphere(60.0 * scale)
    s2 = g.translation(s2, [50.0 * scale, 0.0, 0.0])
    g.union(s1, s2, 1.0)
    return g


def complex_graph(scale=1.0):
    g = SDFGraph()  # generate_complex_object (generation.rs:61-85)
    s = g.sphere(60.0 * scale)
    s = g.translation(s, [50.0 * scale, 0.0, 0.0])
    b = g.box([50.0 * scale, 60.0 * scale, 70.0 * scale])
    b = g.scaling(b, 0.9)
    b = g.rotation_from_axis_angle(b, [0.0, 1.0, 0.0], 10.0)
    g.union(s, b, 1.0)
    return g


def noisy_sphere_graph(radius=80.0, octaves=8):
    g = SDFGraph()  # generate_object_with_multifractal_noise (generation.rs:87-103)
    s = g.sphere(radius)
    g.multifractal_noise(s, octaves, 0.02, 2.0, 0.6, 4.0, 0)
    return g


def noisy_box_graph(extent=246.0, octaves=8):
    """BASELINE config 2: box perturbed by 8-octave gradient noise → 256³ for extent 246."""
    g = SDFGraph()
    b = g.box([extent] * 3)
    g.multifractal_noise(b, octaves, 0.02, 2.0, 0.6, 4.0, 0)
    return g


def csg_zoo_graph():
    """Every node kind once: primitives, all transforms, noise under a scaling + rotation
    (the per-voxel noise path), smooth and hard union / subtraction / intersection, a shared
    sub-graph (DAG unrolling)."""
    g = SDFGraph()
    s = g.sphere(20.0)
    c = g.capsule(24.0, 7.0)
    c = g.rotation_from_axis_angle(c, [1.0, 0.5, 0.25], 0.7)
    c = g.translation(c, [14.0, -3.0, 5.0])
    u = g.union(s, c, 4.0)
    b = g.box([18.0, 30.0, 12.0])
    b = g.translation(b, [-12.0, 6.0, -4.0])
    sub = g.subtraction(u, b, 2.0)
    n = g.multifractal_noise(sub, 3, 0.05, 2.0, 0.5, 1.5, 7)
    n = g.scaling(n, 1.3)
    n = g.rotation_from_axis_angle(n, [0.0, 0.0, 1.0], 0.3)
    s2 = g.sphere(26.0)
    inter = g.intersection(n, s2, 0.0)
    small = g.sphere(5.0)
    t1 = g.translation(small, [0.0, 22.0, 0.0])
    t2 = g.translation(small, [0.0, -22.0, 0.0])  # `small` shared by two parents
    both = g.union(t1, t2, 0.0)
    g.union(inter, both, 3.0)
    return g


def mid_noise_graph():
    """Multi-octave noise in the middle of the program (not the last node, not rotated): the evaluator's
    whole-chunk, octave-by-octave path. Lacunarity 3 makes the last octaves touch more lattice cells per chunk than
    the gradient table holds, so both the table and the direct evaluation run; the translated copy sits far from the
    noise origin (large cell indices)."""
    g = SDFGraph()
    s = g.sphere(26.0)
    n = g.multifractal_noise(s, 4, 0.11, 3.0, 0.55, 3.0, 11)
    n = g.translation(n, [7.5, -3.25, 2.0])
    b = g.box([30.0, 12.0, 44.0])
    u = g.union(n, b, 2.5)
    s2 = g.sphere(9.0)
    n2 = g.multifractal_noise(s2, 2, 0.7, 2.0, 0.5, 1.0, 3)
    n2 = g.translation(n2, [-31.0, 20.0, -18.0])
    g.union(u, n2, 0.0)
    return g


def asteroid_like_graph(n_craters=24, radius=40.0, seed=3):
    """A hand-built stand-in with the asteroid's structure (smooth union of a few spheres, noise,
    smooth subtraction of a balanced union tree of rotated capsules, final noise)."""
    rng = np.random.default_rng(seed)
    g = SDFGraph()
    body = g.sphere(radius)
    for _ in range(3):
        s = g.sphere(float(radius * rng.uniform(0.4, 0.7)))
        s = g.translation(s, list(map(float, rng.uniform(-0.6 * radius, 0.6 * radius, 3))))
        body = g.union(body, s, 0.25 * radius)
    body = g.multifractal_noise(body, 1, 0.02, 2.0, 0.5, 0.1 * radius, 1)
    craters = []
    for _ in range(n_craters):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        c = g.capsule(float(radius * 0.2), float(radius * rng.uniform(0.05, 0.15)))
        c = g.rotation_from_axis_angle(c, list(map(float, rng.normal(size=3))), float(rng.uniform(0, 3.0)))
        c = g.translation(c, list(map(float, d * radius * 0.95)))
        craters.append(c)
    while len(craters) > 1:  # balanced union tree (meta.rs:2390-2409)
        nxt = []
        for i in range(0, len(craters) - 1, 2):
            nxt.append(g.union(craters[i], craters[i + 1], 0.05 * radius))
        if len(craters) % 2:
            nxt.append(craters[-1])
        craters = nxt
    body = g.subtraction(body, craters[0], 0.05 * radius)
    g.multifractal_noise(body, 5, 0.02, 2.0, 0.546, 2.0, 2)
    return g


SAME0 = VoxelTypeGenerator.same(0)
GRADIENT4 = VoxelTypeGenerator.gradient_noise([0, 1, 2, 3], 0.02, 1.0, 0)  # generation.rs:113-124


# ---- comparisons ------------------------------------------------------------------------------

def f32_bits_equal(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Bitwise f32 equality, except that any NaN equals any NaN (x86 and CUDA produce different payloads)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def assert_objects_equal(gpu_chunks, gpu_voxels, orc_chunks, orc_voxels):
    """Per-chunk parity: kind, flags, face distributions, uniform voxel, and all 4096 voxels.

    data_offset is NOT compared: the reference's own value depends on traversal order
    (object.rs:2540-2543); voxels are matched through each side's data_offset."""
    assert len(gpu_chunks) == len(orc_chunks)
    assert np.array_equal(gpu_chunks["kind"], orc_chunks["kind"]), "chunk kinds differ"
    nu = orc_chunks["kind"] == 2
    assert np.array_equal(gpu_chunks["flags"][nu], orc_chunks["flags"][nu]), "chunk flags differ"
    assert np.array_equal(gpu_chunks["face"][nu], orc_chunks["face"][nu]), "face distributions differ"
    un = orc_chunks["kind"] == 1
    for f in ("uniform_type", "uniform_sd", "uniform_flags"):
        assert np.array_equal(gpu_chunks[f][un], orc_chunks[f][un]), f"{f} differs"
    if not nu.any():
        return
    gv = gpu_voxels.reshape(-1, 4096)[gpu_chunks["data_offset"][nu]]
    ov = orc_voxels.reshape(-1, 4096)[orc_chunks["data_offset"][nu]]
    for f in ("sd", "flags", "type"):
        if not np.array_equal(gv[f], ov[f]):
            bad = np.argwhere(gv[f] != ov[f])
            c, v = bad[0]
            cidx = np.flatnonzero(nu)[c]
            raise AssertionError(
                f"voxel field {f!r} differs in {len(bad)} voxels of {len(np.unique(bad[:, 0]))} chunks; first: chunk {cidx} "
                f"voxel {v} (i,j,k={v >> 8},{(v >> 4) & 15},{v & 15}) gpu={gv[f][c, v]} oracle={ov[f][c, v]}")


def assert_meshes_equal(gm: dict, om) -> None:
    """Bit-exact mesh parity incl. order: positions, normals, indices, index materials, submesh table."""
    assert len(gm["positions"]) == om.n_vertices, (len(gm["positions"]), om.n_vertices)
    assert len(gm["indices"]) == om.n_indices, (len(gm["indices"]), om.n_indices)
    assert len(gm["submeshes"]) == om.n_submeshes
    assert np.array_equal(gm["indices"], om.indices), "indices differ"
    assert f32_bits_equal(gm["positions"], om.positions).all(), "positions differ"
    assert f32_bits_equal(gm["normals"], om.normals).all(), "normals differ"
    assert np.array_equal(gm["index_materials"]["indices"], om.index_materials["indices"]), "index material ids differ"
    assert np.array_equal(gm["index_materials"]["weights"], om.index_materials["weights"]), "index material weights differ"
    for f in ("chunk_indices", "index_offset", "index_count", "obscured"):
        assert np.array_equal(gm["submeshes"][f], om.submeshes[f]), f"submesh {f} differs"
    assert np.array_equal(gm["vertex_ranges"], om.vertex_ranges)


def euler_characteristic(positions: np.ndarray, indices: np.ndarray, weld=1e3):
    """V - E + F after welding the vertices chunks duplicate along their seams; also the histogram
    of how many triangles use each edge (closed manifold ⇒ every edge exactly twice)."""
    key = np.round(positions.astype(np.float64) * weld).astype(np.int64)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    tris = inv[indices.reshape(-1, 3).astype(np.int64)]
    edges = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
    edges.sort(axis=1)
    ue, cnt = np.unique(edges, axis=0, return_counts=True)
    v = len(np.unique(inv))
    return v - len(ue) + len(tris), np.bincount(cnt)


# ---- canonical digests (tests/golden/object_digests.json) ----------------------------------------------

def object_digest(chunks, voxels) -> str:
    """sha256 over a canonical serialisation of a voxel object: chunk kinds, flags / face distributions of NonUniform
    chunks, uniform voxels of Uniform chunks, and the voxels of the NonUniform chunks in linear chunk order (the
    reference's own data_offset depends on traversal order, so it is not hashed)."""
    import hashlib

    h = hashlib.sha256()
    kind = np.ascontiguousarray(chunks["kind"], np.uint8)
    h.update(kind.tobytes())
    nu, un = kind == 2, kind == 1
    h.update(np.ascontiguousarray(chunks["flags"][nu], np.uint8).tobytes())
    h.update(np.ascontiguousarray(chunks["face"][nu], np.uint8).tobytes())
    for f in ("uniform_type", "uniform_sd", "uniform_flags"):
        h.update(np.ascontiguousarray(chunks[f][un]).tobytes())
    if nu.any():
        v = voxels.reshape(-1, 4096)[chunks["data_offset"][nu]]
        for f in ("type", "sd", "flags"):
            h.update(np.ascontiguousarray(v[f]).tobytes())
    return h.hexdigest()


def mesh_digest(positions, normals, indices, index_materials, submeshes, vertex_ranges) -> str:
    """sha256 over the mesh buffers in the reference's order (f32 as bits)."""
    import hashlib

    h = hashlib.sha256()
    for a in (np.ascontiguousarray(positions, np.float32), np.ascontiguousarray(normals, np.float32),
              np.ascontiguousarray(indices, np.uint32), np.ascontiguousarray(index_materials),
              np.ascontiguousarray(submeshes), np.ascontiguousarray(vertex_ranges, np.uint32)):
        h.update(a.tobytes())
    return h.hexdigest()


GOLDEN_OBJECTS = {
    "sphere64": (lambda: sphere_graph(31.0), "SAME0"),
    "zoo": (csg_zoo_graph, "GRADIENT4"),
    "asteroid_like": (lambda: asteroid_like_graph(24, 40.0), "GRADIENT4"),
    "mid_noise": (mid_noise_graph, "SAME0"),
    "noisy_box": (lambda: noisy_box_graph(38.0, 8), "SAME0"),
}



# ---- the reference's fake generators (object.rs:3387-3561) as data -------------------------------------

INSIDE = (0, -128, 0)      # Voxel::maximally_inside(VoxelType::default())
OUTSIDE = (255, 127, 1)    # Voxel::maximally_outside()


def _chunks_from_dense(types, sd, flags, grid_shape):
    """Cuts dense (type, sd, flags) grids padded to whole chunks into the per-chunk arrays a ChunkedVoxelGenerator
    returns: (n_chunks, 4096) voxels in linear chunk order + a ChunkSparseness byte per chunk computed the way the
    reference's fixtures do (has_only_empty_voxels: no non-empty voxel; is_void: every voxel maximally outside)."""
    from impact_b200._lib import VOXEL_DTYPE

    cc = [(int(s) + 15) // 16 for s in grid_shape]
    n = cc[0] * cc[1] * cc[2]
    vox = np.zeros((n, 4096), VOXEL_DTYPE)
    sp = np.zeros(n, np.uint8)
    for c in range(n):
        i, j, k = c // (cc[1] * cc[2]), (c // cc[2]) % cc[1], c % cc[2]
        sl = (slice(16 * i, 16 * i + 16), slice(16 * j, 16 * j + 16), slice(16 * k, 16 * k + 16))
        vox[c]["type"] = types[sl].reshape(-1)
        vox[c]["sd"] = sd[sl].reshape(-1)
        vox[c]["flags"] = flags[sl].reshape(-1)
        only_empty = bool(((flags[sl] & 1) != 0).all())
        is_void = bool((sd[sl] == 127).all())
        sp[c] = (1 if only_empty else 0) | (2 if is_void else 0)
    return vox, sp


def offset_box_chunks(shape, offset=(0, 0, 0), voxel=INSIDE):
    """`OffsetBoxVoxelGenerator::new(shape, offset, voxel)` → (voxels, sparseness, grid_shape)."""
    grid = [int(o) + int(s) for o, s in zip(offset, shape)]
    cc = [(g + 15) // 16 for g in grid]
    full = tuple(16 * c for c in cc)
    ty = np.full(full, OUTSIDE[0], np.uint8)
    sd = np.full(full, OUTSIDE[1], np.int8)
    fl = np.full(full, OUTSIDE[2], np.uint8)
    sl = tuple(slice(int(o), int(o) + int(s)) for o, s in zip(offset, shape))
    ty[sl], sd[sl], fl[sl] = voxel
    return (*_chunks_from_dense(ty, sd, fl, grid), grid)


def manual_chunks(cells, offset=(0, 0, 0)):
    """`ManualVoxelGenerator::<N>::with_offset(cells, offset)`: non-zero cells are maximally inside voxels."""
    cells = np.asarray(cells, np.uint8)
    grid = [int(o) + cells.shape[d] for d, o in enumerate(offset)]
    cc = [(g + 15) // 16 for g in grid]
    full = tuple(16 * c for c in cc)
    ty = np.full(full, OUTSIDE[0], np.uint8)
    sd = np.full(full, OUTSIDE[1], np.int8)
    fl = np.full(full, OUTSIDE[2], np.uint8)
    sl = tuple(slice(int(o), int(o) + cells.shape[d]) for d, o in enumerate(offset))
    solid = cells != 0
    ty[sl] = np.where(solid, INSIDE[0], OUTSIDE[0])
    sd[sl] = np.where(solid, INSIDE[1], OUTSIDE[1])
    fl[sl] = np.where(solid, INSIDE[2], OUTSIDE[2])
    return (*_chunks_from_dense(ty, sd, fl, grid), grid)


def random_voxel_chunks(shape, seed, n_types=4, fill=0.5, blobs=True):
    """A fuzz generator: random signed-distance codes (negative = non-empty, the `Voxel` invariant), random types, with
    whole-chunk regions forced solid / void so that Uniform and Void chunks and their conversions occur."""
    rng = np.random.default_rng(seed)
    cc = [(int(s) + 15) // 16 for s in shape]
    full = tuple(16 * c for c in cc)
    if blobs:
        # smooth random field → connected blobs with noisy distance codes
        coarse = rng.normal(size=tuple(c * 2 + 1 for c in cc))
        from scipy.ndimage import zoom
        field = zoom(coarse, [f / s for f, s in zip(full, coarse.shape)], order=1)
        field = field + 0.35 * rng.normal(size=full) + (0.5 - fill)
    else:
        field = rng.normal(size=full) + (0.5 - fill) * 2.0
    sd = np.clip(np.round(field * 60.0), -128, 127).astype(np.int8)
    ty = rng.integers(0, n_types, full).astype(np.uint8)
    # a solid block of whole chunks (Uniform candidates) and a void block
    if cc[0] >= 3:
        sd[16:32, 0:32, 0:32] = -128
        ty[16:32, 0:32, 0:32] = 1
    sd[-16:, -16:, :] = 127
    inside = np.zeros(full, bool)
    inside[tuple(slice(0, int(s)) for s in shape)] = True
    sd[~inside] = 127
    fl = np.where(sd >= 0, 1, 0).astype(np.uint8)
    ty = np.where(sd == 127, 255, ty).astype(np.uint8)
    return (*_chunks_from_dense(ty, sd, fl, shape), list(map(int, shape)))


# ---- mutual absorption fixtures -------------------------------------------------------------------------

def quat_from_axis_angle(axis, angle):
    """Unit quaternion (x, y, z, w) built in float32 like glam's `Quat::from_axis_angle`: (axis * sin(angle / 2),
    cos(angle / 2)) with the half angle, sine and cosine all evaluated in f32."""
    a = np.asarray(axis, np.float64)
    a = (a / np.linalg.norm(a)).astype(np.float32)
    half = np.float32(angle) * np.float32(0.5)
    s, c = np.sin(half), np.cos(half)
    return np.float32([a[0] * s, a[1] * s, a[2] * s, c])


def _rotate(q, v):
    x, y, z, w = [float(c) for c in q]
    b = np.array([x, y, z])
    return v * (w * w - b @ b) + b * (2.0 * (v @ b)) + np.cross(b, v) * (2.0 * w)


def intersection_voxel_ranges(info_a, info_b, q, t):
    """Voxel ranges encompassing the intersection of two objects' occupied boxes — a conservative stand-in for
    `determine_voxel_ranges_encompassing_intersection` (intersection.rs:707-745): each object's occupied box is carried into
    the other's frame, bounded, intersected with that object's occupied range and clamped. None when they miss."""
    ea, eb = float(info_a["voxel_extent"]), float(info_b["voxel_extent"])
    oa, ob = info_a["occupied_voxel_ranges"].astype(np.float64), info_b["occupied_voxel_ranges"].astype(np.float64)
    qc = np.array([-q[0], -q[1], -q[2], q[3]])
    corners = lambda o, e: np.array([[o[0, i] * e, o[1, j] * e, o[2, k] * e] for i in (0, 1) for j in (0, 1) for k in (0, 1)])
    b_in_a = np.array([_rotate(q, c) + np.asarray(t, np.float64) for c in corners(ob, eb)]) / ea
    a_in_b = np.array([_rotate(qc, c - np.asarray(t, np.float64)) for c in corners(oa, ea)]) / eb
    out = []
    for pts, occ in ((b_in_a, oa), (a_in_b, ob)):
        lo = np.maximum(np.floor(pts.min(0)) - 1, occ[:, 0])
        hi = np.minimum(np.ceil(pts.max(0)) + 1, occ[:, 1])
        if (lo >= hi).any():
            return None
        out.append(np.stack([lo, hi], 1).astype(np.uint32))
    return out
