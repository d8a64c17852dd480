"""Canonical sha256 digests of voxel objects and meshes (pure numpy; no device, no oracle).

Used by the parity tests and by bench.py's `parity` block to compare what the CUDA path produced with the
digests of the CPU oracle's output committed under tests/golden/. The object digest is taken per x-plane of
chunks so that the x-slab objects of a multi-GPU run (each rank holds a contiguous range of planes) are
checked against the same committed values as the single-GPU object.

The reference's `data_offset` of a chunk depends on traversal order (object.rs:2540-2543) and is not hashed;
voxels are hashed in linear chunk order (i → j → k, object.rs:3140-3145) through each side's offsets."""
from __future__ import annotations

import hashlib

import numpy as np


def object_plane_digests(chunks, voxels, chunk_counts) -> list:
    """One sha256 per x-plane of chunks: chunk kinds of the plane, flags / face distributions of its NonUniform
    chunks, the uniform voxels of its Uniform chunks, then type / sd / flags of the NonUniform chunks' voxels."""
    cy, cz = int(chunk_counts[1]), int(chunk_counts[2])
    per_plane = cy * cz
    n_planes = len(chunks) // per_plane
    assert n_planes * per_plane == len(chunks)
    vox = voxels.reshape(-1, 4096)
    out = []
    for p in range(n_planes):
        c = chunks[p * per_plane:(p + 1) * per_plane]
        kind = np.ascontiguousarray(c["kind"], np.uint8)
        h = hashlib.sha256()
        h.update(kind.tobytes())
        nu, un = kind == 2, kind == 1
        h.update(np.ascontiguousarray(c["flags"][nu], np.uint8).tobytes())
        h.update(np.ascontiguousarray(c["face"][nu], np.uint8).tobytes())
        for f in ("uniform_type", "uniform_sd", "uniform_flags"):
            h.update(np.ascontiguousarray(c[f][un]).tobytes())
        if nu.any():
            v = vox[c["data_offset"][nu]]
            for f in ("type", "sd", "flags"):
                h.update(np.ascontiguousarray(v[f]).tobytes())
        out.append(h.hexdigest())
    return out


def combine(plane_digests) -> str:
    """Digest of a whole object from its plane digests."""
    return hashlib.sha256("".join(plane_digests).encode()).hexdigest()


def _f32_bits(a) -> np.ndarray:
    """f32 as bits with every NaN mapped to one pattern (x86 and CUDA produce different NaN payloads)."""
    a = np.ascontiguousarray(a, np.float32)
    bits = a.view(np.uint32)
    nan = np.isnan(a)
    if nan.any():
        bits = bits.copy()
        bits[nan] = 0x7FC00000
    return bits


def mesh_digest(positions, normals, indices, index_materials, submeshes, vertex_ranges) -> str:
    """sha256 over the mesh buffers in the reference's order (f32 as bits, NaNs canonical)."""
    h = hashlib.sha256()
    for a in (_f32_bits(positions), _f32_bits(normals),
              np.ascontiguousarray(indices, np.uint32), np.ascontiguousarray(index_materials),
              np.ascontiguousarray(submeshes), np.ascontiguousarray(vertex_ranges, np.uint32)):
        h.update(a.tobytes())
    return h.hexdigest()
