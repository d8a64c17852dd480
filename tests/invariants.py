"""The reference's brute-force `validate_*` invariants (object.rs:1302-1651, object/sdf.rs:510-571,
used by its fuzz targets) restated over the downloaded (chunks, voxels) arrays, so they apply
unchanged to the oracle's and to the GPU library's output."""
from __future__ import annotations

import numpy as np

EMPTY = 1
DN = [1 << 2, 1 << 3, 1 << 4]
UP = [1 << 5, 1 << 6, 1 << 7]


def dense_fields(chunks, voxels, chunk_counts):
    """Dense (sd, flags, type) grids of shape chunk_counts*16 with `get_voxel_if_occupied` semantics
    (object.rs:1040-1075): empty voxels read as `Voxel::maximally_outside()`."""
    cx, cy, cz = chunk_counts
    sd = np.full((cx * 16, cy * 16, cz * 16), 127, np.int8)
    fl = np.full(sd.shape, EMPTY, np.uint8)
    ty = np.full(sd.shape, 255, np.uint8)
    vox = voxels.reshape(-1, 4096)
    for c, ch in enumerate(chunks):
        i, j, k = c // (cy * cz), (c // cz) % cy, c % cz
        sl = (slice(16 * i, 16 * i + 16), slice(16 * j, 16 * j + 16), slice(16 * k, 16 * k + 16))
        if ch["kind"] == 1:
            sd[sl], fl[sl], ty[sl] = ch["uniform_sd"], ch["uniform_flags"], ch["uniform_type"]
        elif ch["kind"] == 2:
            v = vox[ch["data_offset"]].reshape(16, 16, 16)
            sd[sl], fl[sl], ty[sl] = v["sd"], v["flags"], v["type"]
    empty = (fl & EMPTY) != 0
    fl_eff = np.where(empty, np.uint8(EMPTY), fl)
    return sd, fl, fl_eff, ty, empty  # (sd, raw flags, effective flags, type, empty)


def validate_occupied_voxel_ranges(chunks, voxels, chunk_counts, occupied_voxel_ranges):
    _, _, _, _, empty = dense_fields(chunks, voxels, chunk_counts)
    ne = np.argwhere(~empty)
    if len(ne) == 0:
        expected = np.zeros((3, 2), np.uint32)
    else:
        expected = np.stack([ne.min(0), ne.max(0) + 1], 1).astype(np.uint32)
    assert np.array_equal(np.asarray(occupied_voxel_ranges, np.uint32), expected), (occupied_voxel_ranges, expected)


def validate_adjacencies(chunks, voxels, chunk_counts):
    """object.rs:1391-1487."""
    _, _, fl, _, empty = dense_fields(chunks, voxels, chunk_counts)  # fl = EFFECTIVE flags (3rd field)
    for d in range(3):
        # neighbour in +d (outside the grid = maximally outside = empty, no flags)
        up_empty = np.ones_like(empty)
        up_fl = np.full_like(fl, EMPTY)
        src = [slice(None)] * 3
        dst = [slice(None)] * 3
        src[d] = slice(1, None)
        dst[d] = slice(0, -1)
        up_empty[tuple(dst)] = empty[tuple(src)]
        up_fl[tuple(dst)] = fl[tuple(src)]
        # voxel empty → the +d neighbour must not carry the DN flag; else a non-empty neighbour must
        bad = empty & ((up_fl & DN[d]) != 0)
        assert not bad.any(), f"dim {d}: DN flag present above an empty voxel at {np.argwhere(bad)[:5]}"
        bad = ~empty & ~up_empty & ((up_fl & DN[d]) == 0)
        assert not bad.any(), f"dim {d}: DN flag missing above a non-empty voxel at {np.argwhere(bad)[:5]}"
        bad = up_empty & ((fl & UP[d]) != 0)
        assert not bad.any(), f"dim {d}: UP flag present below an empty voxel at {np.argwhere(bad)[:5]}"
        bad = ~up_empty & ~empty & ((fl & UP[d]) == 0)
        assert not bad.any(), f"dim {d}: UP flag missing at {np.argwhere(bad)[:5]}"
        first = [slice(None)] * 3
        first[d] = 0
        bad = (fl[tuple(first)] & DN[d]) != 0
        assert not bad.any(), f"dim {d}: DN flag on the object's lower face"


def _face(ch, dim, side):
    if ch["kind"] == 0:
        return 0
    if ch["kind"] == 1:
        return 1
    return int(ch["face"][2 * dim + side])


def validate_chunk_obscuredness(chunks, chunk_counts):
    """object.rs:1489-1651."""
    cx, cy, cz = chunk_counts
    void = np.zeros(1, chunks.dtype)[0]

    def get(i, j, k):
        if i >= cx or j >= cy or k >= cz:
            return void
        return chunks[(i * cy + j) * cz + k]

    def check(ch, bit, expected, where):
        if ch["kind"] == 0:
            return
        if ch["kind"] == 1:
            assert expected, f"uniform chunk not completely obscured at {where}"
            return
        assert bool(ch["flags"] & bit) == expected, f"obscured bit {bit:#x} should be {expected} at {where}"

    for i in range(cx):
        for j in range(cy):
            for k in range(cz):
                ch = get(i, j, k)
                ups = [get(i + 1, j, k), get(i, j + 1, k), get(i, j, k + 1)]
                for d in range(3):
                    check(ups[d], 1 << d, _face(ch, d, 1) == 1, (i, j, k, d, "upper neighbour"))
                    check(ch, 1 << (3 + d), _face(ups[d], d, 0) == 1, (i, j, k, d))
    for d, rng in enumerate([(1, 2), (0, 2), (0, 1)]):
        for a in range(chunk_counts[rng[0]]):
            for b in range(chunk_counts[rng[1]]):
                idx = [0, 0, 0]
                idx[rng[0]], idx[rng[1]] = a, b
                check(get(*idx), 1 << d, False, (tuple(idx), "lower object face"))


def validate_brick(values, types, adj6, chunk, chunk_idx3, chunks, voxels, chunk_counts, dense=None):
    """object/sdf.rs:510-571 for one chunk's 18³ brick."""
    sd, fl, _, ty, empty = dense if dense is not None else dense_fields(chunks, voxels, chunk_counts)
    v = values.reshape(18, 18, 18)
    t = types.reshape(18, 18, 18)
    lo = np.array(chunk_idx3) * 16 - 1
    shape = np.array(sd.shape)
    for bi in range(18):
        for bj in range(18):
            for bk in range(18):
                g = lo + (bi, bj, bk)
                inside = (g >= 0).all() and (g < shape).all()
                is_empty = True if not inside else bool(empty[tuple(g)])
                neg = np.signbit(v[bi, bj, bk])
                assert neg != is_empty, f"brick sign mismatch at {(bi, bj, bk)}"
                if not is_empty:
                    assert t[bi, bj, bk] == ty[tuple(g)]
