// ORACLE — test infrastructure only; never linked into the product library.
//
// CPU restatement of the disconnected-region extraction of
// engine/crates/impact_voxel/src/object/extraction.rs:
//   extract_any_disconnected_region(_with_property_transferrer)   :78-113
//   extract_smallest_region_with_property_transferrer             :121-281  (region chunk list, bounds)
//   extract_disconnected_region                                   :297-600
//   complete_extracted_voxel_object                               :1902-1973 (NON_EMPTY_VOXEL_THRESHOLD = 8, object.rs:203)
//   create_extracted_voxel_object_in_single_chunk_if_possible     :1976-2141
//   determine_occupied_voxel_ranges                               object.rs:3082-3137
// in terms of the objects' voxel and chunk state (the split detector's own arrays are replaced by
// resolve_connected_regions of split_detection.cpp, which reports the same labels and roots).
#include <algorithm>
#include <cstring>

#include "oracle.hpp"

namespace orc {

namespace {

inline int vlin(int i, int j, int k) { return (i << 8) | (j << 4) | k; }
const Voxel MAXIMALLY_OUTSIDE{255, 127, FLAG_EMPTY};  // Voxel::maximally_outside (lib.rs:300-310)

// determine_occupied_voxel_ranges (object.rs:3082-3137): false if there is no non-empty voxel
bool occupied_voxel_ranges_of(const Object& o, uint32_t r[3][2]) {
    uint32_t lo[3] = {UINT32_MAX, UINT32_MAX, UINT32_MAX}, hi[3] = {0, 0, 0};
    bool any = false;
    for (uint32_t i = 0; i < o.chunk_counts[0]; ++i)
        for (uint32_t j = 0; j < o.chunk_counts[1]; ++j)
            for (uint32_t k = 0; k < o.chunk_counts[2]; ++k) {
                const Chunk& c = o.chunks[o.lin(i, j, k)];
                const uint32_t off[3] = {i * 16, j * 16, k * 16};
                if (c.kind == CK_NONUNIFORM) {
                    const Voxel* v = o.chunk_voxels(c.data_offset);
                    for (int a = 0; a < 16; ++a)
                        for (int b = 0; b < 16; ++b)
                            for (int d = 0; d < 16; ++d)
                                if (!(v[vlin(a, b, d)].flags & FLAG_EMPTY)) {
                                    const uint32_t p[3] = {off[0] + a, off[1] + b, off[2] + d};
                                    for (int q = 0; q < 3; ++q) {
                                        lo[q] = std::min(lo[q], p[q]);
                                        hi[q] = std::max(hi[q], p[q]);
                                    }
                                    any = true;
                                }
                } else if (c.kind == CK_UNIFORM) {
                    for (int q = 0; q < 3; ++q) {
                        lo[q] = std::min(lo[q], off[q]);
                        hi[q] = std::max(hi[q], off[q] + 15);
                    }
                    any = true;
                }
            }
    for (int q = 0; q < 3; ++q) {
        r[q][0] = any ? lo[q] : 0;
        r[q][1] = any ? hi[q] + 1 : 0;
    }
    return any;
}

void finish_extracted(Object& e) {
    // extract_disconnected_region :585-597: cross-chunk state from scratch, then tighten the ranges
    for (int d = 0; d < 3; ++d) {
        e.occ_chunks[d][0] = 0;
        e.occ_chunks[d][1] = e.chunk_counts[d];
        e.occ_voxels[d][0] = 0;
        e.occ_voxels[d][1] = e.chunk_counts[d] * 16;
    }
    update_all_chunk_boundary_adjacencies(e);
    update_occupied_chunk_ranges(e);
    update_occupied_voxel_ranges(e);
}

}  // namespace

void extract_any_disconnected_region(Object& obj, Extraction& out) {
    out = Extraction{};
    SplitDetection sd;
    resolve_connected_regions(obj, sd);
    if (!sd.has_two) return;
    out.found_two = true;
    const uint32_t R = sd.two[sd.smallest];
    out.region_label = R;
    const RegionStats& st = sd.stats[sd.smallest];
    uint32_t r0[3], r1[3];
    for (int d = 0; d < 3; ++d) {
        r0[d] = st.chunk_min[d];
        r1[d] = st.chunk_max[d] + 1;
    }
    Object& e = out.object;
    e.voxel_extent = obj.voxel_extent;
    for (int d = 0; d < 3; ++d) e.chunk_counts[d] = r1[d] - r0[d];
    uint32_t n_uniform = 0, n_non_uniform = 0;

    for (uint32_t i = r0[0]; i < r1[0]; ++i)
        for (uint32_t j = r0[1]; j < r1[1]; ++j)
            for (uint32_t k = r0[2]; k < r1[2]; ++k) {
                const uint32_t c = obj.lin(i, j, k);
                Chunk& src = obj.chunks[c];
                const ChunkRegions& cr = sd.per_chunk[c];
                bool in_region = false, mixed = false;
                bool split_off[256];
                for (uint32_t r = 0; r < cr.region_count; ++r) {
                    split_off[r] = sd.region_root[cr.first_region + r] == R;
                    in_region = in_region || split_off[r];
                    mixed = mixed || !split_off[r];
                }
                if (!in_region || src.kind == CK_VOID) {
                    e.chunks.push_back(Chunk{});  // padding of the extracted object's chunk grid
                    continue;
                }
                if (src.kind == CK_UNIFORM) {
                    e.chunks.push_back(src);
                    src = Chunk{};
                    n_uniform++;
                    continue;
                }
                Chunk dst;
                dst.kind = CK_NONUNIFORM;
                dst.data_offset = n_non_uniform++;
                dst.flags = 0;
                const size_t start = e.voxels.size();
                e.voxels.resize(start + CHUNK_VOXELS);
                Voxel* dv = e.voxels.data() + start;
                Voxel* sv = obj.chunk_voxels(src.data_offset);
                if (mixed) {
                    const uint8_t* labels = sd.voxel_labels.data() + (size_t)src.data_offset * 4096;
                    for (int v = 0; v < 4096; ++v) {
                        if (sv[v].flags & FLAG_EMPTY) {
                            dv[v] = sv[v];  // empty voxels shape the mesh next to the surface: copied unconditionally
                        } else if (split_off[labels[v]]) {
                            dv[v] = sv[v];
                            sv[v] = MAXIMALLY_OUTSIDE;
                        } else {
                            dv[v] = MAXIMALLY_OUTSIDE;
                        }
                    }
                    update_all_internal_state(dst, dv);
                } else {
                    std::memcpy(dv, sv, sizeof(Voxel) * CHUNK_VOXELS);
                    for (int v = 0; v < 4096; ++v) sv[v] = MAXIMALLY_OUTSIDE;
                    std::memcpy(dst.face, src.face, sizeof(dst.face));
                    src = Chunk{};
                }
                e.chunks.push_back(dst);
                if (src.kind == CK_NONUNIFORM) update_all_internal_state(src, sv);
                obj.mark_dirty(c);
            }

    // ---- the original object (:556-571) ----
    update_occupied_chunk_ranges(obj);
    update_occupied_voxel_ranges(obj);
    uint32_t br[3][2];
    for (int d = 0; d < 3; ++d) {
        br[d][0] = r0[d] > 0 ? r0[d] - 1 : 0;
        br[d][1] = r1[d];
    }
    update_upper_boundary_adjacencies_in_ranges(obj, br);

    // ---- complete_extracted_voxel_object ----
    for (int d = 0; d < 3; ++d) out.origin_offset_in_parent[d] = r0[d] * 16;
    if (n_uniform == 0) {
        uint32_t non_empty = 0;
        for (const Voxel& v : e.voxels) {
            if (!(v.flags & FLAG_EMPTY) && ++non_empty >= 8) break;
        }
        if (non_empty < 8) {
            out.discarded = true;  // the voxels are dropped (NotExtracted)
            e = Object{};
            return;
        }
    }
    const uint32_t n_chunks = e.chunk_counts[0] * e.chunk_counts[1] * e.chunk_counts[2];
    if (e.chunk_counts[0] <= 2 && e.chunk_counts[1] <= 2 && e.chunk_counts[2] <= 2 && n_uniform == 0 && n_chunks > 1) {
        uint32_t occ[3][2];
        occupied_voxel_ranges_of(e, occ);
        if (occ[0][1] - occ[0][0] <= 14 && occ[1][1] - occ[1][0] <= 14 && occ[2][1] - occ[2][0] <= 14) {
            uint32_t org[3];
            for (int d = 0; d < 3; ++d) org[d] = occ[d][0] > 0 ? occ[d][0] - 1 : 0;  // room for an empty boundary layer
            std::vector<Voxel> single(CHUNK_VOXELS, MAXIMALLY_OUTSIDE);
            for (uint32_t ci = 0; ci < e.chunk_counts[0]; ++ci)
                for (uint32_t cj = 0; cj < e.chunk_counts[1]; ++cj)
                    for (uint32_t ck = 0; ck < e.chunk_counts[2]; ++ck) {
                        const Chunk& c = e.chunks[e.lin(ci, cj, ck)];
                        if (c.kind != CK_NONUNIFORM) continue;
                        const uint32_t cc[3] = {ci, cj, ck};
                        uint32_t s0[3], span[3], d0[3];
                        for (int d = 0; d < 3; ++d) {
                            const uint32_t base = cc[d] * 16;
                            const uint32_t a = std::min<uint32_t>(org[d] > base ? org[d] - base : 0, 16);
                            const uint32_t b = std::min<uint32_t>(org[d] + 16 > base ? org[d] + 16 - base : 0, 16);
                            s0[d] = a;
                            span[d] = b > a ? b - a : 0;
                            d0[d] = base > org[d] ? base - org[d] : 0;
                        }
                        const Voxel* v = e.chunk_voxels(c.data_offset);
                        for (uint32_t a = 0; a < span[0]; ++a)
                            for (uint32_t b = 0; b < span[1]; ++b)
                                for (uint32_t d = 0; d < span[2]; ++d)
                                    single[vlin(d0[0] + a, d0[1] + b, d0[2] + d)] = v[vlin(s0[0] + a, s0[1] + b, s0[2] + d)];
                    }
            Chunk sc;
            sc.kind = CK_NONUNIFORM;
            sc.data_offset = 0;
            sc.flags = 0;
            for (int d = 0; d < 3; ++d) {
                sc.face[d][0] = org[d] == occ[d][0] ? FD_MIXED : FD_EMPTY;
                sc.face[d][1] = org[d] + 16 == occ[d][1] ? FD_MIXED : FD_EMPTY;
            }
            update_internal_adjacencies(single.data());
            for (int d = 0; d < 3; ++d) {
                out.origin_offset_in_parent[d] += org[d];
                e.chunk_counts[d] = 1;
            }
            e.chunks.assign(1, sc);
            e.voxels = single;
            out.single_chunk = true;
        }
    }
    finish_extracted(e);
    out.extracted = true;
}

}  // namespace orc
