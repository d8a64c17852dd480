// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).
//
// Restates (V = engine/crates/impact_voxel/src):
//   VoxelObject::modify_voxels_within_sphere       V/object/intersection.rs:283-394
//   handle_chunk_voxels_modified                   V/object/intersection.rs:539-598
//   voxel_ranges_touching_aab                      V/object/intersection.rs:766-784
//   VoxelAbsorbingSphere::compute_new_signed_distance  V/interaction/absorption.rs:170-179
//   apply_sphere_absorption closure                V/interaction/absorption.rs:823-843
//   Voxel::set_signed_distance                     V/lib.rs:451-461
#include <algorithm>
#include <cmath>

#include "oracle.hpp"

namespace orc {

static inline int vidx(int i, int j, int k) { return (i << 8) + (j << 4) + k; }

void absorb_sphere(Object& obj, V3 center, float radius, float influence_radius,
                   AbsorbStats* stats) {
    AbsorbStats st{0, 0, 0, 0};
    // Sphere::compute_aabb: centre ± radius
    float lo[3] = {center.x - influence_radius, center.y - influence_radius,
                   center.z - influence_radius};
    float hi[3] = {center.x + influence_radius, center.y + influence_radius,
                   center.z + influence_radius};
    uint32_t tr[3][2];
    bool empty = false;
    for (int d = 0; d < 3; ++d) {
        float fl = std::fmax(std::floor(lo[d]), 0.0f);
        float ce = std::ceil(hi[d]);
        // `as usize` saturating casts
        uint32_t s = fl >= 4294967296.0f ? UINT32_MAX : (uint32_t)fl;
        uint32_t e = !(ce > 0.0f) ? 0u : (ce >= 4294967296.0f ? UINT32_MAX : (uint32_t)ce);
        tr[d][0] = std::max(obj.occ_voxels[d][0], s);
        tr[d][1] = std::min(obj.occ_voxels[d][1], e);
        if (tr[d][0] >= tr[d][1]) empty = true;
    }
    if (empty) {
        if (stats) *stats = st;
        return;
    }
    uint32_t cr[3][2];
    for (int d = 0; d < 3; ++d) {
        cr[d][0] = tr[d][0] / 16;
        cr[d][1] = (tr[d][1] + 15) / 16;
    }
    const float r2 = influence_radius * influence_radius;
    bool removed_chunks = false;

    for (uint32_t ci = cr[0][0]; ci < cr[0][1]; ++ci)
        for (uint32_t cj = cr[1][0]; cj < cr[1][1]; ++cj)
            for (uint32_t ck = cr[2][0]; ck < cr[2][1]; ++ck) {
                uint32_t cidx = obj.lin(ci, cj, ck);
                Chunk& chunk = obj.chunks[cidx];
                if (chunk.kind == CK_VOID) continue;
                if (chunk.kind == CK_UNIFORM) {
                    size_t start = obj.voxels.size();
                    obj.voxels.resize(start + CHUNK_VOXELS, chunk.uniform_voxel);
                    chunk.kind = CK_NONUNIFORM;
                    chunk.data_offset = (uint32_t)(start >> 12);
                    for (int d = 0; d < 3; ++d) chunk.face[d][0] = chunk.face[d][1] = FD_FULL;
                    chunk.flags = CF_OBSCURED_ALL;
                }
                uint32_t cc[3] = {ci, cj, ck};
                uint32_t vr[3][2], tv[3][2];
                for (int d = 0; d < 3; ++d) {
                    vr[d][0] = cc[d] * 16;
                    vr[d][1] = (cc[d] + 1) * 16;
                    tv[d][0] = std::max(vr[d][0], tr[d][0]);
                    tv[d][1] = std::min(vr[d][1], tr[d][1]);
                }
                Voxel* v = obj.chunk_voxels(chunk.data_offset);
                bool touched = false;
                for (uint32_t i = tv[0][0]; i < tv[0][1]; ++i)
                    for (uint32_t j = tv[1][0]; j < tv[1][1]; ++j)
                        for (uint32_t k = tv[2][0]; k < tv[2][1]; ++k) {
                            V3 p = v3((float)i + 0.5f, (float)j + 0.5f, (float)k + 0.5f);
                            V3 df = center - p;
                            float d2 = dot(df, df);
                            if (d2 < r2) {
                                Voxel& vx = v[vidx(i & 15, j & 15, k & 15)];
                                bool was_empty = vx.flags & FLAG_EMPTY;
                                float sphere_sd = std::sqrt(d2) - radius;
                                float nsd = std::fmax(sd_decode(vx.sd), -sphere_sd);
                                vx.sd = sd_encode(nsd);
                                if (!(vx.sd < 0)) {
                                    vx.flags |= FLAG_EMPTY;
                                    if (!was_empty) st.emptied_voxels++;
                                }
                                st.touched_voxels++;
                                touched = true;
                            }
                        }
                if (touched) {
                    st.touched_chunks++;
                    Sparseness sp = update_all_internal_state(chunk, v);
                    if (sp.is_void) {
                        chunk = Chunk{};
                        removed_chunks = true;
                        st.removed_chunks++;
                    }
                    obj.mark_dirty(cidx);
                    for (int d = 0; d < 3; ++d) {
                        if (cc[d] > 0 && tv[d][0] - vr[d][0] < 2) {
                            uint32_t a[3] = {ci, cj, ck};
                            a[d] -= 1;
                            obj.mark_dirty(obj.lin(a[0], a[1], a[2]));
                        }
                        if (cc[d] + 1 < obj.chunk_counts[d] && vr[d][1] - tv[d][1] < 2) {
                            uint32_t a[3] = {ci, cj, ck};
                            a[d] += 1;
                            obj.mark_dirty(obj.lin(a[0], a[1], a[2]));
                        }
                    }
                }
            }
    if (removed_chunks) {
        update_occupied_chunk_ranges(obj);
        update_occupied_voxel_ranges(obj);
    }
    uint32_t br[3][2];
    for (int d = 0; d < 3; ++d) {
        br[d][0] = cr[d][0] > 0 ? cr[d][0] - 1 : 0;
        br[d][1] = cr[d][1];
    }
    update_upper_boundary_adjacencies_in_ranges(obj, br);
    if (stats) *stats = st;
}

}  // namespace orc
