// ORACLE — test infrastructure only (tests/, smoke(), bench.py's CPU arms). Never linked into the product.
// VoxelObjectCollisionProbes (engine/crates/impact_voxel/src/collidable.rs:97-101, 346-780): for every meshed chunk, the
// mesh vertex of lowest (most convex) curvature inside each block of 1^3 .. 8^3 voxels — the points the physics uses to
// probe other objects. MeshedVoxelObject::create and sync_mesh_with_object (mesh.rs:156-205) keep them beside the mesh.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "oracle.hpp"

namespace orc {

// determine_log2_block_size_for_object (collidable.rs:451-471)
uint32_t probes_log2_block_size(const Object& obj) {
    uint32_t min_extent = 0xFFFFFFFFu;
    for (int d = 0; d < 3; ++d) min_extent = std::min(min_extent, obj.occ_voxels[d][1] - obj.occ_voxels[d][0]);
    if (min_extent >= 16) return 3;
    if (min_extent >= 8) return 2;
    if (min_extent >= 4) return 1;
    return 0;
}

// add_points_for_vertices_in_blocks (collidable.rs:614-731). Vector3C::dot is x*x + y*y + z*z in f32 (impact_math
// vector.rs:846-848); every sum below keeps the reference's operation order.
void probes_points_for_chunk(uint32_t log2_block_size, const uint32_t chunk_indices[3], const float* positions, const float* normals,
                             size_t n_vertices, const uint32_t* indices, size_t n_indices, uint32_t start_index,
                             float inverse_voxel_extent, std::vector<float>& points) {
    std::vector<float> curv_sum(n_vertices, 0.0f), curv_count(n_vertices, 0.0f);
    const auto dot = [](const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    for (size_t t = 0; t + 2 < n_indices; t += 3) {
        const size_t i0 = indices[t] - start_index, i1 = indices[t + 1] - start_index, i2 = indices[t + 2] - start_index;
        const float *v0 = positions + 3 * i0, *v1 = positions + 3 * i1, *v2 = positions + 3 * i2;
        const float *n0 = normals + 3 * i0, *n1 = normals + 3 * i1, *n2 = normals + 3 * i2;
        const float e01[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
        const float e12[3] = {v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2]};
        const float e20[3] = {v0[0] - v2[0], v0[1] - v2[1], v0[2] - v2[2]};
        curv_sum[i0] += dot(n0, e01) - dot(n0, e20);
        curv_count[i0] += 2.0f;
        curv_sum[i1] += dot(n1, e12) - dot(n1, e01);
        curv_count[i1] += 2.0f;
        curv_sum[i2] += dot(n2, e20) - dot(n2, e12);
        curv_count[i2] += 2.0f;
    }
    const uint32_t log2_blocks = 4u - log2_block_size, n_blocks = 1u << (3u * log2_blocks);
    float lower[3], upper[3];
    for (int d = 0; d < 3; ++d) {
        lower[d] = (float)(chunk_indices[d] * 16u);
        upper[d] = (float)((chunk_indices[d] + 1u) * 16u);
    }
    std::vector<float> best(4 * (size_t)n_blocks, INFINITY);  // x, y, z, curvature
    for (size_t v = 0; v < n_vertices; ++v) {
        if (curv_count[v] == 0.0f) continue;  // not connected to any edge
        uint32_t block[3];
        for (int d = 0; d < 3; ++d) {
            const float norm = positions[3 * v + d] * inverse_voxel_extent;
            const float clamped = std::fmin(std::fmax(norm, lower[d]), upper[d]);
            const uint32_t voxel = (uint32_t)clamped;               // `as usize`
            block[d] = (voxel & 15u) >> log2_block_size;            // a vertex clamped to the upper face wraps to block 0
        }
        const uint32_t b = (block[0] << (2u * log2_blocks)) + (block[1] << log2_blocks) + block[2];
        const float curvature = curv_sum[v] / curv_count[v];
        if (curvature < best[4 * b + 3]) {
            best[4 * b] = positions[3 * v];
            best[4 * b + 1] = positions[3 * v + 1];
            best[4 * b + 2] = positions[3 * v + 2];
            best[4 * b + 3] = curvature;
        }
    }
    for (uint32_t b = 0; b < n_blocks; ++b)
        if (best[4 * b + 3] != INFINITY) points.insert(points.end(), best.begin() + 4 * b, best.begin() + 4 * b + 3);
}

// recompute_for_all_chunks (collidable.rs:361-392, 473-521)
void probes_compute_for_all_chunks(const Object& obj, const Mesh& mesh, CollisionProbes& pr) {
    pr.points.clear();
    pr.range_of_chunk.clear();
    pr.free_points = RangeAllocator();
    const uint32_t log2_bs = probes_log2_block_size(obj);
    const float inv = 1.0f / obj.voxel_extent;
    for (size_t s = 0; s < mesh.submeshes.size(); ++s) {
        const Submesh& sm = mesh.submeshes[s];
        const uint32_t v0 = mesh.vertex_ranges[2 * s], v1 = mesh.vertex_ranges[2 * s + 1];
        const size_t start = pr.points.size() / 3;
        probes_points_for_chunk(log2_bs, sm.chunk_indices, mesh.positions.data() + 3 * (size_t)v0, mesh.normals.data() + 3 * (size_t)v0,
                                v1 - v0, mesh.indices.data() + sm.index_offset, sm.index_count, v0, inv, pr.points);
        const size_t end = pr.points.size() / 3;
        if (end > start) pr.range_of_chunk[obj.lin(sm.chunk_indices[0], sm.chunk_indices[1], sm.chunk_indices[2])] = {start, end};
    }
}

// sync_with_voxel_object_and_mesh over `dirty` in the given order (the reference iterates a HashSet) + update_for_chunk
// (collidable.rs:394-433, 524-612)
void probes_sync(const Object& obj, const SyncedMesh& sm, const uint32_t* dirty, size_t n_dirty, CollisionProbes& pr) {
    const uint32_t log2_bs = probes_log2_block_size(obj);
    const float inv = 1.0f / obj.voxel_extent;
    const Mesh& mesh = sm.mesh;
    std::vector<float> buffer;
    for (size_t q = 0; q < n_dirty; ++q) {
        const uint32_t c = dirty[q];
        auto old = pr.range_of_chunk.find(c);
        auto row = sm.index_of_chunk.find(c);
        buffer.clear();
        if (row != sm.index_of_chunk.end()) {
            const Submesh& s = mesh.submeshes[row->second];
            const uint32_t v0 = mesh.vertex_ranges[2 * row->second], v1 = mesh.vertex_ranges[2 * row->second + 1];
            probes_points_for_chunk(log2_bs, s.chunk_indices, mesh.positions.data() + 3 * (size_t)v0, mesh.normals.data() + 3 * (size_t)v0,
                                    v1 - v0, mesh.indices.data() + s.index_offset, s.index_count, v0, inv, buffer);
        }
        if (buffer.empty()) {  // no vertices any more, or no points: the chunk goes and its range is freed
            if (old != pr.range_of_chunk.end()) {
                pr.free_points.free_range(old->second.first, old->second.second);
                pr.range_of_chunk.erase(old);
            }
            continue;
        }
        if (old != pr.range_of_chunk.end()) pr.free_points.free_range(old->second.first, old->second.second);
        const size_t count = buffer.size() / 3;
        size_t start = 0;
        if (pr.free_points.allocate_range(count, start)) {
            std::memcpy(pr.points.data() + 3 * start, buffer.data(), buffer.size() * sizeof(float));
        } else {
            start = pr.points.size() / 3;
            pr.points.insert(pr.points.end(), buffer.begin(), buffer.end());
        }
        pr.range_of_chunk[c] = {start, start + count};
    }
    pr.free_points.merge_consecutive_ranges();
}

// ---- mutual voxel-object contacts (collidable.rs:859-1050, 1288-1440) ----------------------------------------------
namespace {

inline bool sign_bit(float f) { return std::signbit(f); }
inline int vidx(int i, int j, int k) { return (i << 8) + (j << 4) + k; }

inline float object_voxel_sd(const Object& obj, uint32_t i, uint32_t j, uint32_t k) {  // voxel_maybe_unchecked(..).signed_distance().to_f32()
    const Chunk& c = obj.chunks[obj.lin(i >> 4, j >> 4, k >> 4)];
    if (c.kind == CK_VOID) return sd_decode(127);
    if (c.kind == CK_UNIFORM) return sd_decode(c.uniform_voxel.sd);
    return sd_decode(obj.chunk_voxels(c.data_offset)[vidx(i & 15, j & 15, k & 15)].sd);
}

// evaluate_sdf_from_corner_samples (object/sdf.rs:579-592)
inline float corner_value(const float d[8], V3 o) {
    const V3 ro = v3(1.0f - o.x, 1.0f - o.y, 1.0f - o.z);
    const float d00 = d[0] * ro.x + d[4] * o.x, d01 = d[1] * ro.x + d[5] * o.x;
    const float d10 = d[2] * ro.x + d[6] * o.x, d11 = d[3] * ro.x + d[7] * o.x;
    const float d0 = d00 * ro.y + d10 * o.y, d1 = d01 * ro.y + d11 * o.y;
    return d0 * ro.z + d1 * o.z;
}
// compute_sdf_gradient_from_corner_samples (object/sdf.rs:603-633)
inline V3 corner_gradient(const float d[8], V3 o) {
    const V3 r = v3(1.0f - o.x, 1.0f - o.y, 1.0f - o.z);
    const V3 d00 = v3(d[4] - d[0], d[2] - d[0], d[1] - d[0]), d01 = v3(d[5] - d[1], d[6] - d[4], d[3] - d[2]);
    const V3 d10 = v3(d[6] - d[2], d[3] - d[1], d[5] - d[4]), d11 = v3(d[7] - d[3], d[7] - d[5], d[7] - d[6]);
    V3 g;
    g.x = (((r.y * r.z) * d00.x + (r.y * o.z) * d01.x) + (o.y * r.z) * d10.x) + (o.y * o.z) * d11.x;
    g.y = (((r.z * r.x) * d00.y + (r.z * o.x) * d01.y) + (o.z * r.x) * d10.y) + (o.z * o.x) * d11.y;
    g.z = (((r.x * r.y) * d00.z + (r.x * o.y) * d01.z) + (o.x * r.y) * d10.z) + (o.x * o.y) * d11.z;
    return g;
}
// UnitVector3::normalized_from_if_above (impact_math vector.rs:1146-1173)
inline bool normalized_if_above(V3 v, float min_norm, V3& out) {
    const float n2 = dot(v, v);
    if (!(n2 > min_norm * min_norm)) return false;
    const float n = std::sqrt(n2);
    out = v3(v.x / n, v.y / n, v.z / n);
    return true;
}
const float MIN_SD_F32 = QUANT_STEP * -128.0f;

// estimate_sdf_value_and_normal_at_point_deep_inside (collidable.rs:1424-1440)
bool deep_inside(V3 norm_center, V3 p, float& sd, V3& normal) {
    sd = MIN_SD_F32;
    return normalized_if_above(p - norm_center, 1e-8f, normal);
}

// determine_sdf_value_and_normal_at_point_if_intersecting (collidable.rs:1288-1422)
bool sdf_value_and_normal_if_intersecting(const Object& obj, const uint32_t dims[3], V3 norm_center, V3 p, float& sd, V3& normal) {
    const float HALF_VOXEL_DIAGONAL = 0.5f * 1.7320508f;
    const V3 lower = v3(p.x - 0.5f, p.y - 0.5f, p.z - 0.5f);
    if (sign_bit(lower.x) || sign_bit(lower.y) || sign_bit(lower.z)) return false;
    const uint64_t li = (uint64_t)lower.x, lj = (uint64_t)lower.y, lk = (uint64_t)lower.z;
    if ((li + 1 >= dims[0]) | (lj + 1 >= dims[1]) | (lk + 1 >= dims[2])) return false;
    const uint32_t ci = (uint32_t)p.x, cj = (uint32_t)p.y, ck = (uint32_t)p.z;
    const Chunk& chunk = obj.chunks[obj.lin(ci >> 4, cj >> 4, ck >> 4)];
    if (chunk.kind == CK_UNIFORM) return deep_inside(norm_center, p, sd, normal);
    if (chunk.kind == CK_VOID) return false;
    const float containing = sd_decode(obj.chunk_voxels(chunk.data_offset)[vidx(ci & 15, cj & 15, ck & 15)].sd);
    if (containing > HALF_VOXEL_DIAGONAL) return false;
    const uint32_t a = (uint32_t)li, b = (uint32_t)lj, c = (uint32_t)lk;
    // (the reference reads the eight samples straight from the chunk when they all lie in it, through the chunk table
    // otherwise: the same values)
    const float d[8] = {object_voxel_sd(obj, a, b, c),         object_voxel_sd(obj, a, b, c + 1),
                        object_voxel_sd(obj, a, b + 1, c),     object_voxel_sd(obj, a, b + 1, c + 1),
                        object_voxel_sd(obj, a + 1, b, c),     object_voxel_sd(obj, a + 1, b, c + 1),
                        object_voxel_sd(obj, a + 1, b + 1, c), object_voxel_sd(obj, a + 1, b + 1, c + 1)};
    const V3 fo = v3(lower.x - std::floor(lower.x), lower.y - std::floor(lower.y), lower.z - std::floor(lower.z));
    sd = corner_value(d, fo);
    if (sd > 0.0f) return false;
    if (std::fabs(sd - MIN_SD_F32) < 1e-3f) return deep_inside(norm_center, p, sd, normal);
    return normalized_if_above(corner_gradient(d, fo), 1e-8f, normal);
}

// one direction of for_each_mutual_voxel_object_contact: the probes of `from` against the distance field of `into`
void probes_against(const Object& from, const CollisionProbes& probes, const Isometry& world_to_from, const Object& into,
                    const InertialMoments& inertial_into, const Isometry& world_to_into, const uint32_t ranges_in_from[3][2],
                    float aabb_margin, bool flip_normal, std::vector<VoxelContact>& out) {
    const uint32_t dims[3] = {into.chunk_counts[0] * 16u, into.chunk_counts[1] * 16u, into.chunk_counts[2] * 16u};
    const float inv_into = 1.0f / into.voxel_extent, inv_from = 1.0f / from.voxel_extent;
    // derive_center_of_mass() * inverse_voxel_extent
    const V3 com = v3(inertial_into.moments[0] / inertial_into.mass, inertial_into.moments[1] / inertial_into.mass,
                      inertial_into.moments[2] / inertial_into.mass);
    const V3 norm_center = v3(com.x * inv_into, com.y * inv_into, com.z * inv_into);
    // aabb_from_voxel_ranges(voxel_extent, ranges).expanded_about_center(margin)
    float lo[3], hi[3];
    uint32_t cr[3][2];
    for (int d = 0; d < 3; ++d) {
        lo[d] = from.voxel_extent * (float)ranges_in_from[d][0] - aabb_margin;
        hi[d] = from.voxel_extent * (float)ranges_in_from[d][1] + aabb_margin;
        cr[d][0] = ranges_in_from[d][0] / 16u;
        cr[d][1] = (ranges_in_from[d][1] + 15u) / 16u;
    }
    const Quat from_inv = quat_conj(world_to_from.q), into_inv = quat_conj(world_to_into.q);
    // chunk_point_ranges() is a HashMap in the reference; here: ascending linear chunk index
    std::vector<uint32_t> keys;
    for (const auto& kv : probes.range_of_chunk) keys.push_back(kv.first);
    std::sort(keys.begin(), keys.end());
    for (uint32_t c : keys) {
        const uint32_t ci = c / (from.chunk_counts[1] * from.chunk_counts[2]), cj = (c / from.chunk_counts[2]) % from.chunk_counts[1],
                       ck = c % from.chunk_counts[2];
        if (ci < cr[0][0] || ci >= cr[0][1] || cj < cr[1][0] || cj >= cr[1][1] || ck < cr[2][0] || ck >= cr[2][1]) continue;
        const auto& range = probes.range_of_chunk.at(c);
        for (size_t q = range.first; q < range.second; ++q) {
            const V3 pf = v3(probes.points[3 * q], probes.points[3 * q + 1], probes.points[3 * q + 2]);
            // contains_point: no sign bit in (p - lower) or (upper - p)
            if (sign_bit(pf.x - lo[0]) || sign_bit(pf.y - lo[1]) || sign_bit(pf.z - lo[2]) || sign_bit(hi[0] - pf.x) ||
                sign_bit(hi[1] - pf.y) || sign_bit(hi[2] - pf.z))
                continue;
            const V3 world = quat_rotate(from_inv, pf - world_to_from.t);  // inverse_transform_point
            const V3 in_into = quat_rotate(world_to_into.q, world) + world_to_into.t;
            const V3 np = v3(in_into.x * inv_into, in_into.y * inv_into, in_into.z * inv_into);
            float sd;
            V3 normal;
            if (!sdf_value_and_normal_if_intersecting(into, dims, norm_center, np, sd, normal)) continue;
            V3 n = quat_rotate(into_inv, normal);
            if (flip_normal) n = v3(-n.x, -n.y, -n.z);
            const float depth = -sd * into.voxel_extent;
            const V3 nf = v3(pf.x * inv_from, pf.y * inv_from, pf.z * inv_from);
            out.push_back(VoxelContact{{(uint32_t)nf.x, (uint32_t)nf.y, (uint32_t)nf.z}, {world.x, world.y, world.z}, {n.x, n.y, n.z}, depth});
        }
    }
}

}  // namespace

// for_each_mutual_voxel_object_contact (collidable.rs:859-1050) given the voxel ranges encompassing the intersection
// (determine_voxel_ranges_encompassing_intersection stays with the caller, like for the mutual absorption): A's probes
// against B's distance field first, then B's against A's. The contact ids are [0, i, j, k] of the probing voxel.
void mutual_voxel_object_contacts(const Object& a, const CollisionProbes& probes_a, const InertialMoments& inertial_a,
                                  const Isometry& world_to_a, const Object& b, const CollisionProbes& probes_b,
                                  const InertialMoments& inertial_b, const Isometry& world_to_b, const uint32_t ranges_in_a[3][2],
                                  const uint32_t ranges_in_b[3][2], std::vector<VoxelContact>& a_against_b,
                                  std::vector<VoxelContact>& b_against_a) {
    a_against_b.clear();
    b_against_a.clear();
    probes_against(a, probes_a, world_to_a, b, inertial_b, world_to_b, ranges_in_a, a.voxel_extent, false, a_against_b);
    // (the reference expands B's box by A's voxel extent too, collidable.rs:969-973)
    probes_against(b, probes_b, world_to_b, a, inertial_a, world_to_a, ranges_in_b, a.voxel_extent, true, b_against_a);
}

}  // namespace orc
