// ORACLE — test infrastructure only (tests/, smoke(), bench.py's CPU arms). Never linked into the product.
// VoxelObjectCollisionProbes (engine/crates/impact_voxel/src/collidable.rs:97-101, 346-780): for every meshed chunk, the
// mesh vertex of lowest (most convex) curvature inside each block of 1^3 .. 8^3 voxels — the points the physics uses to
// probe other objects. MeshedVoxelObject::create and sync_mesh_with_object (mesh.rs:156-205) keep them beside the mesh.
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

}  // namespace orc
