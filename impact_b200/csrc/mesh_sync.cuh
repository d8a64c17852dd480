// Host-side tables of a mesh that is kept in sync with its voxel object (mesh_sync.cu) and of the collision probes that
// follow it (probes.cu): ChunkSubmeshManager (mesh.rs:703-848) and RangeAllocator
// (impact_containers/src/range_allocator.rs). Internal.
#pragma once
#include <unordered_map>

#include "api_internal.cuh"

namespace ivx_ranges {

// the free ranges [first, second) sorted by start, and an upper bound of the longest one: most requests that cannot be
// served (a chunk mesh larger than every hole) are answered without looking at the list
struct Ranges : std::vector<std::pair<uint32_t, uint32_t>> {
    uint32_t longest_at_most = 0;
    void clear() {
        std::vector<std::pair<uint32_t, uint32_t>>::clear();
        longest_at_most = 0;
    }
};

// RangeAllocator::free_range: a range whose start is already free stays as it is (BTreeSet::insert)
inline void release_range(Ranges& fr, uint32_t a, uint32_t b) {
    if (a >= b) return;
    auto it = std::lower_bound(fr.begin(), fr.end(), a, [](const std::pair<uint32_t, uint32_t>& r, uint32_t v) { return r.first < v; });
    if (it != fr.end() && it->first == a) return;
    fr.insert(it, {a, b});
    fr.longest_at_most = std::max(fr.longest_at_most, b - a);
}
// RangeAllocator::allocate_range: the smallest free range that fits (the lowest one of equals), its tail stays free
inline bool take_range(Ranges& fr, uint32_t len, uint32_t& start) {
    if (len > fr.longest_at_most) return false;
    size_t best = fr.size();
    uint32_t best_len = 0xFFFFFFFFu, longest = 0;
    for (size_t q = 0; q < fr.size(); ++q) {
        const uint32_t l = fr[q].second - fr[q].first;
        longest = std::max(longest, l);
        if (l >= len && l < best_len) {
            best = q;
            best_len = l;
            if (l == len) break;  // the lowest exact fit: nothing later can be better
        }
    }
    if (best == fr.size()) {
        fr.longest_at_most = longest;  // (the whole list was seen)
        return false;
    }
    start = fr[best].first;
    if (best_len == len) fr.erase(fr.begin() + best);
    else fr[best].first += len;
    return true;
}
// RangeAllocator::merge_consecutive_ranges
inline void coalesce(Ranges& fr) {
    size_t w = 0;
    uint32_t longest = 0;
    for (size_t q = 0; q < fr.size(); ++q) {
        if (w > 0 && fr[w - 1].second == fr[q].first) fr[w - 1].second = fr[q].second;
        else fr[w++] = fr[q];
        longest = std::max(longest, fr[w - 1].second - fr[w - 1].first);
    }
    fr.resize(w);
    fr.longest_at_most = longest;
}

}  // namespace ivx_ranges

struct ivx_mesh_sync {
    // KeyIndexMapper<[usize; 3]>: table row of a chunk (by linear chunk index) and the chunk of a row
    std::unordered_map<uint32_t, uint32_t> row_of_chunk;
    std::vector<uint32_t> chunk_of_row;
    std::vector<ivx_chunk_submesh> submeshes;  // chunk_submeshes
    std::vector<uint32_t> vertex_ranges;       // chunk_vertex_ranges, 2 words per row
    // RangeAllocator x 2: free ranges [first, second) sorted by start
    ivx_ranges::Ranges free_vertices, free_indices;
    std::vector<uint32_t> updated;  // ChunkSubmeshDataRanges: vertex start, end, index start, end
    bool chunks_were_removed = false;
    uint32_t n_vertices = 0, n_indices = 0;  // lengths of the buffers (holes included)
    std::vector<uint32_t> touched_rows;      // rows written or moved by the current sync (for the device mirror)
    std::vector<uint32_t> last_dirty;        // the invalidated chunks of the last sync, in the order they were taken
                                             // (sync_mesh_with_object hands the same set to the collision probes)
};
