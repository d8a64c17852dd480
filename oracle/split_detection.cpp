// ORACLE — test infrastructure only; never linked into the product library.
//
// CPU restatement of the connected-region ("split") detection of
// engine/crates/impact_voxel/src/object/split_detection.rs:
//   local_regions_for_chunk        update_local_connected_regions_within_occupied_ranges_for_chunk (:662-893),
//                                  full occupied ranges, with the union-find of :1776-1884
//   face pair enumeration          the connection updaters (:1046-1326, :1424-1463) reduced to the SET of
//                                  (region, adjacent region) pairs they record per chunk face
//   resolve_connected_regions      resolve_connected_regions_between_all_chunks (:323-488) + set_root_for_region
//                                  (:1914-1942): same chunk / region visiting order, so the same local region ends
//                                  up as the root (representative) of every global region
//   find_two_disconnected_regions  :193-249; count_regions :255-301
//   smallest-region choice         extraction.rs:121-281 (fewest non-uniform chunks, then fewest chunks, else the
//                                  second region)
// Not restated: the per-region connection-slot limit (:1519-1546), which only drops connections for pathological
// chunks with more regions than slots; the oracle reports such inputs via `overflow` instead.
#include <algorithm>
#include <cstring>

#include "oracle.hpp"

namespace orc {

namespace {

inline int vlin(int i, int j, int k) { return (i << 8) | (j << 4) | k; }

uint16_t find_root(uint16_t* parents, int idx) {
    int r = idx;
    while (parents[r] != r) r = parents[r];
    while (parents[idx] != r) {  // path compression (does not change which voxel is the root)
        int n = parents[idx];
        parents[idx] = (uint16_t)r;
        idx = n;
    }
    return (uint16_t)r;
}

inline void assign_parent(uint16_t* parents, int voxel, int parent) {
    int r = find_root(parents, voxel);
    if (r != parent) parents[r] = (uint16_t)parent;
}

}  // namespace

// labels: 4096 bytes out (255 = empty). Returns false if the chunk has more regions than the reference supports
// (its asserts at :798 and :835 would panic).
bool local_regions_for_chunk(const Voxel* voxels, bool only_empty, uint8_t* labels, uint16_t* boundary_region_count,
                             uint16_t* region_count) {
    static thread_local uint16_t parents[4096];
    for (int i = 0; i < 4096; ++i) parents[i] = (uint16_t)i;
    const auto is_empty = [&](int idx) { return (voxels[idx].flags & 1) != 0; };

    if (!only_empty) {
        for (int i = 0; i < 16; ++i)
            for (int j = 0; j < 16; ++j)
                for (int k = 0; k < 16; ++k) {
                    const int idx = vlin(i, j, k);
                    if (is_empty(idx)) continue;
                    const int root = find_root(parents, idx);
                    const uint8_t f = voxels[idx].flags;
                    if (i < 15 && (f & (1 << 5))) assign_parent(parents, vlin(i + 1, j, k), root);
                    if (j < 15 && (f & (1 << 6))) assign_parent(parents, vlin(i, j + 1, k), root);
                    if (k < 15 && (f & (1 << 7))) assign_parent(parents, vlin(i, j, k + 1), root);
                }
    }

    int current = 0;
    const int MAX_BOUNDARY_LABEL = 255, MAX_LABEL = 255;
    const auto visit_boundary = [&](int i, int j, int k) {
        const int idx = vlin(i, j, k);
        if (is_empty(idx)) {
            labels[idx] = 255;
            return;
        }
        const int set_id = find_root(parents, idx);
        bool take = set_id == idx;
        if (!take) {
            const int si = set_id >> 8, sj = (set_id >> 4) & 15, sk = set_id & 15;
            const bool interior = si > 0 && si < 15 && sj > 0 && sj < 15 && sk > 0 && sk < 15;
            if (interior) {
                parents[set_id] = (uint16_t)idx;  // make_voxel_root
                parents[idx] = (uint16_t)idx;
                take = true;
            }
        }
        if (take) {
            labels[idx] = (uint8_t)current;
            current = std::min(current + 1, MAX_BOUNDARY_LABEL);
        }
    };
    // LoopForChunkVoxels::over_full_boundary (utils.rs:222-300) with Loop3::execute's loop nests (:470-545)
    for (int s = 0; s < 2; ++s)
        for (int j = 0; j < 16; ++j)
            for (int k = 0; k < 16; ++k) visit_boundary(s ? 15 : 0, j, k);
    for (int s = 0; s < 2; ++s)
        for (int i = 1; i < 15; ++i)
            for (int k = 0; k < 16; ++k) visit_boundary(i, s ? 15 : 0, k);
    for (int s = 0; s < 2; ++s)
        for (int i = 1; i < 15; ++i)
            for (int j = 1; j < 15; ++j) visit_boundary(i, j, s ? 15 : 0);
    if (!(current < MAX_BOUNDARY_LABEL)) return false;
    *boundary_region_count = (uint16_t)current;

    for (int i = 1; i < 15; ++i)
        for (int j = 1; j < 15; ++j)
            for (int k = 1; k < 15; ++k) {
                const int idx = vlin(i, j, k);
                if (parents[idx] != idx) continue;
                if (!is_empty(idx)) {
                    labels[idx] = (uint8_t)current;
                    current = std::min(current + 1, MAX_LABEL);
                } else {
                    labels[idx] = 255;
                }
            }
    if (!(current < MAX_LABEL)) return false;
    *region_count = (uint16_t)current;

    for (int idx = 0; idx < 4096; ++idx) {
        if (is_empty(idx)) continue;
        int r = idx;
        while (parents[r] != r) r = parents[r];
        if (r != idx) labels[idx] = labels[r];
    }
    return true;
}

namespace {

struct Resolver {
    std::vector<uint32_t>& parent;  // per global region index: GlobalRegionLabel of the parent
    const std::vector<ChunkRegions>& per_chunk;
    uint32_t index_of(uint32_t label) const { return per_chunk[label >> 8].first_region + (label & 255u); }
    uint32_t find(uint32_t label) {
        uint32_t r = label;
        while (parent[index_of(r)] != r) r = parent[index_of(r)];
        while (parent[index_of(label)] != r) {
            uint32_t n = parent[index_of(label)];
            parent[index_of(label)] = r;
            label = n;
        }
        return r;
    }
};

}  // namespace

void resolve_connected_regions(const Object& obj, SplitDetection& sd) {
    const uint32_t n = (uint32_t)obj.chunks.size();
    sd.overflow = false;
    sd.per_chunk.assign(n, ChunkRegions{0, 0, 0});
    uint32_t n_nu = 0;
    for (const Chunk& c : obj.chunks)
        if (c.kind == CK_NONUNIFORM) n_nu = std::max(n_nu, c.data_offset + 1);
    sd.voxel_labels.assign((size_t)n_nu * 4096, 255);

    // ---- local regions ----
    uint32_t total = 0;
    for (uint32_t c = 0; c < n; ++c) {
        const Chunk& ch = obj.chunks[c];
        ChunkRegions& cr = sd.per_chunk[c];
        cr.first_region = total;
        if (ch.kind == CK_UNIFORM) {
            cr.region_count = cr.boundary_region_count = 1;
        } else if (ch.kind == CK_NONUNIFORM) {
            if (!local_regions_for_chunk(obj.chunk_voxels(ch.data_offset), (ch.flags & CF_ONLY_EMPTY) != 0,
                                         sd.voxel_labels.data() + (size_t)ch.data_offset * 4096, &cr.boundary_region_count,
                                         &cr.region_count))
                sd.overflow = true;
        }
        total += cr.region_count;
    }
    sd.region_parent.assign(total, 0);
    for (uint32_t c = 0; c < n; ++c)
        for (uint32_t r = 0; r < sd.per_chunk[c].region_count; ++r) sd.region_parent[sd.per_chunk[c].first_region + r] = (c << 8) | r;

    // ---- connections across chunk faces: the set of (region, adjacent region) pairs, both directions ----
    std::vector<std::vector<uint32_t>> adj(total);  // adjacent GlobalRegionLabels per region
    const auto connect = [&](uint32_t ca, uint32_t ra, uint32_t cb, uint32_t rb) {
        auto& la = adj[sd.per_chunk[ca].first_region + ra];
        const uint32_t lb_label = (cb << 8) | rb;
        if (std::find(la.begin(), la.end(), lb_label) == la.end()) {
            la.push_back(lb_label);
            adj[sd.per_chunk[cb].first_region + rb].push_back((ca << 8) | ra);
        }
    };
    const uint32_t* cc = obj.chunk_counts;
    for (uint32_t i = 0; i < cc[0]; ++i)
        for (uint32_t j = 0; j < cc[1]; ++j)
            for (uint32_t k = 0; k < cc[2]; ++k) {
                const uint32_t c = obj.lin(i, j, k);
                const Chunk& lo = obj.chunks[c];
                if (lo.kind == CK_VOID) continue;
                for (int d = 0; d < 3; ++d) {
                    uint32_t u[3] = {i, j, k};
                    if (++u[d] >= cc[d]) continue;
                    const uint32_t cu = obj.lin(u[0], u[1], u[2]);
                    const Chunk& up = obj.chunks[cu];
                    if (up.kind == CK_VOID) continue;
                    if (lo.kind == CK_UNIFORM && up.kind == CK_UNIFORM) {
                        connect(c, 0, cu, 0);
                        continue;
                    }
                    const uint8_t* ll = lo.kind == CK_NONUNIFORM ? sd.voxel_labels.data() + (size_t)lo.data_offset * 4096 : nullptr;
                    const uint8_t* lu = up.kind == CK_NONUNIFORM ? sd.voxel_labels.data() + (size_t)up.data_offset * 4096 : nullptr;
                    for (int a = 0; a < 16; ++a)
                        for (int b = 0; b < 16; ++b) {
                            int pl[3], pu[3];
                            pl[d] = 15;
                            pu[d] = 0;
                            pl[(d + 1) % 3] = pu[(d + 1) % 3] = a;
                            pl[(d + 2) % 3] = pu[(d + 2) % 3] = b;
                            const uint32_t la = ll ? ll[vlin(pl[0], pl[1], pl[2])] : 0u;
                            const uint32_t lb = lu ? lu[vlin(pu[0], pu[1], pu[2])] : 0u;
                            if (la == 255u || lb == 255u) continue;
                            connect(c, la, cu, lb);
                        }
                }
            }
    // the reference gives every boundary region 256 / boundary_region_count connection slots and overwrites the
    // last slot beyond that (:1519-1546); flag inputs where that would lose a connection
    for (uint32_t c = 0; c < n; ++c) {
        const ChunkRegions& cr = sd.per_chunk[c];
        const uint32_t cap = obj.chunks[c].kind == CK_UNIFORM ? 256u : 256u / std::max<uint32_t>(1u, cr.boundary_region_count);
        for (uint32_t r = 0; r < cr.region_count; ++r)
            if (adj[cr.first_region + r].size() > cap) sd.overflow = true;
    }

    // ---- global resolution, in the reference's visiting order ----
    Resolver R{sd.region_parent, sd.per_chunk};
    for (uint32_t i = obj.occ_chunks[0][0]; i < obj.occ_chunks[0][1]; ++i)
        for (uint32_t j = obj.occ_chunks[1][0]; j < obj.occ_chunks[1][1]; ++j)
            for (uint32_t k = obj.occ_chunks[2][0]; k < obj.occ_chunks[2][1]; ++k) {
                const uint32_t c = obj.lin(i, j, k);
                const ChunkRegions& cr = sd.per_chunk[c];
                for (uint32_t r = 0; r < cr.boundary_region_count; ++r) {
                    const uint32_t root = R.find((c << 8) | r);
                    for (uint32_t other : adj[cr.first_region + r]) {
                        const uint32_t oroot = R.find(other);
                        if (oroot != root) sd.region_parent[R.index_of(oroot)] = root;
                    }
                }
            }
    sd.region_root.resize(total);
    for (uint32_t c = 0; c < n; ++c)
        for (uint32_t r = 0; r < sd.per_chunk[c].region_count; ++r)
            sd.region_root[sd.per_chunk[c].first_region + r] = R.find((c << 8) | r);

    // ---- count_regions / find_two_disconnected_regions (occupied chunk ranges, chunk then region order) ----
    sd.n_regions = 0;
    sd.has_two = false;
    sd.two[0] = sd.two[1] = 0;
    for (uint32_t i = obj.occ_chunks[0][0]; i < obj.occ_chunks[0][1]; ++i)
        for (uint32_t j = obj.occ_chunks[1][0]; j < obj.occ_chunks[1][1]; ++j)
            for (uint32_t k = obj.occ_chunks[2][0]; k < obj.occ_chunks[2][1]; ++k) {
                const uint32_t c = obj.lin(i, j, k);
                const ChunkRegions& cr = sd.per_chunk[c];
                for (uint32_t r = 0; r < cr.region_count; ++r)
                    if (sd.region_root[cr.first_region + r] == ((c << 8) | r)) {
                        if (sd.n_regions < 2) sd.two[sd.n_regions] = (c << 8) | r;
                        sd.n_regions++;
                    }
            }
    sd.has_two = sd.n_regions >= 2;

    // ---- which of the two would be extracted (extraction.rs:121-281) ----
    std::memset(sd.stats, 0, sizeof(sd.stats));
    sd.smallest = 0;
    if (!sd.has_two) return;
    for (int q = 0; q < 2; ++q)
        for (int d = 0; d < 3; ++d) {
            sd.stats[q].chunk_min[d] = 0xFFFFFFFFu;
            sd.stats[q].chunk_max[d] = 0;
        }
    for (uint32_t i = obj.occ_chunks[0][0]; i < obj.occ_chunks[0][1]; ++i)
        for (uint32_t j = obj.occ_chunks[1][0]; j < obj.occ_chunks[1][1]; ++j)
            for (uint32_t k = obj.occ_chunks[2][0]; k < obj.occ_chunks[2][1]; ++k) {
                const uint32_t c = obj.lin(i, j, k);
                const ChunkRegions& cr = sd.per_chunk[c];
                const uint32_t idx3[3] = {i, j, k};
                bool found[2] = {false, false};
                for (uint32_t r = 0; r < cr.region_count; ++r) {
                    const uint32_t root = sd.region_root[cr.first_region + r];
                    int q;
                    if (root == sd.two[0] && !found[0]) q = 0;
                    else if (root == sd.two[1] && !found[1]) q = 1;
                    else continue;
                    found[q] = true;
                    sd.stats[q].chunk_count++;
                    if (obj.chunks[c].kind == CK_NONUNIFORM) sd.stats[q].non_uniform_chunk_count++;
                    for (int d = 0; d < 3; ++d) {
                        sd.stats[q].chunk_min[d] = std::min(sd.stats[q].chunk_min[d], idx3[d]);
                        sd.stats[q].chunk_max[d] = std::max(sd.stats[q].chunk_max[d], idx3[d]);
                    }
                }
            }
    if (sd.stats[0].non_uniform_chunk_count < sd.stats[1].non_uniform_chunk_count) sd.smallest = 0;
    else if (sd.stats[0].non_uniform_chunk_count > sd.stats[1].non_uniform_chunk_count) sd.smallest = 1;
    else sd.smallest = sd.stats[0].chunk_count < sd.stats[1].chunk_count ? 0 : 1;
}

// count_regions_brute_force (:498-560): plain 6-connected flood fill over all non-empty voxels of the object
uint32_t count_regions_brute_force(const Object& obj) {
    const uint32_t* cc = obj.chunk_counts;
    const uint32_t nx = cc[0] * 16, ny = cc[1] * 16, nz = cc[2] * 16;
    std::vector<uint8_t> filled((size_t)nx * ny * nz, 0);
    for (uint32_t i = 0; i < cc[0]; ++i)
        for (uint32_t j = 0; j < cc[1]; ++j)
            for (uint32_t k = 0; k < cc[2]; ++k) {
                const Chunk& c = obj.chunks[obj.lin(i, j, k)];
                if (c.kind == CK_VOID) continue;
                const Voxel* v = c.kind == CK_NONUNIFORM ? obj.chunk_voxels(c.data_offset) : nullptr;
                for (int a = 0; a < 16; ++a)
                    for (int b = 0; b < 16; ++b)
                        for (int d = 0; d < 16; ++d) {
                            const bool ne = v ? (v[vlin(a, b, d)].flags & 1) == 0 : true;
                            if (ne) filled[((size_t)(i * 16 + a) * ny + (j * 16 + b)) * nz + (k * 16 + d)] = 1;
                        }
            }
    uint32_t count = 0;
    std::vector<size_t> stack;
    for (size_t s = 0; s < filled.size(); ++s) {
        if (filled[s] != 1) continue;
        count++;
        filled[s] = 2;
        stack.push_back(s);
        while (!stack.empty()) {
            const size_t p = stack.back();
            stack.pop_back();
            const size_t z = p % nz, y = (p / nz) % ny, x = p / ((size_t)nz * ny);
            const auto push = [&](size_t q) {
                if (filled[q] == 1) {
                    filled[q] = 2;
                    stack.push_back(q);
                }
            };
            if (x > 0) push(p - (size_t)ny * nz);
            if (x + 1 < nx) push(p + (size_t)ny * nz);
            if (y > 0) push(p - nz);
            if (y + 1 < ny) push(p + nz);
            if (z > 0) push(p - 1);
            if (z + 1 < nz) push(p + 1);
        }
    }
    return count;
}

}  // namespace orc
