// The object mesh kept in sync with a modified voxel object (sm_100a): `VoxelObjectMesh::sync_with_voxel_object`
// (mesh.rs:360-456) with its `ChunkSubmeshManager` (mesh.rs:703-848) and `RangeAllocator`
// (impact_containers/src/range_allocator.rs).
//
// The mesh created by ivx_object_mesh stays on the device. After a modification only the invalidated chunks are
// re-meshed, and each chunk's vertices / indices go where the reference puts them: into the smallest free range that
// fits (ranges freed by chunks that were re-meshed or removed earlier), else at the end of the buffers. Which chunk goes
// where is decided on the host from the chunk meshes' SIZES (one small read-back of the counting pass); the emit pass
// then writes every chunk straight into its place in the persistent buffers, with its indices already offset by its
// vertex range. The submesh table, the vertex ranges and the list of updated ranges (what a renderer has to re-upload,
// `VoxelMeshModifications`, mesh.rs:105-118) are the manager's tables; the table is mirrored to the device after a sync.
//
// The reference walks a HashSet of invalidated chunks, i.e. in no defined order; this implementation walks them in
// ascending linear chunk index, which is one of the orders the reference may take.
#include "mesh_sync.cuh"

namespace {

using namespace ivx_ranges;

// remove_chunk_if_present (mesh.rs:814-826): swap-remove of the row, its ranges become free
void drop_chunk(ivx_mesh_sync& s, uint32_t chunk) {
    auto it = s.row_of_chunk.find(chunk);
    if (it == s.row_of_chunk.end()) return;
    const uint32_t row = it->second, last = (uint32_t)s.chunk_of_row.size() - 1;
    s.row_of_chunk.erase(it);
    release_range(s.free_vertices, s.vertex_ranges[2 * row], s.vertex_ranges[2 * row + 1]);
    release_range(s.free_indices, s.submeshes[row].index_offset, s.submeshes[row].index_offset + s.submeshes[row].index_count);
    if (row != last) {
        s.touched_rows.push_back(row);
        s.chunk_of_row[row] = s.chunk_of_row[last];
        s.row_of_chunk[s.chunk_of_row[row]] = row;
        s.submeshes[row] = s.submeshes[last];
        s.vertex_ranges[2 * row] = s.vertex_ranges[2 * last];
        s.vertex_ranges[2 * row + 1] = s.vertex_ranges[2 * last + 1];
    }
    s.chunk_of_row.pop_back();
    s.submeshes.pop_back();
    s.vertex_ranges.resize(2 * (size_t)last);
    s.chunks_were_removed = true;
}

// rows of the manager's tables that changed → their place in the device mirror (15 words each: submesh + vertex range)
__global__ void k_scatter_rows(const uint32_t* __restrict__ packed, uint32_t n_rows, ivx_chunk_submesh* __restrict__ submeshes,
                               uint32_t* __restrict__ vertex_ranges) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q = t / 16u, wi = t % 16u;
    if (q >= n_rows) return;
    const uint32_t* src = packed + (size_t)q * 16u;
    const uint32_t row = src[0];
    if (wi >= 1u && wi <= 13u) reinterpret_cast<uint32_t*>(&submeshes[row])[wi - 1u] = src[wi];
    else if (wi >= 14u) vertex_ranges[2 * (size_t)row + (wi - 14u)] = src[wi];
}

template <typename T>
int grow(ivx_ctx* ctx, T*& buf, size_t have_elems, size_t keep_elems, size_t want_elems, size_t elem_bytes) {
    if (want_elems <= have_elems && buf) return IVX_OK;
    T* nb = static_cast<T*>(ctx->alloc(std::max<size_t>(1, want_elems) * elem_bytes));
    if (!nb) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh growth: out of device memory");
    if (buf && keep_elems) CU(ctx, cudaMemcpyAsync(nb, buf, keep_elems * elem_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->release(buf);  // (reused by later work on this stream only)
    buf = nb;
    return IVX_OK;
}

}  // namespace

void ivx_mesh_sync_free(ivx_mesh_sync* s) { delete s; }

extern "C" {

int ivx_object_mesh_sync(ivx_ctx* ctx, ivx_object* obj, ivx_mesh_info* out) {
    if (!ctx || !obj || !out) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    std::memset(out, 0, sizeof(*out));
    const uint32_t n = obj->n_chunks;
    if (n == 0) return IVX_OK;
    if (obj->derive_pending) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "slab object: call ivx_object_slab_finalize first");
    if (obj->first_i != 0 || obj->nb[0] != obj->chunk_counts[0])
        IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "the synced mesh is kept for whole objects");
    if (obj->mesh_is_patch || (!obj->mesh.positions && !obj->sync))
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "the object has no full mesh to keep in sync: call ivx_object_mesh first "
                 "(ivx_object_remesh_dirty replaces it by a patch)");
    cudaStream_t st = ctx->stream;
    DeviceMesh& m = obj->mesh;

    // ---- the manager's tables, from the mesh ivx_object_mesh created (VoxelObjectMesh::recreate pushes the chunks) ----
    if (!obj->sync) {
        ivx_mesh_sync* s = new (std::nothrow) ivx_mesh_sync();
        if (!s) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "host allocation failed");
        obj->sync = s;
        s->n_vertices = m.n_vertices;
        s->n_indices = m.n_indices;
        s->submeshes.resize(m.n_submeshes);
        s->vertex_ranges.resize(2 * (size_t)m.n_submeshes);
        if (m.n_submeshes) {
            CU(ctx, cudaMemcpyAsync(s->submeshes.data(), m.submeshes, (size_t)m.n_submeshes * sizeof(ivx_chunk_submesh),
                                    cudaMemcpyDeviceToHost, st));
            CU(ctx, cudaMemcpyAsync(s->vertex_ranges.data(), m.vertex_ranges, (size_t)m.n_submeshes * 8, cudaMemcpyDeviceToHost, st));
            CU(ctx, cudaStreamSynchronize(st));
        }
        s->chunk_of_row.resize(m.n_submeshes);
        for (uint32_t r = 0; r < m.n_submeshes; ++r) {
            const uint32_t* ci = s->submeshes[r].chunk_indices;
            const uint32_t c = (ci[0] * obj->nb[1] + ci[1]) * obj->nb[2] + ci[2];
            s->chunk_of_row[r] = c;
            s->row_of_chunk[c] = r;
        }
        if (m.cap_vertices == 0) {
            m.cap_vertices = m.n_vertices;
            m.cap_indices = m.n_indices;
            m.cap_submeshes = m.n_submeshes;
        }
    }
    ivx_mesh_sync& s = *obj->sync;

    // ---- invalidated chunks (ascending), the exposed ones among them, sizes of their meshes ----
    Tmp tmp(ctx);
    uint32_t* exposed = tmp.get<uint32_t>(n);
    uint32_t* dflag = tmp.get<uint32_t>(n);
    uint32_t* scan = tmp.get<uint32_t>(n);
    uint32_t* work = tmp.get<uint32_t>(n);
    uint32_t* dlist = tmp.get<uint32_t>(n);
    if (!exposed || !dflag || !scan || !work || !dlist) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh sync: out of device memory");
    uint32_t* counters = ctx->d_scratch;
    KL(ctx, launch_flag_dirty_exposed(obj->d_chunks, obj->d_dirty, n, exposed, dflag, st));
    KL(ctx, launch_exclusive_scan(dflag, scan, n, counters + 21, st));
    KL(ctx, launch_scatter_active(dflag, scan, n, dlist, st));
    KL(ctx, launch_exclusive_scan(exposed, scan, n, counters + 16, st));
    KL(ctx, launch_scatter_active(exposed, scan, n, work, st));
    uint32_t words[8];
    if (int rc = ivx_read_words(ctx, counters + 16, 6, words)) return rc;
    const uint32_t n_work = words[0], n_dirty = words[5];
    auto finish = [&]() {
        fill_mesh_info_from(m, out);
        return IVX_OK;
    };
    s.last_dirty.clear();
    if (n_dirty == 0) return finish();
    std::vector<uint32_t> h_dirty(n_dirty), h_work(n_work), h_vc(n_work), h_ic(n_work);
    uint32_t* vcount = tmp.get<uint32_t>(std::max(1u, n_work));
    uint32_t* icount = tmp.get<uint32_t>(std::max(1u, n_work));
    uint32_t* hsub = tmp.get<uint32_t>(std::max(1u, n_work));
    uint32_t* voff = tmp.get<uint32_t>(std::max(1u, n_work));
    uint32_t* ioff = tmp.get<uint32_t>(std::max(1u, n_work));
    uint32_t* sord = tmp.get<uint32_t>(std::max(1u, n_work));
    ivx_chunk_submesh* scratch_sub = tmp.get<ivx_chunk_submesh>(std::max(1u, n_work));
    uint32_t* scratch_vr = tmp.get<uint32_t>(2 * (size_t)std::max(1u, n_work));
    if (!vcount || !icount || !hsub || !voff || !ioff || !sord || !scratch_sub || !scratch_vr)
        IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh sync: out of device memory");
    MeshArgs ma{};
    ma.chunks = obj->d_chunks;
    ma.voxels = obj->d_voxels;
    for (int d = 0; d < 3; ++d) ma.nb[d] = obj->nb[d];
    ma.first_i = obj->first_i;
    ma.voxel_extent = obj->voxel_extent;
    ma.work = work;
    ma.n_work = n_work;
    ma.vertex_count = vcount;
    ma.index_count = icount;
    ma.has_submesh = hsub;
    const uint32_t grid = ivx_persistent_grid(ctx, std::max(1u, n_work), 4);
    if (n_work) KLP(ctx, 4, launch_mesh(false, ma, grid, st));
    CU(ctx, cudaMemcpyAsync(h_dirty.data(), dlist, (size_t)n_dirty * 4, cudaMemcpyDeviceToHost, st));
    if (n_work) {
        CU(ctx, cudaMemcpyAsync(h_work.data(), work, (size_t)n_work * 4, cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaMemcpyAsync(h_vc.data(), vcount, (size_t)n_work * 4, cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaMemcpyAsync(h_ic.data(), icount, (size_t)n_work * 4, cudaMemcpyDeviceToHost, st));
    }
    CU(ctx, cudaStreamSynchronize(st));

    s.last_dirty = h_dirty;
    // ---- placement, chunk by chunk (write_chunk / remove_chunk_if_present, mesh.rs:749-826) ----
    std::vector<uint32_t> h_voff(n_work, 0u), h_ioff(n_work, 0u), row_of_work(n_work, 0xFFFFFFFFu);
    size_t w = 0;
    for (uint32_t q = 0; q < n_dirty; ++q) {
        const uint32_t chunk = h_dirty[q];
        const bool is_work = w < n_work && h_work[w] == chunk;
        const size_t wi = w;
        if (is_work) ++w;
        if (!is_work || h_ic[wi] == 0u) {  // no longer exposed, or exposed with an empty mesh
            drop_chunk(s, chunk);
            continue;
        }
        const uint32_t vc = h_vc[wi], ic = h_ic[wi];
        auto found = s.row_of_chunk.find(chunk);
        if (found != s.row_of_chunk.end()) {
            const uint32_t row = found->second;
            release_range(s.free_vertices, s.vertex_ranges[2 * row], s.vertex_ranges[2 * row + 1]);
            release_range(s.free_indices, s.submeshes[row].index_offset, s.submeshes[row].index_offset + s.submeshes[row].index_count);
        }
        uint32_t v0 = s.n_vertices, i0 = s.n_indices;
        if (!take_range(s.free_vertices, vc, v0)) {
            v0 = s.n_vertices;
            s.n_vertices += vc;
        }
        if (!take_range(s.free_indices, ic, i0)) {
            i0 = s.n_indices;
            s.n_indices += ic;
        }
        uint32_t row;
        if (found != s.row_of_chunk.end()) {
            row = found->second;
        } else {
            row = (uint32_t)s.chunk_of_row.size();
            s.row_of_chunk[chunk] = row;
            s.chunk_of_row.push_back(chunk);
            s.submeshes.emplace_back();
            s.vertex_ranges.resize(s.vertex_ranges.size() + 2);
        }
        s.vertex_ranges[2 * row] = v0;
        s.vertex_ranges[2 * row + 1] = v0 + vc;
        s.touched_rows.push_back(row);
        s.updated.insert(s.updated.end(), {v0, v0 + vc, i0, i0 + ic});
        h_voff[wi] = v0;
        h_ioff[wi] = i0;
        row_of_work[wi] = row;
    }
    // rows move when a later chunk is removed (swap-remove): look the rows up again at the end
    for (size_t wi = 0; wi < n_work; ++wi)
        if (row_of_work[wi] != 0xFFFFFFFFu) row_of_work[wi] = s.row_of_chunk.count(h_work[wi]) ? s.row_of_chunk[h_work[wi]] : 0xFFFFFFFFu;

    // ---- the data: every re-meshed chunk straight into its place ----
    if (int rc = grow(ctx, m.positions, m.cap_vertices, m.n_vertices, s.n_vertices > m.cap_vertices ? s.n_vertices + s.n_vertices / 4 : m.cap_vertices, 12)) return rc;
    if (int rc = grow(ctx, m.normals, m.cap_vertices, m.n_vertices, s.n_vertices > m.cap_vertices ? s.n_vertices + s.n_vertices / 4 : m.cap_vertices, 12)) return rc;
    if (s.n_vertices > m.cap_vertices) m.cap_vertices = s.n_vertices + s.n_vertices / 4;
    if (int rc = grow(ctx, m.indices, m.cap_indices, m.n_indices, s.n_indices > m.cap_indices ? s.n_indices + s.n_indices / 4 : m.cap_indices, 4)) return rc;
    if (int rc = grow(ctx, m.index_materials, m.cap_indices, m.n_indices, s.n_indices > m.cap_indices ? s.n_indices + s.n_indices / 4 : m.cap_indices, 8)) return rc;
    if (s.n_indices > m.cap_indices) m.cap_indices = s.n_indices + s.n_indices / 4;
    if (n_work) {
        std::vector<uint32_t> h_sord(n_work);
        for (uint32_t q = 0; q < n_work; ++q) h_sord[q] = q;  // scratch rows: one per work item
        CU(ctx, cudaMemcpyAsync(voff, h_voff.data(), (size_t)n_work * 4, cudaMemcpyHostToDevice, st));
        CU(ctx, cudaMemcpyAsync(ioff, h_ioff.data(), (size_t)n_work * 4, cudaMemcpyHostToDevice, st));
        CU(ctx, cudaMemcpyAsync(sord, h_sord.data(), (size_t)n_work * 4, cudaMemcpyHostToDevice, st));
        ma.vertex_offset = voff;
        ma.index_offset = ioff;
        ma.submesh_ord = sord;
        ma.positions = m.positions;
        ma.normals = m.normals;
        ma.indices = m.indices;
        ma.index_materials = m.index_materials;
        ma.submeshes = scratch_sub;
        ma.vertex_ranges = scratch_vr;
        uint64_t quads = 0;
        for (uint32_t q = 0; q < n_work; ++q) quads += h_ic[q] / 6u;
        ma.mq_capacity = (uint32_t)std::min<uint64_t>(quads, 0xFFFFFFFFu);
        ma.mq_entries = tmp.get<uint4>(std::max<size_t>(1, (size_t)ma.mq_capacity * 3));
        ma.mq_count = counters + 20;
        if (!ma.mq_entries) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh sync: out of device memory");
        CU(ctx, cudaMemsetAsync(ma.mq_count, 0, 4, st));
        KLP(ctx, 5, launch_mesh(true, ma, grid, st));
        KLP(ctx, 5, launch_mesh_materials(ma.mq_entries, ma.mq_count, ma.mq_capacity, m.index_materials, st));
        // the submesh rows the kernel made (chunk indices, index range, obscuredness table) → the manager's table
        std::vector<ivx_chunk_submesh> h_sub(n_work);
        CU(ctx, cudaMemcpyAsync(h_sub.data(), scratch_sub, (size_t)n_work * sizeof(ivx_chunk_submesh), cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaStreamSynchronize(st));
        for (size_t wi = 0; wi < n_work; ++wi)
            if (row_of_work[wi] != 0xFFFFFFFFu) s.submeshes[row_of_work[wi]] = h_sub[wi];
    }
    // perform_maintainance (mesh.rs:828-831)
    coalesce(s.free_vertices);
    coalesce(s.free_indices);

    // ---- mirror of the tables on the device: only the rows this sync wrote or moved ----
    const uint32_t rows = (uint32_t)s.submeshes.size();
    if (rows > m.cap_submeshes || !m.submeshes) {
        const uint32_t want = rows + rows / 4 + 16;
        ivx_chunk_submesh* ns = static_cast<ivx_chunk_submesh*>(ctx->alloc((size_t)want * sizeof(ivx_chunk_submesh)));
        uint32_t* nv = static_cast<uint32_t*>(ctx->alloc((size_t)want * 8));
        if (!ns || !nv) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh sync: out of device memory");
        const uint32_t keep = std::min(m.n_submeshes, rows);
        if (m.submeshes && keep) {
            CU(ctx, cudaMemcpyAsync(ns, m.submeshes, (size_t)keep * sizeof(ivx_chunk_submesh), cudaMemcpyDeviceToDevice, st));
            CU(ctx, cudaMemcpyAsync(nv, m.vertex_ranges, (size_t)keep * 8, cudaMemcpyDeviceToDevice, st));
        }
        ctx->release(m.submeshes);
        ctx->release(m.vertex_ranges);
        m.submeshes = ns;
        m.vertex_ranges = nv;
        m.cap_submeshes = want;
    }
    std::vector<uint32_t> packed;  // (read by an asynchronous copy: lives until the synchronisation at the end)
    {
        std::sort(s.touched_rows.begin(), s.touched_rows.end());
        s.touched_rows.erase(std::unique(s.touched_rows.begin(), s.touched_rows.end()), s.touched_rows.end());
        packed.reserve(s.touched_rows.size() * 16);
        for (uint32_t row : s.touched_rows) {
            if (row >= rows) continue;  // the row was swap-removed later in this sync
            packed.push_back(row);
            const uint32_t* sw = reinterpret_cast<const uint32_t*>(&s.submeshes[row]);
            packed.insert(packed.end(), sw, sw + 13);
            packed.push_back(s.vertex_ranges[2 * row]);
            packed.push_back(s.vertex_ranges[2 * row + 1]);
        }
        s.touched_rows.clear();
        const uint32_t n_rows = (uint32_t)(packed.size() / 16);
        if (n_rows) {
            uint32_t* d_packed = tmp.get<uint32_t>(packed.size());
            if (!d_packed) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh sync: out of device memory");
            CU(ctx, cudaMemcpyAsync(d_packed, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice, st));
            ctx->launches++;
            k_scatter_rows<<<(n_rows * 16u + 255u) / 256u, 256, 0, st>>>(d_packed, n_rows, m.submeshes, m.vertex_ranges);
            CU(ctx, cudaGetLastError());
        }
    }
    m.n_vertices = s.n_vertices;
    m.n_indices = s.n_indices;
    m.n_submeshes = rows;
    m.n_work = n_work;
    CU(ctx, cudaMemsetAsync(obj->d_dirty, 0, n, st));  // mark_chunk_meshes_synchronized
    CU(ctx, cudaStreamSynchronize(st));
    return finish();
}

int ivx_mesh_modifications(ivx_ctx* ctx, const ivx_object* obj, uint32_t* out_ranges, size_t capacity_records, uint64_t* out_count,
                           int* out_chunks_were_removed) {
    if (!ctx || !obj || !out_count) return IVX_ERR_INVALID_ARGUMENT;
    const ivx_mesh_sync* s = obj->sync;
    const size_t cnt = s ? s->updated.size() / 4 : 0;
    *out_count = cnt;
    if (out_chunks_were_removed) *out_chunks_were_removed = s && s->chunks_were_removed ? 1 : 0;
    if (cnt > capacity_records || (cnt && !out_ranges)) return out_ranges ? IVX_ERR_CAPACITY : IVX_OK;
    if (cnt) std::memcpy(out_ranges, s->updated.data(), cnt * 16);
    return IVX_OK;
}

int ivx_mesh_report_synchronized(ivx_ctx* ctx, ivx_object* obj) {
    if (!ctx || !obj) return IVX_ERR_INVALID_ARGUMENT;
    if (obj->sync) {
        obj->sync->updated.clear();
        obj->sync->chunks_were_removed = false;
    }
    return IVX_OK;
}

}  // extern "C"
