// The mesh buffers a renderer draws from (sm_100a host side + one copy kernel): `VoxelMeshGPUBuffers`
// (engine/crates/impact_voxel/src/gpu_resource.rs:484-600 for_voxel_object, :714-900 sync_with_voxel_object).
//
// The reference keeps five wgpu buffers per meshed voxel object — vertex positions, normal vectors, index materials,
// indices, chunk submeshes — and brings them up to date after every mesh sync by staging the updated ranges from HOST
// memory (`VoxelMeshModifications`, mesh.rs:105-118); a buffer that became too small is re-created with the whole
// slice. Here the mesh never leaves the device, so the buffers are device allocations made with the virtual memory
// API and exported as POSIX file descriptors: a graphics API imports each one ONCE as external memory
// (VK_KHR_external_memory_fd; `cuMemImportFromShareableHandle` for a CUDA consumer) and a sync is one kernel that
// copies the updated ranges device to device on the context's stream, plus the submesh table. Same decisions as the
// reference: updated ranges only while the data fits, whole-slice re-creation (a new descriptor) when it does not,
// the submesh table rewritten whenever something changed, `report_gpu_resources_synchronized` at the end.
//
// The driver entry points are fetched through the runtime (cudaGetDriverEntryPoint): the library keeps its single
// link dependency on cudart.
#include <cuda.h>
#include <unistd.h>

#include "mesh_sync.cuh"

namespace {

struct DriverApi {
    CUresult (*getGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*exportHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
    bool ok = false;
};

template <typename F>
bool entry(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

const DriverApi& driver() {
    static const DriverApi api = [] {
        DriverApi d;
        d.ok = entry("cuMemGetAllocationGranularity", d.getGranularity) && entry("cuMemCreate", d.create) &&
               entry("cuMemAddressReserve", d.reserve) && entry("cuMemMap", d.map) && entry("cuMemSetAccess", d.setAccess) &&
               entry("cuMemExportToShareableHandle", d.exportHandle) && entry("cuMemUnmap", d.unmap) &&
               entry("cuMemRelease", d.release) && entry("cuMemAddressFree", d.addressFree);
        return d;
    }();
    return api;
}

struct SharedBuffer {
    CUmemGenericAllocationHandle handle = 0;
    CUdeviceptr ptr = 0;
    size_t bytes = 0;  // allocation size (a multiple of the granularity)
};

void destroy(SharedBuffer& b) {
    const DriverApi& d = driver();
    if (b.ptr) {
        d.unmap(b.ptr, b.bytes);
        d.addressFree(b.ptr, b.bytes);
    }
    if (b.handle) d.release(b.handle);
    b = SharedBuffer{};
}

// an exportable device allocation of at least `bytes`; *fd: a descriptor of it that the caller owns
const char* create(int device, size_t bytes, SharedBuffer& b, int* fd) {
    const DriverApi& d = driver();
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    if (d.getGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0)
        return "cuMemGetAllocationGranularity";
    b.bytes = (std::max<size_t>(bytes, 1) + gran - 1) / gran * gran;
    if (d.create(&b.handle, b.bytes, &prop, 0) != CUDA_SUCCESS) {
        b = SharedBuffer{};
        return "cuMemCreate (exportable device memory)";
    }
    if (d.reserve(&b.ptr, b.bytes, 0, 0, 0) != CUDA_SUCCESS) {
        b.ptr = 0;
        destroy(b);
        return "cuMemAddressReserve";
    }
    if (d.map(b.ptr, b.bytes, 0, b.handle, 0) != CUDA_SUCCESS) {
        d.addressFree(b.ptr, b.bytes);
        b.ptr = 0;
        destroy(b);
        return "cuMemMap";
    }
    CUmemAccessDesc acc{};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (d.setAccess(b.ptr, b.bytes, &acc, 1) != CUDA_SUCCESS) {
        destroy(b);
        return "cuMemSetAccess";
    }
    int out = -1;
    if (d.exportHandle(&out, b.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) {
        destroy(b);
        return "cuMemExportToShareableHandle";
    }
    *fd = out;
    return nullptr;
}

// one CTA per updated record (vertex start, end, index start, end): the four slices, word by word
__global__ void __launch_bounds__(256) k_copy_updated_ranges(const uint32_t* __restrict__ records, uint32_t n_records,
                                                             const uint32_t* __restrict__ positions, const uint32_t* __restrict__ normals,
                                                             const uint32_t* __restrict__ indices, const uint32_t* __restrict__ materials,
                                                             uint32_t* __restrict__ out_positions, uint32_t* __restrict__ out_normals,
                                                             uint32_t* __restrict__ out_indices, uint32_t* __restrict__ out_materials,
                                                             bool vertices, bool index_data) {
    for (uint32_t r = blockIdx.x; r < n_records; r += gridDim.x) {
        const uint32_t v0 = records[4 * r], v1 = records[4 * r + 1], i0 = records[4 * r + 2], i1 = records[4 * r + 3];
        if (vertices)
            for (uint32_t w = 3u * v0 + threadIdx.x; w < 3u * v1; w += blockDim.x) {
                out_positions[w] = positions[w];
                out_normals[w] = normals[w];
            }
        if (index_data)
            for (uint32_t w = i0 + threadIdx.x; w < i1; w += blockDim.x) {
                out_indices[w] = indices[w];
                out_materials[2 * (size_t)w] = materials[2 * (size_t)w];
                out_materials[2 * (size_t)w + 1] = materials[2 * (size_t)w + 1];
            }
    }
}

}  // namespace

struct ivx_mesh_gpu_buffers {
    SharedBuffer buffer[IVX_MESH_BUFFER_COUNT];
    uint64_t valid_bytes[IVX_MESH_BUFFER_COUNT] = {0, 0, 0, 0, 0};
    uint64_t mesh_serial = 0;  // DeviceMesh::serial of the mesh the buffers hold
};

namespace {

const size_t ELEM_BYTES[IVX_MESH_BUFFER_COUNT] = {12, 12, 8, 4, sizeof(ivx_chunk_submesh)};

void fill_info(const ivx_mesh_gpu_buffers& g, const DeviceMesh& m, ivx_mesh_gpu_buffers_info* info) {
    for (int b = 0; b < IVX_MESH_BUFFER_COUNT; ++b) {
        info->buffer[b].allocation_bytes = g.buffer[b].bytes;
        info->buffer[b].valid_bytes = g.valid_bytes[b];
        info->buffer[b].device_ptr = reinterpret_cast<void*>(g.buffer[b].ptr);
    }
    info->n_vertices = m.n_vertices;
    info->n_indices = m.n_indices;
    info->n_chunks = m.n_submeshes;
}

// (re)creates buffer `b` for `bytes` of data and fills it from `src`
int recreate(ivx_ctx* ctx, ivx_mesh_gpu_buffers& g, int b, const void* src, size_t bytes, ivx_mesh_gpu_buffers_info* info) {
    // the copies of earlier calls into the old allocation must have finished before it is unmapped
    if (g.buffer[b].ptr) CU(ctx, cudaStreamSynchronize(ctx->stream));
    destroy(g.buffer[b]);
    int fd = -1;
    // (room to grow: a buffer is re-created — and has to be imported again — only when the mesh outgrows it)
    if (const char* what = create(ctx->device, bytes + bytes / 4, g.buffer[b], &fd))
        IVX_FAIL(ctx, IVX_ERR_CUDA, "mesh buffers: %s failed", what);
    if (bytes)
        CU(ctx, cudaMemcpyAsync(reinterpret_cast<void*>(g.buffer[b].ptr), src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    g.valid_bytes[b] = bytes;
    info->buffer[b].fd = fd;
    info->buffer[b].recreated = 1;
    info->bytes_copied += bytes;
    return IVX_OK;
}

const void* source(const DeviceMesh& m, int b) {
    switch (b) {
        case IVX_MESH_BUFFER_POSITIONS: return m.positions;
        case IVX_MESH_BUFFER_NORMALS: return m.normals;
        case IVX_MESH_BUFFER_INDEX_MATERIALS: return m.index_materials;
        case IVX_MESH_BUFFER_INDICES: return m.indices;
        default: return m.submeshes;
    }
}
size_t elements(const DeviceMesh& m, int b) {
    return b <= IVX_MESH_BUFFER_NORMALS ? m.n_vertices : (b <= IVX_MESH_BUFFER_INDICES ? m.n_indices : m.n_submeshes);
}

int check(ivx_ctx* ctx, const ivx_object* obj) {
    if (!driver().ok) IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "mesh buffers: the driver's virtual memory entry points are not available");
    if (obj->mesh_is_patch || !obj->mesh.positions)
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "mesh buffers: the object has no full mesh (call ivx_object_mesh first)");
    return IVX_OK;
}

}  // namespace

extern "C" {

int ivx_mesh_gpu_buffers_create(ivx_ctx* ctx, ivx_object* obj, ivx_mesh_gpu_buffers** out, ivx_mesh_gpu_buffers_info* info) {
    if (!ctx || !obj || !out || !info) return IVX_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    cudaSetDevice(ctx->device);
    cudaFree(nullptr);  // (the primary context is current on this thread from here on)
    if (int rc = check(ctx, obj)) return rc;
    std::memset(info, 0, sizeof(*info));
    for (auto& b : info->buffer) b.fd = -1;
    ivx_mesh_gpu_buffers* g = new (std::nothrow) ivx_mesh_gpu_buffers();
    if (!g) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "host allocation failed");
    const DeviceMesh& m = obj->mesh;
    for (int b = 0; b < IVX_MESH_BUFFER_COUNT; ++b)
        if (int rc = recreate(ctx, *g, b, source(m, b), elements(m, b) * ELEM_BYTES[b], info)) {
            for (int q = 0; q < IVX_MESH_BUFFER_COUNT; ++q) {
                if (info->buffer[q].fd >= 0) close(info->buffer[q].fd);
                info->buffer[q].fd = -1;
                destroy(g->buffer[q]);
            }
            delete g;
            return rc;
        }
    g->mesh_serial = m.serial;
    // for_voxel_object uploads the mesh as it is: nothing of it is pending afterwards
    if (obj->sync) ivx_mesh_report_synchronized(ctx, obj);
    fill_info(*g, m, info);
    *out = g;
    return IVX_OK;
}

static int sync_impl(ivx_ctx* ctx, ivx_object* obj, ivx_mesh_gpu_buffers* g, ivx_mesh_gpu_buffers_info* info) {
    const DeviceMesh& m = obj->mesh;
    cudaStream_t st = ctx->stream;
    ivx_mesh_sync* s = obj->sync;
    // a mesh that was created anew since the buffers were filled has no modification list to follow: everything is new
    const bool everything = m.serial != g->mesh_serial;
    const size_t n_records = s ? s->updated.size() / 4 : 0;
    if (!everything && n_records == 0 && !(s && s->chunks_were_removed)) {
        fill_info(*g, m, info);
        return IVX_OK;
    }
    info->n_updated_ranges = (uint32_t)n_records;
    if (everything || n_records) {
        // vertex data, then index data: re-created when the slices outgrew the buffers, else the updated ranges
        bool copy_pair[2] = {false, false};
        for (int pair = 0; pair < 2; ++pair) {
            const int b0 = pair == 0 ? IVX_MESH_BUFFER_POSITIONS : IVX_MESH_BUFFER_INDEX_MATERIALS;
            bool outgrown = false;
            for (int b = b0; b < b0 + 2; ++b) outgrown |= elements(m, b) * ELEM_BYTES[b] > g->buffer[b].bytes;
            if (outgrown) {
                for (int b = b0; b < b0 + 2; ++b)
                    if (int rc = recreate(ctx, *g, b, source(m, b), elements(m, b) * ELEM_BYTES[b], info)) return rc;
            } else if (everything) {
                for (int b = b0; b < b0 + 2; ++b) {
                    const size_t bytes = elements(m, b) * ELEM_BYTES[b];
                    if (bytes)
                        CU(ctx, cudaMemcpyAsync(reinterpret_cast<void*>(g->buffer[b].ptr), source(m, b), bytes, cudaMemcpyDeviceToDevice, st));
                    g->valid_bytes[b] = bytes;
                    info->bytes_copied += bytes;
                }
            } else {
                copy_pair[pair] = true;
                for (int b = b0; b < b0 + 2; ++b) g->valid_bytes[b] = elements(m, b) * ELEM_BYTES[b];
            }
        }
        if (copy_pair[0] || copy_pair[1]) {
            Tmp tmp(ctx);
            uint32_t* records = tmp.get<uint32_t>(4 * n_records);
            if (!records) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh buffers: out of device memory");
            CU(ctx, cudaMemcpyAsync(records, s->updated.data(), n_records * 16, cudaMemcpyHostToDevice, st));
            ctx->launches++;
            k_copy_updated_ranges<<<(unsigned)std::min<size_t>(n_records, 148 * 8), 256, 0, st>>>(
                records, (uint32_t)n_records, reinterpret_cast<const uint32_t*>(m.positions), reinterpret_cast<const uint32_t*>(m.normals),
                m.indices, reinterpret_cast<const uint32_t*>(m.index_materials),
                reinterpret_cast<uint32_t*>(g->buffer[IVX_MESH_BUFFER_POSITIONS].ptr), reinterpret_cast<uint32_t*>(g->buffer[IVX_MESH_BUFFER_NORMALS].ptr),
                reinterpret_cast<uint32_t*>(g->buffer[IVX_MESH_BUFFER_INDICES].ptr),
                reinterpret_cast<uint32_t*>(g->buffer[IVX_MESH_BUFFER_INDEX_MATERIALS].ptr), copy_pair[0], copy_pair[1]);
            CU(ctx, cudaGetLastError());
            for (size_t r = 0; r < n_records; ++r) {
                if (copy_pair[0]) info->bytes_copied += 24ull * (s->updated[4 * r + 1] - s->updated[4 * r]);
                if (copy_pair[1]) info->bytes_copied += 12ull * (s->updated[4 * r + 3] - s->updated[4 * r + 2]);
            }
        }
    }
    // the chunk submesh table: re-created when it outgrew its buffer, else its valid bytes overwritten
    {
        const int b = IVX_MESH_BUFFER_CHUNK_SUBMESHES;
        const size_t bytes = elements(m, b) * ELEM_BYTES[b];
        if (bytes > g->buffer[b].bytes) {
            if (int rc = recreate(ctx, *g, b, source(m, b), bytes, info)) return rc;
        } else {
            if (bytes) CU(ctx, cudaMemcpyAsync(reinterpret_cast<void*>(g->buffer[b].ptr), source(m, b), bytes, cudaMemcpyDeviceToDevice, st));
            g->valid_bytes[b] = bytes;
            info->bytes_copied += bytes;
        }
    }
    g->mesh_serial = m.serial;
    if (s) ivx_mesh_report_synchronized(ctx, obj);  // report_gpu_resources_synchronized
    fill_info(*g, m, info);
    return IVX_OK;
}

int ivx_mesh_gpu_buffers_sync(ivx_ctx* ctx, ivx_object* obj, ivx_mesh_gpu_buffers* g, ivx_mesh_gpu_buffers_info* info) {
    if (!ctx || !obj || !g || !info) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    std::memset(info, 0, sizeof(*info));
    for (auto& b : info->buffer) b.fd = -1;
    if (int rc = check(ctx, obj)) return rc;
    const int rc = sync_impl(ctx, obj, g, info);
    if (rc != IVX_OK)  // no descriptor is handed out by a call that failed
        for (auto& b : info->buffer) {
            if (b.fd >= 0) close(b.fd);
            b.fd = -1;
        }
    return rc;
}

void ivx_mesh_gpu_buffers_destroy(ivx_ctx* ctx, ivx_mesh_gpu_buffers* g) {
    if (!g) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    for (auto& b : g->buffer) destroy(b);
    delete g;
}

}  // extern "C"
