// Multi-GPU plane of the voxel path over PEER MEMORY (NVLink / NVSwitch), sm_100a: halo-plane exchange between the
// x-slab objects of neighbouring ranks and the gather of the per-slab meshes on one rank, with no NCCL message, no
// host-side wait and no host round trip on the data path.
//
// The reference parallelises over contiguous ranges of the x-major linear chunk index (object.rs:423-427); the same
// split over GPUs needs the neighbour's boundary chunk plane for the cross-chunk derived state and the 1-voxel brick
// padding (object.rs:1682-1704, object/sdf.rs:410-428) and one bit per chunk for the quad-ownership rule
// (object/sdf/surface_nets.rs:252-261); the slab meshes concatenate in slab order, which is the reference's order
// (mesh.rs:286-354).
//
// Every rank owns one WINDOW (a cudaMalloc block exported through CUDA IPC and mapped by all ranks):
//
//   header   flags and small records, written by peers with system-scope release stores
//   inboxes  halo messages [side][parity], kinds messages [parity]
//   mesh     (gather rank only) the merged mesh [parity]: positions, normals, indices, index materials, submeshes,
//            vertex ranges, each sized by the capacity given at creation
//
// A step = ivx_object_exchange_halos + ivx_object_mesh_gather on every rank, numbered by an epoch. Producers store
// straight into the consumer's window (k_halo_pack / k_halo_kinds_pack / k_push_dyn run with peer pointers as their
// destination), then publish the epoch in the consumer's header (k_signal); consumers wait ON THE DEVICE (k_wait: one
// thread spinning on an acquire load), so the kernels that need the data are simply queued behind the wait on the
// context's stream. Inboxes and the merged mesh are double buffered by epoch parity; the gather rank publishes
// "step e complete" to every rank and nobody starts the exchange of step e + 2 before seeing it, so a rank is never more
// than one step ahead of another and a buffer is never overwritten while in use.
#include "api_internal.cuh"

namespace {

constexpr uint32_t MAX_WORLD = 64;
constexpr unsigned long long WAIT_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;  // a peer that never arrives

struct CountsRecord {
    uint32_t v, i, s, epoch;
};
struct CommHeader {
    uint32_t halo_flag[2][2];   // [side][parity]: epoch of the message in that inbox
    uint32_t kinds_flag[2];     // [parity]
    uint32_t step_done;         // last epoch the gather rank has completed
    uint32_t error;             // sticky: 1 timeout, 2 merged mesh capacity exceeded
    CountsRecord counts[2][MAX_WORLD];  // [parity][rank]: mesh sizes of that rank's slab
    uint32_t part_done[2][MAX_WORLD];   // gather rank: rank r's part of the merged mesh has landed
    uint32_t bases[2][4];       // local scratch: my vertex / index / submesh base in the merged mesh, ok flag
    uint32_t totals[2][4];      // gather rank: merged sizes
};
constexpr size_t HEADER_BYTES = 8192;
static_assert(sizeof(CommHeader) <= HEADER_BYTES, "header fits its page");

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spins until *flag >= epoch; false (and the sticky error word set) after WAIT_TIMEOUT_NS
// `site` says which wait it was (error word = 1 | site << 8): 1 previous step complete, 2 halo planes, 3 quad-ownership
// bits, 4 mesh sizes of a lower rank, 5 a rank's part of the merged mesh
__device__ bool wait_flag(const uint32_t* flag, uint32_t epoch, uint32_t* error, uint32_t site) {
    const unsigned long long t0 = now_ns();
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(64);
        if (now_ns() - t0 > WAIT_TIMEOUT_NS) {
            *error = 1u | (site << 8);
            return false;
        }
    }
    return true;
}

// publishes `epoch` in up to two peer flag words after everything this stream stored before
__global__ void k_signal(uint32_t* flag_a, uint32_t* flag_b, uint32_t epoch) {
    __threadfence_system();
    if (flag_a) st_release_sys(flag_a, epoch);
    if (flag_b) st_release_sys(flag_b, epoch);
}
__global__ void k_wait(const uint32_t* flag_a, const uint32_t* flag_b, uint32_t epoch, uint32_t* error, uint32_t site) {
    if (flag_a) wait_flag(flag_a, epoch, error, site);
    if (flag_b) wait_flag(flag_b, epoch, error, site);
}

struct PeerHeaders {
    CommHeader* h[MAX_WORLD];
};
// my mesh sizes → every rank's header (record first, epoch last)
__global__ void k_publish_counts(PeerHeaders peers, uint32_t world, uint32_t rank, uint32_t parity, uint32_t epoch,
                                 const uint32_t* __restrict__ counts3) {
    const uint32_t r = threadIdx.x;
    if (r >= world) return;
    CountsRecord* rec = &peers.h[r]->counts[parity][rank];
    rec->v = counts3 ? counts3[0] : 0u;
    rec->i = counts3 ? counts3[1] : 0u;
    rec->s = counts3 ? counts3[2] : 0u;
    __threadfence_system();
    st_release_sys(&rec->epoch, epoch);
}
// waits for the sizes of the lower ranks, leaves my bases (exclusive prefix) in bases[parity]; ok = the part fits
__global__ void k_prefix_counts(CommHeader* mine, uint32_t rank, uint32_t parity, uint32_t epoch, uint32_t cap_v, uint32_t cap_i,
                                uint32_t cap_s, uint32_t* gather_error) {
    uint64_t v = 0, i = 0, s = 0;
    bool ok = true;
    for (uint32_t r = 0; r <= rank && ok; ++r) {
        const CountsRecord* rec = &mine->counts[parity][r];
        ok = wait_flag(&rec->epoch, epoch, &mine->error, 4u);
        if (r < rank) {
            v += rec->v;
            i += rec->i;
            s += rec->s;
        } else if (v + rec->v > cap_v || i + rec->i > cap_i || s + rec->s > cap_s) {
            ok = false;
            mine->error = 2u;
            st_release_sys(gather_error, 2u);
        }
    }
    mine->bases[parity][0] = (uint32_t)v;
    mine->bases[parity][1] = (uint32_t)i;
    mine->bases[parity][2] = (uint32_t)s;
    mine->bases[parity][3] = ok ? 1u : 0u;
}
// One field of my slab's mesh → its place in the merged mesh in the gather rank's window: word q of `src` lands at
// dst[base * elem_words + q], plus bases[add_sel] when q % period == col (period 1: every word — the u32 vertex
// indices and vertex ranges; period 13, col 3: ChunkSubmesh::index_offset; period 0: plain copy). 16-byte accesses when
// both sides are aligned.
__global__ void __launch_bounds__(256) k_push_dyn(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst_field,
                                                  size_t n_words, uint32_t elem_words, const uint32_t* __restrict__ bases,
                                                  uint32_t base_sel, uint32_t add_sel, uint32_t period, uint32_t col) {
    if (bases[3] == 0u) return;
    uint32_t* dst = dst_field + (size_t)bases[base_sel] * elem_words;
    const uint32_t add = bases[add_sel];
    const size_t n4 = n_words / 4;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto fix = [&](uint32_t w, size_t q) { return (period == 1u || (period > 1u && q % period == col)) ? w + add : w; };
    if (aligned) {
        for (size_t q = t; q < n4; q += stride) {
            uint4 v = reinterpret_cast<const uint4*>(src)[q];
            v.x = fix(v.x, 4 * q);
            v.y = fix(v.y, 4 * q + 1);
            v.z = fix(v.z, 4 * q + 2);
            v.w = fix(v.w, 4 * q + 3);
            reinterpret_cast<uint4*>(dst)[q] = v;
        }
        for (size_t q = 4 * n4 + t; q < n_words; q += stride) dst[q] = fix(src[q], q);
    } else {
        for (size_t q = t; q < n_words; q += stride) dst[q] = fix(src[q], q);
    }
}
// The mesh stays where it is (ivx_object_mesh_distributed): indices, submesh index offsets and vertex ranges become the
// ones the part has in the mesh of the whole job; my bases go to the host
__global__ void __launch_bounds__(256) k_rebase_local(uint32_t* __restrict__ indices, size_t n_indices, uint32_t* __restrict__ submesh_words,
                                                      size_t n_submeshes, uint32_t* __restrict__ vertex_ranges,
                                                      const uint32_t* __restrict__ bases, uint32_t* __restrict__ host_words) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    if (t == 0) {
        host_words[4] = bases[0];
        host_words[5] = bases[1];
        host_words[6] = bases[2];
    }
    if (bases[3] == 0u) return;
    const uint32_t vb = bases[0], ib = bases[1];
    for (size_t q = t; q < n_indices; q += stride) indices[q] += vb;
    for (size_t q = t; q < n_submeshes; q += stride) submesh_words[13 * q + 3] += ib;  // ChunkSubmesh::index_offset
    for (size_t q = t; q < 2 * n_submeshes; q += stride) vertex_ranges[q] += vb;
}
// gather rank: all parts have landed → merged sizes for the host, "step complete" to everybody
__global__ void k_complete_step(CommHeader* mine, PeerHeaders peers, uint32_t world, uint32_t parity, uint32_t epoch,
                                uint32_t* __restrict__ host_words) {
    uint64_t v = 0, i = 0, s = 0;
    bool ok = true;
    for (uint32_t r = 0; r < world && ok; ++r) {
        ok = wait_flag(&mine->part_done[parity][r], epoch, &mine->error, 5u);
        const CountsRecord* rec = &mine->counts[parity][r];
        v += rec->v;
        i += rec->i;
        s += rec->s;
    }
    mine->totals[parity][0] = (uint32_t)v;
    mine->totals[parity][1] = (uint32_t)i;
    mine->totals[parity][2] = (uint32_t)s;
    host_words[0] = (uint32_t)v;
    host_words[1] = (uint32_t)i;
    host_words[2] = (uint32_t)s;
    host_words[3] = mine->error;
    __threadfence_system();
    for (uint32_t r = 0; r < world; ++r) st_release_sys(&peers.h[r]->step_done, epoch);
}
__global__ void k_store_error(const CommHeader* mine, uint32_t* __restrict__ host_word) {
    *host_word = mine->error;
    __threadfence_system();
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

struct ivx_comm {
    ivx_comm_config cfg{};
    unsigned char* base = nullptr;            // my window
    unsigned char* peer[MAX_WORLD] = {};      // every rank's window as mapped here (peer[rank] == base)
    bool connected = false;
    bool local = false;                       // peers are plain pointers of this process (ivx_comm_connect_local)
    uint32_t epoch = 0;                       // of the current step (0 = none yet)
    bool exchanged = false;                   // exchange_halos ran for the current epoch
    size_t halo_bytes = 0, kinds_bytes = 0;
    size_t off_halo[2][2] = {}, off_kinds[2] = {}, off_mesh[2][6] = {};
    size_t window_bytes = 0;
    uint32_t* h_words = nullptr;              // pinned + mapped: totals and error of the last step
    uint32_t* h_words_dev = nullptr;

    CommHeader* header(uint32_t r) const { return reinterpret_cast<CommHeader*>(peer[r]); }
    // offsets are the same in every window (the mesh region exists in the gather rank's only)
    void layout() {
        size_t o = HEADER_BYTES;
        halo_bytes = halo_message_bytes(cfg.plane_chunks);
        kinds_bytes = cfg.plane_chunks;
        for (int side = 0; side < 2; ++side)
            for (int par = 0; par < 2; ++par) {
                off_halo[side][par] = o;
                o += align256(halo_bytes);
            }
        for (int par = 0; par < 2; ++par) {
            off_kinds[par] = o;
            o += align256(kinds_bytes);
        }
        const size_t field[6] = {(size_t)cfg.mesh_vertices * 12, (size_t)cfg.mesh_vertices * 12, (size_t)cfg.mesh_indices * 4,
                                 (size_t)cfg.mesh_indices * 8,  (size_t)cfg.mesh_submeshes * 52, (size_t)cfg.mesh_submeshes * 8};
        size_t mesh_end = o;
        for (int par = 0; par < 2; ++par)
            for (int f = 0; f < 6; ++f) {
                off_mesh[par][f] = mesh_end;
                mesh_end += align256(field[f]);
            }
        window_bytes = cfg.rank == cfg.gather_rank ? mesh_end : o;
    }
};

extern "C" {

int ivx_comm_create(ivx_ctx* ctx, const ivx_comm_config* config, ivx_comm** out_comm, unsigned char out_handle[64]) {
    if (!ctx || !config || !out_comm || !out_handle) return IVX_ERR_INVALID_ARGUMENT;
    *out_comm = nullptr;
    if (config->world == 0 || config->world > MAX_WORLD || config->rank >= config->world || config->gather_rank >= config->world)
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "communicator of %u ranks (at most %u), rank %u, gather rank %u", config->world,
                 MAX_WORLD, config->rank, config->gather_rank);
    if (config->mesh_vertices > 0xFFFFFFFFull || config->mesh_indices > 0xFFFFFFFFull || config->mesh_submeshes > 0xFFFFFFFFull)
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "merged mesh capacities are 32-bit counts (mesh.rs:87: u32 indices)");
    cudaSetDevice(ctx->device);
    ivx_comm* c = new (std::nothrow) ivx_comm();
    if (!c) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "host allocation failed");
    c->cfg = *config;
    c->layout();
    void* p = nullptr;
    if (int rc = ivx_peer_alloc(ctx, c->window_bytes, &p, out_handle)) {
        delete c;
        return rc;
    }
    c->base = static_cast<unsigned char*>(p);
    c->peer[config->rank] = c->base;
    if (cudaMemsetAsync(c->base, 0, HEADER_BYTES, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
        cudaHostAlloc(&c->h_words, 16 * sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&c->h_words_dev, c->h_words, 0) != cudaSuccess) {
        cudaGetLastError();
        ivx_comm_destroy(ctx, c);
        IVX_FAIL(ctx, IVX_ERR_CUDA, "communicator window setup failed");
    }
    std::memset(c->h_words, 0, 16 * sizeof(uint32_t));
    *out_comm = c;
    return IVX_OK;
}

int ivx_comm_connect(ivx_ctx* ctx, ivx_comm* c, const unsigned char* all_handles) {
    if (!ctx || !c || !all_handles) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    for (uint32_t r = 0; r < c->cfg.world; ++r) {
        if (r == c->cfg.rank || c->peer[r]) continue;
        void* p = nullptr;
        if (int rc = ivx_peer_open(ctx, all_handles + (size_t)r * 64, &p)) return rc;
        c->peer[r] = static_cast<unsigned char*>(p);
    }
    c->connected = true;
    return IVX_OK;
}

// ranks of one process (several contexts, on one device or several): the windows are plain pointers
int ivx_comm_connect_local(ivx_ctx* ctx, ivx_comm* c, ivx_comm* const* all) {
    if (!ctx || !c || !all) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    for (uint32_t r = 0; r < c->cfg.world; ++r) {
        if (!all[r] || !all[r]->base) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "communicator of rank %u is missing", r);
        if (r == c->cfg.rank) continue;
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, all[r]->base) == cudaSuccess && at.device != ctx->device) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "no peer access from device %d to device %d: %s", ctx->device, at.device,
                         cudaGetErrorString(e));
            }
            cudaGetLastError();
        }
        c->peer[r] = all[r]->base;
    }
    c->local = true;
    c->connected = true;
    return IVX_OK;
}

void ivx_comm_destroy(ivx_ctx* ctx, ivx_comm* c) {
    if (!ctx || !c) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (uint32_t r = 0; r < c->cfg.world; ++r)
        if (!c->local && r != c->cfg.rank && c->peer[r]) ivx_peer_close(ctx, c->peer[r]);
    if (c->base) ivx_peer_free(ctx, c->base);
    if (c->h_words) cudaFreeHost(c->h_words);
    delete c;
}

// generate_slab → [this call] → the object's owned planes carry the whole object's derived state
int ivx_object_exchange_halos(ivx_ctx* ctx, ivx_comm* c, ivx_object* obj, int lower_rank, int upper_rank) {
    if (!ctx || !c || !obj) return IVX_ERR_INVALID_ARGUMENT;
    if (!c->connected) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "call ivx_comm_connect first");
    cudaSetDevice(ctx->device);
    const int world = (int)c->cfg.world, rank = (int)c->cfg.rank;
    if (lower_rank >= rank || (upper_rank >= 0 && upper_rank <= rank) || upper_rank >= world)
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "neighbour ranks %d / %d of rank %d", lower_rank, upper_rank, rank);
    const int nbr[2] = {lower_rank, upper_rank};
    for (int side = 0; side < 2; ++side)
        if ((nbr[side] >= 0) != (obj->n_chunks != 0 && obj->halo_present[side]))
            IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "object %s a neighbour slab on side %d but rank %d was given",
                     obj->halo_present[side] ? "has" : "has no", side, nbr[side]);
    if (obj->n_chunks && obj->nb[1] * obj->nb[2] != c->cfg.plane_chunks)
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "communicator was created for planes of %u chunks, object has %u",
                 c->cfg.plane_chunks, obj->nb[1] * obj->nb[2]);
    if (obj->n_chunks && !obj->derive_pending) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "object is not a slab with pending derived state");
    cudaStream_t st = ctx->stream;
    const uint32_t epoch = ++c->epoch;
    const uint32_t par = epoch & 1u;
    c->exchanged = true;
    CommHeader* mine = c->header(rank);
    // nobody runs more than one step ahead: the buffers of parity `par` were last used by step epoch - 2
    if (epoch > 2) {
        ctx->launches++;
        k_wait<<<1, 1, 0, st>>>(&mine->step_done, nullptr, epoch - 2, &mine->error, 1u);
        CU(ctx, cudaGetLastError());
    }
    if (obj->n_chunks == 0) return IVX_OK;
    const uint32_t plane = c->cfg.plane_chunks;
    const uint32_t lo = obj->own_begin - obj->first_i, hi = obj->own_end - obj->first_i;  // local own planes [lo, hi)

    // exchange A: my boundary planes → straight into the neighbours' inboxes (their side is the opposite one)
    for (int side = 0; side < 2; ++side) {
        if (nbr[side] < 0) continue;
        unsigned char* dst = c->peer[nbr[side]] + c->off_halo[1 - side][par];
        const uint32_t own_first = (side == 0 ? lo : hi - 1) * plane;
        KL(ctx, launch_halo_pack(obj->d_chunks, own_first, plane, side == 0 ? 0u : 15u, obj->d_voxels, dst, st));
    }
    if (nbr[0] >= 0 || nbr[1] >= 0) {
        ctx->launches++;
        k_signal<<<1, 1, 0, st>>>(nbr[0] >= 0 ? &c->header(nbr[0])->halo_flag[1][par] : nullptr,
                                  nbr[1] >= 0 ? &c->header(nbr[1])->halo_flag[0][par] : nullptr, epoch);
        CU(ctx, cudaGetLastError());
        ctx->launches++;
        k_wait<<<1, 1, 0, st>>>(nbr[0] >= 0 ? &mine->halo_flag[0][par] : nullptr, nbr[1] >= 0 ? &mine->halo_flag[1][par] : nullptr,
                                epoch, &mine->error, 2u);
        CU(ctx, cudaGetLastError());
    }
    for (int side = 0; side < 2; ++side) {
        if (nbr[side] < 0) continue;
        if (int rc = ivx_object_halo_import(ctx, obj, side, c->base + c->off_halo[side][par], c->halo_bytes)) return rc;
    }
    if (int rc = ivx_object_slab_classify(ctx, obj)) return rc;
    // exchange B: which chunks of my lowest plane end up non-uniform → the lower neighbour
    if (nbr[0] >= 0) {
        KL(ctx, launch_halo_kinds_pack(obj->d_chunks, obj->d_convert_flag, lo * plane, plane,
                                       c->peer[nbr[0]] + c->off_kinds[par], st));
        ctx->launches++;
        k_signal<<<1, 1, 0, st>>>(&c->header(nbr[0])->kinds_flag[par], nullptr, epoch);
        CU(ctx, cudaGetLastError());
    }
    if (nbr[1] >= 0) {
        ctx->launches++;
        k_wait<<<1, 1, 0, st>>>(&mine->kinds_flag[par], nullptr, epoch, &mine->error, 3u);
        CU(ctx, cudaGetLastError());
        if (int rc = ivx_object_halo_kinds_import(ctx, obj, 1, c->base + c->off_kinds[par], plane)) return rc;
    }
    return ivx_internal_slab_finalize(ctx, obj, /*sync=*/false);
}

static int mesh_gather_impl(ivx_ctx* ctx, ivx_comm* c, ivx_object* obj, ivx_mesh_info* out_local, ivx_gathered_mesh* out_merged,
                            bool distributed, uint64_t* out_bases) {
    if (!ctx || !c || !obj || !out_local) return IVX_ERR_INVALID_ARGUMENT;
    if (!c->connected) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "call ivx_comm_connect first");
    cudaSetDevice(ctx->device);
    std::memset(out_local, 0, sizeof(*out_local));
    if (out_merged) std::memset(out_merged, 0, sizeof(*out_merged));
    if (!c->exchanged) {  // a step without a halo exchange (world of one rank, or the host ran the slab protocol itself)
        ++c->epoch;
        if (c->epoch > 2) {
            ctx->launches++;
            k_wait<<<1, 1, 0, ctx->stream>>>(&c->header(c->cfg.rank)->step_done, nullptr, c->epoch - 2, &c->header(c->cfg.rank)->error, 1u);
            CU(ctx, cudaGetLastError());
        }
    }
    c->exchanged = false;
    const uint32_t epoch = c->epoch, par = epoch & 1u;
    const uint32_t world = c->cfg.world, rank = c->cfg.rank, dst = c->cfg.gather_rank;
    cudaStream_t st = ctx->stream;
    CommHeader* mine = c->header(rank);

    // my slab's mesh (no host round trip when the object's plan knows its sizes)
    uint32_t counts[4] = {0, 0, 0, 0};
    if (int rc = ivx_internal_mesh(ctx, obj, /*sync=*/false, counts, out_local)) return rc;
    const DeviceMesh& m = obj->mesh;

    PeerHeaders ph{};
    for (uint32_t r = 0; r < world; ++r) ph.h[r] = c->header(r);
    // sizes → everybody (device counters of mesh_impl: d_scratch[17..19]); an empty slab publishes zeros
    ctx->launches++;
    k_publish_counts<<<1, MAX_WORLD, 0, st>>>(ph, world, rank, par, epoch, counts[0] ? ctx->d_scratch + 17 : nullptr);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    k_prefix_counts<<<1, 1, 0, st>>>(mine, rank, par, epoch, (uint32_t)c->cfg.mesh_vertices, (uint32_t)c->cfg.mesh_indices,
                                     (uint32_t)c->cfg.mesh_submeshes, &c->header(dst)->error);
    CU(ctx, cudaGetLastError());
    // my part → its place in the merged mesh, rebased on the way (VoxelObjectMesh layout, mesh.rs:50-103)
    const uint32_t* bases = mine->bases[par];
    unsigned char* gw = c->peer[dst];
    const uint32_t grid = (uint32_t)ctx->sm_count * 4u;
    struct Field {
        const void* src;
        size_t words;
        uint32_t elem_words, base_sel, add_sel, period, col;
    };
    const size_t nv = m.n_vertices, ni = m.n_indices, ns = m.n_submeshes;
    const Field fields[6] = {
        {m.positions, nv * 3, 3, 0, 0, 0, 0},       {m.normals, nv * 3, 3, 0, 0, 0, 0},
        {m.indices, ni, 1, 1, 0, 1, 0},             {m.index_materials, ni * 2, 2, 1, 0, 0, 0},
        {m.submeshes, ns * 13, 13, 2, 1, 13, 3},    {m.vertex_ranges, ns * 2, 2, 2, 0, 1, 0},
    };
    if (distributed) {
        // every rank keeps its part; only the sizes travelled
        ctx->launches++;
        k_rebase_local<<<grid, 256, 0, st>>>(m.indices, ni, reinterpret_cast<uint32_t*>(m.submeshes), ns, m.vertex_ranges, bases,
                                             c->h_words_dev);
        CU(ctx, cudaGetLastError());
    }
    for (int f = 0; f < 6 && !distributed; ++f) {
        if (fields[f].words == 0) continue;
        ctx->launches++;
        k_push_dyn<<<grid, 256, 0, st>>>(static_cast<const uint32_t*>(fields[f].src),
                                         reinterpret_cast<uint32_t*>(gw + c->off_mesh[par][f]), fields[f].words,
                                         fields[f].elem_words, bases, fields[f].base_sel, fields[f].add_sel, fields[f].period,
                                         fields[f].col);
        CU(ctx, cudaGetLastError());
    }
    ctx->launches++;
    k_signal<<<1, 1, 0, st>>>(&c->header(dst)->part_done[par][rank], nullptr, epoch);
    CU(ctx, cudaGetLastError());
    if (rank == dst) {
        ctx->launches++;
        k_complete_step<<<1, 1, 0, st>>>(mine, ph, world, par, epoch, c->h_words_dev);
        CU(ctx, cudaGetLastError());
    } else {
        ctx->launches++;
        k_store_error<<<1, 1, 0, st>>>(mine, c->h_words_dev + 3);
        CU(ctx, cudaGetLastError());
    }
    CU(ctx, cudaStreamSynchronize(st));
    if (int rc = ivx_internal_take_plan_error(ctx, obj)) return rc;
    const uint32_t err = c->h_words[3];
    if ((err & 0xFFu) == 1u)
        IVX_FAIL(ctx, IVX_ERR_CUDA, "multi-GPU step %u: a peer rank did not arrive within %llu s (wait site %u: 1 previous step, "
                 "2 halo planes, 3 quad-ownership bits, 4 mesh sizes, 5 mesh parts)", epoch, WAIT_TIMEOUT_NS / 1000000000ull, err >> 8);
    if (err == 2u) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "multi-GPU step %u: the merged mesh does not fit the communicator's capacity "
                            "(%llu vertices, %llu indices, %llu submeshes)", epoch, (unsigned long long)c->cfg.mesh_vertices,
                            (unsigned long long)c->cfg.mesh_indices, (unsigned long long)c->cfg.mesh_submeshes);
    if (out_bases)
        for (int q = 0; q < 3; ++q) out_bases[q] = c->h_words[4 + q];
    if (rank == dst && out_merged && !distributed) {
        out_merged->n_vertices = c->h_words[0];
        out_merged->n_indices = c->h_words[1];
        out_merged->n_submeshes = c->h_words[2];
        out_merged->d_positions = c->base + c->off_mesh[par][0];
        out_merged->d_normals = c->base + c->off_mesh[par][1];
        out_merged->d_indices = c->base + c->off_mesh[par][2];
        out_merged->d_index_materials = c->base + c->off_mesh[par][3];
        out_merged->d_submeshes = c->base + c->off_mesh[par][4];
        out_merged->d_vertex_ranges = c->base + c->off_mesh[par][5];
    }
    return IVX_OK;
}

int ivx_object_mesh_gather(ivx_ctx* ctx, ivx_comm* c, ivx_object* obj, ivx_mesh_info* out_local, ivx_gathered_mesh* out_merged) {
    return mesh_gather_impl(ctx, c, obj, out_local, out_merged, /*distributed=*/false, nullptr);
}

int ivx_object_mesh_distributed(ivx_ctx* ctx, ivx_comm* c, ivx_object* obj, ivx_mesh_info* out_local, uint64_t out_bases[3]) {
    if (!out_bases) return IVX_ERR_INVALID_ARGUMENT;
    return mesh_gather_impl(ctx, c, obj, out_local, nullptr, /*distributed=*/true, out_bases);
}

}  // extern "C"
