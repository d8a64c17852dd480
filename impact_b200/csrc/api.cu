// C ABI of libimpact_voxel_cuda.so: context, device memory pool, and the
// host-side orchestration of the generate → derive → mesh → modify kernels.
// See include/impact_voxel_cuda.h for the contract of each entry point.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "api_internal.cuh"

using namespace ivx;

namespace {

// device words → mapped pinned host words, by a kernel: a device→host memcpy on the compute stream would queue on
// the copy engine behind the bulk transfers of a streamed generation and stall every kernel after it
__global__ void k_store_words(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t n) {
    if (threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}
cudaError_t store_words(ivx_ctx* ctx, const uint32_t* d_src, uint32_t host_word, uint32_t n) {
    k_store_words<<<1, 64, 0, ctx->stream>>>(d_src, ctx->h_pinned_dev + host_word, n);
    return cudaGetLastError();
}

int read_words(ivx_ctx* ctx, const uint32_t* d_src, uint32_t n, uint32_t* out) {
    if (n > 32) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "read_words: at most 32 words");
    CU(ctx, store_words(ctx, d_src, 0, n));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(out, ctx->h_pinned, n * sizeof(uint32_t));
    return IVX_OK;
}

// ---- plans (api_internal.cuh GenPlan) ----
struct PlanWords {
    uint32_t idx[12];
    uint32_t want[12];
    uint32_t n;
};
// compares device counters with what the plan promised; a mismatch sets the context's sticky plan-error word
__global__ void k_check_plan(const uint32_t* __restrict__ counters, PlanWords w, uint32_t* __restrict__ err_word) {
    if (threadIdx.x < w.n && counters[w.idx[threadIdx.x]] != w.want[threadIdx.x]) {
        *err_word = 1u + w.idx[threadIdx.x];
        __threadfence_system();
    }
}
cudaError_t check_plan(ivx_ctx* ctx, const uint32_t* d_counters, const PlanWords& w) {
    k_check_plan<<<1, 32, 0, ctx->stream>>>(d_counters, w, ctx->h_pinned_dev + ivx_ctx::PLAN_ERROR_WORD);
    return cudaGetLastError();
}
// after a stream synchronisation: did any plan check fail since the last look?
int take_plan_error(ivx_ctx* ctx) {
    const uint32_t e = ctx->h_pinned[ivx_ctx::PLAN_ERROR_WORD];
    if (!e) return IVX_OK;
    ctx->h_pinned[ivx_ctx::PLAN_ERROR_WORD] = 0;
    ctx->plans.clear();
    IVX_FAIL(ctx, IVX_ERR_CUDA, "a cached generation plan did not match the device's counters (word %u); the plans were "
             "dropped, repeat the call", e - 1u);
}
GenPlan* lookup_plan(ivx_ctx* ctx, const ivx_program* prog, float voxel_extent, const ivx_type_generator* tg, uint32_t i_begin,
                     uint32_t i_end, bool whole, bool streamed) {
    if (std::getenv("IVX_NO_PLANS")) return nullptr;
    for (auto& pl : ctx->plans)
        if (pl.prog_uid == prog->uid && pl.voxel_extent == voxel_extent && std::memcmp(&pl.types, tg, sizeof(*tg)) == 0 &&
            pl.i_begin == i_begin && pl.i_end == i_end && pl.whole == whole && pl.streamed == streamed)
            return &pl;
    return nullptr;
}
GenPlan* new_plan(ivx_ctx* ctx, const ivx_program* prog, float voxel_extent, const ivx_type_generator* tg, uint32_t i_begin,
                  uint32_t i_end, bool whole, bool streamed) {
    if (ctx->plans.size() >= ivx_ctx::MAX_PLANS) ctx->plans.erase(ctx->plans.begin());
    GenPlan pl;
    pl.serial = ctx->next_serial++;
    pl.prog_uid = prog->uid;
    pl.voxel_extent = voxel_extent;
    pl.types = *tg;
    pl.i_begin = i_begin;
    pl.i_end = i_end;
    pl.whole = whole;
    pl.streamed = streamed;
    ctx->plans.push_back(pl);
    return &ctx->plans.back();
}

// plans are keyed by the program's CONTENT (a host that rebuilds the same graph every frame still hits its plan)
uint64_t program_content_hash(const HostProgram& h) {
    uint64_t x = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; ++i) x = (x ^ b[i]) * 1099511628211ull;
    };
    if (!h.nodes.empty()) mix(h.nodes.data(), h.nodes.size() * sizeof(ivx_node));
    mix(&h.stack_depth, sizeof(h.stack_depth));
    mix(h.domain_lo, sizeof(h.domain_lo));
    mix(h.domain_hi, sizeof(h.domain_hi));
    return x | 1ull;
}

void free_mesh(ivx_ctx* ctx, DeviceMesh& m) {
    ctx->release(m.positions);
    ctx->release(m.normals);
    ctx->release(m.indices);
    ctx->release(m.index_materials);
    ctx->release(m.submeshes);
    ctx->release(m.vertex_ranges);
    const uint64_t serial = m.serial;
    m = DeviceMesh{};
    m.serial = serial;
}

// SDFVoxelGenerator::new (generation.rs:207-258)
void derive_grid(const HostProgram& p, uint32_t grid_shape[3], float shifted_center[3]) {
    float ext[3];
    for (int d = 0; d < 3; ++d) ext[d] = p.domain_hi[d] - p.domain_lo[d];
    if (p.nodes.empty() || ext[0] == 0.0f || ext[1] == 0.0f || ext[2] == 0.0f) {
        for (int d = 0; d < 3; ++d) {
            grid_shape[d] = 0;
            shifted_center[d] = -0.5f;
        }
        return;
    }
    for (int d = 0; d < 3; ++d) {
        float c = std::ceil(ext[d]);
        uint32_t n = c > 0.0f ? (uint32_t)c : 0u;
        grid_shape[d] = n + 2;
        float half = 0.5f * (float)grid_shape[d];
        float dc = 0.5f * (p.domain_lo[d] + p.domain_hi[d]);
        shifted_center[d] = (half - dc) - 0.5f;
    }
}

int upload_program(ivx_ctx* ctx, ivx_program* prog) {
    const auto& nodes = prog->host.nodes;
    std::vector<Instr> root;
    root.reserve(nodes.size());
    for (uint32_t i = 0; i < nodes.size(); ++i) {
        uint32_t op;
        switch (nodes[i].kind) {
            case IVX_SPHERE:
            case IVX_CAPSULE:
            case IVX_BOX: op = OP_LEAF; break;
            case IVX_TRANSLATION:
            case IVX_ROTATION: continue;  // baked into the leaf transforms (atomic.rs:745)
            case IVX_SCALING: op = OP_SCALE; break;
            case IVX_MULTIFRACTAL_NOISE: op = OP_NOISE; break;
            default: op = OP_COMBINE; break;
        }
        root.push_back(Instr{(op << 28) | i, 0.0f});
    }
    prog->root_len = (uint32_t)root.size();
    prog->d_nodes = static_cast<ivx_node*>(ctx->alloc(std::max<size_t>(1, nodes.size()) * sizeof(ivx_node)));
    prog->d_root = static_cast<Instr*>(ctx->alloc(std::max<size_t>(1, root.size()) * sizeof(Instr)));
    prog->d_root_meta = static_cast<uint32_t*>(ctx->alloc(2 * sizeof(uint32_t)));
    if (!prog->d_nodes || !prog->d_root || !prog->d_root_meta) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "program upload: out of device memory");
    if (!nodes.empty())
        CU(ctx, cudaMemcpyAsync(prog->d_nodes, nodes.data(), nodes.size() * sizeof(ivx_node), cudaMemcpyHostToDevice, ctx->stream));
    if (!root.empty())
        CU(ctx, cudaMemcpyAsync(prog->d_root, root.data(), root.size() * sizeof(Instr), cudaMemcpyHostToDevice, ctx->stream));
    uint32_t meta[2] = {0, prog->root_len};
    CU(ctx, cudaMemcpyAsync(prog->d_root_meta, meta, sizeof(meta), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IVX_OK;
}

uint32_t persistent_grid(ivx_ctx* ctx, uint32_t n_work, int blocks_per_sm) {
    uint32_t g = (uint32_t)(ctx->sm_count * std::max(1, blocks_per_sm));
    return std::max(1u, std::min(n_work, g));
}

// Chooses the evaluator's stack placement: levels 0..smem_levels-1 in shared
// memory (16 KiB each), deeper levels in a global spill buffer.
int plan_eval_stack(ivx_ctx* ctx, uint32_t max_depth, Tmp& tmp, uint32_t n_active, EvalArgs& ea, uint32_t& grid) {
    const int need = max_depth > 0 ? (int)max_depth - 1 : 0;
    // most specialised programs need one or two operand levels; keeping the shared-memory share small
    // lets several CTAs share an SM, the rare deeper levels go to an L2-resident spill buffer
#ifndef IVX_EVAL_SMEM_LEVELS
#define IVX_EVAL_SMEM_LEVELS 2
#endif
    const int smem_levels = std::min(need, IVX_EVAL_SMEM_LEVELS);
    ea.smem_levels = smem_levels;
    ea.spill_levels = need - smem_levels;
    int bps = eval_max_blocks_per_sm(smem_levels);
    if (bps < 1) bps = 1;
    grid = persistent_grid(ctx, n_active, bps);
    ea.spill = nullptr;
    if (ea.spill_levels > 0) {
        ea.spill = tmp.get<float>((size_t)grid * ea.spill_levels * 4096);
        if (!ea.spill) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "evaluator spill buffer: out of device memory");
    }
    return IVX_OK;
}

}  // namespace

int ivx_read_words(ivx_ctx* ctx, const uint32_t* d_src, uint32_t n, uint32_t* out) { return read_words(ctx, d_src, n, out); }
uint32_t ivx_persistent_grid(ivx_ctx* ctx, uint32_t n_work, int blocks_per_sm) { return persistent_grid(ctx, n_work, blocks_per_sm); }

namespace {

#ifndef IVX_STREAM_PARTS
#define IVX_STREAM_PARTS 16u  // at most this many parts of at least 3 chunk planes
#endif
// events that are destroyed on every exit path
struct EventList {
    std::vector<cudaEvent_t> ev;
    ~EventList() {
        for (cudaEvent_t e : ev)
            if (e) cudaEventDestroy(e);
    }
    cudaError_t create(size_t n) {
        ev.assign(n, nullptr);
        for (auto& e : ev) {
            cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            if (rc != cudaSuccess) return rc;
        }
        return cudaSuccess;
    }
    cudaEvent_t operator[](size_t i) const { return ev[i]; }
};

// host destination of a streamed generation (ivx_object_generate_streamed)
struct StreamOut {
    ivx_chunk_desc* h_chunks;
    size_t chunk_capacity;
    ivx_voxel* h_voxels;
    size_t voxel_capacity;   // voxels
    uint64_t n_non_uniform;  // out
};

__global__ void k_plane_sum_flags(const uint32_t* __restrict__ flag, uint32_t n, uint32_t plane, uint32_t* __restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && flag[c]) atomicAdd(&out[c / plane], 1u);
}
__global__ void k_plane_sum_stored(const DevChunk* __restrict__ chunks, uint32_t n, uint32_t plane, uint32_t* __restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && chunks[c].kind != 0) atomicAdd(&out[c / plane], 1u);
}

// `plane_stats` (planning only, ivx_program_plane_work): [p] = chunks of local plane p the SDF program was evaluated on,
// [planes + p] = chunks of it that ended up stored (Uniform or NonUniform: the ones that needed voxel types)
int generate_impl(ivx_ctx* ctx, const ivx_program* prog, float voxel_extent, const ivx_type_generator* tg,
                  uint32_t i_begin, uint32_t i_end, bool whole, ivx_object** out, StreamOut* so = nullptr,
                  std::vector<uint32_t>* plane_stats = nullptr) {
    if (!(voxel_extent > 0.0f)) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "voxel_extent must be > 0");
    if (tg->kind > 1) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "unknown voxel type generator kind %u", tg->kind);
    if (tg->kind == 1 && (tg->n_types == 0 || tg->n_types > 255))
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "GradientNoise needs 1..255 voxel types");

    ivx_object* obj = new (std::nothrow) ivx_object();
    if (!obj) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "host allocation failed");
    struct Guard {
        ivx_ctx* c;
        ivx_object* o;
        ~Guard() {
            if (o) ivx_object_free(c, o);
        }
    } guard{ctx, obj};

    GenParams gp{};
    derive_grid(prog->host, gp.grid_shape, gp.shifted_center);
    for (int d = 0; d < 3; ++d) gp.chunk_counts[d] = (gp.grid_shape[d] + 15) / 16;
    if (whole) {
        i_begin = 0;
        i_end = gp.chunk_counts[0];
    }
    if (i_begin > i_end || i_end > gp.chunk_counts[0])
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "chunk plane range [%u, %u) outside [0, %u)", i_begin, i_end, gp.chunk_counts[0]);
    gp.ci_begin = i_begin;
    gp.ci_end = i_end;
    gp.types = *tg;
    // a plan of an earlier generation of the same object: no host round trips below
    const GenPlan* plan = plane_stats ? nullptr : lookup_plan(ctx, prog, voxel_extent, tg, i_begin, i_end, whole, so != nullptr);
    const GenPlan plan_copy = plan ? *plan : GenPlan{};  // ctx->plans may be reshuffled by nested calls
    if (plan) plan = &plan_copy;
    gp.n_nodes = (uint32_t)prog->host.nodes.size();
    gp.stack_depth = prog->host.stack_depth;

    obj->voxel_extent = voxel_extent;
    for (int d = 0; d < 3; ++d) {
        obj->grid_shape[d] = gp.grid_shape[d];
        obj->chunk_counts[d] = gp.chunk_counts[d];
    }
    // a slab keeps one extra chunk plane per inner side for the neighbouring rank's boundary plane
    const uint32_t halo_lo = (!whole && i_end > i_begin && i_begin > 0) ? 1u : 0u;
    const uint32_t halo_hi = (!whole && i_end > i_begin && i_end < gp.chunk_counts[0]) ? 1u : 0u;
    obj->halo_present[0] = halo_lo != 0;
    obj->halo_present[1] = halo_hi != 0;
    obj->first_i = i_begin - halo_lo;
    obj->own_begin = i_begin;
    obj->own_end = i_end;
    obj->nb[0] = (i_end - i_begin) + halo_lo + halo_hi;
    obj->nb[1] = gp.chunk_counts[1];
    obj->nb[2] = gp.chunk_counts[2];
    const uint32_t n = obj->nb[0] * obj->nb[1] * obj->nb[2];
    obj->n_chunks = n;
    if (n == 0) {
        guard.o = nullptr;
        *out = obj;
        return IVX_OK;
    }

    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    obj->d_chunks = static_cast<DevChunk*>(ctx->alloc((size_t)n * sizeof(DevChunk)));
    obj->d_dirty = static_cast<uint8_t*>(ctx->alloc(n));
    if (!obj->d_chunks || !obj->d_dirty) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "chunk table: out of device memory");
    CU(ctx, cudaMemsetAsync(obj->d_dirty, 0, n, st));

    uint32_t* counters = ctx->d_scratch;  // [0] err [1] max_depth [2..7] occ [8] total caps [9] n_active [10] n_slots
    {
        uint32_t init[16] = {0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        CU(ctx, cudaMemcpyAsync(counters, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }

    // ---- program specialisation, coarse to fine ----
    // Conservative folds over nested super-blocks (8³, 4³, 2³ chunks) shorten the program every chunk
    // below them has to look at; the exact fold per chunk then makes the reference's own decisions.
    uint32_t words[16];
    std::vector<uint32_t> sizes;
    {
        const uint32_t mx = std::max(obj->nb[0], std::max(obj->nb[1], obj->nb[2]));
        for (uint32_t sz = 8; sz >= 2; sz >>= 1)
            if (mx >= 2 * sz) sizes.push_back(sz);
    }
    FoldArgs fa{};
    fa.nodes = prog->d_nodes;
    fa.gp = gp;
    fa.first_chunk[0] = obj->first_i;
    fa.first_chunk[1] = fa.first_chunk[2] = 0;
    fa.explicit_origins = nullptr;
    fa.chunks = nullptr;
    fa.max_depth = counters + 1;
    fa.occ = nullptr;
    fa.error_flag = counters;
    fa.prune = 1;
    fa.saturate = 1;
    fa.own_lo = halo_lo;
    fa.own_hi = obj->nb[0] - halo_hi;
    // the "parent" of the coarsest level is the whole program
    const Instr* par_instrs = prog->d_root;
    const uint32_t* par_off = prog->d_root_meta;
    const uint32_t* par_len = prog->d_root_meta + 1;
    uint32_t par_nb[3] = {0, 0, 0};
    uint32_t par_size = 0;
    sizes.push_back(1);  // the exact level
    uint64_t arena_bound = 0;  // host-side upper bound of the previous level's arena, in instructions
    Instr* ch_instrs = nullptr;
    uint32_t* ch_off = nullptr;
    uint32_t* ch_len = nullptr;
    for (size_t li = 0; li < sizes.size(); ++li) {
        const uint32_t sz = sizes[li];
        const bool exact = sz == 1;
        uint32_t lnb[3];
        for (int d = 0; d < 3; ++d) lnb[d] = (obj->nb[d] + sz - 1) / sz;
        const uint32_t nblk = lnb[0] * lnb[1] * lnb[2];
        uint32_t* caps = tmp.get<uint32_t>(nblk);
        uint32_t* off = tmp.get<uint32_t>(nblk);
        uint32_t* len = tmp.get<uint32_t>(nblk);
        if (!caps || !off || !len) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "fold buffers: out of device memory");
        if (par_size == 0) {
            KL(ctx, launch_fill_u32(caps, nblk, prog->root_len, st));
        } else {
            KL(ctx, launch_child_caps(par_len, nblk, lnb, par_nb, par_size / sz, caps, st));
        }
        KL(ctx, launch_exclusive_scan(caps, off, nblk, counters + 8, st));
        // The arena of this level holds sum(caps) instructions. That sum is only known on the device; a host-side
        // bound (every block's cap is at most the root program's length, and at most its parent's share of the parent
        // arena) avoids a host round trip per level as long as it stays small next to HBM — the arena is scratch.
        uint64_t bound = (uint64_t)nblk * prog->root_len;
        if (par_size != 0) {
            const uint64_t r = par_size / sz;
            bound = std::min(bound, arena_bound * r * r * r);
        }
        arena_bound = bound;
        size_t arena_instrs;
        if (bound * sizeof(Instr) <= ((size_t)1 << 30)) {
            arena_instrs = (size_t)bound;
        } else {
            if (int rc = read_words(ctx, counters, 16, words)) return rc;
            if (words[0]) IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "SDF program needs an operand stack deeper than 64");
            arena_instrs = words[8];
            arena_bound = words[8];
        }
        Instr* instrs = tmp.get<Instr>(std::max<size_t>(1, arena_instrs));
        if (!instrs) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "program arena (%zu instructions): out of device memory", arena_instrs);
        fa.n_blocks = nblk;
        for (int d = 0; d < 3; ++d) {
            fa.nb[d] = lnb[d];
            fa.parent_nb[d] = par_nb[d];
        }
        fa.block_chunks = sz;
        fa.ratio = par_size ? par_size / sz : 1;
        fa.parent_instrs = par_instrs;
        fa.parent_off = par_off;
        fa.parent_len = par_len;
        fa.out_instrs = instrs;
        fa.out_off = off;
        fa.out_len = len;
        if (exact) {
            fa.chunks = obj->d_chunks;
            fa.occ = counters + 2;
        }
        KLP(ctx, exact ? 1 : 0, launch_fold(exact, fa, st));
        par_instrs = instrs;
        par_off = off;
        par_len = len;
        for (int d = 0; d < 3; ++d) par_nb[d] = lnb[d];
        par_size = sz;
        if (exact) {
            ch_instrs = instrs;
            ch_off = off;
            ch_len = len;
        }
    }

    // ---- slot planning ----
    uint32_t* active_flag = tmp.get<uint32_t>(n);
    uint32_t* slot_flag = tmp.get<uint32_t>(n);
    uint32_t* active_scan = tmp.get<uint32_t>(n);
    uint32_t* slot_of = whole ? tmp.get<uint32_t>(n) : static_cast<uint32_t*>(ctx->alloc((size_t)n * 4));
    if (!whole) obj->d_slot_of = slot_of;
    uint32_t* active_list = tmp.get<uint32_t>(n);
    if (!active_flag || !slot_flag || !active_scan || !slot_of || !active_list)
        IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "slot planning: out of device memory");
    KL(ctx, launch_plan_slots(obj->d_chunks, n, obj->nb, active_flag, slot_flag, st));
    uint32_t* d_plane_stats = nullptr;
    if (plane_stats) {
        d_plane_stats = tmp.get<uint32_t>(2 * (size_t)obj->nb[0]);
        if (!d_plane_stats) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "plane statistics: out of device memory");
        CU(ctx, cudaMemsetAsync(d_plane_stats, 0, 2 * (size_t)obj->nb[0] * 4, st));
        ctx->launches++;
        k_plane_sum_flags<<<(n + 255) / 256, 256, 0, st>>>(active_flag, n, obj->nb[1] * obj->nb[2], d_plane_stats);
        CU(ctx, cudaGetLastError());
    }
    KL(ctx, launch_exclusive_scan(active_flag, active_scan, n, counters + 9, st));
    KL(ctx, launch_exclusive_scan(slot_flag, slot_of, n, counters + 10, st));
    KL(ctx, launch_scatter_active(active_flag, active_scan, n, active_list, st));
    if (plan) {
        words[0] = 0;
        words[9] = plan->n_active;
        words[10] = plan->n_slots;
        words[1] = plan->max_depth;
    } else {
        if (int rc = read_words(ctx, counters, 16, words)) return rc;
        if (int rc = take_plan_error(ctx)) return rc;
    }
    if (words[0]) IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "SDF program needs an operand stack deeper than 64");
    const uint32_t n_active = words[9], n_slots = words[10], max_depth = words[1];
    if (!plan && std::getenv("IVX_DEBUG")) {
        std::vector<uint32_t> hc(n), hact(n);
        cudaMemcpy(hc.data(), ch_len, n * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hact.data(), active_flag, n * 4, cudaMemcpyDeviceToHost);
        uint64_t sc = 0, sact = 0, mx = 0;
        for (uint32_t c = 0; c < n; ++c)
            if (hact[c]) { sc += hc[c]; sact++; mx = std::max<uint64_t>(mx, hc[c]); }
        std::fprintf(stderr, "[ivx] root_len %u | chunks %u parent arena %u | active %llu mean_len %.1f max %llu | max_depth %u slots %u\n",
                     prog->root_len, n, words[8], (unsigned long long)sact, sact ? (double)sc / sact : 0.0,
                     (unsigned long long)mx, max_depth, n_slots);
    }
    // halo planes get their slots up front so that importing them never moves the pool
    obj->slot_capacity = n_slots + (halo_lo + halo_hi) * obj->nb[1] * obj->nb[2];
    obj->slots_used = n_slots;
    obj->d_voxels = static_cast<unsigned char*>(ctx->alloc(std::max<size_t>(1, (size_t)obj->slot_capacity) * SLOT_BYTES));
    if (!obj->d_voxels)
        IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "voxel storage (%u chunks): out of device memory", obj->slot_capacity);
    KL(ctx, launch_set_reserved_slots(obj->d_chunks, n, slot_flag, slot_of, slot_of, st));

    // ---- evaluate active chunks ----
    EvalArgs ea{};
    ea.neg_zero = -0.0f;
    ea.nodes = prog->d_nodes;
    ea.gp = gp;
    ea.n_active = n_active;
    ea.active = active_list;
    for (int d = 0; d < 3; ++d) ea.nb[d] = obj->nb[d];
    ea.first_i = obj->first_i;
    ea.explicit_origins = nullptr;
    ea.instrs = ch_instrs;
    ea.off = ch_off;
    ea.len = ch_len;
    ea.slot_of = slot_of;
    ea.voxels = obj->d_voxels;
    ea.chunks = obj->d_chunks;
    ea.occ = counters + 2;
    ea.raw_out = nullptr;
    ea.saturate_final_noise = 1;
    uint32_t egrid = 1;
    if (int rc = plan_eval_stack(ctx, max_depth, tmp, n_active, ea, egrid)) return rc;
    TypesArgs ta{};
    ta.gp = gp;
    for (int d = 0; d < 3; ++d) ta.nb[d] = obj->nb[d];
    ta.first_i = obj->first_i;
    ta.slot_of = slot_of;
    ta.voxels = obj->d_voxels;
    ta.chunks = obj->d_chunks;
    ta.occ = counters + 2;
    ta.neg_zero = -0.0f;
    ta.noise_evaluations = ctx->profiling ? ctx->d_counters64 : nullptr;
    const int types_bps = std::max(1, types_max_blocks_per_sm());

    // A streamed generation (ivx_object_generate_streamed) cuts the chunk planes into parts: each part is evaluated,
    // typed, gets its cross-chunk state as soon as the next plane is typed, is packed to the reference's voxel layout
    // and handed to the copy stream, so the device→host transfer of part p runs under the arithmetic of parts > p.
    // The ordinary generation is the one-part case without the pack / copy tail.
    const uint32_t plane = obj->nb[1] * obj->nb[2];
    const uint32_t P = (so && whole) ? std::min<uint32_t>(IVX_STREAM_PARTS, std::max<uint32_t>(1u, obj->nb[0] / 3u)) : 1u;
    std::vector<uint32_t> xb(P + 1), ab(P + 1);
    for (uint32_t q = 0; q <= P; ++q) xb[q] = (uint32_t)((uint64_t)q * obj->nb[0] / P);
    ab[0] = 0;
    ab[P] = n_active;
    if (P > 1 && plan && plan->part_active.size() == P + 1) {
        ab = plan->part_active;
    } else if (P > 1) {
        // active_scan[c] = number of active chunks before chunk c
        for (uint32_t q = 1; q < P; ++q)
            CU(ctx, store_words(ctx, active_scan + (size_t)xb[q] * plane, 32 + q, 1));
        CU(ctx, cudaStreamSynchronize(st));
        for (uint32_t q = 1; q < P; ++q) ab[q] = ctx->h_pinned[32 + q];
    }
    uint32_t* convert_flag = nullptr;
    if (whole) {
        convert_flag = tmp.get<uint32_t>(n);
        if (!convert_flag) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "derive: out of device memory");
    } else {
        // a slab waits for its halo planes: ivx_object_halo_* → ivx_object_slab_classify → ivx_object_slab_finalize
        obj->derive_pending = true;
    }
    uint32_t* pk_flag = nullptr;
    uint32_t* pk_ord = nullptr;
    uint32_t* part_counts = counters + 44;  // [P] NonUniform chunks per packed plane range
    EventList part_done;
    std::vector<uint32_t> part_lo(P), part_hi(P);
    if (so) {
        if (so->chunk_capacity < n) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %u chunk descriptors", n);
        obj->d_stage_voxels = static_cast<ivx_voxel*>(ctx->alloc(std::max<size_t>(1, (size_t)n_slots) * 12288));
        obj->d_stage_chunks = static_cast<ivx_chunk_desc*>(ctx->alloc((size_t)n * sizeof(ivx_chunk_desc)));
        pk_flag = tmp.get<uint32_t>(n);
        pk_ord = tmp.get<uint32_t>(n);
        if (!obj->d_stage_voxels || !obj->d_stage_chunks || !pk_flag || !pk_ord)
            IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "streamed generation: out of device memory");
        CU(ctx, cudaMemsetAsync(part_counts, 0, 16 * sizeof(uint32_t), st));
        CU(ctx, part_done.create(P));
    }
    // Streamed: the evaluation of all parts runs ahead on the compute stream while typing, cross-chunk state and
    // packing of each evaluated part follow on a second, higher-priority compute stream, so the tail of one kernel
    // is filled by the CTAs of the other instead of idling eight times.
    cudaStream_t st2 = so ? ctx->aux_stream : st;
    EventList evaluated;
    if (so) {
        CU(ctx, evaluated.create(P));
    }
    for (uint32_t q = 0; q < P; ++q) {
        const uint32_t cnt = ab[q + 1] - ab[q];
        if (cnt) {
            ea.active = active_list + ab[q];
            ea.n_active = cnt;
            KLP(ctx, 2, launch_eval(ea, std::min(egrid, cnt), st));
        }
        if (so) CU(ctx, cudaEventRecord(evaluated[q], st));
    }
    for (uint32_t q = 0; q < P; ++q) {
        const uint32_t cnt = ab[q + 1] - ab[q];
        if (so) CU(ctx, cudaStreamWaitEvent(st2, evaluated[q], 0));
        if (cnt) {
            ta.active = active_list + ab[q];
            ta.n_active = cnt;
            if (so) {
                KL(ctx, launch_types(ta, persistent_grid(ctx, cnt, types_bps), st2));
            } else {
                KLP(ctx, 7, launch_types(ta, persistent_grid(ctx, cnt, types_bps), st2));
            }
        }
        // ---- cross-chunk derived state of the planes whose six neighbours are typed ----
        const uint32_t lo = q == 0 ? 0u : xb[q] - 1u, hi = q + 1 == P ? obj->nb[0] : xb[q + 1] - 1u;
        part_lo[q] = lo;
        part_hi[q] = std::max(lo, hi);
        if (whole && hi > lo) {
            if (so) {
                KL(ctx, launch_boundary_classify(obj->d_chunks, n, obj->nb, nullptr, convert_flag, lo, hi, st2));
                KL(ctx, launch_boundary_apply(obj->d_chunks, n, obj->nb, nullptr, convert_flag, slot_of, obj->d_voxels, nullptr,
                                              n, lo, hi, persistent_grid(ctx, n, 8), st2));
            } else {
                KLP(ctx, 3, launch_boundary_classify(obj->d_chunks, n, obj->nb, nullptr, convert_flag, lo, hi, st2));
                KLP(ctx, 3, launch_boundary_apply(obj->d_chunks, n, obj->nb, nullptr, convert_flag, slot_of, obj->d_voxels, nullptr,
                                                  n, lo, hi, persistent_grid(ctx, n, 8), st2));
            }
        }
        if (so) {
            if (hi > lo) {
                const uint32_t c0 = lo * plane, nc = (hi - lo) * plane;
                KL(ctx, launch_nonuniform_flags(obj->d_chunks + c0, nc, pk_flag, st2));
                KL(ctx, launch_exclusive_scan(pk_flag, pk_ord, nc, part_counts + q, st2));
                KL(ctx, launch_pack_voxels(obj->d_chunks + c0, nc, pk_ord, part_counts, q, obj->d_voxels, obj->d_stage_voxels,
                                           obj->d_stage_chunks + c0, persistent_grid(ctx, nc, 8), st2));
                k_store_words<<<1, 64, 0, st2>>>(part_counts + q, ctx->h_pinned_dev + 40 + q, 1);
                CU(ctx, cudaGetLastError());
            }
            CU(ctx, cudaEventRecord(part_done[q], st2));
        }
    }
    if (so) {
        CU(ctx, cudaStreamWaitEvent(st, part_done[P - 1], 0));  // later work on the compute stream sees the finished object
    }
    if (so) {
        // the host follows the parts as they finish and queues their transfers; the compute stream never waits
        uint64_t base = 0;
        int rc = IVX_OK;
        for (uint32_t q = 0; q < P; ++q) {
            cudaError_t e = cudaEventSynchronize(part_done[q]);
            if (e == cudaSuccess && part_hi[q] > part_lo[q] && rc == IVX_OK) {
                const uint64_t cnt = ctx->h_pinned[40 + q];
                const size_t c0 = (size_t)part_lo[q] * plane, nc = (size_t)(part_hi[q] - part_lo[q]) * plane;
                if ((base + cnt) * 4096 > so->voxel_capacity) {
                    rc = IVX_ERR_CAPACITY;
                } else {
                    e = cudaMemcpyAsync(so->h_chunks + c0, obj->d_stage_chunks + c0, nc * sizeof(ivx_chunk_desc),
                                        cudaMemcpyDeviceToHost, ctx->copy_stream);
                    if (e == cudaSuccess && cnt)
                        e = cudaMemcpyAsync(reinterpret_cast<unsigned char*>(so->h_voxels) + base * 12288,
                                            reinterpret_cast<unsigned char*>(obj->d_stage_voxels) + base * 12288, cnt * 12288,
                                            cudaMemcpyDeviceToHost, ctx->copy_stream);
                    base += cnt;
                }
            }
            if (e != cudaSuccess && rc == IVX_OK) rc = IVX_ERR_CUDA;
        }
        if (rc == IVX_ERR_CAPACITY) IVX_FAIL(ctx, rc, "streamed generation: voxel buffer too small");
        if (rc != IVX_OK) IVX_FAIL(ctx, rc, "streamed generation: %s", cudaGetErrorString(cudaGetLastError()));
        so->n_non_uniform = base;
    }

    if (plane_stats) {
        ctx->launches++;
        k_plane_sum_stored<<<(n + 255) / 256, 256, 0, st>>>(obj->d_chunks, n, obj->nb[1] * obj->nb[2], d_plane_stats + obj->nb[0]);
        CU(ctx, cudaGetLastError());
        plane_stats->assign(2 * (size_t)obj->nb[0], 0u);
        CU(ctx, cudaMemcpyAsync(plane_stats->data(), d_plane_stats, plane_stats->size() * 4, cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaStreamSynchronize(st));
    }
    if (plan) {
        // the device's counters against the plan, checked at the next synchronisation (take_plan_error)
        PlanWords pw{};
        const uint32_t idx[10] = {0, 1, 9, 10, 2, 3, 4, 5, 6, 7};
        const uint32_t want[10] = {0, plan->max_depth, plan->n_active, plan->n_slots, plan->occ[0], plan->occ[1], plan->occ[2],
                                   plan->occ[3], plan->occ[4], plan->occ[5]};
        pw.n = 10;
        for (int q = 0; q < 10; ++q) {
            pw.idx[q] = idx[q];
            pw.want[q] = want[q];
        }
        CU(ctx, check_plan(ctx, counters, pw));
        for (int q = 0; q < 6; ++q) words[2 + q] = plan->occ[q];
        obj->plan_serial = plan->serial;
    } else {
        if (int rc = read_words(ctx, counters, 16, words)) return rc;
        if (int rc = take_plan_error(ctx)) return rc;
        if (!std::getenv("IVX_NO_PLANS")) {
            GenPlan* np = new_plan(ctx, prog, voxel_extent, tg, i_begin, i_end, whole, so != nullptr);
            np->n_active = n_active;
            np->n_slots = n_slots;
            np->max_depth = max_depth;
            for (int q = 0; q < 6; ++q) np->occ[q] = words[2 + q];
            if (P > 1) np->part_active = ab;
            obj->plan_serial = np->serial;
        }
    }
    const bool any = words[2] != 0xFFFFFFFFu;
    for (int d = 0; d < 3; ++d) {
        obj->occ_voxels[d] = any ? words[2 + d] : 0u;
        obj->occ_voxels[3 + d] = any ? words[5 + d] + 1u : 0u;
    }
    guard.o = nullptr;
    *out = obj;
    return IVX_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// small kernels that only the API layer needs
namespace ivx {

__global__ void k_set_reserved_slots(DevChunk* chunks, uint32_t n, const uint32_t* slot_flag, const uint32_t* slot_scan,
                                     uint32_t* slot_of) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const uint32_t s = slot_flag[c] ? slot_scan[c] : 0xFFFFFFFFu;
    slot_of[c] = s;
    chunks[c].slot = s;
}
cudaError_t launch_set_reserved_slots(DevChunk* chunks, uint32_t n, const uint32_t* slot_flag, const uint32_t* slot_scan,
                                      uint32_t* slot_of, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_set_reserved_slots<<<(n + 255) / 256, 256, 0, st>>>(chunks, n, slot_flag, slot_scan, slot_of);
    return cudaGetLastError();
}

__global__ void k_nonuniform_flags(const DevChunk* chunks, uint32_t n, uint32_t* flag) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) flag[c] = chunks[c].kind == 2 ? 1u : 0u;
}
cudaError_t launch_nonuniform_flags(const DevChunk* chunks, uint32_t n, uint32_t* flag, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_nonuniform_flags<<<(n + 255) / 256, 256, 0, st>>>(chunks, n, flag);
    return cudaGetLastError();
}

// planes → the reference's 3-byte AoS voxels, NonUniform chunks in linear chunk
// order; chunk descriptors with data_offset = that ordinal (object.rs:574-577).
// `ordinal` counts NonUniform chunks inside this call's chunk range; `part_counts[0..part)` are the counts of
// the ranges packed before it (streamed generation), so the global ordinal is their sum + ordinal[c].
__global__ void __launch_bounds__(256) k_pack_voxels(const DevChunk* __restrict__ chunks, uint32_t n,
                                                     const uint32_t* __restrict__ ordinal,
                                                     const uint32_t* __restrict__ part_counts, uint32_t part,
                                                     const unsigned char* __restrict__ voxels, ivx_voxel* __restrict__ out,
                                                     ivx_chunk_desc* __restrict__ out_chunks) {
    __shared__ __align__(16) uint32_t s_out[3 * 1024];  // one chunk of interleaved voxels
    const int tid = threadIdx.x;
    uint32_t base = 0;
    for (uint32_t q = 0; q < part; ++q) base += part_counts[q];
    for (uint32_t c = blockIdx.x; c < n; c += gridDim.x) {
        const DevChunk ch = chunks[c];
        const uint32_t ord = ch.kind == 2 ? base + ordinal[c] : 0u;
        if (tid == 0) {
            ivx_chunk_desc d;
            d.kind = ch.kind;
            d.flags = ch.kind == 2 ? ch.flags : 0;
            for (int q = 0; q < 6; ++q) d.face[q] = ch.kind == 2 ? ch.face[q] : 0;
            d.uniform_voxel.voxel_type = ch.kind == 1 ? ch.u_type : 0;
            d.uniform_voxel.signed_distance = ch.kind == 1 ? ch.u_sd : 0;
            d.uniform_voxel.flags = ch.kind == 1 ? ch.u_flags : 0;
            d._pad = 0;
            d.data_offset = ord;
            out_chunks[c] = d;
        }
        if (ch.kind != 2 || out == nullptr) continue;
        const unsigned char* slot = voxels + (size_t)ch.slot * SLOT_BYTES;
        // this thread's 16 voxels: one 16-byte row of each plane → 48 interleaved bytes {type, sd, flags}
        const uint4 wt = *reinterpret_cast<const uint4*>(slot + PLANE_TYPE + tid * 16);
        const uint4 wd = *reinterpret_cast<const uint4*>(slot + PLANE_SD + tid * 16);
        const uint4 wf = *reinterpret_cast<const uint4*>(slot + PLANE_FLAGS + tid * 16);
        const uint32_t pt[4] = {wt.x, wt.y, wt.z, wt.w}, pd[4] = {wd.x, wd.y, wd.z, wd.w}, pf[4] = {wf.x, wf.y, wf.z, wf.w};
        uint32_t w12[12];
#pragma unroll
        for (int w = 0; w < 12; ++w) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = 4 * w + j, v = b / 3, f = b % 3;
                const uint32_t src = f == 0 ? pt[v >> 2] : (f == 1 ? pd[v >> 2] : pf[v >> 2]);
                acc |= ((src >> (8 * (v & 3))) & 0xFFu) << (8 * j);
            }
            w12[w] = acc;
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
            *reinterpret_cast<uint4*>(&s_out[tid * 12 + 4 * q]) = make_uint4(w12[4 * q], w12[4 * q + 1], w12[4 * q + 2], w12[4 * q + 3]);
        __syncthreads();
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(out) + (size_t)ord * 12288);
#pragma unroll
        for (int q = 0; q < 3; ++q) o[q * 256 + tid] = *reinterpret_cast<const uint4*>(&s_out[(q * 256 + tid) * 4]);
        __syncthreads();
    }
}
cudaError_t launch_pack_voxels(const DevChunk* chunks, uint32_t n, const uint32_t* ordinal, const uint32_t* part_counts,
                               uint32_t part, const unsigned char* voxels, ivx_voxel* out, ivx_chunk_desc* out_chunks,
                               uint32_t grid, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_pack_voxels<<<grid, 256, 0, st>>>(chunks, n, ordinal, part_counts, part, voxels, out, out_chunks);
    return cudaGetLastError();
}

__global__ void k_flag_dirty_exposed(const DevChunk* chunks, const uint8_t* dirty, uint32_t n, uint32_t* exposed_flag,
                                     uint32_t* dirty_flag) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const bool d = dirty[c] != 0;
    const DevChunk ch = chunks[c];
    dirty_flag[c] = d ? 1u : 0u;
    exposed_flag[c] = (d && ch.kind == 2 && (ch.flags & 0x3F) != 0x3F) ? 1u : 0u;
}
cudaError_t launch_flag_dirty_exposed(const DevChunk* chunks, const uint8_t* dirty, uint32_t n, uint32_t* exposed_flag,
                                      uint32_t* dirty_flag, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_flag_dirty_exposed<<<(n + 255) / 256, 256, 0, st>>>(chunks, dirty, n, exposed_flag, dirty_flag);
    return cudaGetLastError();
}

__global__ void k_count_kinds(const DevChunk* chunks, uint32_t n, uint32_t lo, uint32_t hi, uint32_t* out3) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n || c < lo || c >= hi) return;
    atomicAdd(&out3[chunks[c].kind], 1u);
}

}  // namespace ivx

namespace {

// meshes the chunks flagged in `work_flag` (ascending linear order) into `m`.
// `plan`: the mesh counts of an earlier identical call (GenPlan::mesh_counts) — buffers and launches are sized from it
// and the device's own counts are only compared with it afterwards, so the call has no host round trip before its end.
// `sync`: wait for the mesh (and report a plan mismatch) before returning; with false the caller does both later.
int mesh_impl(ivx_ctx* ctx, ivx_object* obj, const uint32_t* work_flag, DeviceMesh& m, const uint32_t* plan = nullptr,
              bool sync = true, uint32_t counts_out[4] = nullptr) {
    free_mesh(ctx, m);
    const uint32_t n = obj->n_chunks;
    if (n == 0) return IVX_OK;
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    uint32_t* counters = ctx->d_scratch;
    uint32_t* scan = tmp.get<uint32_t>(n);
    uint32_t* work = tmp.get<uint32_t>(n);
    if (!scan || !work) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh: out of device memory");
    KL(ctx, launch_exclusive_scan(work_flag, scan, n, counters + 16, st));
    KL(ctx, launch_scatter_active(work_flag, scan, n, work, st));
    uint32_t words[8];
    if (plan) {
        words[0] = plan[0];
    } else {
        if (int rc = read_words(ctx, counters + 16, 1, words)) return rc;
    }
    const uint32_t n_work = words[0];
    m.n_work = n_work;
    if (counts_out) counts_out[0] = n_work, counts_out[1] = counts_out[2] = counts_out[3] = 0;
    if (n_work == 0) {
        if (plan) {
            PlanWords pw{};
            pw.n = 1;
            pw.idx[0] = 16;
            pw.want[0] = 0;
            CU(ctx, check_plan(ctx, counters, pw));
        }
        return IVX_OK;
    }

    uint32_t* vcount = tmp.get<uint32_t>(n_work);
    uint32_t* icount = tmp.get<uint32_t>(n_work);
    uint32_t* hsub = tmp.get<uint32_t>(n_work);
    uint32_t* voff = tmp.get<uint32_t>(n_work);
    uint32_t* ioff = tmp.get<uint32_t>(n_work);
    uint32_t* sord = tmp.get<uint32_t>(n_work);
    if (!vcount || !icount || !hsub || !voff || !ioff || !sord) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh: out of device memory");
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
    const uint32_t grid = persistent_grid(ctx, n_work, 4);
    KLP(ctx, 4, launch_mesh(false, ma, grid, st));
    KL(ctx, launch_exclusive_scan(vcount, voff, n_work, counters + 17, st));
    KL(ctx, launch_exclusive_scan(icount, ioff, n_work, counters + 18, st));
    KL(ctx, launch_exclusive_scan(hsub, sord, n_work, counters + 19, st));
    if (plan) {
        words[0] = plan[1];
        words[1] = plan[2];
        words[2] = plan[3];
        PlanWords pw{};
        pw.n = 4;
        for (uint32_t q = 0; q < 4; ++q) {
            pw.idx[q] = 16 + q;
            pw.want[q] = plan[q];
        }
        CU(ctx, check_plan(ctx, counters, pw));
        ma.cap_vertices = plan[1];
        ma.cap_indices = plan[2];
        ma.cap_submeshes = plan[3];
    } else {
        if (int rc = read_words(ctx, counters + 17, 3, words)) return rc;
    }
    m.n_vertices = words[0];
    m.n_indices = words[1];
    m.n_submeshes = words[2];
    if (counts_out) counts_out[1] = words[0], counts_out[2] = words[1], counts_out[3] = words[2];
    m.positions = static_cast<float*>(ctx->alloc(std::max<size_t>(1, (size_t)m.n_vertices) * 12));
    m.normals = static_cast<float*>(ctx->alloc(std::max<size_t>(1, (size_t)m.n_vertices) * 12));
    m.indices = static_cast<uint32_t*>(ctx->alloc(std::max<size_t>(1, (size_t)m.n_indices) * 4));
    m.index_materials = static_cast<ivx_index_materials*>(ctx->alloc(std::max<size_t>(1, (size_t)m.n_indices) * 8));
    m.submeshes = static_cast<ivx_chunk_submesh*>(ctx->alloc(std::max<size_t>(1, (size_t)m.n_submeshes) * sizeof(ivx_chunk_submesh)));
    m.vertex_ranges = static_cast<uint32_t*>(ctx->alloc(std::max<size_t>(1, (size_t)m.n_submeshes) * 8));
    if (!m.positions || !m.normals || !m.indices || !m.index_materials || !m.submeshes || !m.vertex_ranges)
        IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh buffers: out of device memory");
    ma.vertex_offset = voff;
    ma.index_offset = ioff;
    ma.submesh_ord = sord;
    ma.positions = m.positions;
    ma.normals = m.normals;
    ma.indices = m.indices;
    ma.index_materials = m.index_materials;
    ma.submeshes = m.submeshes;
    ma.vertex_ranges = m.vertex_ranges;
    // room to record every quad as multi-material (scratch; the usual share is a few per cent)
    ma.mq_capacity = m.n_indices / 6u;
    ma.mq_entries = tmp.get<uint4>(std::max<size_t>(1, (size_t)ma.mq_capacity * 3));
    ma.mq_count = counters + 20;
    if (!ma.mq_entries) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh: out of device memory");
    CU(ctx, cudaMemsetAsync(ma.mq_count, 0, 4, st));
    KLP(ctx, 5, launch_mesh(true, ma, grid, st));
    KLP(ctx, 5, launch_mesh_materials(ma.mq_entries, ma.mq_count, ma.mq_capacity, m.index_materials, st));
    if (sync) {
        CU(ctx, cudaStreamSynchronize(st));
        if (int rc = take_plan_error(ctx)) return rc;
    }
    return IVX_OK;
}

void fill_mesh_info(const DeviceMesh& m, ivx_mesh_info* out) {
    out->n_vertices = m.n_vertices;
    out->n_indices = m.n_indices;
    out->n_submeshes = m.n_submeshes;
    out->n_exposed_chunks = m.n_work;
    out->d_positions = m.positions;
    out->d_normals = m.normals;
    out->d_index_materials = m.index_materials;
    out->d_indices = m.indices;
    out->d_submeshes = m.submeshes;
    out->d_vertex_ranges = m.vertex_ranges;
}

}  // namespace
extern "C" void fill_mesh_info_from(const DeviceMesh& m, ivx_mesh_info* out) { fill_mesh_info(m, out); }
namespace {

// grows the voxel pool so that `extra` more slots fit
int ensure_slots(ivx_ctx* ctx, ivx_object* obj, uint32_t extra) {
    if (obj->slots_used + extra <= obj->slot_capacity) return IVX_OK;
    const uint32_t want = std::max(obj->slots_used + extra, obj->slot_capacity + obj->slot_capacity / 2 + 64);
    unsigned char* nv = static_cast<unsigned char*>(ctx->alloc((size_t)want * SLOT_BYTES));
    if (!nv) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "voxel storage growth (%u chunks): out of device memory", want);
    if (obj->slots_used)
        CU(ctx, cudaMemcpyAsync(nv, obj->d_voxels, (size_t)obj->slots_used * SLOT_BYTES, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->release(obj->d_voxels);
    obj->d_voxels = nv;
    obj->slot_capacity = want;
    return IVX_OK;
}

}  // namespace

// ===========================================================================
extern "C" {

uint32_t ivx_abi_version(void) { return IVX_ABI_VERSION; }

int ivx_create(const ivx_config* config, ivx_ctx** out_ctx) {
    if (!config || !out_ctx) return IVX_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    if (config->abi_version != IVX_ABI_VERSION) return IVX_ERR_INVALID_ARGUMENT;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return IVX_ERR_NO_DEVICE;  // no CPU fallback
    }
    if (config->device < 0 || config->device >= count) return IVX_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(config->device) != cudaSuccess) return IVX_ERR_CUDA;
    ivx_ctx* ctx = new (std::nothrow) ivx_ctx();
    if (!ctx) return IVX_ERR_OUT_OF_MEMORY;
    ctx->device = config->device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, config->device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (config->stream) {
        ctx->stream = static_cast<cudaStream_t>(config->stream);
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return IVX_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
            ivx_destroy(ctx);
            return IVX_ERR_CUDA;
        }
    }
    if (cudaHostAlloc(&ctx->h_pinned, 64 * sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&ctx->h_pinned_dev, ctx->h_pinned, 0) != cudaSuccess ||
        cudaMalloc(&ctx->d_scratch, 128 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&ctx->d_counters64, 8 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(ctx->d_counters64, 0, 8 * sizeof(unsigned long long)) != cudaSuccess) {
        ivx_destroy(ctx);
        return IVX_ERR_OUT_OF_MEMORY;
    }
    std::memset(ctx->h_pinned, 0, 64 * sizeof(uint32_t));
    *out_ctx = ctx;
    return IVX_OK;
}

void ivx_destroy(ivx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->aux_stream) {
        cudaStreamSynchronize(ctx->aux_stream);
        cudaStreamDestroy(ctx->aux_stream);
    }
    for (auto& e : ctx->prof_events) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    for (auto& b : ctx->pool) cudaFree(b.ptr);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->d_counters64) cudaFree(ctx->d_counters64);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* ivx_last_error(const ivx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t ivx_kernel_launch_count(const ivx_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ivx_synchronize(ivx_ctx* ctx) {
    if (!ctx) return IVX_ERR_INVALID_ARGUMENT;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->aux_stream));
    CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
    return take_plan_error(ctx);
}

static void drain_profile(ivx_ctx* ctx) {
    for (auto& e : ctx->prof_events) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess && e.id < 12) {
            ctx->prof_ms[e.id] += ms;
            ctx->prof_launches[e.id] += 1;
        }
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    ctx->prof_events.clear();
}

int ivx_profile_enable(ivx_ctx* ctx, int enabled) {
    if (!ctx) return IVX_ERR_INVALID_ARGUMENT;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    drain_profile(ctx);
    ctx->profiling = enabled != 0;
    return IVX_OK;
}
int ivx_profile_reset(ivx_ctx* ctx) {
    if (!ctx) return IVX_ERR_INVALID_ARGUMENT;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    drain_profile(ctx);
    for (int i = 0; i < 12; ++i) {
        ctx->prof_ms[i] = 0;
        ctx->prof_launches[i] = 0;
    }
    CU(ctx, cudaMemsetAsync(ctx->d_counters64, 0, 8 * sizeof(unsigned long long), ctx->stream));
    return IVX_OK;
}
int ivx_profile_counter(ivx_ctx* ctx, uint32_t counter_id, uint64_t* out_value) {
    if (!ctx || !out_value || counter_id >= 8) return IVX_ERR_INVALID_ARGUMENT;
    unsigned long long v = 0;
    CU(ctx, cudaMemcpyAsync(&v, ctx->d_counters64 + counter_id, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *out_value = v;
    return IVX_OK;
}
int ivx_profile_get(ivx_ctx* ctx, uint32_t kernel_id, double* out_total_ms, uint64_t* out_launches) {
    if (!ctx || kernel_id >= 12) return IVX_ERR_INVALID_ARGUMENT;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    drain_profile(ctx);
    if (out_total_ms) *out_total_ms = ctx->prof_ms[kernel_id];
    if (out_launches) *out_launches = ctx->prof_launches[kernel_id];
    return IVX_OK;
}

int ivx_program_build(ivx_ctx* ctx, const ivx_sdf_node* nodes, uint32_t n_nodes, uint32_t root, ivx_program** out) {
    if (!ctx || !out || (n_nodes && !nodes)) return IVX_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    cudaSetDevice(ctx->device);
    ivx_program* p = new (std::nothrow) ivx_program();
    if (!p) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "host allocation failed");
    std::string e = compile_program(nodes, n_nodes, root, p->host);
    if (!e.empty()) {
        delete p;
        IVX_FAIL(ctx, IVX_ERR_GRAPH, "%s", e.c_str());
    }
    if (int rc = upload_program(ctx, p)) {
        ivx_program_free(ctx, p);
        return rc;
    }
    p->uid = program_content_hash(p->host);
    *out = p;
    return IVX_OK;
}

int ivx_program_upload(ivx_ctx* ctx, const ivx_node* nodes, uint32_t n_nodes, uint32_t stack_depth,
                       const float domain_lo[3], const float domain_hi[3], ivx_program** out) {
    if (!ctx || !out || (n_nodes && !nodes) || !domain_lo || !domain_hi) return IVX_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    cudaSetDevice(ctx->device);
    // The list is a post-order program: walk it like the evaluators will (primitives push, transforms and noise replace the
    // top, combinations pop two and push one) instead of trusting the caller — an operator without its operands would make
    // the kernels read below their operand stacks.
    uint32_t depth = 0, max_depth = 0;
    for (uint32_t i = 0; i < n_nodes; ++i) {
        const uint32_t kind = nodes[i].kind;
        if (kind > IVX_INTERSECTION) IVX_FAIL(ctx, IVX_ERR_GRAPH, "Invalid SDF node kind %u", kind);
        const uint32_t operands = kind <= IVX_BOX ? 0u : (kind <= IVX_MULTIFRACTAL_NOISE ? 1u : 2u);
        if (depth < operands) IVX_FAIL(ctx, IVX_ERR_GRAPH, "SDF program node %u (kind %u) has %u of its %u operands", i, kind, depth, operands);
        depth = depth - operands + 1u;
        max_depth = std::max(max_depth, depth);
        if (kind == IVX_MULTIFRACTAL_NOISE && nodes[i].octaves > 64u)
            IVX_FAIL(ctx, IVX_ERR_GRAPH, "SDF program node %u: %u noise octaves", i, nodes[i].octaves);
    }
    if (n_nodes && depth != 1u) IVX_FAIL(ctx, IVX_ERR_GRAPH, "SDF program leaves %u values instead of one", depth);
    ivx_program* p = new (std::nothrow) ivx_program();
    if (!p) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "host allocation failed");
    p->host.nodes.assign(nodes, nodes + n_nodes);
    p->host.stack_depth = std::max(stack_depth, max_depth);
    for (int d = 0; d < 3; ++d) {
        p->host.domain_lo[d] = domain_lo[d];
        p->host.domain_hi[d] = domain_hi[d];
    }
    if (int rc = upload_program(ctx, p)) {
        ivx_program_free(ctx, p);
        return rc;
    }
    p->uid = program_content_hash(p->host);
    *out = p;
    return IVX_OK;
}

int ivx_program_info_get(ivx_ctx* ctx, const ivx_program* p, ivx_program_info* out) {
    if (!ctx || !p || !out) return IVX_ERR_INVALID_ARGUMENT;
    out->node_count = (uint32_t)p->host.nodes.size();
    out->stack_depth = p->host.stack_depth;
    for (int d = 0; d < 3; ++d) {
        out->domain_lo[d] = p->host.domain_lo[d];
        out->domain_hi[d] = p->host.domain_hi[d];
    }
    return IVX_OK;
}

int ivx_program_nodes(ivx_ctx* ctx, const ivx_program* p, ivx_node* out, uint32_t capacity) {
    if (!ctx || !p || (!out && capacity)) return IVX_ERR_INVALID_ARGUMENT;
    if (capacity < p->host.nodes.size()) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %zu nodes", p->host.nodes.size());
    if (!p->host.nodes.empty()) std::memcpy(out, p->host.nodes.data(), p->host.nodes.size() * sizeof(ivx_node));
    return IVX_OK;
}

void ivx_program_free(ivx_ctx* ctx, ivx_program* p) {
    if (!ctx || !p) return;
    ctx->release(p->d_nodes);
    ctx->release(p->d_root);
    ctx->release(p->d_root_meta);
    delete p;
}

int ivx_program_eval_chunks(ivx_ctx* ctx, const ivx_program* prog, const float* origins, uint32_t n_chunks, float* out) {
    if (!ctx || !prog || (n_chunks && (!origins || !out))) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (n_chunks == 0) return IVX_OK;
    if (prog->host.nodes.empty()) {
        for (size_t i = 0; i < (size_t)n_chunks * 4096; ++i) out[i] = 0.02f * 127.0f;  // atomic.rs:642-645
        return IVX_OK;
    }
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    uint32_t* counters = ctx->d_scratch;
    uint32_t init[2] = {0, 0};
    CU(ctx, cudaMemcpyAsync(counters, init, sizeof(init), cudaMemcpyHostToDevice, st));
    float* d_org = tmp.get<float>((size_t)n_chunks * 3);
    float* d_out = tmp.get<float>((size_t)n_chunks * 4096);
    Instr* instrs = tmp.get<Instr>((size_t)n_chunks * std::max(1u, prog->root_len));
    uint32_t* off = tmp.get<uint32_t>(n_chunks);
    uint32_t* len = tmp.get<uint32_t>(n_chunks);
    if (!d_org || !d_out || !instrs || !off || !len) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "eval_chunks: out of device memory");
    std::vector<uint32_t> h(n_chunks);
    for (uint32_t b = 0; b < n_chunks; ++b) h[b] = b * prog->root_len;
    CU(ctx, cudaMemcpyAsync(off, h.data(), n_chunks * 4, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(d_org, origins, (size_t)n_chunks * 12, cudaMemcpyHostToDevice, st));
    FoldArgs fa{};
    fa.nodes = prog->d_nodes;
    fa.gp.n_nodes = (uint32_t)prog->host.nodes.size();
    fa.n_blocks = n_chunks;
    fa.block_chunks = 1;
    fa.explicit_origins = d_org;
    fa.ratio = 1;
    fa.parent_instrs = prog->d_root;
    fa.parent_off = prog->d_root_meta;
    fa.parent_len = prog->d_root_meta + 1;
    fa.out_instrs = instrs;
    fa.out_off = off;
    fa.out_len = len;
    fa.chunks = nullptr;
    fa.max_depth = counters + 1;
    fa.occ = nullptr;
    fa.error_flag = counters;
    fa.prune = 1;     // value-preserving, so the raw distances stay bit-exact
    fa.saturate = 0;  // raw f32 distances are wanted here, not quantised codes
    KL(ctx, launch_fold(true, fa, st));
    uint32_t words[2];
    if (int rc = read_words(ctx, counters, 2, words)) return rc;
    if (words[0]) IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "SDF program needs an operand stack deeper than 64");
    EvalArgs ea{};
    ea.neg_zero = -0.0f;
    ea.nodes = prog->d_nodes;
    ea.n_active = n_chunks;
    ea.active = nullptr;
    ea.explicit_origins = d_org;
    ea.instrs = instrs;
    ea.off = off;
    ea.len = len;
    ea.raw_out = d_out;
    uint32_t grid = 1;
    if (int rc = plan_eval_stack(ctx, words[1], tmp, n_chunks, ea, grid)) return rc;
    KL(ctx, launch_eval(ea, grid, st));
    CU(ctx, cudaMemcpyAsync(out, d_out, (size_t)n_chunks * 4096 * 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return IVX_OK;
}

int ivx_program_eval_blocks(ivx_ctx* ctx, const ivx_program* prog, const float* origins, uint32_t n_blocks, uint32_t size,
                            float* out) {
    if (!ctx || !prog || (n_blocks && (!origins || !out))) return IVX_ERR_INVALID_ARGUMENT;
    if (size != 1 && size != 2) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "block size must be 1 or 2");
    cudaSetDevice(ctx->device);
    if (n_blocks == 0) return IVX_OK;
    if (prog->host.stack_depth > 64) IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "SDF program needs an operand stack deeper than 64");
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    const size_t count = (size_t)n_blocks * size * size * size;
    float* d_org = tmp.get<float>((size_t)n_blocks * 3);
    float* d_out = tmp.get<float>(count);
    if (!d_org || !d_out) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "eval_blocks: out of device memory");
    CU(ctx, cudaMemcpyAsync(d_org, origins, (size_t)n_blocks * 12, cudaMemcpyHostToDevice, st));
    KL(ctx, launch_eval_blocks(prog->d_nodes, (uint32_t)prog->host.nodes.size(), d_org, n_blocks, (int)size, d_out, st));
    CU(ctx, cudaMemcpyAsync(out, d_out, count * 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return IVX_OK;
}

int ivx_object_generate(ivx_ctx* ctx, const ivx_program* program, float voxel_extent, const ivx_type_generator* tg,
                        ivx_object** out) {
    if (!ctx || !program || !tg || !out) return IVX_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    cudaSetDevice(ctx->device);
    return generate_impl(ctx, program, voxel_extent, tg, 0, 0, true, out);
}

int ivx_object_generate_streamed(ivx_ctx* ctx, const ivx_program* program, float voxel_extent, const ivx_type_generator* tg,
                                 ivx_chunk_desc* host_chunks, size_t chunk_capacity, ivx_voxel* host_voxels,
                                 size_t voxel_capacity, ivx_object** out, uint64_t* out_non_uniform_chunks) {
    if (!ctx || !program || !tg || !out || !host_chunks || !host_voxels) return IVX_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (out_non_uniform_chunks) *out_non_uniform_chunks = 0;
    cudaSetDevice(ctx->device);
    StreamOut so{host_chunks, chunk_capacity, host_voxels, voxel_capacity, 0};
    const int rc = generate_impl(ctx, program, voxel_extent, tg, 0, 0, true, out, &so);
    if (rc == IVX_OK && out_non_uniform_chunks) *out_non_uniform_chunks = so.n_non_uniform;
    return rc;
}

// Work estimate per chunk plane for a balanced slab partition: the conservative fold levels of generate_impl
// over the whole grid, then every finest super-block is classified as void / inside / undecided.
namespace ivx {
__global__ void k_plane_work(const Instr* __restrict__ instrs, const uint32_t* __restrict__ off, const uint32_t* __restrict__ len,
                             uint32_t n_blocks, uint3 lnb, uint32_t sz, uint3 nb, uint32_t w_inside, uint32_t w_active,
                             uint32_t* __restrict__ out) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint32_t bk = b % lnb.z, bj = (b / lnb.z) % lnb.y, bi = b / (lnb.z * lnb.y);
    uint32_t w = w_active;
    if (len[b] == 1u) {
        const Instr in = instrs[off[b]];
        if ((in.op_node >> 28) == OP_CONST) {
            if (in.value >= 2.03f) w = 0u;
            else if (in.value <= -2.5601f) w = w_inside;
        }
    }
    if (w == 0u) return;
    const uint32_t cy = min(sz, nb.y - bj * sz), cz = min(sz, nb.z - bk * sz);
    for (uint32_t p = bi * sz; p < min(nb.x, (bi + 1u) * sz); ++p) atomicAdd(&out[p], w * cy * cz);
}
}  // namespace ivx

int ivx_program_plane_work(ivx_ctx* ctx, const ivx_program* prog, float voxel_extent, const ivx_type_generator* tg,
                           uint32_t* out_work, uint32_t capacity, uint32_t* out_planes) {
    if (!ctx || !prog || !tg || !out_planes) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (!(voxel_extent > 0.0f)) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "voxel_extent must be > 0");
    GenParams gp{};
    derive_grid(prog->host, gp.grid_shape, gp.shifted_center);
    uint32_t nb[3];
    for (int d = 0; d < 3; ++d) nb[d] = gp.chunk_counts[d] = (gp.grid_shape[d] + 15) / 16;
    *out_planes = nb[0];
    if (!out_work) return IVX_OK;
    if (capacity < nb[0]) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %u planes", nb[0]);
    // Planning by doing: the first request for a (program, extent, type generator) generates the whole object once on
    // this device and counts, per chunk plane, the chunks the SDF program ran on and the chunks that needed voxel
    // types — the two costs of generation (k_eval ~45 ns, k_types ~30 ns per voxel type per chunk on a B200) — plus a
    // small per-chunk share for the fold and the cross-chunk pass. It is kept with the context (keyed by the program's
    // content), so partitioning the same program again costs nothing. IVX_PLANE_WORK=estimate (or an object whose
    // storage bound does not fit the device) selects the cheap estimate below instead: conservative folds only.
    {
        for (const auto& wp : ctx->work_plans)
            if (wp.prog_uid == prog->uid && wp.voxel_extent == voxel_extent && std::memcmp(&wp.types, tg, sizeof(*tg)) == 0 &&
                wp.work.size() == nb[0]) {
                std::memcpy(out_work, wp.work.data(), nb[0] * 4);
                return IVX_OK;
            }
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const char* mode = std::getenv("IVX_PLANE_WORK");
        const uint64_t bound = (uint64_t)nb[0] * nb[1] * nb[2] * SLOT_BYTES;
        if (!(mode && std::strcmp(mode, "estimate") == 0) && !prog->host.nodes.empty() && bound < free_b / 2) {
            ivx_object* dry = nullptr;
            std::vector<uint32_t> stats;
            if (int rc = generate_impl(ctx, prog, voxel_extent, tg, 0, 0, true, &dry, nullptr, &stats)) return rc;
            ivx_object_free(ctx, dry);
            const uint32_t w_eval = 45u, w_types = tg->kind == 1 ? 30u * std::max(1u, tg->n_types) : 4u, w_chunk = 8u;
            ivx_ctx::WorkPlan wp;
            wp.prog_uid = prog->uid;
            wp.voxel_extent = voxel_extent;
            wp.types = *tg;
            wp.work.resize(nb[0]);
            for (uint32_t p = 0; p < nb[0]; ++p)
                wp.work[p] = 1u + w_eval * stats[p] + w_types * stats[nb[0] + p] + w_chunk * nb[1] * nb[2];
            std::memcpy(out_work, wp.work.data(), nb[0] * 4);
            if (ctx->work_plans.size() >= ivx_ctx::MAX_PLANS) ctx->work_plans.erase(ctx->work_plans.begin());
            ctx->work_plans.push_back(std::move(wp));
            return IVX_OK;
        }
    }
    // relative cost of a chunk (measured on the 1024^3 asteroid: k_types 0.14 us, k_eval 0.18 us per chunk):
    // an inside chunk needs a type per voxel under GradientNoise and nothing under Same; an undecided chunk needs
    // the SDF program and, when it is not void, types
    const bool noise_types = tg->kind == 1;
    const uint32_t w_inside = noise_types ? 10u : 1u, w_active = noise_types ? 19u : 14u;
    const uint32_t n = nb[0] * nb[1] * nb[2];
    std::vector<uint32_t> host(nb[0], 1u);
    std::vector<uint32_t> sizes;
    const uint32_t mx = std::max(nb[0], std::max(nb[1], nb[2]));
    for (uint32_t sz = 8; sz >= 2; sz >>= 1)
        if (mx >= 2 * sz) sizes.push_back(sz);
    if (n == 0 || sizes.empty() || prog->host.nodes.empty()) {
        std::memcpy(out_work, host.data(), nb[0] * 4);
        return IVX_OK;
    }
    gp.ci_begin = 0;
    gp.ci_end = nb[0];
    gp.types = *tg;
    gp.n_nodes = (uint32_t)prog->host.nodes.size();
    gp.stack_depth = prog->host.stack_depth;
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    uint32_t* counters = ctx->d_scratch;
    {
        uint32_t init[16] = {0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        CU(ctx, cudaMemcpyAsync(counters, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    FoldArgs fa{};
    fa.nodes = prog->d_nodes;
    fa.gp = gp;
    fa.first_chunk[0] = fa.first_chunk[1] = fa.first_chunk[2] = 0;
    fa.max_depth = counters + 1;
    fa.error_flag = counters;
    fa.prune = 1;
    fa.saturate = 1;
    fa.own_lo = 0;
    fa.own_hi = nb[0];
    const Instr* par_instrs = prog->d_root;
    const uint32_t* par_off = prog->d_root_meta;
    const uint32_t* par_len = prog->d_root_meta + 1;
    uint32_t par_nb[3] = {0, 0, 0}, par_size = 0, lnb[3] = {0, 0, 0}, nblk = 0;
    uint32_t words[16];
    for (uint32_t sz : sizes) {
        for (int d = 0; d < 3; ++d) lnb[d] = (nb[d] + sz - 1) / sz;
        nblk = lnb[0] * lnb[1] * lnb[2];
        uint32_t* caps = tmp.get<uint32_t>(nblk);
        uint32_t* off = tmp.get<uint32_t>(nblk);
        uint32_t* len = tmp.get<uint32_t>(nblk);
        if (!caps || !off || !len) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "plane work: out of device memory");
        if (par_size == 0) KL(ctx, launch_fill_u32(caps, nblk, prog->root_len, st));
        else KL(ctx, launch_child_caps(par_len, nblk, lnb, par_nb, par_size / sz, caps, st));
        KL(ctx, launch_exclusive_scan(caps, off, nblk, counters + 8, st));
        if (int rc = read_words(ctx, counters, 16, words)) return rc;
        Instr* instrs = tmp.get<Instr>(std::max<size_t>(1, words[8]));
        if (!instrs) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "plane work: out of device memory");
        fa.n_blocks = nblk;
        for (int d = 0; d < 3; ++d) {
            fa.nb[d] = lnb[d];
            fa.parent_nb[d] = par_nb[d];
        }
        fa.block_chunks = sz;
        fa.ratio = par_size ? par_size / sz : 1;
        fa.parent_instrs = par_instrs;
        fa.parent_off = par_off;
        fa.parent_len = par_len;
        fa.out_instrs = instrs;
        fa.out_off = off;
        fa.out_len = len;
        KLP(ctx, 0, launch_fold(false, fa, st));
        par_instrs = instrs;
        par_off = off;
        par_len = len;
        for (int d = 0; d < 3; ++d) par_nb[d] = lnb[d];
        par_size = sz;
    }
    uint32_t* d_work = tmp.get<uint32_t>(nb[0]);
    if (!d_work) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "plane work: out of device memory");
    CU(ctx, cudaMemsetAsync(d_work, 0, nb[0] * 4, st));
    ctx->launches++;
    k_plane_work<<<(nblk + 255) / 256, 256, 0, st>>>(par_instrs, par_off, par_len, nblk, make_uint3(lnb[0], lnb[1], lnb[2]), par_size,
                                                     make_uint3(nb[0], nb[1], nb[2]), w_inside, w_active, d_work);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(host.data(), d_work, nb[0] * 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    if (words[0]) IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "SDF program needs an operand stack deeper than 64");
    for (uint32_t p = 0; p < nb[0]; ++p) out_work[p] = host[p] + 1u;  // every plane costs something
    return IVX_OK;
}

int ivx_object_generate_slab(ivx_ctx* ctx, const ivx_program* program, float voxel_extent, const ivx_type_generator* tg,
                             uint32_t chunk_i_begin, uint32_t chunk_i_end, ivx_object** out) {
    if (!ctx || !program || !tg || !out) return IVX_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    cudaSetDevice(ctx->device);
    return generate_impl(ctx, program, voxel_extent, tg, chunk_i_begin, chunk_i_end, false, out);
}

// ---- slab protocol (multi-GPU) ------------------------------------------------
namespace {
struct HaloPlane {
    uint32_t plane_chunks, own_first, halo_first;
};
// side 0 = towards lower chunk-i, side 1 = towards higher
bool halo_plane(const ivx_object* obj, int side, HaloPlane& hp) {
    if (side < 0 || side > 1 || !obj->halo_present[side]) return false;
    hp.plane_chunks = obj->nb[1] * obj->nb[2];
    const uint32_t lo = obj->own_begin - obj->first_i, hi = obj->own_end - obj->first_i;  // local own planes [lo, hi)
    hp.own_first = (side == 0 ? lo : hi - 1) * hp.plane_chunks;
    hp.halo_first = (side == 0 ? lo - 1 : hi) * hp.plane_chunks;
    return true;
}
}  // namespace

int ivx_object_halo_capacity(ivx_ctx* ctx, const ivx_object* obj, size_t* out_bytes) {
    if (!ctx || !obj || !out_bytes) return IVX_ERR_INVALID_ARGUMENT;
    *out_bytes = halo_message_bytes(obj->nb[1] * obj->nb[2]);
    return IVX_OK;
}

int ivx_object_halo_export(ivx_ctx* ctx, const ivx_object* obj, int side, void* d_buf, size_t capacity, size_t* out_bytes) {
    if (!ctx || !obj || !d_buf || !out_bytes) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    HaloPlane hp;
    if (!halo_plane(obj, side, hp)) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "object has no neighbour slab on side %d", side);
    const size_t bytes = halo_message_bytes(hp.plane_chunks);
    if (capacity < bytes) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "halo export needs %zu bytes, buffer has %zu", bytes, capacity);
    // the layer of the boundary plane that touches the neighbour: i = 0 of my lowest plane, i = 15 of my highest
    KL(ctx, launch_halo_pack(obj->d_chunks, hp.own_first, hp.plane_chunks, side == 0 ? 0u : 15u, obj->d_voxels,
                             static_cast<unsigned char*>(d_buf), ctx->stream));
    *out_bytes = bytes;
    return IVX_OK;
}

int ivx_object_halo_import(ivx_ctx* ctx, ivx_object* obj, int side, const void* d_buf, size_t bytes) {
    if (!ctx || !obj || !d_buf) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    HaloPlane hp;
    if (!halo_plane(obj, side, hp)) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "object has no neighbour slab on side %d", side);
    if (!obj->derive_pending) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "halo import after ivx_object_slab_finalize");
    if (bytes != halo_message_bytes(hp.plane_chunks))
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "halo buffer of %zu bytes does not hold a plane of %u chunks (%zu bytes)", bytes,
                 hp.plane_chunks, halo_message_bytes(hp.plane_chunks));
    // one slot per chunk of the plane was reserved at generation, so importing never moves the pool
    if (int rc = ensure_slots(ctx, obj, hp.plane_chunks)) return rc;
    // my lower halo plane holds the neighbour's highest plane (its layer i = 15), my upper one its layer i = 0
    KL(ctx, launch_halo_unpack(obj->d_chunks, hp.halo_first, hp.plane_chunks, obj->slots_used, side == 0 ? 15u : 0u,
                               obj->d_voxels, static_cast<const unsigned char*>(d_buf), ctx->stream));
    obj->slots_used += hp.plane_chunks;
    return IVX_OK;
}

int ivx_object_slab_classify(ivx_ctx* ctx, ivx_object* obj) {
    if (!ctx || !obj) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (!obj->derive_pending) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "object is not a slab with pending derived state");
    if (obj->n_chunks == 0) return IVX_OK;
    if (!obj->d_convert_flag) obj->d_convert_flag = static_cast<uint32_t*>(ctx->alloc((size_t)obj->n_chunks * 4));
    if (!obj->d_convert_flag) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "derive: out of device memory");
    KLP(ctx, 3, launch_boundary_classify(obj->d_chunks, obj->n_chunks, obj->nb, nullptr, obj->d_convert_flag,
                                         obj->own_begin - obj->first_i, obj->own_end - obj->first_i, ctx->stream));
    return IVX_OK;
}

int ivx_object_halo_kinds_export(ivx_ctx* ctx, const ivx_object* obj, int side, void* d_buf, size_t capacity) {
    if (!ctx || !obj || !d_buf) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    HaloPlane hp;
    if (!halo_plane(obj, side, hp)) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "object has no neighbour slab on side %d", side);
    if (!obj->d_convert_flag) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "call ivx_object_slab_classify first");
    if (capacity < hp.plane_chunks) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "kinds export needs %u bytes", hp.plane_chunks);
    KL(ctx, launch_halo_kinds_pack(obj->d_chunks, obj->d_convert_flag, hp.own_first, hp.plane_chunks,
                                   static_cast<uint8_t*>(d_buf), ctx->stream));
    return IVX_OK;
}

int ivx_object_halo_kinds_import(ivx_ctx* ctx, ivx_object* obj, int side, const void* d_buf, size_t bytes) {
    if (!ctx || !obj || !d_buf) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    HaloPlane hp;
    if (!halo_plane(obj, side, hp)) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "object has no neighbour slab on side %d", side);
    if (bytes < hp.plane_chunks) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "kinds buffer needs %u bytes", hp.plane_chunks);
    KL(ctx, launch_halo_kinds_unpack(obj->d_chunks, hp.halo_first, hp.plane_chunks, static_cast<const uint8_t*>(d_buf),
                                     ctx->stream));
    return IVX_OK;
}

int ivx_object_slab_finalize(ivx_ctx* ctx, ivx_object* obj) {
    if (!ctx || !obj) return IVX_ERR_INVALID_ARGUMENT;
    return ivx_internal_slab_finalize(ctx, obj, true);
}
int ivx_internal_slab_finalize(ivx_ctx* ctx, ivx_object* obj, bool sync) {
    cudaSetDevice(ctx->device);
    if (!obj->derive_pending) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "object is not a slab with pending derived state");
    if (obj->n_chunks) {
        if (!obj->d_convert_flag)
            if (int rc = ivx_object_slab_classify(ctx, obj)) return rc;
        KLP(ctx, 3, launch_boundary_apply(obj->d_chunks, obj->n_chunks, obj->nb, nullptr, obj->d_convert_flag, obj->d_slot_of,
                                          obj->d_voxels, nullptr, obj->n_chunks, obj->own_begin - obj->first_i,
                                          obj->own_end - obj->first_i, persistent_grid(ctx, obj->n_chunks, 8), ctx->stream));
        if (sync) CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // (released blocks are reused by later work on this same stream only, so no wait is needed for that)
    ctx->release(obj->d_slot_of);
    ctx->release(obj->d_convert_flag);
    obj->d_slot_of = obj->d_convert_flag = nullptr;
    obj->derive_pending = false;
    return IVX_OK;
}

int ivx_object_info_get(ivx_ctx* ctx, const ivx_object* obj, ivx_object_info* out) {
    if (!ctx || !obj || !out) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    std::memset(out, 0, sizeof(*out));
    out->voxel_extent = obj->voxel_extent;
    for (int d = 0; d < 3; ++d) {
        out->grid_shape[d] = obj->grid_shape[d];
        out->chunk_counts[d] = obj->chunk_counts[d];
    }
    out->chunk_i_begin = obj->own_begin;
    out->chunk_i_end = obj->own_end;
    if (obj->n_chunks) {
        uint32_t* c3 = ctx->d_scratch + 24;
        CU(ctx, cudaMemsetAsync(c3, 0, 12, ctx->stream));
        ctx->launches++;
        const uint32_t plane = obj->nb[1] * obj->nb[2];
        k_count_kinds<<<(obj->n_chunks + 255) / 256, 256, 0, ctx->stream>>>(
            obj->d_chunks, obj->n_chunks, (obj->own_begin - obj->first_i) * plane, (obj->own_end - obj->first_i) * plane, c3);
        CU(ctx, cudaGetLastError());
        uint32_t w[3];
        if (int rc = read_words(ctx, c3, 3, w)) return rc;
        out->n_void = w[0];
        out->n_uniform = w[1];
        out->n_non_uniform = w[2];
    }
    for (int d = 0; d < 3; ++d) {
        out->occupied_voxel_ranges[2 * d] = obj->occ_voxels[d];
        out->occupied_voxel_ranges[2 * d + 1] = obj->occ_voxels[3 + d];
        out->occupied_chunk_ranges[2 * d] = obj->occ_voxels[d] / 16;
        out->occupied_chunk_ranges[2 * d + 1] = (obj->occ_voxels[3 + d] + 15) / 16;
    }
    out->device_bytes = (uint64_t)obj->slot_capacity * SLOT_BYTES + (uint64_t)obj->n_chunks * (sizeof(DevChunk) + 1);
    return IVX_OK;
}

int ivx_object_download(ivx_ctx* ctx, const ivx_object* obj, ivx_chunk_desc* chunks, size_t chunk_capacity,
                        ivx_voxel* voxels, size_t voxel_capacity) {
    if (!ctx || !obj) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    // a slab downloads its own planes; halo planes belong to the neighbouring ranks
    const uint32_t plane = obj->nb[1] * obj->nb[2];
    const uint32_t n = (obj->own_end - obj->own_begin) * plane;
    const DevChunk* own_chunks = obj->d_chunks + (size_t)(obj->own_begin - obj->first_i) * plane;
    if (n == 0) return IVX_OK;
    if (!chunks || chunk_capacity < n) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %u chunk descriptors", n);
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    uint32_t* flag = tmp.get<uint32_t>(n);
    uint32_t* ord = tmp.get<uint32_t>(n);
    ivx_chunk_desc* d_desc = tmp.get<ivx_chunk_desc>(n);
    if (!flag || !ord || !d_desc) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "download: out of device memory");
    KL(ctx, launch_nonuniform_flags(own_chunks, n, flag, st));
    KL(ctx, launch_exclusive_scan(flag, ord, n, ctx->d_scratch + 28, st));
    uint32_t nnu;
    if (int rc = read_words(ctx, ctx->d_scratch + 28, 1, &nnu)) return rc;
    ivx_voxel* d_vox = nullptr;
    if (voxels) {
        if (voxel_capacity < (size_t)nnu * 4096) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %zu voxels", (size_t)nnu * 4096);
        d_vox = tmp.get<ivx_voxel>(std::max<size_t>(1, (size_t)nnu * 4096));
        if (!d_vox) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "download: out of device memory");
    }
    KL(ctx, launch_pack_voxels(own_chunks, n, ord, nullptr, 0, obj->d_voxels, d_vox, d_desc, persistent_grid(ctx, n, 8), st));
    CU(ctx, cudaMemcpyAsync(chunks, d_desc, (size_t)n * sizeof(ivx_chunk_desc), cudaMemcpyDeviceToHost, st));
    if (d_vox && nnu) CU(ctx, cudaMemcpyAsync(voxels, d_vox, (size_t)nnu * 4096 * sizeof(ivx_voxel), cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return IVX_OK;
}

// ivx_object_download with the device→host transfer left running on the context's copy stream: the voxels are packed
// into a staging block on the compute stream, the copies follow on the copy stream, and later work on the compute stream
// (meshing, the mesh gather) overlaps them. The host buffers are complete after ivx_synchronize.
int ivx_object_download_async(ivx_ctx* ctx, ivx_object* obj, ivx_chunk_desc* chunks, size_t chunk_capacity, ivx_voxel* voxels,
                              size_t voxel_capacity, uint64_t* out_non_uniform_chunks) {
    if (!ctx || !obj) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (out_non_uniform_chunks) *out_non_uniform_chunks = 0;
    const uint32_t plane = obj->nb[1] * obj->nb[2];
    const uint32_t n = (obj->own_end - obj->own_begin) * plane;
    const DevChunk* own_chunks = obj->d_chunks + (size_t)(obj->own_begin - obj->first_i) * plane;
    if (n == 0) return IVX_OK;
    if (!chunks || chunk_capacity < n) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %u chunk descriptors", n);
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    uint32_t* flag = tmp.get<uint32_t>(n);
    uint32_t* ord = tmp.get<uint32_t>(n);
    if (!flag || !ord) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "download: out of device memory");
    KL(ctx, launch_nonuniform_flags(own_chunks, n, flag, st));
    KL(ctx, launch_exclusive_scan(flag, ord, n, ctx->d_scratch + 28, st));
    uint32_t nnu;
    if (int rc = read_words(ctx, ctx->d_scratch + 28, 1, &nnu)) return rc;
    if (voxels && voxel_capacity < (size_t)nnu * 4096) IVX_FAIL(ctx, IVX_ERR_CAPACITY, "need room for %zu voxels", (size_t)nnu * 4096);
    // the staging blocks belong to the object (freed with it): the copy stream reads them after this call returns
    CU(ctx, cudaStreamSynchronize(ctx->copy_stream));  // an earlier transfer out of the old staging blocks
    ctx->release(obj->d_stage_voxels);
    ctx->release(obj->d_stage_chunks);
    obj->d_stage_voxels = voxels ? static_cast<ivx_voxel*>(ctx->alloc(std::max<size_t>(1, (size_t)nnu) * 12288)) : nullptr;
    obj->d_stage_chunks = static_cast<ivx_chunk_desc*>(ctx->alloc((size_t)n * sizeof(ivx_chunk_desc)));
    if ((voxels && !obj->d_stage_voxels) || !obj->d_stage_chunks) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "download: out of device memory");
    KL(ctx, launch_pack_voxels(own_chunks, n, ord, nullptr, 0, obj->d_voxels, obj->d_stage_voxels, obj->d_stage_chunks,
                               persistent_grid(ctx, n, 8), st));
    EventList packed;
    CU(ctx, packed.create(1));
    CU(ctx, cudaEventRecord(packed[0], st));
    CU(ctx, cudaStreamWaitEvent(ctx->copy_stream, packed[0], 0));
    CU(ctx, cudaMemcpyAsync(chunks, obj->d_stage_chunks, (size_t)n * sizeof(ivx_chunk_desc), cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (voxels && nnu)
        CU(ctx, cudaMemcpyAsync(voxels, obj->d_stage_voxels, (size_t)nnu * 12288, cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (out_non_uniform_chunks) *out_non_uniform_chunks = nnu;
    return IVX_OK;
}

void ivx_object_free(ivx_ctx* ctx, ivx_object* obj) {
    if (!ctx || !obj) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->aux_stream);
    cudaStreamSynchronize(ctx->copy_stream);
    free_mesh(ctx, obj->mesh);
    ivx_mesh_sync_free(obj->sync);
    obj->sync = nullptr;
    ctx->release(obj->d_stage_voxels);
    ctx->release(obj->d_stage_chunks);
    ctx->release(obj->d_chunks);
    ctx->release(obj->d_voxels);
    ctx->release(obj->d_dirty);
    ctx->release(obj->d_slot_of);
    ctx->release(obj->d_convert_flag);
    ctx->release(obj->d_labels);
    ctx->release(obj->d_regions);
    ctx->release(obj->d_label_stale);
    ctx->release(obj->d_region_first);
    ctx->release(obj->d_region_label);
    ivx_probes_free(ctx, obj->probes);
    ctx->release(obj->d_region_root);
    delete obj;
}

// ivx_object_mesh, optionally without the final synchronisation (comm.cu queues the gather behind the mesh kernels)
int ivx_internal_mesh(ivx_ctx* ctx, ivx_object* obj, bool sync, uint32_t counts[4], ivx_mesh_info* out) {
    std::memset(out, 0, sizeof(*out));
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    if (obj->n_chunks == 0) return IVX_OK;
    if (obj->derive_pending)
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "slab object: call ivx_object_slab_finalize before meshing");
    // VoxelObjectMesh::recreate clears the chunk submesh manager (mesh.rs:286-296)
    ivx_mesh_sync_free(obj->sync);
    obj->sync = nullptr;
    obj->mesh_is_patch = false;
    obj->mesh.serial++;
    Tmp tmp(ctx);
    uint32_t* flag = tmp.get<uint32_t>(obj->n_chunks);
    if (!flag) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mesh: out of device memory");
    KL(ctx, launch_exposed_flags(obj->d_chunks, obj->n_chunks, obj->nb, obj->own_begin - obj->first_i,
                                 obj->own_end - obj->first_i, flag, ctx->stream));
    // an object that has not been modified since its generation meshes exactly like the last object generated from
    // the same plan (for a slab: given the same neighbours — the device's counts are checked against the plan)
    GenPlan* gp = obj->plan_serial ? ctx->find_plan(obj->plan_serial) : nullptr;
    uint32_t plan_counts[4];
    const bool planned = gp && gp->mesh_valid;
    if (planned) std::memcpy(plan_counts, gp->mesh_counts, sizeof(plan_counts));
    int rc = mesh_impl(ctx, obj, flag, obj->mesh, planned ? plan_counts : nullptr, sync || !planned, counts);
    if (rc != IVX_OK && planned && sync) {
        // stale plan (a slab whose neighbours changed): once more, from the device's own counts
        if ((gp = ctx->find_plan(obj->plan_serial))) gp->mesh_valid = false;
        rc = mesh_impl(ctx, obj, flag, obj->mesh, nullptr, true, counts);
    }
    if (rc) return rc;
    if (!planned && (gp = obj->plan_serial ? ctx->find_plan(obj->plan_serial) : nullptr)) {
        gp->mesh_valid = true;
        std::memcpy(gp->mesh_counts, counts, sizeof(uint32_t) * 4);
    }
    fill_mesh_info(obj->mesh, out);
    return IVX_OK;
}
// after the caller's own stream synchronisation: a plan mismatch flagged by the queued checks
int ivx_internal_take_plan_error(ivx_ctx* ctx, ivx_object* obj) {
    const int rc = take_plan_error(ctx);
    if (rc && obj) obj->plan_serial = 0;
    return rc;
}

int ivx_object_mesh(ivx_ctx* ctx, ivx_object* obj, ivx_mesh_info* out) {
    if (!ctx || !obj || !out) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    uint32_t counts[4];
    return ivx_internal_mesh(ctx, obj, true, counts, out);
}

int ivx_mesh_download_checked(ivx_ctx* ctx, const ivx_object* obj, uint32_t n_vertices, uint32_t n_indices, uint32_t n_submeshes,
                              float* positions, float* normals, ivx_index_materials* index_materials, uint32_t* indices,
                              ivx_chunk_submesh* submeshes, uint32_t* vertex_ranges) {
    if (!ctx || !obj) return IVX_ERR_INVALID_ARGUMENT;
    const DeviceMesh& m = obj->mesh;
    if (m.n_vertices != n_vertices || m.n_indices != n_indices || m.n_submeshes != n_submeshes)
        IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "the object's mesh (%u vertices, %u indices, %u submeshes) is not the one these buffers "
                 "were sized for (%u, %u, %u): it was re-created or patched since", m.n_vertices, m.n_indices, m.n_submeshes,
                 n_vertices, n_indices, n_submeshes);
    return ivx_mesh_download(ctx, obj, positions, normals, index_materials, indices, submeshes, vertex_ranges);
}

int ivx_mesh_download(ivx_ctx* ctx, const ivx_object* obj, float* positions, float* normals,
                      ivx_index_materials* index_materials, uint32_t* indices, ivx_chunk_submesh* submeshes,
                      uint32_t* vertex_ranges) {
    if (!ctx || !obj) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    const DeviceMesh& m = obj->mesh;
    cudaStream_t st = ctx->stream;
    if (positions && m.n_vertices) CU(ctx, cudaMemcpyAsync(positions, m.positions, (size_t)m.n_vertices * 12, cudaMemcpyDeviceToHost, st));
    if (normals && m.n_vertices) CU(ctx, cudaMemcpyAsync(normals, m.normals, (size_t)m.n_vertices * 12, cudaMemcpyDeviceToHost, st));
    if (index_materials && m.n_indices) CU(ctx, cudaMemcpyAsync(index_materials, m.index_materials, (size_t)m.n_indices * 8, cudaMemcpyDeviceToHost, st));
    if (indices && m.n_indices) CU(ctx, cudaMemcpyAsync(indices, m.indices, (size_t)m.n_indices * 4, cudaMemcpyDeviceToHost, st));
    if (submeshes && m.n_submeshes) CU(ctx, cudaMemcpyAsync(submeshes, m.submeshes, (size_t)m.n_submeshes * sizeof(ivx_chunk_submesh), cudaMemcpyDeviceToHost, st));
    if (vertex_ranges && m.n_submeshes) CU(ctx, cudaMemcpyAsync(vertex_ranges, m.vertex_ranges, (size_t)m.n_submeshes * 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return IVX_OK;
}

// voxel ranges given by the caller instead of derived from a shape (modify_voxels_within_ranges), with the closure's inputs
struct ExplicitRanges {
    uint32_t v0[3], v1[3];
    ivx::MutualArgs mutual;
};
static int absorb_impl(ivx_ctx* ctx, ivx_object* obj, const ivx::AbsorbShape& shape, ivx_absorb_stats* out_stats,
                       const InertialUpdate* upd = nullptr, const ExplicitRanges* er = nullptr);

// ---- mesh gather over peer memory (multi-GPU) ------------------------------------------------------------
int ivx_peer_alloc(ivx_ctx* ctx, size_t bytes, void** out_ptr, unsigned char out_handle[64]) {
    if (!ctx || !out_ptr || !out_handle) return IVX_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ivx_peer_alloc hands out 64-byte handles");
    cudaSetDevice(ctx->device);
    *out_ptr = nullptr;
    void* p = nullptr;
    // a dedicated cudaMalloc (not the context pool): an IPC handle names a whole allocation
    CU(ctx, cudaMalloc(&p, std::max<size_t>(bytes, 256)));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        cudaGetLastError();
        IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    std::memcpy(out_handle, &h, 64);
    *out_ptr = p;
    return IVX_OK;
}
int ivx_peer_free(ivx_ctx* ctx, void* ptr) {
    if (!ctx) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (ptr) CU(ctx, cudaFree(ptr));
    return IVX_OK;
}
int ivx_peer_open(ivx_ctx* ctx, const unsigned char handle[64], void** out_ptr) {
    if (!ctx || !handle || !out_ptr) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    *out_ptr = nullptr;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    }
    *out_ptr = p;
    return IVX_OK;
}
int ivx_peer_close(ivx_ctx* ctx, void* ptr) {
    if (!ctx) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (ptr) CU(ctx, cudaIpcCloseMemHandle(ptr));
    return IVX_OK;
}

int ivx_mesh_push(ivx_ctx* ctx, const ivx_object* obj, void* dst_base, const uint64_t field_offsets[6], uint32_t vertex_base,
                  uint32_t index_base, uint32_t submesh_base) {
    if (!ctx || !obj || !dst_base || !field_offsets) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    const DeviceMesh& m = obj->mesh;
    unsigned char* d = static_cast<unsigned char*>(dst_base);
    cudaStream_t st = ctx->stream;
    const uint32_t grid = (uint32_t)ctx->sm_count * 4u;
    const size_t nv = m.n_vertices, ni = m.n_indices, ns = m.n_submeshes;
    static_assert(sizeof(ivx_chunk_submesh) == 52, "13 words per submesh");
    // positions, normals, indices (+ vertex base), index materials, submeshes (index_offset + index base),
    // vertex ranges (+ vertex base): the layout of VoxelObjectMesh (mesh.rs:50-103), concatenated in slab order
    KL(ctx, launch_push_words(m.positions, d + field_offsets[0] + (size_t)vertex_base * 12, nv * 3, 0, 0, 0, grid, st));
    KL(ctx, launch_push_words(m.normals, d + field_offsets[1] + (size_t)vertex_base * 12, nv * 3, 0, 0, 0, grid, st));
    KL(ctx, launch_push_words(m.indices, d + field_offsets[2] + (size_t)index_base * 4, ni, vertex_base, 1, 0, grid, st));
    KL(ctx, launch_push_words(m.index_materials, d + field_offsets[3] + (size_t)index_base * 8, ni * 2, 0, 0, 0, grid, st));
    KL(ctx, launch_push_words(m.submeshes, d + field_offsets[4] + (size_t)submesh_base * 52, ns * 13, index_base, 13, 3, grid, st));
    KL(ctx, launch_push_words(m.vertex_ranges, d + field_offsets[5] + (size_t)submesh_base * 8, ns * 2, vertex_base, 1, 0, grid, st));
    return IVX_OK;
}

static ivx::AbsorbShape sphere_shape(const float center[3], float radius, float influence_radius) {
    ivx::AbsorbShape s{};
    s.capsule = 0;
    for (int d = 0; d < 3; ++d) s.center[d] = center[d];
    s.radius = radius;
    s.influence_radius = influence_radius;
    s.influence_radius_sq = influence_radius * influence_radius;  // Sphere::radius_squared = radius.powi(2)
    return s;
}
static ivx::AbsorbShape capsule_shape(const float segment_start[3], const float segment_vector[3], float radius,
                                      float influence_radius) {
    ivx::AbsorbShape s{};
    s.capsule = 1;
    for (int d = 0; d < 3; ++d) {
        s.center[d] = segment_start[d];
        s.seg[d] = segment_vector[d];
    }
    // Capsule::create_point_containment_tester (capsule.rs:168-181)
    const float len2 = (s.seg[0] * s.seg[0] + s.seg[1] * s.seg[1]) + s.seg[2] * s.seg[2];
    for (int d = 0; d < 3; ++d) s.seg_over_len2[d] = len2 > 1e-8f ? s.seg[d] / len2 : 0.0f;
    s.radius = radius;
    s.influence_radius = influence_radius;
    s.influence_radius_sq = influence_radius * influence_radius;
    return s;
}

int ivx_object_absorb_sphere(ivx_ctx* ctx, ivx_object* obj, const float center[3], float radius, float influence_radius,
                             ivx_absorb_stats* out_stats) {
    if (!ctx || !obj || !center) return IVX_ERR_INVALID_ARGUMENT;
    if (!(radius >= 0.0f) || !(influence_radius >= 0.0f)) return IVX_ERR_INVALID_ARGUMENT;  // Sphere::new asserts
    return absorb_impl(ctx, obj, sphere_shape(center, radius, influence_radius), out_stats);
}

int ivx_object_absorb_sphere_inertial(ivx_ctx* ctx, ivx_object* obj, const float center[3], float radius,
                                      float influence_radius, const float* voxel_type_densities, uint32_t n_densities,
                                      ivx_inertial_moments* inout_moments, ivx_absorb_stats* out_stats) {
    if (!ctx || !obj || !center || !inout_moments || (!voxel_type_densities && n_densities) || n_densities > 256)
        return IVX_ERR_INVALID_ARGUMENT;
    if (!(radius >= 0.0f) || !(influence_radius >= 0.0f)) return IVX_ERR_INVALID_ARGUMENT;
    const InertialUpdate upd{voxel_type_densities, n_densities, inout_moments};
    return absorb_impl(ctx, obj, sphere_shape(center, radius, influence_radius), out_stats, &upd);
}

int ivx_object_absorb_capsule_inertial(ivx_ctx* ctx, ivx_object* obj, const float segment_start[3],
                                       const float segment_vector[3], float radius, float influence_radius,
                                       const float* voxel_type_densities, uint32_t n_densities,
                                       ivx_inertial_moments* inout_moments, ivx_absorb_stats* out_stats) {
    if (!ctx || !obj || !segment_start || !segment_vector || !inout_moments || (!voxel_type_densities && n_densities) ||
        n_densities > 256)
        return IVX_ERR_INVALID_ARGUMENT;
    if (!(radius >= 0.0f) || !(influence_radius >= 0.0f)) return IVX_ERR_INVALID_ARGUMENT;
    const InertialUpdate upd{voxel_type_densities, n_densities, inout_moments};
    return absorb_impl(ctx, obj, capsule_shape(segment_start, segment_vector, radius, influence_radius), out_stats, &upd);
}

int ivx_objects_absorb_mutually(ivx_ctx* ctx, ivx_object* a, ivx_object* b, const ivx_isometry* transform_from_b_to_a,
                                float smoothness, const uint32_t ranges_in_a[6], const uint32_t ranges_in_b[6],
                                const float* voxel_type_densities, uint32_t n_densities, ivx_inertial_moments* inout_a,
                                ivx_inertial_moments* inout_b, ivx_absorb_stats* stats_a, ivx_absorb_stats* stats_b) {
    if (!ctx || !a || !b || a == b || !transform_from_b_to_a || !ranges_in_a || !ranges_in_b) return IVX_ERR_INVALID_ARGUMENT;
    const bool inertial = inout_a || inout_b;
    if (inertial && (!inout_a || !inout_b || (!voxel_type_densities && n_densities) || n_densities > 256))
        return IVX_ERR_INVALID_ARGUMENT;
    if (!(smoothness >= 0.0f)) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (stats_a) std::memset(stats_a, 0, sizeof(*stats_a));
    if (stats_b) std::memset(stats_b, 0, sizeof(*stats_b));
    if (b->first_i != 0 || b->nb[0] != b->chunk_counts[0] || a->first_i != 0 || a->nb[0] != a->chunk_counts[0])
        IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "mutual absorption on a slab-partitioned object is not supported");
    const float ea = a->voxel_extent, eb = b->voxel_extent;
    const float inv_ea = 1.0f / ea, inv_eb = 1.0f / eb;  // inverse_voxel_extent = voxel_extent.recip() (object.rs:348)
    const float b_dist_to_a = eb * inv_ea, a_dist_to_b = ea * inv_eb;
    // the snapshot of A's signed distances: the intersection ranges padded by ceil(b_dist_to_a) voxels (absorption.rs:927-944)
    const float pad_f = std::ceil(b_dist_to_a);
    const uint32_t pad = !(pad_f > 0.0f) ? 0u : (pad_f >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)pad_f);
    ExplicitRanges ra{}, rb{};
    uint64_t n_snap = 1;
    for (int d = 0; d < 3; ++d) {
        const uint32_t dim_a = a->chunk_counts[d] * 16u;
        if (ranges_in_a[2 * d + 1] > dim_a || ranges_in_b[2 * d + 1] > b->chunk_counts[d] * 16u)
            IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "intersection voxel range beyond the grid");
        ra.v0[d] = ranges_in_a[2 * d] > pad ? ranges_in_a[2 * d] - pad : 0u;
        ra.v1[d] = (uint32_t)std::min<uint64_t>((uint64_t)ranges_in_a[2 * d + 1] + pad, dim_a);
        rb.v0[d] = ranges_in_b[2 * d];
        rb.v1[d] = ranges_in_b[2 * d + 1];
        n_snap *= ra.v1[d] > ra.v0[d] ? ra.v1[d] - ra.v0[d] : 0u;
    }
    if (n_snap > 0xFFFFFFFFull) IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "snapshot of more than 2^32 voxels");
    Tmp tmp(ctx);
    float* snapshot = tmp.get<float>((size_t)std::max<uint64_t>(n_snap, 1));
    if (!snapshot) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "mutual absorption: out of device memory for the snapshot");
    {
        const float max_f32 = 127.0f * 0.02f;  // VoxelSignedDistance::MAX_F32
        uint32_t bits;
        std::memcpy(&bits, &max_f32, 4);
        if (n_snap) KL(ctx, launch_fill_u32(reinterpret_cast<uint32_t*>(snapshot), (uint32_t)n_snap, bits, ctx->stream));
    }
    ivx::MutualArgs m{};
    for (int q = 0; q < 4; ++q) m.q[q] = transform_from_b_to_a->rotation[q];
    for (int d = 0; d < 3; ++d) {
        m.t[d] = transform_from_b_to_a->translation[d];
        m.s0[d] = ra.v0[d];
        m.s1[d] = ra.v1[d];
        m.o_nb[d] = b->nb[d];
    }
    m.smoothness = smoothness;
    m.qik = 0.25f / smoothness;
    m.snapshot = snapshot;
    m.o_chunks = b->d_chunks;
    m.o_voxels = b->d_voxels;
    ivx::AbsorbShape shape{};
    // object A: every voxel of the padded ranges that is not maximally outside samples B's signed distance field
    m.extent = ea;
    m.inv_extent_other = inv_eb;
    m.dist_scale = b_dist_to_a;
    ra.mutual = m;
    shape.capsule = 2;
    const InertialUpdate ua{voxel_type_densities, n_densities, inout_a}, ub{voxel_type_densities, n_densities, inout_b};
    if (int rc = absorb_impl(ctx, a, shape, stats_a, inertial ? &ua : nullptr, &ra)) return rc;
    // object B: samples A's snapshot
    m.extent = eb;
    m.inv_extent_other = inv_ea;
    m.dist_scale = a_dist_to_b;
    rb.mutual = m;
    shape.capsule = 3;
    return absorb_impl(ctx, b, shape, stats_b, inertial ? &ub : nullptr, &rb);
}

int ivx_object_absorb_capsule(ivx_ctx* ctx, ivx_object* obj, const float segment_start[3], const float segment_vector[3],
                              float radius, float influence_radius, ivx_absorb_stats* out_stats) {
    if (!ctx || !obj || !segment_start || !segment_vector) return IVX_ERR_INVALID_ARGUMENT;
    if (!(radius >= 0.0f) || !(influence_radius >= 0.0f)) return IVX_ERR_INVALID_ARGUMENT;  // Capsule::new asserts
    return absorb_impl(ctx, obj, capsule_shape(segment_start, segment_vector, radius, influence_radius), out_stats);
}

}  // extern "C"

static int absorb_impl(ivx_ctx* ctx, ivx_object* obj, const ivx::AbsorbShape& shape, ivx_absorb_stats* out_stats,
                       const InertialUpdate* upd, const ExplicitRanges* er) {
    using namespace ivx;
    const float influence_radius = shape.influence_radius;
    cudaSetDevice(ctx->device);
    if (out_stats) std::memset(out_stats, 0, sizeof(*out_stats));
    if (obj->first_i != 0 || obj->nb[0] != obj->chunk_counts[0])
        IVX_FAIL(ctx, IVX_ERR_UNSUPPORTED, "absorption on a slab-partitioned object is not supported");
    const uint32_t n = obj->n_chunks;
    if (n == 0) return IVX_OK;
    // voxel_ranges_touching_aab (intersection.rs:766-784) on the occupied voxel ranges
    AbsorbRange r{};
    bool empty = false;
    for (int d = 0; er && d < 3; ++d) {
        // modify_voxels_within_ranges (intersection.rs:167-261): the ranges as given, chunk ranges encompassing them
        if (er->v1[d] > obj->chunk_counts[d] * 16u) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "voxel range beyond the grid");
        r.v0[d] = er->v0[d];
        r.v1[d] = er->v1[d];
        if (r.v0[d] >= r.v1[d]) empty = true;
        r.c0[d] = r.v0[d] / 16;
        r.c1[d] = (r.v1[d] + 15) / 16;
    }
    for (int d = 0; !er && d < 3; ++d) {
        // Sphere::compute_aabb / Capsule::compute_aabb (the union of the two end spheres' boxes)
        float lo = shape.center[d] - influence_radius, hi = shape.center[d] + influence_radius;
        if (shape.capsule) {
            const float e = shape.center[d] + shape.seg[d];
            lo = std::fmin(lo, e - influence_radius);
            hi = std::fmax(hi, e + influence_radius);
        }
        const float fl = std::fmax(std::floor(lo), 0.0f);
        const float ce = std::ceil(hi);
        const uint32_t s = fl >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)fl;
        const uint32_t e = !(ce > 0.0f) ? 0u : (ce >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)ce);
        r.v0[d] = std::max(obj->occ_voxels[d], s);
        r.v1[d] = std::min(obj->occ_voxels[3 + d], e);
        if (r.v0[d] >= r.v1[d]) empty = true;
        r.c0[d] = r.v0[d] / 16;
        r.c1[d] = (r.v1[d] + 15) / 16;
    }
    if (empty) return IVX_OK;
    const uint32_t n_range = (r.c1[0] - r.c0[0]) * (r.c1[1] - r.c0[1]) * (r.c1[2] - r.c0[2]);
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    uint32_t* counters = ctx->d_scratch + 32;  // [0] new slots [1..4] stats [5] new slots (boundary) [6..11] occ
    uint32_t* need = tmp.get<uint32_t>(n_range);
    uint32_t* ord = tmp.get<uint32_t>(n_range);
    if (!need || !ord) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "absorb: out of device memory");
    {
        uint32_t init[12] = {0, 0, 0, 0, 0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0};
        CU(ctx, cudaMemcpyAsync(counters, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    KL(ctx, launch_absorb_plan(obj->d_chunks, obj->nb, r, shape, need, n_range, st));
    KL(ctx, launch_exclusive_scan(need, ord, n_range, counters, st));
    // Slots for Uniform chunks that become NonUniform. How many is known on the device only; as long as the pool holds
    // the worst case (every chunk of the range here, every chunk of the refreshed box below) nothing has to be read
    // back before the end of the call — the usual case, since the pool grows by half whenever it grows.
    uint32_t w[16] = {0};
    // (the cross-chunk pass below refreshes the upper faces of the chunks [start - 1, end): it can convert chunks up to `end`)
    uint64_t n_box = 1;
    for (int d = 0; d < 3; ++d) n_box *= std::min(obj->nb[d], r.c1[d] + 1u) - (r.c0[d] > 0 ? r.c0[d] - 1u : 0u);
    const bool roomy = !upd && (uint64_t)obj->slots_used + n_range + n_box <= obj->slot_capacity;
    if (!roomy) {
        if (int rc = read_words(ctx, counters, 1, w)) return rc;
        // (room for the worst case of a range like this one: the following calls then run without this read-back)
        if (int rc = ensure_slots(ctx, obj, upd ? w[0] : (uint32_t)std::min<uint64_t>(n_range + n_box, 0x7FFFFFFFu))) return rc;
    }
    const uint32_t slots_at_entry = obj->slots_used;
    AbsorbArgs aa{};
    aa.chunks = obj->d_chunks;
    aa.nb = make_uint3(obj->nb[0], obj->nb[1], obj->nb[2]);
    aa.range = r;
    aa.n_range = n_range;
    aa.voxels = obj->d_voxels;
    aa.shape = shape;
    aa.first_new_slot = obj->slots_used;
    aa.new_slot_ord = ord;
    aa.dirty = obj->d_dirty;
    aa.label_stale = obj->d_label_stale;
    aa.stats = counters + 1;
    if (er) aa.mutual = er->mutual;
    if (upd) {
        aa.removed_cols = tmp.get<uint16_t>((size_t)n_range * 256);
        aa.removed_info = tmp.get<uint32_t>((size_t)n_range * 2);
        if (!aa.removed_cols || !aa.removed_info) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "absorb: out of device memory");
        CU(ctx, cudaMemsetAsync(aa.removed_info, 0, (size_t)n_range * 8, st));  // chunks the kernel skips removed nothing
    }
    KLP(ctx, 6, launch_absorb_apply(aa, persistent_grid(ctx, n_range, 4), st));
    if (!roomy) obj->slots_used += w[0];
    obj->split_valid = false;  // the voxels changed: labels / roots downloaded from now on must come from a new resolve
    obj->plan_serial = 0;      // ... and the object no longer is what its generation plan describes
    // The voxels are modified from here on. If the inertial update fails (an emptied voxel's type has no density) the
    // error is reported only after the boundary refresh, occupied ranges and stale-label marks below have run, so the
    // object the caller keeps is consistent.
    int upd_rc = IVX_OK;
    std::string upd_err;
    if (upd) {
        upd_rc = ivx_apply_removed_voxels(ctx, obj, r, n_range, aa.removed_info, aa.removed_cols, *upd);
        if (upd_rc != IVX_OK && upd_rc != IVX_ERR_INVALID_ARGUMENT) return upd_rc;  // device failure: nothing to salvage
        if (upd_rc) upd_err = ivx_last_error(ctx);
    }

    // boundary refresh over chunk range [start-1, end) (intersection.rs:391-393). Every chunk with a face in one of
    // those pairs lies in the box [start-1, end+1) clipped to the grid: the pass is local to it (`ChunkBox`).
    AbsorbRange b = r;
    for (int d = 0; d < 3; ++d) b.c0[d] = r.c0[d] > 0 ? r.c0[d] - 1 : 0;
    ChunkBox box{};
    for (int d = 0; d < 3; ++d) {
        box.c0[d] = b.c0[d];
        box.d[d] = std::min(obj->nb[d], r.c1[d] + 1u) - b.c0[d];
    }
    const uint32_t n_in_box = box.d[0] * box.d[1] * box.d[2];
    uint8_t* face_mask = tmp.get<uint8_t>(n_in_box);
    uint32_t* convert_flag = tmp.get<uint32_t>(n_in_box);
    uint32_t* slot_of = tmp.get<uint32_t>(n_in_box);
    const bool large_box = n_in_box > BOUNDARY_BOX_ONE_CTA;  // (a mutual absorption over most of an object, say)
    uint32_t* need2 = large_box ? tmp.get<uint32_t>(n_in_box) : nullptr;
    uint32_t* ord2 = large_box ? tmp.get<uint32_t>(n_in_box) : nullptr;
    if (!face_mask || !convert_flag || !slot_of || (large_box && (!need2 || !ord2)))
        IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "absorb: out of device memory");
    // the conversions of the boundary pass take the slots after those of the absorption itself
    KL(ctx, launch_boundary_refresh_box(obj->d_chunks, n, obj->nb, box, b, obj->slots_used, roomy ? counters : nullptr, face_mask,
                                        convert_flag, slot_of, obj->d_label_stale, counters + 5, need2, ord2, true, false, nullptr,
                                        0, st));
    if (!roomy) {
        if (int rc = read_words(ctx, counters, 12, w)) return rc;
        if (int rc = ensure_slots(ctx, obj, w[5])) return rc;
        obj->slots_used += w[5];
    }
    KL(ctx, launch_boundary_refresh_box(obj->d_chunks, n, obj->nb, box, b, 0, nullptr, face_mask, convert_flag, slot_of, nullptr,
                                        nullptr, nullptr, nullptr, false, true, obj->d_voxels, persistent_grid(ctx, n_in_box, 8),
                                        st));
    // everything the host wants to know, in one read: slots handed out, statistics, invalidated chunks, and — computed
    // on the device only if chunks were removed (`counters[4]`) — the occupied ranges (intersection.rs:387-389)
    CU(ctx, cudaMemsetAsync(counters + 12, 0, 4, st));
    KL(ctx, launch_count_nonzero_u8(obj->d_dirty, n, counters + 12, st));
    KL(ctx, launch_occupied_ranges(obj->d_chunks, n, obj->nb, obj->first_i, obj->d_voxels, counters + 6, ctx->d_scratch + 64,
                                   counters + 4, st));
    if (int rc = read_words(ctx, counters, 13, w)) return rc;
    if (roomy) obj->slots_used = slots_at_entry + w[0] + w[5];
    if (w[4]) {
        const bool any = w[6] != 0xFFFFFFFFu;
        for (int d = 0; d < 3; ++d) {
            obj->occ_voxels[d] = any ? w[6 + d] : 0u;
            obj->occ_voxels[3 + d] = any ? w[9 + d] + 1u : 0u;
        }
    }
    if (out_stats) {
        out_stats->touched_chunks = w[1];
        out_stats->touched_voxels = w[2];
        out_stats->emptied_voxels = w[3];
        out_stats->removed_chunks = w[4];
        out_stats->dirty_chunks = w[12];
    }
    if (upd_rc) IVX_FAIL(ctx, upd_rc, "%s", upd_err.c_str());
    return IVX_OK;
}

// ---- disconnected-region extraction (object/extraction.rs) -------------------------------------------------
namespace {

// update_all_chunk_boundary_adjacencies (`range` null) or update_upper_boundary_adjacencies_for_chunks_in_ranges over
// the chunk range [c0, c1) of `range` (object.rs:1659-1785): conversions get slots, then adjacency bits / obscuredness
int refresh_boundaries(ivx_ctx* ctx, ivx_object* obj, const AbsorbRange* range) {
    const uint32_t n = obj->n_chunks;
    if (n == 0) return IVX_OK;
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    uint32_t* counters = ctx->d_scratch + 32;
    uint8_t* face_mask = range ? tmp.get<uint8_t>(n) : nullptr;
    uint32_t* convert_flag = tmp.get<uint32_t>(n);
    uint32_t* need = tmp.get<uint32_t>(n);
    uint32_t* ord = tmp.get<uint32_t>(n);
    uint32_t* slot_of = tmp.get<uint32_t>(n);
    if ((range && !face_mask) || !convert_flag || !need || !ord || !slot_of)
        IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "boundary refresh: out of device memory");
    if (range) KL(ctx, launch_absorb_face_mask(obj->nb, *range, face_mask, n, st));
    KL(ctx, launch_boundary_classify(obj->d_chunks, n, obj->nb, face_mask, convert_flag, 0, obj->nb[0], st));
    KL(ctx, launch_need_slot_for_convert(obj->d_chunks, convert_flag, n, need, obj->d_label_stale, st));
    KL(ctx, launch_exclusive_scan(need, ord, n, counters + 5, st));
    uint32_t w;
    if (int rc = read_words(ctx, counters + 5, 1, &w)) return rc;
    if (int rc = ensure_slots(ctx, obj, w)) return rc;
    KL(ctx, launch_assign_slots(obj->d_chunks, need, ord, obj->slots_used, nullptr, n, slot_of, st));
    obj->slots_used += w;
    KL(ctx, launch_boundary_apply(obj->d_chunks, n, obj->nb, face_mask, convert_flag, slot_of, obj->d_voxels, nullptr, n,
                                  range ? range->c0[0] : 0u, range ? std::min(obj->nb[0], range->c1[0] + 1u) : obj->nb[0],
                                  persistent_grid(ctx, n, 8), st));
    return IVX_OK;
}

// update_occupied_ranges (object.rs:1149-1280)
int refresh_occupied_ranges(ivx_ctx* ctx, ivx_object* obj) {
    const uint32_t n = obj->n_chunks;
    uint32_t* occ = ctx->d_scratch + 38;
    const uint32_t init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0};
    CU(ctx, cudaMemcpyAsync(occ, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    KL(ctx, launch_occupied_ranges(obj->d_chunks, n, obj->nb, obj->first_i, obj->d_voxels, occ, ctx->d_scratch + 64, nullptr,
                                   ctx->stream));
    uint32_t o[6];
    if (int rc = read_words(ctx, occ, 6, o)) return rc;
    const bool any = o[0] != 0xFFFFFFFFu;
    for (int d = 0; d < 3; ++d) {
        obj->occ_voxels[d] = any ? o[d] : 0u;
        obj->occ_voxels[3 + d] = any ? o[3 + d] + 1u : 0u;
    }
    return IVX_OK;
}

// a fresh whole object of `nb` chunks with room for `slots` stored chunks
int new_plain_object(ivx_ctx* ctx, float voxel_extent, const uint32_t nb[3], uint32_t slots, ivx_object** out) {
    ivx_object* o = new (std::nothrow) ivx_object();
    if (!o) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "host allocation failed");
    o->voxel_extent = voxel_extent;
    for (int d = 0; d < 3; ++d) {
        o->grid_shape[d] = nb[d] * 16u;
        o->chunk_counts[d] = o->nb[d] = nb[d];
    }
    o->first_i = 0;
    o->own_begin = 0;
    o->own_end = nb[0];
    o->n_chunks = nb[0] * nb[1] * nb[2];
    o->slot_capacity = std::max(1u, slots);
    o->slots_used = 0;
    o->d_chunks = static_cast<DevChunk*>(ctx->alloc((size_t)o->n_chunks * sizeof(DevChunk)));
    o->d_dirty = static_cast<uint8_t*>(ctx->alloc(o->n_chunks));
    o->d_voxels = static_cast<unsigned char*>(ctx->alloc((size_t)o->slot_capacity * SLOT_BYTES));
    if (!o->d_chunks || !o->d_dirty || !o->d_voxels) {
        ivx_object_free(ctx, o);
        IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "extracted object: out of device memory");
    }
    cudaMemsetAsync(o->d_dirty, 0, o->n_chunks, ctx->stream);
    *out = o;
    return IVX_OK;
}

}  // namespace

extern "C" {

int ivx_object_from_generated_chunks(ivx_ctx* ctx, float voxel_extent, const uint32_t grid_shape[3], const ivx_voxel* voxels,
                                     const uint8_t* sparseness, ivx_object** out_object) {
    if (!ctx || !grid_shape || !out_object) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    *out_object = nullptr;
    if (!(voxel_extent > 0.0f)) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "voxel_extent must be positive");
    uint32_t cc[3];
    uint64_t n64 = 1;
    for (int d = 0; d < 3; ++d) {
        cc[d] = (grid_shape[d] + 15u) / 16u;
        n64 *= cc[d];
    }
    if (n64 > 0x7FFFFFFFull / 10) IVX_FAIL(ctx, IVX_ERR_INVALID_ARGUMENT, "grid too large");
    const uint32_t n = (uint32_t)n64;
    if (n && (!voxels || !sparseness)) return IVX_ERR_INVALID_ARGUMENT;
    uint32_t stored_bound = 0;
    for (uint32_t c = 0; c < n; ++c) stored_bound += (sparseness[c] & 2u) ? 0u : 1u;
    ivx_object* o = nullptr;
    if (int rc = new_plain_object(ctx, voxel_extent, cc, stored_bound, &o)) return rc;
    for (int d = 0; d < 3; ++d) o->grid_shape[d] = grid_shape[d];
    if (n == 0) {
        *out_object = o;
        return IVX_OK;
    }
    auto fail = [&](int rc) {
        ivx_object_free(ctx, o);
        return rc;
    };
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    unsigned char* d_src = tmp.get<unsigned char>((size_t)n * SLOT_BYTES);
    uint8_t* d_sp = tmp.get<uint8_t>(n);
    uint32_t* counter = ctx->d_scratch + 50;
    if (!d_src || !d_sp) {
        ctx->err = "ingest: out of device memory";
        return fail(IVX_ERR_OUT_OF_MEMORY);
    }
    static_assert(sizeof(ivx_voxel) == 3, "Voxel is three bytes");
    cudaError_t e = cudaMemcpyAsync(d_src, voxels, (size_t)n * SLOT_BYTES, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_sp, sparseness, n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(counter, 0, 8, st);
    if (e == cudaSuccess) {
        ctx->launches++;
        e = launch_ingest_chunks(d_src, d_sp, n, o->d_chunks, o->d_voxels, counter, persistent_grid(ctx, n, 4), st);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        ctx->err = std::string("ingest failed: ") + cudaGetErrorString(e);
        return fail(IVX_ERR_CUDA);
    }
    uint32_t used[2] = {0, 0};
    if (int rc = read_words(ctx, counter, 2, used)) return fail(rc);
    if (used[1]) {
        ctx->err = "ingest: a voxel's EMPTY flag disagrees with the sign of its signed distance (Voxel invariant, lib.rs:300-348)";
        return fail(IVX_ERR_INVALID_ARGUMENT);
    }
    o->slots_used = used[0];
    // update_occupied_voxel_ranges, then compute_all_derived_state's cross-chunk pass (object.rs:239-263, 1659-1785)
    if (int rc = refresh_occupied_ranges(ctx, o)) return fail(rc);
    if (int rc = refresh_boundaries(ctx, o, nullptr)) return fail(rc);
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        ctx->err = "ingest: stream synchronisation failed";
        return fail(IVX_ERR_CUDA);
    }
    *out_object = o;
    return IVX_OK;
}

int ivx_object_extract_disconnected_region(ivx_ctx* ctx, ivx_object* obj, ivx_extraction_info* info, ivx_object** out_extracted) {
    if (!ctx || !obj || !info || !out_extracted) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    std::memset(info, 0, sizeof(*info));
    *out_extracted = nullptr;
    // find_two_disconnected_regions on freshly resolved regions (labels of unmodified chunks are reused)
    ivx_split_info si;
    if (int rc = ivx_object_resolve_connected_regions(ctx, obj, &si)) return rc;
    info->n_regions_before = si.n_regions;
    if (!si.has_two) return IVX_OK;
    info->found_two = 1;
    const ivx_region_candidate& cand = si.candidates[si.smallest];
    const uint32_t R = cand.label;
    info->region_label = R;
    uint32_t r0[3], nbE[3];
    for (int d = 0; d < 3; ++d) {
        r0[d] = cand.chunk_min[d];
        nbE[d] = cand.chunk_max[d] + 1 - r0[d];
        info->origin_offset_in_parent[d] = r0[d] * 16u;
    }
    const uint32_t nE = nbE[0] * nbE[1] * nbE[2];

    // ---- the region's chunks inside its bounding box, in linear order (extraction.rs:137-245, 339-349): classified on
    //      the device from the roots the resolve left there ----
    Tmp tmp(ctx);
    cudaStream_t st = ctx->stream;
    uint8_t* d_mode = tmp.get<uint8_t>(nE);
    uint32_t* d_src = tmp.get<uint32_t>(nE);
    uint32_t* d_first = tmp.get<uint32_t>(nE);
    uint32_t* d_slot = tmp.get<uint32_t>(nE);
    uint32_t* d_nu_flag = tmp.get<uint32_t>(nE);
    uint8_t* d_is_r = tmp.get<uint8_t>(std::max<size_t>(1, obj->region_total));
    uint32_t* d_count = ctx->d_scratch + 44;  // [0] non-empty voxels moved [1] uniform chunks [2] non-uniform chunks
    if (!d_mode || !d_src || !d_first || !d_slot || !d_nu_flag || !d_is_r) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "extraction: out of device memory");
    CU(ctx, cudaMemsetAsync(d_count, 0, 12, st));
    ExtractPlanArgs pa{};
    pa.regions = obj->d_regions;
    pa.first = obj->d_region_first;
    pa.root = obj->d_region_root;
    pa.total = obj->region_total;
    pa.region_root = obj->region_two[si.smallest];
    pa.n_ext = nE;
    for (int d = 0; d < 3; ++d) {
        pa.ext[d] = nbE[d];
        pa.lo[d] = r0[d];
    }
    pa.nb1 = obj->nb[1];
    pa.nb2 = obj->nb[2];
    pa.mode = d_mode;
    pa.src_index = d_src;
    pa.first_region = d_first;
    pa.non_uniform_flag = d_nu_flag;
    pa.dst_slot = d_slot;
    pa.uniform_count = d_count + 1;
    pa.is_member = d_is_r;
    ctx->launches += 4;
    CU(ctx, launch_extract_plan(pa, d_count + 2, st));
    uint32_t plan_words[3];
    if (int rc = read_words(ctx, d_count, 3, plan_words)) return rc;
    const uint32_t n_uniform = plan_words[1], n_non_uniform = plan_words[2];
    info->region_chunks = n_uniform + n_non_uniform;

    ivx_object* ext = nullptr;
    // uniform chunks may be converted by the cross-chunk pass of the extracted object: room for them too
    if (int rc = new_plain_object(ctx, obj->voxel_extent, nbE, n_non_uniform + n_uniform, &ext)) return rc;
    struct Guard {
        ivx_ctx* c;
        ivx_object* o;
        ~Guard() {
            if (o) ivx_object_free(c, o);
        }
    } guard{ctx, ext};
    ext->slots_used = n_non_uniform;

    ExtractArgs xa{};
    xa.src_chunks = obj->d_chunks;
    xa.src_voxels = obj->d_voxels;
    xa.src_labels = obj->d_labels;
    xa.src_dirty = obj->d_dirty;
    xa.src_label_stale = obj->d_label_stale;
    xa.n_ext = nE;
    xa.mode = d_mode;
    xa.src_index = d_src;
    xa.first_region = d_first;
    xa.dst_slot = d_slot;
    xa.region_is_r = d_is_r;
    xa.dst_chunks = ext->d_chunks;
    xa.dst_voxels = ext->d_voxels;
    xa.non_empty_count = d_count;
    KL(ctx, launch_extract_chunks(xa, persistent_grid(ctx, nE, 4), st));
    obj->split_valid = false;
    obj->plan_serial = 0;

    // ---- the object the region left (extraction.rs:556-575) ----
    if (int rc = refresh_occupied_ranges(ctx, obj)) return rc;
    AbsorbRange b{};
    for (int d = 0; d < 3; ++d) {
        b.c0[d] = r0[d] > 0 ? r0[d] - 1 : 0;
        b.c1[d] = r0[d] + nbE[d];
    }
    if (int rc = refresh_boundaries(ctx, obj, &b)) return rc;

    // ---- complete_extracted_voxel_object (extraction.rs:1902-1973) ----
    uint32_t moved;
    if (int rc = read_words(ctx, d_count, 1, &moved)) return rc;
    info->moved_non_empty_voxels = moved;
    if (n_uniform == 0 && moved < 8u) {  // NON_EMPTY_VOXEL_THRESHOLD (object.rs:203): dropped
        info->discarded = 1;
        return IVX_OK;  // guard frees the extracted object
    }
    if (nbE[0] <= 2 && nbE[1] <= 2 && nbE[2] <= 2 && n_uniform == 0 && nE > 1) {
        if (int rc = refresh_occupied_ranges(ctx, ext)) return rc;  // determine_occupied_voxel_ranges
        const uint32_t* ov = ext->occ_voxels;
        if (ov[3] - ov[0] <= 14u && ov[4] - ov[1] <= 14u && ov[5] - ov[2] <= 14u) {
            uint32_t org[3], one3[3] = {1, 1, 1};
            for (int d = 0; d < 3; ++d) org[d] = ov[d] > 0 ? ov[d] - 1 : 0;  // room for one empty boundary layer
            ivx_object* one = nullptr;
            if (int rc = new_plain_object(ctx, obj->voxel_extent, one3, 1, &one)) return rc;
            one->slots_used = 1;
            cudaError_t le = launch_repack_single(ext->d_chunks, ext->nb, ext->d_voxels, org, ov, ov + 3, one->d_chunks, one->d_voxels, st);
            ctx->launches++;
            if (le != cudaSuccess) {
                ivx_object_free(ctx, one);
                IVX_FAIL(ctx, IVX_ERR_CUDA, "repack: %s", cudaGetErrorString(le));
            }
            for (int d = 0; d < 3; ++d) info->origin_offset_in_parent[d] += org[d];
            ivx_object_free(ctx, ext);
            ext = one;
            guard.o = one;
            info->single_chunk = 1;
        }
    }
    // derived state of the extracted object from scratch (extraction.rs:585-597)
    if (int rc = refresh_boundaries(ctx, ext, nullptr)) return rc;
    if (int rc = refresh_occupied_ranges(ctx, ext)) return rc;
    CU(ctx, cudaStreamSynchronize(st));
    info->extracted = 1;
    guard.o = nullptr;
    *out_extracted = ext;
    return IVX_OK;
}

int ivx_object_dirty_chunks(ivx_ctx* ctx, const ivx_object* obj, uint32_t* out, uint32_t capacity, uint32_t* out_count) {
    if (!ctx || !obj || !out_count) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    *out_count = 0;
    const uint32_t n = obj->n_chunks;
    if (n == 0) return IVX_OK;
    // the marks are compacted on the device (ascending linear chunk index); only the list crosses the bus
    Tmp tmp(ctx);
    uint32_t* flag = tmp.get<uint32_t>(n);
    uint32_t* scan = tmp.get<uint32_t>(n);
    uint32_t* list = tmp.get<uint32_t>(n);
    uint32_t* unused = tmp.get<uint32_t>(n);
    if (!flag || !scan || !list || !unused) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "dirty chunks: out of device memory");
    cudaStream_t st = ctx->stream;
    KL(ctx, launch_flag_dirty_exposed(obj->d_chunks, obj->d_dirty, n, unused, flag, st));
    KL(ctx, launch_exclusive_scan(flag, scan, n, ctx->d_scratch + 28, st));
    KL(ctx, launch_scatter_active(flag, scan, n, list, st));
    uint32_t cnt = 0;
    if (int rc = read_words(ctx, ctx->d_scratch + 28, 1, &cnt)) return rc;
    *out_count = cnt;
    if (out && cnt) {
        CU(ctx, cudaMemcpyAsync(out, list, (size_t)std::min(cnt, capacity) * 4, cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaStreamSynchronize(st));
    }
    return IVX_OK;
}

int ivx_object_remesh_dirty(ivx_ctx* ctx, ivx_object* obj, ivx_mesh_info* out) {
    if (!ctx || !obj || !out) return IVX_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    std::memset(out, 0, sizeof(*out));
    const uint32_t n = obj->n_chunks;
    if (n == 0) return IVX_OK;
    Tmp tmp(ctx);
    uint32_t* exposed = tmp.get<uint32_t>(n);
    uint32_t* dflag = tmp.get<uint32_t>(n);
    if (!exposed || !dflag) IVX_FAIL(ctx, IVX_ERR_OUT_OF_MEMORY, "remesh: out of device memory");
    KL(ctx, launch_flag_dirty_exposed(obj->d_chunks, obj->d_dirty, n, exposed, dflag, ctx->stream));
    // the patch takes the place of the object's mesh (ivx_object_mesh_sync is the call that keeps the mesh)
    ivx_mesh_sync_free(obj->sync);
    obj->sync = nullptr;
    obj->mesh_is_patch = true;
    obj->mesh.serial++;
    if (int rc = mesh_impl(ctx, obj, exposed, obj->mesh)) return rc;
    CU(ctx, cudaMemsetAsync(obj->d_dirty, 0, n, ctx->stream));  // mark_chunk_meshes_synchronized
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    fill_mesh_info(obj->mesh, out);
    return IVX_OK;
}

}  // extern "C"
