// Internal state of libimpact_voxel_cuda.so shared by the translation units that implement the C ABI
// (api.cu, split.cu): context with its device memory pool, program / object records, error and launch
// macros. Not part of the public ABI.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

using namespace ivx;

// ---------------------------------------------------------------------------
struct PoolBlock {
    void* ptr;
    size_t size;
    bool used;
};

// Everything generate_impl / mesh_impl read back from the device for one (program, voxel extent, type generator, plane
// range): counts that size the next allocations and launches. They are a pure function of those inputs, so — like an
// FFT plan — they are kept, and the next generation of the same object runs without a host round trip. A kernel
// compares the device's own counters with the plan afterwards (plan_error word), so a wrong plan is an error, never a
// wrong object.
struct GenPlan {
    uint64_t serial = 0;
    uint64_t prog_uid = 0;
    float voxel_extent = 0.0f;
    ivx_type_generator types{};
    uint32_t i_begin = 0, i_end = 0;
    bool whole = false, streamed = false;
    uint32_t n_active = 0, n_slots = 0, max_depth = 0, occ[6] = {0, 0, 0, 0, 0, 0};
    std::vector<uint32_t> part_active;  // streamed generation: active-chunk boundaries of the parts
    // mesh of the object as generated (for a slab: after the halo exchange with the same neighbours)
    bool mesh_valid = false;
    uint32_t mesh_counts[4] = {0, 0, 0, 0};  // exposed chunks, vertices, indices, submeshes
};

struct ivx_ctx {
    int device = 0;
    std::vector<GenPlan> plans;  // most recent last; at most MAX_PLANS
    struct WorkPlan {            // ivx_program_plane_work results
        uint64_t prog_uid = 0;
        float voxel_extent = 0.0f;
        ivx_type_generator types{};
        std::vector<uint32_t> work;
    };
    std::vector<WorkPlan> work_plans;
    uint64_t next_serial = 1;
    static constexpr size_t MAX_PLANS = 32;
    static constexpr uint32_t PLAN_ERROR_WORD = 63;  // h_pinned word set by k_check_plan on a mismatch
    GenPlan* find_plan(uint64_t serial) {
        for (auto& pl : plans)
            if (pl.serial == serial) return &pl;
        return nullptr;
    }
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;  // device→host copies of streamed generation, beside the compute stream
    cudaStream_t aux_stream = nullptr;   // second compute stream of streamed generation (typing / packing of part p
                                         // under the evaluation of part p + 1)
    std::string err;
    uint64_t launches = 0;
    int sm_count = 148;
    std::vector<PoolBlock> pool;
    // optional per-kernel timing
    bool profiling = false;
    struct ProfEvent { cudaEvent_t a, b; uint32_t id; };
    std::vector<ProfEvent> prof_events;
    double prof_ms[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t prof_launches[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t* h_pinned = nullptr;  // 64 words of pinned scratch for counter read-back
    uint32_t* h_pinned_dev = nullptr;  // the same words as the device sees them (mapped): counters are stored by a
                                       // kernel, not by the copy engine, which may be busy with a bulk transfer
    uint32_t* d_scratch = nullptr; // 64 words of device counters
    unsigned long long* d_counters64 = nullptr;  // profiling counters: [0] 4-D simplex evaluations of k_types

    void* alloc(size_t bytes) {
        if (bytes == 0) bytes = 16;
        bytes = (bytes + 255) & ~(size_t)255;
        int best = -1;
        for (int i = 0; i < (int)pool.size(); ++i)
            if (!pool[i].used && pool[i].size >= bytes && pool[i].size <= bytes * 2 + (1 << 20) &&
                (best < 0 || pool[i].size < pool[best].size))
                best = i;
        if (best >= 0) {
            pool[best].used = true;
            return pool[best].ptr;
        }
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            // release cached blocks and retry once
            for (auto it = pool.begin(); it != pool.end();) {
                if (!it->used) {
                    cudaFree(it->ptr);
                    it = pool.erase(it);
                } else {
                    ++it;
                }
            }
            if (cudaMalloc(&p, bytes) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
        }
        pool.push_back({p, bytes, true});
        return p;
    }
    void release(void* p) {
        if (!p) return;
        for (auto& b : pool)
            if (b.ptr == p) {
                b.used = false;
                return;
            }
    }
};

struct ivx_program {
    uint64_t uid = 0;  // unique per built program (plans are keyed by it, not by the pointer)
    HostProgram host;
    ivx_node* d_nodes = nullptr;
    Instr* d_root = nullptr;     // the whole program as an instruction list
    uint32_t root_len = 0;
    uint32_t* d_root_meta = nullptr;  // [0] = offset (0), [1] = length
};

struct ivx_mesh_sync;  // mesh_sync.cu: the host-side ChunkSubmeshManager of a mesh that is kept in sync
void ivx_mesh_sync_free(ivx_mesh_sync* s);
struct ivx_probes;     // probes.cu: collision probes beside the mesh
struct ivx_ctx;
void ivx_probes_free(ivx_ctx* ctx, ivx_probes* p);

struct DeviceMesh {
    uint32_t n_vertices = 0, n_indices = 0, n_submeshes = 0, n_work = 0;
    uint32_t cap_vertices = 0, cap_indices = 0, cap_submeshes = 0;  // allocated elements once the mesh is kept in sync (0: exact)
    uint64_t serial = 0;  // counts the times the mesh was created anew (ivx_object_mesh, ivx_object_remesh_dirty)
    float* positions = nullptr;
    float* normals = nullptr;
    uint32_t* indices = nullptr;
    ivx_index_materials* index_materials = nullptr;
    ivx_chunk_submesh* submeshes = nullptr;
    uint32_t* vertex_ranges = nullptr;
};

struct ivx_object {
    float voxel_extent = 1.0f;
    uint32_t grid_shape[3] = {0, 0, 0};
    uint32_t chunk_counts[3] = {0, 0, 0};  // full grid
    uint32_t nb[3] = {0, 0, 0};            // locally stored chunk planes (slab)
    uint32_t first_i = 0;                  // global chunk-i of local plane 0
    uint32_t own_begin = 0, own_end = 0;   // owned planes, global chunk-i
    uint32_t n_chunks = 0;
    DevChunk* d_chunks = nullptr;
    unsigned char* d_voxels = nullptr;
    uint32_t slot_capacity = 0, slots_used = 0;
    uint8_t* d_dirty = nullptr;
    uint32_t occ_voxels[6] = {0, 0, 0, 0, 0, 0};  // lo xyz, hi xyz (exclusive)
    uint32_t n_void = 0, n_uniform = 0, n_non_uniform = 0;
    DeviceMesh mesh;
    ivx_mesh_sync* sync = nullptr;  // set by ivx_object_mesh_sync
    ivx_probes* probes = nullptr;   // set by ivx_object_collision_probes
    bool mesh_is_patch = false;     // `mesh` holds the patch of ivx_object_remesh_dirty, not the object's mesh
    uint64_t plan_serial = 0;  // GenPlan this object was generated with; 0 once the object has been modified
    // slab protocol (multi-GPU): derived state is pending until the halo planes are imported
    bool derive_pending = false;
    uint32_t* d_slot_of = nullptr;       // slot reserved for each chunk (conversion target), 0xFFFFFFFF = none
    uint32_t* d_convert_flag = nullptr;  // result of ivx_object_slab_classify
    bool halo_present[2] = {false, false};
    // connected regions (split.cu): device labels per slot, host results of the last resolve
    uint8_t* d_labels = nullptr;
    uint32_t label_slots = 0;
    uint32_t* d_regions = nullptr;      // per chunk: kind << 16 | boundary_region_count << 8 | region_count
    uint8_t* d_label_stale = nullptr;   // per chunk: modified since its labels were computed
    // results of the last resolve (regions.cu), valid while split_valid
    uint32_t* d_region_first = nullptr;  // per chunk: index of its region 0 among all local regions
    uint32_t* d_region_label = nullptr;  // per local region: chunk << 8 | region
    uint32_t* d_region_root = nullptr;   // per local region: index of the root of its connected region
    uint32_t region_cap = 0;             // capacity of the per-region arrays
    uint32_t region_total = 0;           // local regions
    uint32_t region_records = 0;         // connection records of the last resolve (sizes the next one)
    uint32_t region_two[2] = {0, 0};     // region indices of the two roots find_two_disconnected_regions reports
    bool split_valid = false;
    // streamed generation: packed voxels / descriptors staged for the copy stream
    ivx_voxel* d_stage_voxels = nullptr;
    ivx_chunk_desc* d_stage_chunks = nullptr;
};

#define IVX_FAIL(ctx, code, ...)                              \
    do {                                                      \
        char _b[512];                                         \
        std::snprintf(_b, sizeof(_b), __VA_ARGS__);           \
        (ctx)->err = _b;                                      \
        return (code);                                        \
    } while (0)

#define CU(ctx, expr)                                                                               \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            cudaGetLastError();                                                                     \
            IVX_FAIL(ctx, _e == cudaErrorMemoryAllocation ? IVX_ERR_OUT_OF_MEMORY : IVX_ERR_CUDA,   \
                     "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);   \
        }                                                                                           \
    } while (0)

// kernel launch through a launch_* wrapper: counts the launch
#define KL(ctx, expr)      \
    do {                   \
        (ctx)->launches++; \
        CU(ctx, expr);     \
    } while (0)

// kernel launch with optional event timing under kernel id `kid`
#define KLP(ctx, kid, expr)                                              \
    do {                                                                 \
        (ctx)->launches++;                                               \
        ivx_ctx::ProfEvent _pe{nullptr, nullptr, (uint32_t)(kid)};       \
        if ((ctx)->profiling) {                                          \
            cudaEventCreate(&_pe.a);                                     \
            cudaEventCreate(&_pe.b);                                     \
            cudaEventRecord(_pe.a, (ctx)->stream);                       \
        }                                                                \
        cudaError_t _le = (expr);                                        \
        if ((ctx)->profiling) {                                          \
            cudaEventRecord(_pe.b, (ctx)->stream);                       \
            (ctx)->prof_events.push_back(_pe);                           \
        }                                                                \
        CU(ctx, _le);                                                    \
    } while (0)

struct Tmp {
    ivx_ctx* ctx;
    std::vector<void*> ptrs;
    explicit Tmp(ivx_ctx* c) : ctx(c) {}
    ~Tmp() {
        for (void* p : ptrs) ctx->release(p);
    }
    template <typename T>
    T* get(size_t count) {
        void* p = ctx->alloc(count * sizeof(T));
        if (p) ptrs.push_back(p);
        return static_cast<T*>(p);
    }
};


// inertia.cu: VoxelObjectInertialPropertyUpdater::remove_voxel (object/inertia.rs:374-395) for every voxel an absorption
// emptied, applied to `inout` in the reference's visiting order
struct InertialUpdate {
    const float* densities;
    uint32_t n_densities;
    ivx_inertial_moments* inout;
};
int ivx_apply_removed_voxels(ivx_ctx* ctx, const ivx_object* obj, const AbsorbRange& r, uint32_t n_range,
                             const uint32_t* removed_info, const uint16_t* removed_cols, const InertialUpdate& upd);

// api.cu
#define IVX_HIDDEN __attribute__((visibility("hidden")))
extern "C" {
IVX_HIDDEN int ivx_internal_mesh(ivx_ctx* ctx, ivx_object* obj, bool sync, uint32_t counts[4], ivx_mesh_info* out);
IVX_HIDDEN int ivx_internal_take_plan_error(ivx_ctx* ctx, ivx_object* obj);
IVX_HIDDEN void fill_mesh_info_from(const DeviceMesh& m, ivx_mesh_info* out);
IVX_HIDDEN int ivx_internal_slab_finalize(ivx_ctx* ctx, ivx_object* obj, bool sync);
}
int ivx_read_words(ivx_ctx* ctx, const uint32_t* d_src, uint32_t n, uint32_t* out);
uint32_t ivx_persistent_grid(ivx_ctx* ctx, uint32_t n_work, int blocks_per_sm);
