// Test harness (host only): drives the library's host-side range allocator (impact_b200/csrc/mesh_sync.cuh,
// ivx_ranges) through a C interface so that tests/test_host_range_allocator.py can hold it against the oracle's
// RangeAllocator restatement step by step. Built by the test with nvcc; not part of the product library.
#include "../../impact_b200/csrc/mesh_sync.cuh"

extern "C" {
void* rh_create() { return new ivx_ranges::Ranges(); }
void rh_destroy(void* h) { delete static_cast<ivx_ranges::Ranges*>(h); }
void rh_free_range(void* h, uint32_t a, uint32_t b) { ivx_ranges::release_range(*static_cast<ivx_ranges::Ranges*>(h), a, b); }
// → 1 and *start when a free range fits, else 0
int rh_allocate_range(void* h, uint32_t len, uint32_t* start) {
    uint32_t s = 0;
    const bool ok = ivx_ranges::take_range(*static_cast<ivx_ranges::Ranges*>(h), len, s);
    *start = s;
    return ok ? 1 : 0;
}
void rh_merge(void* h) { ivx_ranges::coalesce(*static_cast<ivx_ranges::Ranges*>(h)); }
uint32_t rh_count(void* h) { return (uint32_t)static_cast<ivx_ranges::Ranges*>(h)->size(); }
void rh_ranges(void* h, uint32_t* out) {
    const auto& r = *static_cast<ivx_ranges::Ranges*>(h);
    for (size_t q = 0; q < r.size(); ++q) {
        out[2 * q] = r[q].first;
        out[2 * q + 1] = r[q].second;
    }
}
}
