#!/bin/bash
# Builds the CPU oracle with AddressSanitizer + UndefinedBehaviorSanitizer and runs the oracle-side test files against
# it (the reference's CI runs ASan and Miri on its CPU code, .github/workflows/impact.yml:219-308). No GPU needed.
set -e
cd "$(dirname "$0")/.."
mkdir -p /tmp/orc_asan
g++ -O1 -g -march=x86-64-v3 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -fsanitize=address,undefined \
    -fno-omit-frame-pointer -pthread -shared -o /tmp/orc_asan/liboracle.so oracle/*.cpp
cp oracle/_build/liboracle.so /tmp/orc_asan/liboracle.orig.so
trap 'cp /tmp/orc_asan/liboracle.orig.so oracle/_build/liboracle.so; touch oracle/_build/liboracle.so' EXIT
cp /tmp/orc_asan/liboracle.so oracle/_build/liboracle.so
touch oracle/_build/liboracle.so
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    python -m pytest tests/test_oracle_*.py tests/test_surface_voxels.py -x -q -m "not gpu" -p no:cacheprovider

# The product's host-only code (graph compiler program.cpp, geometry.cpp) under the same sanitizers: built without CUDA,
# the few other symbols _lib.py touches at load time are stubbed, the host-only tests run against it.
cat > /tmp/orc_asan/stubs.cpp <<'STUB'
#include <cstdint>
extern "C" {
const char* ivx_last_error(const void*) { return ""; }
uint32_t ivx_abi_version(void) { return 1; }
uint64_t ivx_kernel_launch_count(const void*) { return 0; }
void ivx_destroy(void*) {}
void ivx_program_free(void*, void*) {}
void ivx_object_free(void*, void*) {}
}
STUB
g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off -fsanitize=address,undefined -fno-omit-frame-pointer -shared \
    -o /tmp/orc_asan/libhost.so impact_b200/csrc/program.cpp impact_b200/csrc/geometry.cpp /tmp/orc_asan/stubs.cpp
cp /tmp/orc_asan/liboracle.orig.so oracle/_build/liboracle.so; touch oracle/_build/liboracle.so
IMPACT_VOXEL_CUDA_LIB=/tmp/orc_asan/libhost.so ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    python -m pytest tests/test_host_geometry.py tests/test_host_and_cabi.py -q -p no:cacheprovider \
    -k "(box or random_boxes or encompass or compile or unrolls or errors) and not plane"
