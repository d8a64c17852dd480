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
