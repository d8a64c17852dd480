"""The library's host-side range allocator (impact_b200/csrc/mesh_sync.cuh, `ivx_ranges`: what places re-meshed chunks and
collision-probe points into freed ranges) against the oracle's `RangeAllocator` restatement
(impact_containers/src/range_allocator.rs, pinned by the reference's own tests in tests/test_oracle_synced_mesh.py):
random sequences of free / allocate / merge must give the same ranges, call by call. The library's version answers
requests no hole can serve from a bound on the longest hole and stops at the lowest exact fit — shortcuts that must not
change a single placement. CPU only: a small harness (tests/native/ranges_harness.cu) is built with nvcc, host code only."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "ranges_harness.cu")
SO = os.path.join(HERE, "native", "_ranges_harness.so")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def harness():
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    deps = [SRC, os.path.join(HERE, "..", "impact_b200", "csrc", "mesh_sync.cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call([NVCC, "-std=c++17", "-O2", "-fmad=false", "-DIVX_FMAD_OFF", "--expt-relaxed-constexpr", "-w",
                               "-shared", "-Xcompiler", "-fPIC", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.rh_create.restype = C.c_void_p
    lib.rh_count.restype = C.c_uint32
    for f in (lib.rh_destroy, lib.rh_free_range, lib.rh_allocate_range, lib.rh_merge, lib.rh_count, lib.rh_ranges):
        f.argtypes = None
    return lib


class Ours:
    def __init__(self, lib):
        self.lib, self.h = lib, C.c_void_p(lib.rh_create())

    def free_range(self, a, b):
        self.lib.rh_free_range(self.h, C.c_uint32(a), C.c_uint32(b))

    def allocate_range(self, n):
        s = C.c_uint32()
        ok = self.lib.rh_allocate_range(self.h, C.c_uint32(n), C.byref(s))
        return (s.value, s.value + n) if ok else None

    def merge_consecutive_ranges(self):
        self.lib.rh_merge(self.h)

    def ranges(self):
        n = self.lib.rh_count(self.h)
        out = np.zeros((max(1, n), 2), np.uint32)
        self.lib.rh_ranges(self.h, out.ctypes.data_as(C.c_void_p))
        return out[:n]

    def __del__(self):
        self.lib.rh_destroy(self.h)


def test_the_reference_known_answers(harness):
    # impact_containers/src/range_allocator.rs:153-250, the reference's own tests, case by case
    a = Ours(harness)
    assert a.allocate_range(1) is None                                    # allocates_nothing_before_freed
    a = Ours(harness)                                                     # frees_and_allocates_single_range
    a.free_range(2, 6)
    assert a.allocate_range(4) == (2, 6) and a.allocate_range(1) is None
    a = Ours(harness)                                                     # allocates_range_in_smallest_slot
    a.free_range(2, 6)
    a.free_range(10, 12)
    assert a.allocate_range(2) == (10, 12) and a.allocate_range(4) == (2, 6)
    a = Ours(harness)                                                     # uses_parts_of_larger_slots
    a.free_range(2, 12)
    assert a.allocate_range(4) == (2, 6) and a.allocate_range(4) == (6, 10)
    assert a.allocate_range(4) is None
    assert a.allocate_range(2) == (10, 12) and a.allocate_range(1) is None
    a = Ours(harness)                                                     # does_not_merge_two_disconnected_free_ranges
    a.free_range(2, 5)
    a.free_range(6, 9)
    a.merge_consecutive_ranges()
    assert a.allocate_range(6) is None
    for pieces, want in ((((2, 6), (6, 8)), (2, 8)),                      # merges_two / three / four_consecutive_free_ranges
                         (((2, 6), (6, 8), (8, 42)), (2, 42)),
                         (((2, 6), (6, 8), (8, 42), (42, 50)), (2, 50))):
        a = Ours(harness)
        for lo, hi in pieces:
            a.free_range(lo, hi)
        a.merge_consecutive_ranges()
        assert a.allocate_range(want[1] - want[0]) == want and a.allocate_range(1) is None


@pytest.mark.parametrize("seed", range(12))
def test_random_sequences_match_the_oracle_call_by_call(harness, oracle, seed):
    rng = np.random.default_rng(seed)
    ours, ref = Ours(harness), oracle.RangeAllocator()
    live = []      # allocated ranges (from either source: both must agree)
    end = 0        # length of the buffer so far
    sizes = (1, 2, 3, 5, 8, 13, 40, 100, 400) if seed % 2 else tuple(range(1, 30))
    for step in range(3000):
        op = rng.integers(0, 100)
        if op < 45 or not live:                                  # allocate, else append (what mesh_sync does)
            n = int(rng.choice(sizes))
            got, want = ours.allocate_range(n), ref.allocate_range(n)
            assert got == want, (seed, step, n, got, want)
            if got is None:
                got = (end, end + n)
                end += n
            live.append(got)
        elif op < 90:                                            # free a live range
            a, b = live.pop(int(rng.integers(0, len(live))))
            ours.free_range(a, b)
            ref.free_range(a, b)
            if rng.integers(0, 4) == 0:                          # the same start again: stays as it is (BTreeSet::insert)
                ours.free_range(a, b)
                ref.free_range(a, b)
        else:
            ours.merge_consecutive_ranges()
            ref.merge_consecutive_ranges()
    # drain: every remaining hole is handed out in the same order by both
    ours.merge_consecutive_ranges()
    ref.merge_consecutive_ranges()
    holes = ours.ranges()
    assert len(holes) > 0 and (holes[:, 0] < holes[:, 1]).all() and (holes[1:, 0] > holes[:-1, 1]).all()
    for n in sorted((int(b - a) for a, b in holes), reverse=True):
        assert ours.allocate_range(n) == ref.allocate_range(n)
    assert ours.allocate_range(1) == ref.allocate_range(1)
