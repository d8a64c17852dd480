"""The BASELINE.json configurations as concrete inputs: graph builders shared by
bench.py and the tests. Pure host code (numpy); no device needed."""
from __future__ import annotations

import numpy as np

from .graph import SDFGraph, VoxelTypeGenerator


def sphere(radius: float = 31.0) -> SDFGraph:
    """Config 1: `Sphere(r = 31)` → 64³ (also the engine-bench shape r = 100 → 202³)."""
    g = SDFGraph()
    g.sphere(radius)
    return g


def noisy_box(extent: float = 246.0, octaves: int = 8) -> SDFGraph:
    """Config 2: box perturbed by multi-octave gradient noise (parameters of the reference bench,
    engine/src/benchmark/benchmarks/generation.rs:87-92); extent 246 → 256³."""
    g = SDFGraph()
    b = g.box([extent] * 3)
    g.multifractal_noise(b, octaves, 0.02, 2.0, 0.6, 4.0, 0)
    return g


def asteroid_stand_in(scale: float = 1.0, seed: int = 0, n_craters=(40, 150, 250)) -> SDFGraph:
    """Hand-built atomic graph with the structure and node counts of
    engine/benches/data/asteroid.vgen.ron (3-6 smooth-unioned spheres + 1-octave noise; three crater
    passes of 40 / 150 / 250 rotated capsules in balanced smooth-union trees, smooth-subtracted; final
    5-octave noise). Used until the meta-graph compiler places the craters by sphere-casting; crater
    positions here are drawn on the body's bounding sphere instead."""
    rng = np.random.default_rng(seed)
    g = SDFGraph()

    def balanced_union(ids, k):
        ids = list(ids)
        while len(ids) > 1:
            nxt = [g.union(ids[i], ids[i + 1], k) for i in range(0, len(ids) - 1, 2)]
            if len(ids) % 2:
                nxt.append(ids[-1])
            ids = nxt
        return ids[0]

    spheres = []
    for _ in range(5):
        s = g.sphere(float(rng.uniform(30.0, 60.0) * rng.uniform(0.5, 2.0) * scale))
        s = g.translation(s, [float(x) for x in rng.uniform(-30.0, 30.0, 3) * scale])
        spheres.append(s)
    body = balanced_union(spheres, 25.0 * scale)
    body = g.multifractal_noise(body, 1, 0.01 / scale, 2.0, 0.5, 8.0 * scale, 0)
    body_radius = 110.0 * scale
    for count, (rmin, rmax), k in zip(n_craters, [(50.0, 80.0), (15.0, 55.0), (5.0, 25.0)], [8.0, 3.0, 3.0]):
        caps = []
        for _ in range(count):
            r = float(rng.uniform(rmin, rmax) ** 0.5 * rmin ** 0.5 * scale * 0.35)
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            c = g.capsule(r, r)
            c = g.rotation_from_axis_angle(c, [float(x) for x in rng.normal(size=3)], float(rng.uniform(0.0, 3.1)))
            c = g.translation(c, [float(x) for x in d * body_radius * rng.uniform(0.75, 1.05)])
            caps.append(c)
        body = g.subtraction(body, balanced_union(caps, k * scale), k * scale)
    g.multifractal_noise(body, 5, 0.02 / scale, 2.0, 0.546, 2.0 * scale, 0)
    return g


def gradient_noise_types() -> VoxelTypeGenerator:
    """`GradientNoise(types [0,1,2,3], noise_freq 0.02, type_freq 1.0, seed 0)` (generation.rs:113-124)."""
    return VoxelTypeGenerator.gradient_noise([0, 1, 2, 3], 0.02, 1.0, 0)


def same_type(t: int = 0) -> VoxelTypeGenerator:
    return VoxelTypeGenerator.same(t)


def grid_shape_of(graph: SDFGraph):
    """`SDFVoxelGenerator::new` grid shape (generation.rs:230-234) from the host-compiled domain."""
    from .voxel import compile_program_host

    _, _, lo, hi = compile_program_host(graph)
    return tuple(int(np.ceil(np.float32(h) - np.float32(l))) + 2 for l, h in zip(lo, hi))


def scale_to_max_dim(make, target_lo: int, target_hi: int, scale0: float = 1.0):
    """Finds `scale` with max(grid_shape(make(scale))) in (target_lo, target_hi]."""
    lo_s, hi_s = None, None
    s = scale0
    for _ in range(60):
        m = max(grid_shape_of(make(s)))
        if target_lo < m <= target_hi:
            return s
        if m > target_hi:
            hi_s = s
        else:
            lo_s = s
        if lo_s is None:
            s *= 0.5
        elif hi_s is None:
            s *= 2.0
        else:
            s = 0.5 * (lo_s + hi_s)
    raise RuntimeError("could not scale the graph into the requested grid size")
