"""Host-side mirror of the reference's atomic SDF graph API.

Mirrors `SDFGraph` / `SDFNode` from
engine/crates/impact_voxel/src/generation/sdf/atomic.rs:55-81, 1019-1128 (same
constructor names, argument meaning and assertion behaviour). The graph is kept
as a numpy structured array with the exact layout of `ivx_sdf_node`
(include/impact_voxel_cuda.h), so it crosses the C ABI without conversion.
"""
from __future__ import annotations

import math

import numpy as np

# enum ivx_node_kind (include/impact_voxel_cuda.h) == SDFNode variant order (atomic.rs:63-81)
SPHERE, CAPSULE, BOX, TRANSLATION, ROTATION, SCALING, NOISE, UNION, SUBTRACTION, INTERSECTION = range(10)

SDF_NODE_DTYPE = np.dtype(
    [("kind", "<u4"), ("child", "<u4", (2,)), ("octaves", "<u4"), ("seed", "<u4"), ("p", "<f4", (8,))],
    align=True,
)
assert SDF_NODE_DTYPE.itemsize == 52

PROG_NODE_DTYPE = np.dtype(
    [
        ("kind", "<u4"),
        ("octaves", "<u4"),
        ("seed", "<u4"),
        ("leaf_count", "<u4"),
        ("p", "<f4", (8,)),
        ("transform", "<f4", (16,)),
        ("dom_lo", "<f4", (3,)),
        ("dom_hi", "<f4", (3,)),
        ("margin", "<f4"),
        ("_pad", "<u4"),
    ],
    align=True,
)
assert PROG_NODE_DTYPE.itemsize == 144


def _f32(x) -> np.float32:
    return np.float32(x)


class SDFGraph:
    """`SDFGraph` (atomic.rs:1019-1058): node list + root id; `add_node` makes the new node the root."""

    def __init__(self):
        self._nodes: list[tuple] = []
        self.root_node_id = 0

    # -- SDFGraph ---------------------------------------------------------------
    def add_node(self, node: tuple) -> int:
        node_id = len(self._nodes)
        self._nodes.append(node)
        self.root_node_id = node_id
        return node_id

    def set_root_node(self, node_id: int) -> None:
        assert node_id < len(self._nodes)
        self.root_node_id = node_id

    def __len__(self) -> int:
        return len(self._nodes)

    def nodes(self) -> np.ndarray:
        arr = np.zeros(len(self._nodes), dtype=SDF_NODE_DTYPE)
        for i, (kind, child, octaves, seed, p) in enumerate(self._nodes):
            arr[i]["kind"] = kind
            arr[i]["child"] = child
            arr[i]["octaves"] = octaves
            arr[i]["seed"] = seed
            pp = np.zeros(8, dtype=np.float32)
            pp[: len(p)] = np.asarray(p, dtype=np.float32)
            arr[i]["p"] = pp
        return arr

    # -- SDFNode constructors (atomic.rs:1060-1128) -------------------------------
    def sphere(self, radius: float) -> int:
        assert radius >= 0.0
        return self.add_node((SPHERE, (0, 0), 0, 0, [radius]))

    def capsule(self, segment_length: float, radius: float) -> int:
        assert segment_length >= 0.0 and radius >= 0.0
        return self.add_node((CAPSULE, (0, 0), 0, 0, [segment_length, radius]))

    def box(self, extents) -> int:
        assert all(not math.copysign(1.0, e) < 0 for e in extents)
        return self.add_node((BOX, (0, 0), 0, 0, list(extents)))

    def translation(self, child_id: int, translation) -> int:
        return self.add_node((TRANSLATION, (child_id, 0), 0, 0, list(translation)))

    def rotation(self, child_id: int, quaternion_xyzw) -> int:
        return self.add_node((ROTATION, (child_id, 0), 0, 0, list(quaternion_xyzw)))

    def rotation_from_axis_angle(self, child_id: int, axis, angle: float) -> int:
        """`SDFRotation::from_axis_angle` (atomic.rs:1296-1299) → glam `Quat::from_axis_angle`."""
        a = np.asarray(axis, dtype=np.float32)
        a = a / np.float32(np.sqrt(np.float32((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2])))
        half = np.float32(angle) * np.float32(0.5)
        s, c = np.float32(math.sin(float(half))), np.float32(math.cos(float(half)))
        v = a * s
        return self.rotation(child_id, [v[0], v[1], v[2], c])

    def scaling(self, child_id: int, scaling: float) -> int:
        assert scaling > 0.0
        return self.add_node((SCALING, (child_id, 0), 0, 0, [scaling]))

    def multifractal_noise(self, child_id, octaves, frequency, lacunarity, persistence, amplitude, seed) -> int:
        return self.add_node(
            (NOISE, (child_id, 0), int(octaves), int(seed), [frequency, lacunarity, persistence, amplitude])
        )

    def union(self, child_1_id: int, child_2_id: int, smoothness: float) -> int:
        assert smoothness >= 0.0
        return self.add_node((UNION, (child_1_id, child_2_id), 0, 0, [smoothness]))

    def subtraction(self, child_1_id: int, child_2_id: int, smoothness: float) -> int:
        assert smoothness >= 0.0
        return self.add_node((SUBTRACTION, (child_1_id, child_2_id), 0, 0, [smoothness]))

    def intersection(self, child_1_id: int, child_2_id: int, smoothness: float) -> int:
        assert smoothness >= 0.0
        return self.add_node((INTERSECTION, (child_1_id, child_2_id), 0, 0, [smoothness]))


class VoxelTypeGenerator:
    """`VoxelTypeGenerator` (generation/voxel_type.rs:9-36): `Same` or `GradientNoise`."""

    DTYPE = np.dtype(
        [("kind", "<u4"), ("same_type", "<u4"), ("n_types", "<u4"), ("noise_frequency", "<f4"),
         ("voxel_type_frequency", "<f4"), ("seed", "<u4")],
        align=True,
    )

    def __init__(self, kind, same_type=0, n_types=1, noise_frequency=0.0, voxel_type_frequency=0.0, seed=0):
        self.kind, self.same_type, self.n_types = kind, same_type, n_types
        self.noise_frequency, self.voxel_type_frequency, self.seed = noise_frequency, voxel_type_frequency, seed

    @classmethod
    def same(cls, voxel_type: int) -> "VoxelTypeGenerator":
        return cls(0, same_type=voxel_type)

    @classmethod
    def gradient_noise(cls, voxel_types, noise_frequency, voxel_type_frequency, seed) -> "VoxelTypeGenerator":
        assert len(voxel_types) > 0
        return cls(1, n_types=len(voxel_types), noise_frequency=noise_frequency,
                   voxel_type_frequency=voxel_type_frequency, seed=seed)

    def pod(self) -> np.ndarray:
        a = np.zeros(1, dtype=self.DTYPE)
        a[0] = (self.kind, self.same_type, self.n_types, self.noise_frequency, self.voxel_type_frequency, self.seed)
        return a
