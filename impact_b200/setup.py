"""Host-side mirror of the reference's voxel object setup (engine/crates/impact_voxel/src/setup.rs): the setup components
that describe an object as a shape + optional noise modification + voxel types, and `setup_voxel_object`, which turns a
generator into a meshed object with its collision probes. Everything heavy happens in the library (ivx_program_build,
ivx_object_generate, ivx_object_mesh, ivx_object_collision_probes); this file only keeps the reference's names, argument
meaning and assertions so that callers and tests read like the reference's.

    VoxelBox / VoxelSphere / VoxelCapsule / VoxelSphereUnion .add(graph)      setup.rs:329-527
    MultifractalNoiseSDFModification + apply_modifications                   setup.rs:286-327, 529-553
    SameVoxelType / GradientNoiseVoxelTypes → VoxelTypeGenerator             setup.rs:184-284
    GeneratedVoxelObject (meta graph + scale factor + seed)                  setup.rs:44-51, 169-182
    setup_voxel_object                                                       setup.rs:555-579
"""
from __future__ import annotations

import numpy as np

from .graph import SDFGraph, VoxelTypeGenerator


class VoxelBox:
    """A box with the given voxel extent and number of voxels in each direction (setup.rs:329-372)."""

    def __init__(self, voxel_extent: float, extent_x: float, extent_y: float, extent_z: float):
        assert voxel_extent > 0.0 and extent_x >= 0.0 and extent_y >= 0.0 and extent_z >= 0.0
        self.voxel_extent, self.extents = voxel_extent, [extent_x, extent_y, extent_z]

    def extents_in_voxels(self):
        return list(self.extents)

    def add(self, graph: SDFGraph) -> int:
        return graph.box(self.extents_in_voxels())


class VoxelSphere:
    """A sphere with the given voxel extent and number of voxels across its radius (setup.rs:374-412)."""

    def __init__(self, voxel_extent: float, radius: float):
        assert voxel_extent > 0.0 and radius >= 0.0
        self.voxel_extent, self.radius = voxel_extent, radius

    def radius_in_voxels(self):
        return self.radius

    def add(self, graph: SDFGraph) -> int:
        return graph.sphere(self.radius_in_voxels())


class VoxelCapsule:
    """A capsule with the given voxel extent, segment length and radius in voxels (setup.rs:414-462)."""

    def __init__(self, voxel_extent: float, segment_length: float, radius: float):
        assert voxel_extent > 0.0 and segment_length >= 0.0 and radius > 0.0
        self.voxel_extent, self.segment_length, self.radius = voxel_extent, segment_length, radius

    def add(self, graph: SDFGraph) -> int:
        return graph.capsule(self.segment_length, self.radius)


class VoxelSphereUnion:
    """The smooth union of two spheres, the second one offset (in voxels) from the first (setup.rs:464-527)."""

    def __init__(self, voxel_extent: float, radius_1: float, radius_2: float, center_offsets, smoothness: float):
        assert voxel_extent > 0.0 and radius_1 >= 0.0 and radius_2 >= 0.0
        self.voxel_extent, self.radius_1, self.radius_2 = voxel_extent, radius_1, radius_2
        self.center_offsets, self.smoothness = [float(x) for x in center_offsets], smoothness

    def add(self, graph: SDFGraph) -> int:
        sphere_1 = graph.sphere(self.radius_1)
        sphere_2 = graph.translation(graph.sphere(self.radius_2), self.center_offsets)
        return graph.union(sphere_1, sphere_2, self.smoothness)


class MultifractalNoiseSDFModification:
    """A multifractal noise perturbation of the shape's distance field (setup.rs:286-327)."""

    def __init__(self, octaves: int, frequency: float, lacunarity: float, persistence: float, amplitude: float, seed: int):
        self.octaves, self.frequency, self.lacunarity = octaves, frequency, lacunarity
        self.persistence, self.amplitude, self.seed = persistence, amplitude, seed


def apply_modifications(graph: SDFGraph, node_id: int, multifractal_noise_modification=None) -> None:
    """setup.rs:529-553: the modification becomes the node above `node_id` (and the graph's root)."""
    m = multifractal_noise_modification
    if m is not None:
        graph.multifractal_noise(node_id, m.octaves, m.frequency, m.lacunarity, m.persistence, m.amplitude, m.seed)


class SameVoxelType:
    """One voxel type for the whole object (setup.rs:53-62, 184-207); the type by its index in the registry."""

    def __init__(self, voxel_type: int):
        self.voxel_type = voxel_type

    def create_generator(self) -> VoxelTypeGenerator:
        return VoxelTypeGenerator.same(self.voxel_type)


class GradientNoiseVoxelTypes:
    """Voxel types distributed by a 4-D gradient noise pattern (setup.rs:64-80, 209-284)."""

    VOXEL_TYPE_ARRAY_SIZE = 256  # VoxelTypeRegistry::max_n_voxel_types()

    def __init__(self, voxel_types, noise_frequency: float, voxel_type_frequency: float, seed: int):
        voxel_types = list(voxel_types)
        assert 0 < len(voxel_types) <= self.VOXEL_TYPE_ARRAY_SIZE
        self.voxel_types, self.noise_frequency = voxel_types, noise_frequency
        self.voxel_type_frequency, self.seed = voxel_type_frequency, seed

    def create_generator(self) -> VoxelTypeGenerator:
        return VoxelTypeGenerator.gradient_noise(self.voxel_types, self.noise_frequency, self.voxel_type_frequency, self.seed)


class GeneratedVoxelObject:
    """An object generated from a meta SDF graph (setup.rs:44-51, 169-182): the generator's meta nodes, the voxel extent,
    the scale factor the graph is compiled with and the seed."""

    def __init__(self, meta_nodes, voxel_extent: float, scale_factor: float, seed: int):
        assert voxel_extent > 0.0 and scale_factor > 0.0
        self.meta_nodes, self.voxel_extent, self.scale_factor, self.seed = meta_nodes, voxel_extent, scale_factor, seed

    def build_graph(self, ctx=None) -> SDFGraph:
        from .meta import compile_meta_nodes

        return compile_meta_nodes(self.meta_nodes, self.scale_factor, self.seed, ctx)


def create_sdf_generator(ctx, shape, multifractal_noise_modification=None):
    """The graph of a shape component with its modification, compiled (scene setup, engine/src/setup/scene/voxel.rs:87-118)."""
    graph = SDFGraph()
    node_id = shape.add(graph)
    apply_modifications(graph, node_id, multifractal_noise_modification)
    return ctx.build_generator(graph)


def setup_voxel_object(generator):
    """`setup_voxel_object` (setup.rs:555-579): `VoxelObject::generate` + `MeshedVoxelObject::create` (mesh and collision
    probes) → (object, mesh, probes). The reference then hands the meshed object to its `VoxelObjectManager`."""
    from .voxel import VoxelObject, VoxelObjectMesh

    obj = VoxelObject.generate(generator)
    mesh = VoxelObjectMesh.create(obj)
    probes = mesh.collision_probes()
    return obj, mesh, probes
