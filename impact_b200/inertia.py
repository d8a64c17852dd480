"""Host-side mirror of `VoxelObjectInertialPropertyManager` (engine/crates/impact_voxel/src/object/inertia.rs).

The ten sums (mass, moments, moments of inertia, products of inertia about the voxel-grid origin) come from the GPU
(`ivx_object_inertial_moments`, bit for bit the reference's f32 sums); everything this module does with them is a
handful of f32 operations in the reference's order:

  initialized_from                inertia.rs:125-137   → VoxelObjectInertialPropertyManager.initialized_from
  derive_center_of_mass           inertia.rs:169-172
  derive_inertial_properties      inertia.rs:160-167, 293-326 (+ impact_physics/src/inertia.rs:511-546)
  add / offset_reference_point_by inertia.rs:243-268 (+ impact_physics/src/inertia.rs:548-587)
  begin_update().remove_voxel     inertia.rs:374-395, 591-625 (single voxels, host side)

`Matrix3::inverse` is glam's `Mat3A::inverse` (third party, cross-product form); it is restated here from its published
algorithm and only enters the inverse inertia tensor.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _v3(x, y, z) -> np.ndarray:
    return np.array([x, y, z], F)


def compute_delta_to_com_moments_and_products_of_inertia(mass, d: np.ndarray):
    """impact_physics/src/inertia.rs:534-546 (parallel axis theorem)."""
    mass = F(mass)
    sq = d * d
    moi = -mass * (sq[[1, 2, 0]] + sq[[2, 0, 1]])
    poi = -mass * (d * d[[1, 2, 0]])
    return moi.astype(F), poi.astype(F)


def compute_delta_to_com_inertia_matrix(mass, d: np.ndarray) -> np.ndarray:
    """impact_physics/src/inertia.rs:511-526; returns the 3 x 3 matrix (symmetric)."""
    moi, poi = compute_delta_to_com_moments_and_products_of_inertia(mass, d)
    sxy, syz, szx = -poi
    return np.array([[moi[0], sxy, szx], [sxy, moi[1], syz], [szx, syz, moi[2]]], F)


def _mat3_inverse(m: np.ndarray) -> np.ndarray:
    """glam `Mat3A::inverse`: cross products of the columns over the determinant, transposed."""
    x, y, z = m[:, 0].astype(F), m[:, 1].astype(F), m[:, 2].astype(F)
    t0, t1, t2 = np.cross(y, z).astype(F), np.cross(z, x).astype(F), np.cross(x, y).astype(F)
    det = F(np.dot(z, t2))
    inv_det = F(1.0) / det
    return np.stack([t0 * inv_det, t1 * inv_det, t2 * inv_det], axis=0).astype(F)  # rows = transposed columns


class InertialProperties:
    """`impact_physics::inertia::InertialProperties`: mass, centre of mass, inertia tensor about it (+ inverse)."""

    def __init__(self, mass, center_of_mass, inertia_tensor, inverse_inertia_tensor):
        self.mass = F(mass)
        self.center_of_mass = np.asarray(center_of_mass, F)
        self.inertia_tensor = np.asarray(inertia_tensor, F)
        self.inverse_inertia_tensor = np.asarray(inverse_inertia_tensor, F)

    @classmethod
    def of_uniform_box(cls, ex, ey, ez, density) -> "InertialProperties":
        """impact_physics/src/inertia.rs:87-99 (centre of mass at the origin)."""
        ex, ey, ez, density = F(ex), F(ey), F(ez), F(density)
        mass = (ex * ey * ez) * density
        c = F(1.0 / 12.0)
        diag = _v3(c * mass * (ey * ey + ez * ez), c * mass * (ex * ex + ez * ez), c * mass * (ex * ex + ey * ey))
        return cls(mass, _v3(0, 0, 0), np.diag(diag), np.diag(F(1.0) / diag))

    def translated(self, t) -> "InertialProperties":
        """`transform` with a pure translation (inertia.rs:248-262): only the centre of mass moves."""
        return InertialProperties(self.mass, self.center_of_mass + np.asarray(t, F), self.inertia_tensor,
                                  self.inverse_inertia_tensor)


class VoxelObjectInertialPropertyUpdater:
    """`VoxelObjectInertialPropertyUpdater` (inertia.rs:27-35): removes single voxels from the sums."""

    def __init__(self, parent: "VoxelObjectInertialPropertyManager", voxel_extent, voxel_type_densities):
        self.parent = parent
        self.e = F(voxel_extent)
        self.e2 = self.e * self.e
        self.e3 = self.e2 * self.e
        self.densities = np.asarray(voxel_type_densities, F)

    def remove_voxel(self, object_voxel_indices, voxel_type: int):
        m = compute_moments_for_voxel(self.e, self.e2, self.e3, self.densities, object_voxel_indices, voxel_type)
        self.parent.m = (self.parent.m - m).astype(F)


def compute_moments_for_voxel(e, e2, e3, densities, ijk, voxel_type: int) -> np.ndarray:
    """inertia.rs:591-625 → the voxel's ten terms."""
    density = F(densities[voxel_type])
    lo = (e * np.asarray(ijk, F)).astype(F)
    hi = lo + e
    lo2, hi2 = lo * lo, hi * hi
    h2 = hi2 - lo2
    h3 = hi2 * hi - lo2 * lo
    out = np.zeros(10, F)
    out[0] = e3 * density
    out[1:4] = (F(0.5) * e2 * density) * h2
    out[4:7] = (F(1.0 / 3.0) * e2 * density) * (h3[[1, 0, 0]] + h3[[2, 2, 1]])
    out[7:10] = (F(0.25) * e * density) * (h2 * h2[[1, 2, 0]])
    return out


class VoxelObjectInertialPropertyManager:
    """`VoxelObjectInertialPropertyManager` (inertia.rs:19-25). `m` = the ten f32 sums."""

    def __init__(self, moments: np.ndarray):
        self.m = np.asarray(moments, F).copy()
        assert self.m.shape == (10,)

    @classmethod
    def zeroed(cls) -> "VoxelObjectInertialPropertyManager":
        return cls(np.zeros(10, F))

    @classmethod
    def initialized_from(cls, voxel_object, voxel_type_densities) -> "VoxelObjectInertialPropertyManager":
        """inertia.rs:125-137: integrates the object's voxels on the GPU."""
        return cls(voxel_object.inertial_moments(voxel_type_densities))

    mass = property(lambda self: self.m[0])
    moments = property(lambda self: self.m[1:4])
    moments_of_inertia = property(lambda self: self.m[4:7])
    products_of_inertia = property(lambda self: self.m[7:10])

    def derive_center_of_mass(self) -> np.ndarray:
        return (self.moments / self.mass).astype(F)

    def begin_update(self, voxel_extent, voxel_type_densities) -> VoxelObjectInertialPropertyUpdater:
        return VoxelObjectInertialPropertyUpdater(self, voxel_extent, voxel_type_densities)

    def add(self, other: "VoxelObjectInertialPropertyManager") -> "VoxelObjectInertialPropertyManager":
        return VoxelObjectInertialPropertyManager(self.m + other.m)

    def offset_reference_point_by(self, offset) -> None:
        """inertia.rs:255-268 with compute_delta_to_moments_and_products_of_inertia_defined_relative_to_point
        (impact_physics/src/inertia.rs:568-587)."""
        offset = np.asarray(offset, F)
        com = self.derive_center_of_mass()
        moi_a, poi_a = compute_delta_to_com_moments_and_products_of_inertia(self.mass, com)
        moi_b, poi_b = compute_delta_to_com_moments_and_products_of_inertia(self.mass, (offset - com).astype(F))
        self.m[1:4] = self.moments - offset * self.mass
        self.m[4:7] = self.moments_of_inertia + (moi_a + (-moi_b))
        self.m[7:10] = self.products_of_inertia + (poi_a + (-poi_b))

    def derive_inertial_properties(self) -> InertialProperties:
        """inertia.rs:293-326: centre of mass and the inertia tensor about it (and its inverse, computed on the
        mass-normalised matrix like the reference)."""
        mass = self.mass
        inv_mass = F(1.0) / mass
        com = (self.moments * inv_mass).astype(F)
        jx, jy, jz = self.moments_of_inertia
        pxy, pyz, pzx = self.products_of_inertia
        origin_tensor = np.array([[jx, -pxy, -pzx], [-pxy, jy, -pyz], [-pzx, -pyz, jz]], F)
        com_tensor = (origin_tensor + compute_delta_to_com_inertia_matrix(mass, com)).astype(F)
        scaled = (inv_mass * com_tensor).astype(F)
        inverse = (inv_mass * _mat3_inverse(scaled)).astype(F)
        return InertialProperties(mass, com, com_tensor, inverse)
