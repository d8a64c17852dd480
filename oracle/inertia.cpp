// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).
//
// Restates (V = engine/crates/impact_voxel/src):
//   compute_moments_for_voxel                          V/object/inertia.rs:591-625
//   compute_moments_for_non_uniform_chunk              V/object/inertia.rs:629-706
//   compute_moments_for_uniform_chunk                  V/object/inertia.rs:710-752
//   compute_inertial_property_moments_for_object       V/object/inertia.rs:754-789
//   VoxelObjectInertialPropertyUpdater::remove_voxel   V/object/inertia.rs:377-394
//   VoxelObjectInertialPropertyManager::derive_inertial_properties  V/object/inertia.rs:160-167, 293-326
//   InertiaTensor::compute_delta_to_com_*              impact_physics/src/inertia.rs:511-546
// All sums are sequential f32 sums in the reference's visiting order; nothing is reassociated.
#include <cmath>

#include "oracle.hpp"

namespace orc {

void moments_for_voxel(float e, float e2, float e3, const float* densities, const uint32_t ijk[3], uint8_t type,
                       InertialMoments& out) {
    const float density = densities[type];
    float h2[3], h3[3];
    for (int d = 0; d < 3; ++d) {
        const float lo = e * (float)ijk[d];
        const float hi = lo + e;
        const float lo2 = lo * lo, hi2 = hi * hi;
        const float lo3 = lo2 * lo, hi3 = hi2 * hi;
        h2[d] = hi2 - lo2;
        h3[d] = hi3 - lo3;
    }
    out.mass = e3 * density;
    const float fm = (0.5f * e2) * density;
    const float fi = ((1.0f / 3.0f) * e2) * density;
    const float fp = (0.25f * e) * density;
    out.moments[0] = fm * h2[0];
    out.moments[1] = fm * h2[1];
    out.moments[2] = fm * h2[2];
    out.moi[0] = fi * (h3[1] + h3[2]);
    out.moi[1] = fi * (h3[0] + h3[2]);
    out.moi[2] = fi * (h3[0] + h3[1]);
    out.poi[0] = fp * (h2[0] * h2[1]);
    out.poi[1] = fp * (h2[1] * h2[2]);
    out.poi[2] = fp * (h2[2] * h2[0]);
}

void moments_for_non_uniform_chunk(float e, const Voxel* voxels, const float* densities, const uint32_t cc[3],
                                   InertialMoments& out) {
    float mass = 0.0f, m[3] = {0, 0, 0}, mi[3] = {0, 0, 0}, pi[3] = {0, 0, 0};
    const float x0 = (float)(cc[0] * CHUNK_SIZE) * e;
    const float y0 = (float)(cc[1] * CHUNK_SIZE) * e;
    const float z0 = (float)(cc[2] * CHUNK_SIZE) * e;
    int idx = 0;
    float xl = x0, xh = xl + e;
    for (int i = 0; i < CHUNK_SIZE; ++i) {
        float yl = y0, yh = yl + e;
        const float xl2 = xl * xl, xh2 = xh * xh;
        const float xl3 = xl2 * xl, xh3 = xh2 * xh;
        const float h2x = xh2 - xl2, h3x = xh3 - xl3;
        for (int j = 0; j < CHUNK_SIZE; ++j) {
            float zl = z0, zh = zl + e;
            const float yl2 = yl * yl, yh2 = yh * yh;
            const float yl3 = yl2 * yl, yh3 = yh2 * yh;
            const float h2y = yh2 - yl2, h3y = yh3 - yl3;
            for (int k = 0; k < CHUNK_SIZE; ++k) {
                const Voxel& v = voxels[idx];
                if (!(v.flags & FLAG_EMPTY)) {
                    const float zl2 = zl * zl, zh2 = zh * zh;
                    const float zl3 = zl2 * zl, zh3 = zh2 * zh;
                    const float h2z = zh2 - zl2, h3z = zh3 - zl3;
                    const float d = densities[v.type];
                    mass += d;
                    m[0] += d * h2x;
                    m[1] += d * h2y;
                    m[2] += d * h2z;
                    mi[0] += d * (h3y + h3z);
                    mi[1] += d * (h3x + h3z);
                    mi[2] += d * (h3x + h3y);
                    pi[0] += d * (h2x * h2y);
                    pi[1] += d * (h2y * h2z);
                    pi[2] += d * (h2z * h2x);
                }
                ++idx;
                zl = zh;
                zh += e;
            }
            yl = yh;
            yh += e;
        }
        xl = xh;
        xh += e;
    }
    const float e2 = e * e, e3 = e2 * e;
    const float fm = 0.5f * e2, fi = (1.0f / 3.0f) * e2, fp = 0.25f * e;
    out.mass = mass * e3;
    for (int d = 0; d < 3; ++d) {
        out.moments[d] = m[d] * fm;
        out.moi[d] = mi[d] * fi;
        out.poi[d] = pi[d] * fp;
    }
}

void moments_for_uniform_chunk(float e, const float* densities, uint8_t type, const uint32_t cc[3],
                               InertialMoments& out) {
    const float density = densities[type];
    const float ce = (float)CHUNK_SIZE * e;
    float h2[3], h3[3];
    for (int d = 0; d < 3; ++d) {
        const float lo = (float)cc[d] * ce;
        const float hi = lo + ce;
        const float lo2 = lo * lo, hi2 = hi * hi;
        const float lo3 = lo2 * lo, hi3 = hi2 * hi;
        h2[d] = hi2 - lo2;
        h3[d] = hi3 - lo3;
    }
    const float ce2 = ce * ce, ce3 = ce2 * ce;
    out.mass = ce3 * density;
    const float fm = (0.5f * ce2) * density;
    const float fi = ((1.0f / 3.0f) * ce2) * density;
    const float fp = (0.25f * ce) * density;
    out.moments[0] = fm * h2[0];
    out.moments[1] = fm * h2[1];
    out.moments[2] = fm * h2[2];
    out.moi[0] = fi * (h3[1] + h3[2]);
    out.moi[1] = fi * (h3[0] + h3[2]);
    out.moi[2] = fi * (h3[0] + h3[1]);
    out.poi[0] = fp * (h2[0] * h2[1]);
    out.poi[1] = fp * (h2[1] * h2[2]);
    out.poi[2] = fp * (h2[2] * h2[0]);
}

static inline void add_to(InertialMoments& a, const InertialMoments& b) {
    a.mass += b.mass;
    for (int d = 0; d < 3; ++d) {
        a.moments[d] += b.moments[d];
        a.moi[d] += b.moi[d];
        a.poi[d] += b.poi[d];
    }
}

// VoxelObjectInertialPropertyManager::initialized_from: chunks of the occupied chunk range, i → j → k
void inertial_moments_for_object(const Object& obj, const float* densities, InertialMoments& out,
                                 InertialMoments* per_chunk) {
    out = InertialMoments{};
    for (uint32_t i = obj.occ_chunks[0][0]; i < obj.occ_chunks[0][1]; ++i)
        for (uint32_t j = obj.occ_chunks[1][0]; j < obj.occ_chunks[1][1]; ++j)
            for (uint32_t k = obj.occ_chunks[2][0]; k < obj.occ_chunks[2][1]; ++k) {
                const uint32_t cidx = obj.lin(i, j, k);
                const Chunk& c = obj.chunks[cidx];
                const uint32_t cc[3] = {i, j, k};
                InertialMoments part{};
                if (c.kind == CK_NONUNIFORM)
                    moments_for_non_uniform_chunk(obj.voxel_extent, obj.chunk_voxels(c.data_offset), densities, cc, part);
                else if (c.kind == CK_UNIFORM)
                    moments_for_uniform_chunk(obj.voxel_extent, densities, c.uniform_voxel.type, cc, part);
                else
                    continue;
                add_to(out, part);
                if (per_chunk) per_chunk[cidx] = part;
            }
}

void InertialUpdater::remove_voxel(const uint32_t ijk[3], uint8_t type) {
    InertialMoments v;
    moments_for_voxel(e, e2, e3, densities, ijk, type, v);
    parent->mass -= v.mass;
    for (int d = 0; d < 3; ++d) {
        parent->moments[d] -= v.moments[d];
        parent->moi[d] -= v.moi[d];
        parent->poi[d] -= v.poi[d];
    }
    removed++;
}

}  // namespace orc
