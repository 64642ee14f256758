// back_core.h -- per-voxel cores of the two backprojectors (__host__ __device__, see ray_core.h for
// why: tests/emu runs the same code on the CPU against the oracle).
#pragma once
#include <math.h>
#include <stddef.h>
#include "tomo_common.h"
#include "ray_core.h"   // TOMO_HD, TOMO_LDG

TOMO_HD float tomo_tent(float d) { return fmaxf(0.f, 1.f - fabsf(d)); }

// Contribution of one view to voxel (x, y, z) of vol = A^T y: gather over the sample lattice.
// The trilinear weight a sample at p gives voxel v is prod_axis tent(p_a - v_a)
// (src/ray_wt_grad.f90:35-89 read column-wise), so with q = Linv (v - P00) the lattice coordinates
// of the voxel, n0 = rint(q), rho = q - n0, every contributing lattice point n0 + m satisfies
// |m_k - rho_k| < sum_a |Linv[k][a]| = RB_k.  Distances d = L (m - rho) involve only small numbers.
TOMO_HD float adjoint_gather_view(const float* __restrict__ P, const double* __restrict__ V,
                                  int ndx, int ndz, int x, int y, int z)
{
    const double vx = (double)x - V[V_P00 + 0], vy = (double)y - V[V_P00 + 1], vz = (double)z - V[V_P00 + 2];
    int n0[3], mlo[3], mhi[3];
    float rho[3];
    const int lim[3] = {ndx, ndz, (int)V[V_N]};     // valid lattice: 0 <= ix < ndx, 0 <= iz < ndz, 0 <= j < n
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double q = V[V_LINV + 3 * k] * vx + V[V_LINV + 3 * k + 1] * vy + V[V_LINV + 3 * k + 2] * vz;
        const double qr = rint(fmin(fmax(q, -1.0e9), 1.0e9));
        n0[k] = (int)qr;
        rho[k] = (float)(q - qr);
        const float rb = (float)V[V_RB + k] * 1.0001f + 1e-4f;
        mlo[k] = (int)ceilf(rho[k] - rb);
        mhi[k] = (int)floorf(rho[k] + rb);
        if (mlo[k] < -n0[k]) mlo[k] = -n0[k];
        if (mhi[k] > lim[k] - 1 - n0[k]) mhi[k] = lim[k] - 1 - n0[k];
    }
    const float Ux = (float)V[V_U], Uy = (float)V[V_U + 1], Uz = (float)V[V_U + 2];
    const float Wx = (float)V[V_W], Wy = (float)V[V_W + 1], Wz = (float)V[V_W + 2];
    const float Dx = (float)V[V_D], Dy = (float)V[V_D + 1], Dz = (float)V[V_D + 2];
    float acc = 0.f;
    for (int mi = mlo[0]; mi <= mhi[0]; ++mi) {
        const float ci = (float)mi - rho[0];
        const float* __restrict__ Prow = P + (size_t)(n0[0] + mi) * ndz + n0[1];
        for (int mj = mlo[2]; mj <= mhi[2]; ++mj) {
            const float cj = (float)mj - rho[2];
            const float bx = fmaf(ci, Ux, cj * Dx), by = fmaf(ci, Uy, cj * Dy), bz = fmaf(ci, Uz, cj * Dz);
            for (int mk = mlo[1]; mk <= mhi[1]; ++mk) {
                const float ck = (float)mk - rho[1];
                const float w = tomo_tent(fmaf(ck, Wx, bx)) * tomo_tent(fmaf(ck, Wy, by)) * tomo_tent(fmaf(ck, Wz, bz));
                if (w > 0.f) acc = fmaf(w, TOMO_LDG(Prow + mk), acc);
            }
        }
    }
    return acc;
}

// Contribution of one view to a voxel under the orphan voxel-driven backprojector:
// x' = Ry (Rx Rz x + t) (src/external_back_projection.f90:17-25), 4 independently bounds-checked
// bilinear taps at (x'_x - origin_x, x'_z - origin_z); y' is never used (:47-66).
// (cx, cy, cz) is the physical voxel centre.
TOMO_HD float voxel_bilinear_view(const float* __restrict__ P, const double* __restrict__ V,
                                  int ndx, int ndz, const double origin[3], double cx, double cy, double cz)
{
    const double ux = V[V_VROT + 0] * cx + V[V_VROT + 1] * cy + V[V_VROT + 2] * cz + V[V_VTR + 0] - origin[0];
    const double uz = V[V_VROT + 6] * cx + V[V_VROT + 7] * cy + V[V_VROT + 8] * cz + V[V_VTR + 2] - origin[2];
    const double flx = floor(ux), flz = floor(uz);
    const float ax = (float)(ux - flx), az = (float)(uz - flz);
    const int fx = (int)fmin(fmax(flx, -2.0), 1.0e9), fz = (int)fmin(fmax(flz, -2.0), 1.0e9);
    const bool x0 = (fx >= 0 && fx < ndx), x1 = (fx + 1 >= 0 && fx + 1 < ndx);
    const bool z0 = (fz >= 0 && fz < ndz), z1 = (fz + 1 >= 0 && fz + 1 < ndz);
    const float* __restrict__ c = P + (ptrdiff_t)fx * ndz + fz;
    float v = 0.f;
    if (x0 && z0) v = fmaf(TOMO_LDG(c), (1.f - ax) * (1.f - az), v);
    if (x1 && z0) v = fmaf(TOMO_LDG(c + ndz), ax * (1.f - az), v);
    if (x0 && z1) v = fmaf(TOMO_LDG(c + 1), (1.f - ax) * az, v);
    if (x1 && z1) v = fmaf(TOMO_LDG(c + ndz + 1), ax * az, v);
    return v;
}
