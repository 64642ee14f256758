// sep_core.h -- cores of the separable (untilted-view) forward projector.
//
// When alpha = beta = 0 (the default pose set of ProjectionMatrix.projection_matrix,
// utilities/projection_operators.py:30-34, and every "known geometry" reconstruction) the sample lattice
// p = P00 + ix U + iz W + j D has W = (0, 0, W_z) and U_z = D_z = 0: the z coordinate of a sample depends on iz
// only, its (x, y) on (ix, j) only.  The trilinear sum then factors exactly:
//     proj[ix, iz] = (1 - wz) S[ix, fz] + wz S[ix, fz + 1],   S[ix, z] = sum_j bilinear_xy(vol[:, :, z]; x_j, y_j)
// (same terms as src/ray_wt_grad.f90:20-91, regrouped; the zero border of the padded volume again stands in for
// the per-corner bounds checks).  S is evaluated for four z planes per thread with 128-bit loads: cell and
// weights are computed once per (ix, j) instead of once per sample.
#pragma once
#include "ray_core.h"

TOMO_HD float tomo_tent_f(float d) { return fmaxf(0.f, 1.f - fabsf(d)); }

#define SEP_CHUNK 128                    // z planes staged per warp
#define SEP_OUT   (SEP_CHUNK - 4)        // planes a chunk produces outputs for (needs S[z] and S[z+1])

struct SepSetup {
    double p[2], D[2];
    int sg[2], st[2];
    int j0, j1, stepoff;
};

// xy-only clip of ray ix: samples with -1 <= x < nx and -1 <= y < ny (any z)
TOMO_HD void sep_setup(const double* __restrict__ V, const RayDims dm, int ix, SepSetup& r)
{
    const double N[2] = {(double)dm.nx, (double)dm.ny};
    const int ust[2] = {dm.sxp, dm.syp};
    double jlo = 0.0, jhi = V[V_N];
    bool empty = false;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        r.p[a] = V[V_P00 + a] + (double)ix * V[V_U + a];
        r.D[a] = V[V_D + a];
        if (r.D[a] > 0.0) {
            jlo = fmax(jlo, (-1.0 - r.p[a]) * V[V_INVD + a]);
            jhi = fmin(jhi, (N[a] - r.p[a]) * V[V_INVD + a]);
        } else if (r.D[a] < 0.0) {
            jlo = fmax(jlo, (N[a] - r.p[a]) * V[V_INVD + a]);
            jhi = fmin(jhi, (-1.0 - r.p[a]) * V[V_INVD + a]);
        } else if (r.p[a] <= -1.0 || r.p[a] >= N[a]) {
            empty = true;
        }
        r.sg[a] = (r.D[a] < 0.0) ? -1 : 1;
        r.st[a] = r.sg[a] * ust[a];
    }
    jlo = fmin(fmax(jlo, 0.0), V[V_N]);
    jhi = fmin(fmax(jhi, -1.0), V[V_N]);
    r.j0 = (int)ceil(jlo);
    r.j1 = (int)floor(jhi) + 1;
    if (r.j1 > (int)V[V_N]) r.j1 = (int)V[V_N];
    if (empty) r.j1 = r.j0;
    r.stepoff = (int)floor(fabs(r.D[0])) * r.st[0] + (int)floor(fabs(r.D[1])) * r.st[1];
}

struct f4 { float x, y, z, w; };

TOMO_HD f4 sep_ld4(const float* p)
{
#if defined(__CUDA_ARCH__)
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    f4 r; r.x = p[0]; r.y = p[1]; r.z = p[2]; r.w = p[3]; return r;
#endif
}

// S[ix, zq .. zq+3] for padded planes zq .. zq+3 (zq a multiple of 4)
TOMO_HD void sep_march_xy(const float* __restrict__ vol, const double* __restrict__ V, const RayDims dm,
                          const SepSetup& r, int zq, float S[4])
{
    const int ust[2] = {dm.sxp, dm.syp};
    float df[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) { const double ad = fabs(r.D[a]); df[a] = (float)(ad - floor(ad)); }
    const int o01 = r.st[1], o10 = r.st[0], o11 = r.st[0] + r.st[1];
    f2 a01 = f2_make(0.f, 0.f), a23 = f2_make(0.f, 0.f);
    for (int jc = r.j0; jc < r.j1; jc += RAY_REBASE) {
        float f[2];
        int off = zq;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const double q = (double)r.sg[a] * (r.p[a] + (double)jc * r.D[a]);
            const double qi = floor(q);
            f[a] = (float)(q - qi);
            int i = (int)qi;
            if (f[a] >= 1.0f) { f[a] -= 1.0f; i += 1; }
            off += (TOMO_PAD + r.sg[a] * i) * ust[a];
        }
        const int jend = (jc + RAY_REBASE < r.j1) ? jc + RAY_REBASE : r.j1;
        for (int j = jc; j < jend; ++j) {
            const float* __restrict__ c = vol + off;
            const f4 v00 = sep_ld4(c), v01 = sep_ld4(c + o01), v10 = sep_ld4(c + o10), v11 = sep_ld4(c + o11);
            // bilinear weights of the four (x, y) corners: (1-fx)(1-fy), (1-fx)fy, fx(1-fy), fx fy
            const float w11 = f[0] * f[1], w10 = f[0] - w11, w01 = f[1] - w11, w00 = (1.0f - f[0]) - w01;
            a01 = f2_fma(f2_make(w00, w00), f2_make(v00.x, v00.y), a01); a23 = f2_fma(f2_make(w00, w00), f2_make(v00.z, v00.w), a23);
            a01 = f2_fma(f2_make(w01, w01), f2_make(v01.x, v01.y), a01); a23 = f2_fma(f2_make(w01, w01), f2_make(v01.z, v01.w), a23);
            a01 = f2_fma(f2_make(w10, w10), f2_make(v10.x, v10.y), a01); a23 = f2_fma(f2_make(w10, w10), f2_make(v10.z, v10.w), a23);
            a01 = f2_fma(f2_make(w11, w11), f2_make(v11.x, v11.y), a01); a23 = f2_fma(f2_make(w11, w11), f2_make(v11.z, v11.w), a23);
            f[0] += df[0]; f[1] += df[1];
            off += r.stepoff;
            if (f[0] >= 1.0f) { f[0] -= 1.0f; off += r.st[0]; }
            if (f[1] >= 1.0f) { f[1] -= 1.0f; off += r.st[1]; }
        }
    }
    S[0] = a01.x; S[1] = a01.y; S[2] = a23.x; S[3] = a23.y;
}

// z cell (padded plane index of the floor corner) and ceil weight of detector row iz
TOMO_HD void sep_zcell(const double* __restrict__ V, int iz, int& fzp, float& wz)
{
    const double zs = V[V_P00 + 2] + (double)iz * V[V_W + 2];
    const double fl = floor(zs);
    wz = (float)(zs - fl);
    fzp = (int)fmin(fmax(fl, -1.0e6), 1.0e6) + TOMO_PAD;
}

// ---- separable adjoint (untilted views) -----------------------------------------------------------------------
// vol[x, y, z] += sum_ix K(ix; x, y) * Yz[ix, z]   with
//   Yz[ix, z]    = sum_iz tent(z0 + iz W_z - z) * y[ix, iz]          (transpose of the 2-tap z interpolation)
//   K(ix; x, y)  = sum_j tent(x_j - x) tent(y_j - y)                 (transpose of the bilinear (x, y) part)
// -- the same terms as the generic adjoint (back_core.h), regrouped; exact because z decouples when W = (0,0,W_z).

// Yz[ix, z] for one voxel plane z: the (at most 3 for W_z > 2/3) detector rows within one voxel of it
TOMO_HD float sep_zgather(const float* __restrict__ Prow, const double* __restrict__ V, int ndz, int z)
{
    const double z0 = V[V_P00 + 2], wz = V[V_W + 2];
    const double lo = ((double)z - 1.0 - z0) / wz, hi = ((double)z + 1.0 - z0) / wz;
    int i0 = (int)fmin(fmax(ceil(lo), 0.0), (double)ndz), i1 = (int)fmin(fmax(floor(hi), -1.0), (double)(ndz - 1));
    float acc = 0.f;
    for (int iz = i0; iz <= i1; ++iz) {
        const float d = (float)(z0 + (double)iz * wz - (double)z);
        acc = fmaf(tomo_tent_f(d), TOMO_LDG(Prow + iz), acc);
    }
    return acc;
}

// Candidate ix range and the in-plane lattice coordinates of voxel column (x, y)
struct SepColumn {
    int n0i, n0j, milo, mihi, mjlo, mjhi;
    float rhoi, rhoj;
};

TOMO_HD void sep_column_setup(const double* __restrict__ V, int ndx, int x, int y, SepColumn& c)
{
    const double vx = (double)x - V[V_P00 + 0], vy = (double)y - V[V_P00 + 1];
    const double qi = V[V_LINV + 0] * vx + V[V_LINV + 1] * vy;          // lattice ix of the column
    const double qj = V[V_LINV + 6] * vx + V[V_LINV + 7] * vy;          // lattice j
    const double ri = rint(fmin(fmax(qi, -1.0e9), 1.0e9)), rj = rint(fmin(fmax(qj, -1.0e9), 1.0e9));
    c.n0i = (int)ri; c.n0j = (int)rj;
    c.rhoi = (float)(qi - ri); c.rhoj = (float)(qj - rj);
    const float rbi = (float)(fabs(V[V_LINV + 0]) + fabs(V[V_LINV + 1])) * 1.0001f + 1e-4f;
    const float rbj = (float)(fabs(V[V_LINV + 6]) + fabs(V[V_LINV + 7])) * 1.0001f + 1e-4f;
    c.milo = (int)ceilf(c.rhoi - rbi); c.mihi = (int)floorf(c.rhoi + rbi);
    c.mjlo = (int)ceilf(c.rhoj - rbj); c.mjhi = (int)floorf(c.rhoj + rbj);
    const int nj = (int)V[V_N];
    if (c.milo < -c.n0i) c.milo = -c.n0i;
    if (c.mihi > ndx - 1 - c.n0i) c.mihi = ndx - 1 - c.n0i;
    if (c.mjlo < -c.n0j) c.mjlo = -c.n0j;
    if (c.mjhi > nj - 1 - c.n0j) c.mjhi = nj - 1 - c.n0j;
}

// K(ix = n0i + mi; x, y)
TOMO_HD float sep_column_weight(const double* __restrict__ V, const SepColumn& c, int mi)
{
    const float Ux = (float)V[V_U], Uy = (float)V[V_U + 1], Dx = (float)V[V_D], Dy = (float)V[V_D + 1];
    const float ci = (float)mi - c.rhoi;
    float K = 0.f;
    for (int mj = c.mjlo; mj <= c.mjhi; ++mj) {
        const float cj = (float)mj - c.rhoj;
        K = fmaf(tomo_tent_f(fmaf(ci, Ux, cj * Dx)), tomo_tent_f(fmaf(ci, Uy, cj * Dy)), K);
    }
    return K;
}

// ---- separable projection + gradient (untilted views) --------------------------------------------------------
// Per (ix, z plane) the march accumulates six moments of the in-plane bilinear interpolant B and its one-sided
// in-plane derivatives (same cells as src/ray_wt_grad.f90:142-220: the (x, y) fraction is carried as 64-bit fixed
// point so every cell is the float64 one):
//     S = sum B,  T = sum j B,  Gx = sum dB/dx,  Gy = sum dB/dy,  TGx = sum j dB/dx,  TGy = sum j dB/dy
// and a ray (ix, iz) with z cell fz and weight wz gets
//     proj = lerp(S),  S0 = (lerp(Gx), lerp(Gy), S[fz+1] - S[fz]),  S1 = (lerp(TGx), lerp(TGy), T[fz+1] - T[fz]).
struct SepMoments { float S[4], T[4], Gx[4], Gy[4], TGx[4], TGy[4]; };

TOMO_HD f2 f2_add(f2 a, f2 b) { return f2_fma(a, f2_make(1.f, 1.f), b); }

TOMO_HD void sep_march_xy_grad(const float* __restrict__ vol, const double* __restrict__ V, const RayDims dm,
                               const SepSetup& r, int zq, SepMoments& m)
{
    const int ust[2] = {dm.sxp, dm.syp};
    const int o01 = r.st[1], o10 = r.st[0], o11 = r.st[0] + r.st[1];
    unsigned fh[2], fl[2], dh[2], dl[2];
    int off = zq;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const double q = (double)r.sg[a] * (r.p[a] + (double)r.j0 * r.D[a]);
        const double qi = floor(q);
        const unsigned long long f64 = (unsigned long long)((q - qi) * 18446744073709551616.0);
        const double ad = fabs(r.D[a]);
        const unsigned long long d64 = (unsigned long long)((ad - floor(ad)) * 18446744073709551616.0);
        fh[a] = (unsigned)(f64 >> 32); fl[a] = (unsigned)f64;
        dh[a] = (unsigned)(d64 >> 32); dl[a] = (unsigned)d64;
        off += (TOMO_PAD + r.sg[a] * (int)qi) * ust[a];
    }
    f2 S0 = f2_make(0.f, 0.f), S1 = S0, T0 = S0, T1 = S0, X0 = S0, X1 = S0, Y0 = S0, Y1 = S0, TX0 = S0, TX1 = S0, TY0 = S0, TY1 = S0;
    float fj = (float)r.j0;
    for (int j = r.j0; j < r.j1; ++j) {
        const float fx = fix_to_float(fh[0]), fy = fix_to_float(fh[1]);
        const float* __restrict__ c = vol + off;
        const f4 v00 = sep_ld4(c), v01 = sep_ld4(c + o01), v10 = sep_ld4(c + o10), v11 = sep_ld4(c + o11);
        const f2 fy2 = f2_make(fy, fy), fx2 = f2_make(fx, fx), fj2 = f2_make(fj, fj);
        // planes 0,1 (.x, .y) and planes 2,3 (.z, .w) as two register pairs
#define SEP_GRAD_HALF(a00, a01, a10, a11, Sq, Tq, Xq, Yq, TXq, TYq)                                   \
        {                                                                                             \
            const f2 dy0 = f2_sub(a01, a00), dy1 = f2_sub(a11, a10);                                  \
            const f2 y0 = f2_fma(fy2, dy0, a00), y1 = f2_fma(fy2, dy1, a10);                          \
            const f2 gx = f2_sub(y1, y0);                                                             \
            const f2 val = f2_fma(fx2, gx, y0);                                                       \
            const f2 gy = f2_fma(fx2, f2_sub(dy1, dy0), dy0);                                         \
            Sq = f2_add(val, Sq);  Tq = f2_fma(fj2, val, Tq);                                         \
            Xq = f2_add(gx, Xq);   TXq = f2_fma(fj2, gx, TXq);                                        \
            Yq = f2_add(gy, Yq);   TYq = f2_fma(fj2, gy, TYq);                                        \
        }
        SEP_GRAD_HALF(f2_make(v00.x, v00.y), f2_make(v01.x, v01.y), f2_make(v10.x, v10.y), f2_make(v11.x, v11.y), S0, T0, X0, Y0, TX0, TY0)
        SEP_GRAD_HALF(f2_make(v00.z, v00.w), f2_make(v01.z, v01.w), f2_make(v10.z, v10.w), f2_make(v11.z, v11.w), S1, T1, X1, Y1, TX1, TY1)
#undef SEP_GRAD_HALF
        fj += 1.0f;
        off += r.stepoff;
        off += (int)fix64_add(fh[0], fl[0], dh[0], dl[0]) * r.st[0];
        off += (int)fix64_add(fh[1], fl[1], dh[1], dl[1]) * r.st[1];
    }
    const float sx = (float)r.sg[0], sy = (float)r.sg[1];      // derivatives w.r.t. the real (un-mirrored) coordinates
    m.S[0] = S0.x; m.S[1] = S0.y; m.S[2] = S1.x; m.S[3] = S1.y;
    m.T[0] = T0.x; m.T[1] = T0.y; m.T[2] = T1.x; m.T[3] = T1.y;
    m.Gx[0] = sx * X0.x; m.Gx[1] = sx * X0.y; m.Gx[2] = sx * X1.x; m.Gx[3] = sx * X1.y;
    m.Gy[0] = sy * Y0.x; m.Gy[1] = sy * Y0.y; m.Gy[2] = sy * Y1.x; m.Gy[3] = sy * Y1.y;
    m.TGx[0] = sx * TX0.x; m.TGx[1] = sx * TX0.y; m.TGx[2] = sx * TX1.x; m.TGx[3] = sx * TX1.y;
    m.TGy[0] = sy * TY0.x; m.TGy[1] = sy * TY0.y; m.TGy[2] = sy * TY1.x; m.TGy[3] = sy * TY1.y;
}

// Ray sums from the six plane arrays (each indexed by plane; k = floor plane index into them)
TOMO_HD void sep_ray_sums(const float* S, const float* T, const float* Gx, const float* Gy, const float* TGx,
                          const float* TGy, int k, float wz, RaySums& out)
{
    out.acc = fmaf(wz, S[k + 1] - S[k], S[k]);
    out.s0[0] = fmaf(wz, Gx[k + 1] - Gx[k], Gx[k]);
    out.s0[1] = fmaf(wz, Gy[k + 1] - Gy[k], Gy[k]);
    out.s0[2] = S[k + 1] - S[k];
    out.s1[0] = fmaf(wz, TGx[k + 1] - TGx[k], TGx[k]);
    out.s1[1] = fmaf(wz, TGy[k + 1] - TGy[k], TGy[k]);
    out.s1[2] = T[k + 1] - T[k];
}
