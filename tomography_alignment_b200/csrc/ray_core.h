// ray_core.h -- per-ray core of the ray-driven forward projector / gradient.
//
// __host__ __device__ so that tests/emu can run exactly this code on the CPU against the oracle
// (test harness only; the product always runs it inside ray_kernels.cu on the GPU).
//
// What it computes, per ray (view, ix, iz):
//   acc = sum_j trilinear(vol, p_j),   p_j = P00 + ix*U + iz*W + j*D          (ray_voxel_utilities.py:89-94,
//                                                                             src/ray_wt_grad.f90:20-91)
//   S0  = sum_j G_j,  S1 = sum_j j*G_j,  G = spatial gradient of the interpolant (src/ray_wt_grad.f90:142-220)
// How:
//   * the ray is clipped to the samples with -1 <= p < N on every axis (all others have no
//     in-bounds corner); j keeps the reference's phase;
//   * the volume is the zero-bordered copy (TOMO_PAD), so no per-corner bounds checks;
//   * positions are carried as (cell offset, fraction): float32 fraction re-based from float64
//     every RAY_REBASE samples for the forward projector, exact 64-bit fixed point for the gradient;
//   * axes along which the ray runs backwards are mirrored (q = -p) so the carry is one-sided.
#pragma once
#include <math.h>
#include "tomo_common.h"

#ifdef __CUDACC__
#define TOMO_HD __host__ __device__ __forceinline__
#else
#define TOMO_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define TOMO_LDG(p) __ldg(p)
#else
#define TOMO_LDG(p) (*(p))
#endif

#ifndef RAY_GRAD_UNROLL
#define RAY_GRAD_UNROLL 1
#endif
#ifndef RAY_REBASE
#define RAY_REBASE 64
#endif

struct RayDims {
    int nx, ny, nz;      // volume shape
    int sxp, syp;        // padded strides (floats) along x and y; z stride is 1
};

struct RaySums {
    float acc;           // projection value
    float s0[3], s1[3];  // gradient moments (real, un-mirrored coordinates)
};

// Clipped sample range and mirrored-frame constants of one ray.
struct RaySetup {
    double p[3], D[3];   // p0 of the ray and the step, voxel-index coordinates
    int sg[3], st[3];    // mirror sign per axis, signed padded stride per axis
    int j0, j1;          // samples j0 <= j < j1 can touch the volume
    int stepoff;         // integer part of |D| (step_size > 1) folded into the address step
};

TOMO_HD void ray_setup(const double* __restrict__ V, const RayDims dm, int ix, int iz, RaySetup& r)
{
    const double N[3] = {(double)dm.nx, (double)dm.ny, (double)dm.nz};
    const int ust[3] = {dm.sxp, dm.syp, 1};
    double jlo = 0.0, jhi = V[V_N];
    bool empty = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        r.p[a] = V[V_P00 + a] + (double)ix * V[V_U + a] + (double)iz * V[V_W + a];
        r.D[a] = V[V_D + a];
        // a sample has an in-bounds corner iff -1 < p < N on every axis
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
    // clamp in float64 first: 1/D is astronomically large for near-axis-parallel rays
    jlo = fmin(fmax(jlo, 0.0), V[V_N]);
    jhi = fmin(fmax(jhi, -1.0), V[V_N]);
    r.j0 = (int)ceil(jlo);
    r.j1 = (int)floor(jhi) + 1;
    if (r.j1 > (int)V[V_N]) r.j1 = (int)V[V_N];
    if (empty) r.j1 = r.j0;
    r.stepoff = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) r.stepoff += (int)floor(fabs(r.D[a])) * r.st[a];
}

// Mirrored-frame cell and fraction of sample j on axis a, from float64:
// q = sg*(p + j*D), cell = floor(q), frac = q - cell in [0, 1).
TOMO_HD void ray_cell(const RaySetup& r, int a, int j, int ustride, int& off, double& frac)
{
    const double q = (double)r.sg[a] * (r.p[a] + (double)j * r.D[a]);
    const double qi = floor(q);
    frac = q - qi;
    off += (TOMO_PAD + r.sg[a] * (int)qi) * ustride;
}

// frac (64-bit fixed point, hi:lo) += d; returns the carry out of bit 63.
TOMO_HD unsigned fix64_add(unsigned& hi, unsigned& lo, unsigned dhi, unsigned dlo)
{
#if defined(__CUDA_ARCH__)
    unsigned c;
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, 0, 0;"
        : "+r"(lo), "+r"(hi), "=r"(c) : "r"(dlo), "r"(dhi));
    return c;
#else
    const unsigned long long f = ((unsigned long long)hi << 32) | lo, d = ((unsigned long long)dhi << 32) | dlo;
    const unsigned long long s = f + d;
    hi = (unsigned)(s >> 32); lo = (unsigned)s;
    return s < f ? 1u : 0u;
#endif
}

// top 23 bits of a 0.32 fixed-point fraction as a float in [0, 1)
TOMO_HD float fix_to_float(unsigned hi)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__funnelshift_r(hi, 0x7Fu, 9)) - 1.0f;
#else
    union { unsigned u; float f; } c; c.u = 0x3f800000u | (hi >> 9); return c.f - 1.0f;
#endif
}

// Packed float32x2 arithmetic: Blackwell issues one FFMA2 / FADD2 for two float32 lanes of a register pair
// (sm_100 intrinsics __ffma2_rn / __fadd2_rn); the kernels are issue-bound, so pairing the two x-corners of
// every (y, z) corner halves the interpolation instructions.  Host build (tests/emu): same IEEE operations.
struct f2 { float x, y; };
TOMO_HD f2 f2_make(float a, float b) { f2 r; r.x = a; r.y = b; return r; }
TOMO_HD f2 f2_fma(f2 a, f2 b, f2 c)
{
#if defined(__CUDA_ARCH__)
    const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
    return f2_make(r.x, r.y);
#else
    return f2_make(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
// b - a
TOMO_HD f2 f2_sub(f2 b, f2 a)
{
#if defined(__CUDA_ARCH__)
    const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(-1.f, -1.f), make_float2(b.x, b.y));
    return f2_make(r.x, r.y);
#else
    return f2_make(b.x - a.x, b.y - a.y);
#endif
}

// The 8 zero-padded corner loads and the trilinear interpolant in nested-lerp form, the two x-corners of each
// (y, z) corner packed in one register pair (.x = m-floor x, .y = m-ceil x).  Leaves behind the pieces the
// spatial gradient is built from: dzA/dzB (z differences at y-floor / y-ceil), dy (y differences), gx, val.
#if defined(TOMO_PROBE_LDG64) && defined(__CUDA_ARCH__)
// TIMING PROBE ONLY (wrong values for odd offsets): the z pair of every (x, y) corner as one aligned 64-bit load,
// lerps ordered x, y, z so that the packed pairs are the z pairs.
#define RAY_SAMPLE(c, fx_, fy2_, fz2_)                                                                \
    const float2* cq = (const float2*)((unsigned long long)(c) & ~7ull);                               \
    const float2 q00 = __ldg(cq), q10 = __ldg((const float2*)((const float*)cq + o10));                 \
    const float2 q01 = __ldg((const float2*)((const float*)cq + o01)), q11 = __ldg((const float2*)((const float*)cq + o11)); \
    const f2 fx2_ = f2_make(fx_, fx_);                                                                  \
    const f2 v00 = f2_make(q00.x, q00.y), v10 = f2_make(q10.x, q10.y), v01 = f2_make(q01.x, q01.y), v11 = f2_make(q11.x, q11.y); \
    const f2 dzA = f2_sub(v10, v00), dzB = f2_sub(v11, v01);                                            \
    const f2 aA = f2_fma(fx2_, dzA, v00), aB = f2_fma(fx2_, dzB, v01);                                  \
    const f2 dy = f2_sub(aB, aA);                                                                       \
    const f2 b = f2_fma(fy2_, dy, aA);                                                                  \
    const float gx = b.y - b.x;                                                                         \
    const float val = fmaf((fz2_).x, gx, b.x);
#else
#define RAY_SAMPLE(c, fx_, fy2_, fz2_)                                                                \
    const f2 loA = f2_make(TOMO_LDG(c),            TOMO_LDG(c + o10));                                  \
    const f2 hiA = f2_make(TOMO_LDG(c + oz),       TOMO_LDG(c + o10 + oz));                             \
    const f2 loB = f2_make(TOMO_LDG(c + o01),      TOMO_LDG(c + o11));                                  \
    const f2 hiB = f2_make(TOMO_LDG(c + o01 + oz), TOMO_LDG(c + o11 + oz));                             \
    const f2 dzA = f2_sub(hiA, loA), dzB = f2_sub(hiB, loB);                                            \
    const f2 aA = f2_fma(fz2_, dzA, loA), aB = f2_fma(fz2_, dzB, loB);                                  \
    const f2 dy = f2_sub(aB, aA);                                                                       \
    const f2 b = f2_fma(fy2_, dy, aA);                                                                  \
    const float gx = b.y - b.x;                                                                         \
    const float val = fmaf(fx_, gx, b.x);
#endif

// Forward only: (cell offset, float32 fraction) marching, re-based from float64 every RAY_REBASE
// samples.  The interpolant is continuous, so a cell decision that is off by float32 rounding next
// to a lattice plane changes nothing.
// SXP / SYP != 0: the padded strides (and with SGX / SGY the signed corner offsets) are compile-time constants, so all eight
// corner loads address off ONE 64-bit register with immediate offsets (otherwise every corner costs a 64-bit add: 6 of the
// ~42 instructions of a sample).  SXP = SYP = 0: strides and signs at run time (SGX, SGY unused).
template <int SGX, int SGY, int SGZ, int SXP, int SYP>
TOMO_HD void ray_march_forward(const float* __restrict__ vol, const double* __restrict__ V, const RayDims dm,
                               int ix, int iz, RaySums& out)
{
    RaySetup r;
    ray_setup(V, dm, ix, iz, r);
    const int ust[3] = {dm.sxp, dm.syp, 1};
    float df[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { const double ad = fabs(r.D[a]); df[a] = (float)(ad - floor(ad)); }
    // the z stride (+-1) is a template constant: the z-ceil corner of each pair is an immediate offset
    const int o10 = SXP ? SGX * SXP : r.st[0], o01 = SYP ? SGY * SYP : r.st[1], o11 = o10 + o01;
    constexpr int oz = SGZ;
    float acc = 0.f;
    for (int jc = r.j0; jc < r.j1; jc += RAY_REBASE) {
        float f[3];
        int off = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double fr;
            ray_cell(r, a, jc, ust[a], off, fr);
            f[a] = (float)fr;
            if (f[a] >= 1.0f) { f[a] -= 1.0f; off += (a == 0) ? o10 : (a == 1) ? o01 : oz; }
        }
        const int jend = (jc + RAY_REBASE < r.j1) ? jc + RAY_REBASE : r.j1;
#pragma unroll 2
        for (int j = jc; j < jend; ++j) {
            const float* __restrict__ c = vol + off;
            RAY_SAMPLE(c, f[0], f2_make(f[1], f[1]), f2_make(f[2], f[2]))
            acc += val;
            f[0] += df[0]; f[1] += df[1]; f[2] += df[2];
            off += r.stepoff;
            if (f[0] >= 1.0f) { f[0] -= 1.0f; off += o10; }
            if (f[1] >= 1.0f) { f[1] -= 1.0f; off += o01; }
            if (f[2] >= 1.0f) { f[2] -= 1.0f; off += oz; }
        }
    }
    out.acc = acc;
}

// Projection + gradient moments.  The spatial gradient of the trilinear interpolant jumps across
// lattice planes (one-sided differences, src/ray_wt_grad.f90:142-220), so the cell of every sample
// must be the float64 one: the fraction is carried as 64-bit fixed point, which accumulates j*D
// exactly (no re-basing, no branches); only the interpolation weights are rounded to float32.
template <int SGX, int SGY, int SGZ, int SXP, int SYP>
TOMO_HD void ray_march_gradient(const float* __restrict__ vol, const double* __restrict__ V, const RayDims dm,
                                int ix, int iz, RaySums& out)
{
    RaySetup r;
    ray_setup(V, dm, ix, iz, r);
    const int ust[3] = {dm.sxp, dm.syp, 1};
    const int o10 = SXP ? SGX * SXP : r.st[0], o01 = SYP ? SGY * SYP : r.st[1], o11 = o10 + o01;
    constexpr int oz = SGZ;
    unsigned fh[3], fl[3], dh[3], dl[3];
    int off = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double fr;
        ray_cell(r, a, r.j0, ust[a], off, fr);
        const unsigned long long f64 = (unsigned long long)(fr * 18446744073709551616.0);   // fr <= 1 - 2^-53
        const double ad = fabs(r.D[a]);
        const unsigned long long d64 = (unsigned long long)((ad - floor(ad)) * 18446744073709551616.0);
        fh[a] = (unsigned)(f64 >> 32); fl[a] = (unsigned)f64;
        dh[a] = (unsigned)(d64 >> 32); dl[a] = (unsigned)d64;
    }
    float acc = 0.f, s0z = 0.f, s1z = 0.f;
    f2 s0xy = f2_make(0.f, 0.f), s1xy = f2_make(0.f, 0.f);
    // the sample index is needed as a float only (moment weights): it is also the loop counter (exact below 2^24)
    const float fjend = (float)r.j1;
    constexpr int kUnroll = RAY_GRAD_UNROLL;
#pragma unroll kUnroll
    for (float fj = (float)r.j0; fj < fjend; fj += 1.0f) {
        const float fx = fix_to_float(fh[0]), fy = fix_to_float(fh[1]), fz = fix_to_float(fh[2]);
        const float* __restrict__ c = vol + off;
        const f2 fy2 = f2_make(fy, fy);
        RAY_SAMPLE(c, fx, fy2, f2_make(fz, fz))
        acc += val;
        const float gy = fmaf(fx, dy.y - dy.x, dy.x);
        const f2 e = f2_fma(fy2, f2_sub(dzB, dzA), dzA);
        const float gz = fmaf(fx, e.y - e.x, e.x);
        // moments: (gx, gy) packed, gz scalar
        const f2 gxy = f2_make(gx, gy);
        s0xy = f2_fma(gxy, f2_make(1.f, 1.f), s0xy); s0z += gz;
        s1xy = f2_fma(f2_make(fj, fj), gxy, s1xy); s1z = fmaf(fj, gz, s1z);
        off += r.stepoff;
        if (fix64_add(fh[0], fl[0], dh[0], dl[0])) off += o10;
        if (fix64_add(fh[1], fl[1], dh[1], dl[1])) off += o01;
        if (fix64_add(fh[2], fl[2], dh[2], dl[2])) off += oz;
    }
    out.acc = acc;
    out.s0[0] = s0xy.x * (float)r.sg[0]; out.s0[1] = s0xy.y * (float)r.sg[1]; out.s0[2] = s0z * (float)r.sg[2];
    out.s1[0] = s1xy.x * (float)r.sg[0]; out.s1[1] = s1xy.y * (float)r.sg[1]; out.s1[2] = s1z * (float)r.sg[2];
}

template <bool GRAD, int SGX, int SGY, int SGZ, int SXP, int SYP>
TOMO_HD void ray_march_one(const float* __restrict__ vol, const double* __restrict__ V, const RayDims dm,
                           int ix, int iz, RaySums& out)
{
    if (GRAD) ray_march_gradient<SGX, SGY, SGZ, SXP, SYP>(vol, V, dm, ix, iz, out);
    else      ray_march_forward<SGX, SGY, SGZ, SXP, SYP>(vol, V, dm, ix, iz, out);
}

// Signs of the step D: uniform per view, so these branches never diverge.
template <bool GRAD, int SXP, int SYP>
TOMO_HD void ray_march_fixed(const float* __restrict__ vol, const double* __restrict__ V, const RayDims dm,
                             int ix, int iz, RaySums& out)
{
    const int sg = (V[V_D + 0] < 0.0 ? 4 : 0) | (V[V_D + 1] < 0.0 ? 2 : 0) | (V[V_D + 2] < 0.0 ? 1 : 0);
    switch (sg) {
        case 0:  ray_march_one<GRAD,  1,  1,  1, SXP, SYP>(vol, V, dm, ix, iz, out); break;
        case 1:  ray_march_one<GRAD,  1,  1, -1, SXP, SYP>(vol, V, dm, ix, iz, out); break;
        case 2:  ray_march_one<GRAD,  1, -1,  1, SXP, SYP>(vol, V, dm, ix, iz, out); break;
        case 3:  ray_march_one<GRAD,  1, -1, -1, SXP, SYP>(vol, V, dm, ix, iz, out); break;
        case 4:  ray_march_one<GRAD, -1,  1,  1, SXP, SYP>(vol, V, dm, ix, iz, out); break;
        case 5:  ray_march_one<GRAD, -1,  1, -1, SXP, SYP>(vol, V, dm, ix, iz, out); break;
        case 6:  ray_march_one<GRAD, -1, -1,  1, SXP, SYP>(vol, V, dm, ix, iz, out); break;
        default: ray_march_one<GRAD, -1, -1, -1, SXP, SYP>(vol, V, dm, ix, iz, out); break;
    }
}

// Padded strides of an N^3 volume: syp = tomo_nzp(N), sxp = (N + 2 TOMO_PAD) * syp.  The cubes of BASELINE.json's configurations
// (and 128^3) get the compile-time-stride variants; every other shape takes the run-time-stride march.  Same arithmetic, same bits.
#define RAY_CUBE_SYP(N) ((((N) + 2 * TOMO_PAD + 31) / 32) * 32)
#define RAY_CUBE_SXP(N) (((N) + 2 * TOMO_PAD) * RAY_CUBE_SYP(N))
template <bool GRAD>
TOMO_HD void ray_march(const float* __restrict__ vol, const double* __restrict__ V, const RayDims dm,
                       int ix, int iz, RaySums& out)
{
#ifndef RAY_NO_FIXED_STRIDES
#define RAY_TRY_CUBE(N)                                                                                  \
    if (dm.syp == RAY_CUBE_SYP(N) && dm.sxp == RAY_CUBE_SXP(N)) {                                        \
        ray_march_fixed<GRAD, RAY_CUBE_SXP(N), RAY_CUBE_SYP(N)>(vol, V, dm, ix, iz, out);                \
        return;                                                                                          \
    }
    RAY_TRY_CUBE(512) RAY_TRY_CUBE(256) RAY_TRY_CUBE(1024) RAY_TRY_CUBE(128) RAY_TRY_CUBE(64)
#undef RAY_TRY_CUBE
#endif
    if (V[V_D + 2] < 0.0) ray_march_one<GRAD, 1, 1, -1, 0, 0>(vol, V, dm, ix, iz, out);
    else                  ray_march_one<GRAD, 1, 1,  1, 0, 0>(vol, V, dm, ix, iz, out);
}

// d proj / d theta_k for one ray, API order [tx, ty, tz, phi, alpha, beta]
// (utilities/ray_voxel_utilities.py:37-48; src/ray_wt_grad.f90:136-141):
//   d p_j / d t_k     = M[:,k]
//   d p_j / d angle_k = E_k + ix F_k + iz H_k + j K_k
TOMO_HD void ray_gradient(const double* __restrict__ V, int ix, int iz, const RaySums& s, float dp[6])
{
    const double S0[3] = {s.s0[0], s.s0[1], s.s0[2]}, S1[3] = {s.s1[0], s.s1[1], s.s1[2]};
#pragma unroll
    for (int k = 0; k < 3; ++k)
        dp[k] = (float)(S0[0] * V[V_M + 3 * k] + S0[1] * V[V_M + 3 * k + 1] + S0[2] * V[V_M + 3 * k + 2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double d = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double g0 = V[V_E + 3 * k + a] + (double)ix * V[V_F + 3 * k + a] + (double)iz * V[V_H + 3 * k + a];
            d += S0[a] * g0 + S1[a] * V[V_K + 3 * k + a];
        }
        dp[3 + k] = (float)d;
    }
}
