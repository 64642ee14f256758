// zq_core.h -- "z-quad" core of the ray-driven forward projector / gradient: one thread owns FOUR z-adjacent rays.
//
// Same sums as ray_core.h (src/ray_wt_grad.f90:20-91 forward, :121-222 gradient; sample positions
// p_j = p0 + j*step*r_hat of utilities/ray_voxel_utilities.py:89-94), reorganised for views whose detector rows map almost
// one-to-one onto volume z planes (W = d p / d iz ~ (0, 0, 1): every pose of examples/generate_data.py):
//   * rays iz0 .. iz0+3 of one detector column sit in the SAME (x, y) cell and in four CONSECUTIVE z cells for ~98 % of
//     their samples.  The thread marches ray iz0 exactly (64-bit fixed-point fractions, no mirroring: a negative step is
//     "floor(D) + fraction"), loads the 8 z planes [4*floor(zc/4), +8) of the cell's four (x, y) corner columns with
//     eight 128-bit loads, and interpolates all four rays from them: 2 loads per sample instead of 8, one march per four
//     samples.  The position of the window inside the aligned 8 planes (s = zc mod 4) selects one of four copies of the
//     interpolation code whose register indices are compile-time constants.
//   * the fractions of rays 1..3 are the base ray's plus k*W (float32).  A ray whose fraction leaves [0, 1) on some axis --
//     it sits in a neighbouring cell, an "irregular" sample -- is NOT interpolated from the window: the sample is recorded
//     (16 bits: j and a 3-bit ray mask) in a per-thread list in shared memory and evaluated later, exactly, from float64
//     (zq_exact_sample, the arithmetic of ray_core.h's ray_cell).  For the gradient the test has a margin of 1e-5, so the
//     cell every regular sample is differentiated in is the float64 one (the gradient jumps across lattice planes).
//     Lists are drained when one runs full (warp-collective decision) and at the end; every ray's samples are added in a
//     fixed order, so results are bitwise reproducible.
//   * the clip range marched is the union of the four rays' ranges (ray_setup); a ray outside its own range is masked (its
//     samples have no in-bounds corner).  The window may then lie up to 3 planes beside the zero border: it stays inside the
//     padded buffer because tomo_padded_volume_bytes() ends with TOMO_PAD_TAIL floats of slack.
// __host__ __device__: tests/emu runs this file on the CPU against the oracle.
#pragma once
#include "ray_core.h"

#ifndef ZQ_CAP
#define ZQ_CAP 32                    // events per thread list
#endif
#define ZQ_G 4                       // rays per thread
#define ZQ_MAX_SAMPLES 8191          // j must fit 13 bits of an event

struct ZqView {                      // per-view constants in registers
    float wx, wy, wz1;               // W_x, W_y, W_z - 1
};

struct ZqSums {                      // per-ray accumulators (s0, s1 used by the gradient only)
    float acc[ZQ_G];
    float s0[ZQ_G][3], s1[ZQ_G][3];
};

#if defined(__CUDA_ARCH__)
#define ZQ_LD4(p) __ldg(reinterpret_cast<const float4*>(p))
struct zq_f4 { float4 v; TOMO_HD float get(int i) const { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; } };
TOMO_HD zq_f4 zq_ld4(const float* p) { zq_f4 r; r.v = ZQ_LD4(p); return r; }
#else
struct zq_f4 { float v[4]; TOMO_HD float get(int i) const { return v[i]; } };
TOMO_HD zq_f4 zq_ld4(const float* p) { zq_f4 r; r.v[0] = p[0]; r.v[1] = p[1]; r.v[2] = p[2]; r.v[3] = p[3]; return r; }
#endif

// plane I (compile-time) of an aligned 8-plane window held as two float4
template <int I> TOMO_HD float zq_plane(const zq_f4& lo, const zq_f4& hi)
{
    static_assert(I >= 0 && I < 8, "window index");
    return I < 4 ? lo.get(I & 3) : hi.get(I & 3);
}

// trilinear value (and spatial gradient) from the 8 corner values of one ray: c[column][z], columns 00, 10, 01, 11 (x, y)
template <bool GRAD>
TOMO_HD void zq_interp(const float c00z0, const float c00z1, const float c10z0, const float c10z1,
                       const float c01z0, const float c01z1, const float c11z0, const float c11z1,
                       float fx, float fy, float fz, float& val, float& gx, float& gy, float& gz)
{
    const float d00 = c00z1 - c00z0, d10 = c10z1 - c10z0, d01 = c01z1 - c01z0, d11 = c11z1 - c11z0;
    const float a00 = fmaf(fz, d00, c00z0), a10 = fmaf(fz, d10, c10z0), a01 = fmaf(fz, d01, c01z0), a11 = fmaf(fz, d11, c11z0);
    const float dx0 = a10 - a00, dx1 = a11 - a01;
    const float b0 = fmaf(fx, dx0, a00), b1 = fmaf(fx, dx1, a01);
    const float dy = b1 - b0;
    val = fmaf(fy, dy, b0);
    if (GRAD) {
        gx = fmaf(fy, dx1 - dx0, dx0);
        gy = dy;
        const float e0 = fmaf(fx, d10 - d00, d00), e1 = fmaf(fx, d11 - d01, d01);
        gz = fmaf(fy, e1 - e0, e0);
    }
}

// One sample of ray (ix, iz) at index j evaluated from float64 (no marching state): the arithmetic of ray_cell / RAY_SAMPLE
// without mirroring.  j must lie inside the ray's clip range (ray_setup), so all 8 corners are inside the padded volume.
template <bool GRAD>
TOMO_HD void zq_exact_sample(const float* __restrict__ vol, const double* __restrict__ V, const RayDims dm,
                             int ix, int iz, int j, float& acc, float s0[3], float s1[3])
{
    const int ust[3] = {dm.sxp, dm.syp, 1};
    float f[3];
    int off = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double p = V[V_P00 + a] + (double)ix * V[V_U + a] + (double)iz * V[V_W + a];      // as ray_setup
        const double q = p + (double)j * V[V_D + a];
        const double qi = floor(q);
        // the cell is the float64 one (the gradient jumps across lattice planes): a fraction of 1 - 2e-8 must not round up
        // into the next cell, it is clamped to the largest float32 below 1 instead
        f[a] = fminf((float)(q - qi), 0.99999994f);
        off += (TOMO_PAD + (int)qi) * ust[a];
    }
    const float* __restrict__ c = vol + off;
    const int o10 = dm.sxp, o01 = dm.syp, o11 = dm.sxp + dm.syp;
    float val, gx = 0.f, gy = 0.f, gz = 0.f;
    zq_interp<GRAD>(TOMO_LDG(c), TOMO_LDG(c + 1), TOMO_LDG(c + o10), TOMO_LDG(c + o10 + 1),
                    TOMO_LDG(c + o01), TOMO_LDG(c + o01 + 1), TOMO_LDG(c + o11), TOMO_LDG(c + o11 + 1),
                    f[0], f[1], f[2], val, gx, gy, gz);
    acc += val;
    if (GRAD) {
        const float fj = (float)j;
        s0[0] += gx; s0[1] += gy; s0[2] += gz;
        s1[0] = fmaf(fj, gx, s1[0]); s1[1] = fmaf(fj, gy, s1[1]); s1[2] = fmaf(fj, gz, s1[2]);
    }
}

// true iff 0 <= f < 1 (forward) / MARGIN <= f < 1 - MARGIN (gradient): one unsigned compare on the float's bits
template <bool GRAD> TOMO_HD bool zq_frac_ok(float f)
{
#if defined(__CUDA_ARCH__)
    const unsigned b = __float_as_uint(f);
#else
    union { float f; unsigned u; } cv; cv.f = f; const unsigned b = cv.u;
#endif
    // positive floats order like their bit patterns; negative ones (sign bit) and NaN compare high
    constexpr unsigned LO = GRAD ? 0x3727c5acu /* 1e-5f */ : 0u;
    constexpr unsigned HI = GRAD ? 0x3f7fff58u /* 1 - 1e-5f */ : 0x3f800000u;
    return (b - LO) < (HI - LO);
}

// The four rays of a thread at one sample, window position S = (z cell of the base ray) mod 4.
// w[column][half]: aligned 8-plane windows of the columns (x,y), (x+1,y), (x,y+1), (x+1,y+1).
// allow: bit k set iff ray k is inside its own clip range at this sample (others contribute nothing).
// bad (out): bit k-1 set iff ray k is allowed but irregular (another cell): the caller records an event.
template <bool GRAD, int S>
TOMO_HD void zq_group(const zq_f4 (&w)[4][2], float fx0, float fy0, float fz0, const ZqView& zv, float fj,
                      unsigned allow, unsigned& bad, ZqSums& s)
{
    bad = 0u;
#define ZQ_RAY(K)                                                                                                   \
    {                                                                                                               \
        const float fx = (K) ? fmaf((float)(K), zv.wx, fx0) : fx0, fy = (K) ? fmaf((float)(K), zv.wy, fy0) : fy0,     \
                    fz = (K) ? fmaf((float)(K), zv.wz1, fz0) : fz0;                                                 \
        const bool in = (allow >> (K)) & 1u;                                                                        \
        const bool reg = (K) == 0 || (zq_frac_ok<GRAD>(fx) && zq_frac_ok<GRAD>(fy) && zq_frac_ok<GRAD>(fz));          \
        float val, gx = 0.f, gy = 0.f, gz = 0.f;                                                                    \
        zq_interp<GRAD>(zq_plane<S + (K)>(w[0][0], w[0][1]), zq_plane<S + (K) + 1>(w[0][0], w[0][1]),               \
                        zq_plane<S + (K)>(w[1][0], w[1][1]), zq_plane<S + (K) + 1>(w[1][0], w[1][1]),               \
                        zq_plane<S + (K)>(w[2][0], w[2][1]), zq_plane<S + (K) + 1>(w[2][0], w[2][1]),               \
                        zq_plane<S + (K)>(w[3][0], w[3][1]), zq_plane<S + (K) + 1>(w[3][0], w[3][1]),               \
                        fx, fy, fz, val, gx, gy, gz);                                                               \
        if (in && reg) {                                                                                            \
            s.acc[K] += val;                                                                                        \
            if (GRAD) {                                                                                             \
                s.s0[K][0] += gx; s.s0[K][1] += gy; s.s0[K][2] += gz;                                               \
                s.s1[K][0] = fmaf(fj, gx, s.s1[K][0]); s.s1[K][1] = fmaf(fj, gy, s.s1[K][1]);                       \
                s.s1[K][2] = fmaf(fj, gz, s.s1[K][2]);                                                              \
            }                                                                                                       \
        }                                                                                                           \
        if ((K) != 0 && in && !reg) bad |= 1u << ((K) ? (K) - 1 : 0);                                                        \
    }
    ZQ_RAY(0) ZQ_RAY(1) ZQ_RAY(2) ZQ_RAY(3)
#undef ZQ_RAY
}

// Evaluate and clear the thread's event list.  An event is (j << 3) | m with bit k-1 of m set for ray k (k = 1..3; the base
// ray is always regular).  `live` masks rays that exist (iz < ndz).  On the device the loop trip count is the warp's maximum
// so that the lanes stay converged (no warp-collective operation is needed inside).
template <bool GRAD>
TOMO_HD void zq_drain(const float* __restrict__ vol, const double* __restrict__ V, const RayDims dm, int ix, int iz0,
                      const unsigned short* ev, int ev_stride, int& cnt, ZqSums& s)
{
    for (int i = 0; i < cnt; ++i) {
        const unsigned e = ev[(size_t)i * ev_stride];
        const int j = (int)(e >> 3);
#pragma unroll
        for (int k = 1; k < ZQ_G; ++k)
            if (e & (1u << (k - 1)))
                zq_exact_sample<GRAD>(vol, V, dm, ix, iz0 + k, j, s.acc[k], s.s0[k], s.s1[k]);
    }
    cnt = 0;
}

// Keep a value in a register: without this ptxas, short of registers, re-derives per-view constants (float64 -> fixed
// point conversions!) inside the sample loop instead of keeping them.
#if defined(__CUDA_ARCH__)
#define ZQ_PIN_U(x) asm volatile("" : "+r"(x))
#define ZQ_PIN_F(x) asm volatile("" : "+f"(x))
#else
#define ZQ_PIN_U(x)
#define ZQ_PIN_F(x)
#endif

// Whole march of the rays (ix, iz0 .. iz0 + nrays - 1).
//   ev : this thread's event list (ZQ_CAP entries) followed by 2 * ZQ_G entries for the rays' clip ranges, stride ev_stride
// `vol` must be 16-byte aligned (the padded volume is).
template <bool GRAD>
TOMO_HD void zq_march(const float* __restrict__ vol, const double* __restrict__ V, const RayDims dm,
                      int ix, int iz0, int nrays, unsigned short* ev, int ev_stride, ZqSums& s)
{
#if defined(__CUDA_ARCH__)
    constexpr unsigned FULL = 0xffffffffu;
#endif
    const int ust[3] = {dm.sxp, dm.syp, 1};
    unsigned short* rng = ev + (size_t)ZQ_CAP * ev_stride;           // j0[k], j1[k] of the four rays (13-bit sample indices)
    // clip ranges: per ray, their union [ju0, ju1) (what is marched) and the intersection [ji0, ji1) of the rays that touch the
    // volume at all ("live": the others contribute nothing anywhere and are masked)
    int ju0 = 0x7fffffff, ju1 = -0x7fffffff, ji0 = -0x7fffffff, ji1 = 0x7fffffff;
    unsigned live = 0u;
#pragma unroll
    for (int k = 0; k < ZQ_G; ++k) {
        int j0 = 0, j1 = 0;
        if (k < nrays) {
            RaySetup r;
            ray_setup(V, dm, ix, iz0 + k, r);
            j0 = r.j0; j1 = r.j1;
        }
        if (j1 > j0) {
            live |= 1u << k;
            ju0 = j0 < ju0 ? j0 : ju0; ju1 = j1 > ju1 ? j1 : ju1;
            ji0 = j0 > ji0 ? j0 : ji0; ji1 = j1 < ji1 ? j1 : ji1;
        } else {
            j0 = j1 = 0;
        }
        rng[(size_t)(2 * k) * ev_stride] = (unsigned short)j0;
        rng[(size_t)(2 * k + 1) * ev_stride] = (unsigned short)j1;
    }
    if (ju1 <= ju0) { ju0 = ju1 = 0; ji0 = ji1 = 0; }
#pragma unroll
    for (int k = 0; k < ZQ_G; ++k) {
        s.acc[k] = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) { s.s0[k][a] = 0.f; s.s1[k][a] = 0.f; }
    }

    // base ray (iz0) at sample ju0: un-mirrored cell offset + 64-bit fixed-point fractions; step = floor(D) + fraction
    unsigned fh[3], fl[3], dh[3], dl[3];
    int off = 0, stepoff = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double p = V[V_P00 + a] + (double)ix * V[V_U + a] + (double)iz0 * V[V_W + a];
        const double q = p + (double)ju0 * V[V_D + a];
        double qi = floor(q), fr = q - qi;
        if (fr >= 1.0) { fr = 0.0; qi += 1.0; }              // q = -1e-17: q - floor(q) rounds up to 1
        const unsigned long long f64 = (unsigned long long)(fr * 18446744073709551616.0);
        double di = floor(V[V_D + a]), df = V[V_D + a] - di;
        if (df >= 1.0) { df = 0.0; di += 1.0; }              // D = -5e-17 (a view at phi = pi): step 0, not -1 + 1
        const unsigned long long d64 = (unsigned long long)(df * 18446744073709551616.0);
        fh[a] = (unsigned)(f64 >> 32); fl[a] = (unsigned)f64;
        dh[a] = (unsigned)(d64 >> 32); dl[a] = (unsigned)d64;
        ZQ_PIN_U(dh[a]); ZQ_PIN_U(dl[a]);
        off += (TOMO_PAD + (int)fmin(fmax(qi, -1.0e6), 1.0e6)) * ust[a];
        stepoff += (int)di * ust[a];
    }
    ZQ_PIN_U(stepoff);
    ZqView zv;
    zv.wx = (float)V[V_W + 0]; zv.wy = (float)V[V_W + 1]; zv.wz1 = (float)(V[V_W + 2] - 1.0);
    ZQ_PIN_F(zv.wx); ZQ_PIN_F(zv.wy); ZQ_PIN_F(zv.wz1);
    const int o10 = dm.sxp, o01 = dm.syp, o11 = dm.sxp + dm.syp;
    int cnt = 0;

#if defined(__CUDA_ARCH__)
    const int jw0 = __reduce_min_sync(FULL, ju0 < ju1 ? ju0 : 0x7fffffff), jw1 = __reduce_max_sync(FULL, ju0 < ju1 ? ju1 : -0x7fffffff);
#else
    const int jw0 = ju0, jw1 = ju1;
#endif
    const unsigned ulen = (unsigned)(ju1 - ju0), ilen = (ji1 > ji0) ? (unsigned)(ji1 - ji0) : 0u;
    for (int j = jw0; j < jw1; ++j) {
        if ((unsigned)(j - ju0) < ulen) {
            unsigned allow = live;
            if (!((unsigned)(j - ji0) < ilen)) {           // near the ends of the march: per-ray range tests
                allow = 0u;
#pragma unroll
                for (int k = 0; k < ZQ_G; ++k)
                    allow |= (j >= (int)rng[(size_t)(2 * k) * ev_stride] && j < (int)rng[(size_t)(2 * k + 1) * ev_stride]) ? (1u << k) : 0u;
            }
            const float fx0 = fix_to_float(fh[0]), fy0 = fix_to_float(fh[1]), fz0 = fix_to_float(fh[2]);
            const float* __restrict__ c = vol + (off & ~3);
            zq_f4 w[4][2];
            w[0][0] = zq_ld4(c);        w[0][1] = zq_ld4(c + 4);
            w[1][0] = zq_ld4(c + o10);  w[1][1] = zq_ld4(c + o10 + 4);
            w[2][0] = zq_ld4(c + o01);  w[2][1] = zq_ld4(c + o01 + 4);
            w[3][0] = zq_ld4(c + o11);  w[3][1] = zq_ld4(c + o11 + 4);
            const float fj = (float)j;
            unsigned bad;
            switch (off & 3) {
                case 0:  zq_group<GRAD, 0>(w, fx0, fy0, fz0, zv, fj, allow, bad, s); break;
                case 1:  zq_group<GRAD, 1>(w, fx0, fy0, fz0, zv, fj, allow, bad, s); break;
                case 2:  zq_group<GRAD, 2>(w, fx0, fy0, fz0, zv, fj, allow, bad, s); break;
                default: zq_group<GRAD, 3>(w, fx0, fy0, fz0, zv, fj, allow, bad, s); break;
            }
            if (bad) { ev[(size_t)cnt * ev_stride] = (unsigned short)(((unsigned)j << 3) | bad); ++cnt; }
            // advance the base ray one sample: exact fixed-point fractions, carries move the cell
            off += stepoff;
            off += (int)fix64_add(fh[0], fl[0], dh[0], dl[0]) * dm.sxp;
            off += (int)fix64_add(fh[1], fl[1], dh[1], dl[1]) * dm.syp;
            off += (int)fix64_add(fh[2], fl[2], dh[2], dl[2]);
        }
#if defined(__CUDA_ARCH__)
        if (__any_sync(FULL, cnt >= ZQ_CAP)) zq_drain<GRAD>(vol, V, dm, ix, iz0, ev, ev_stride, cnt, s);
#else
        if (cnt >= ZQ_CAP) zq_drain<GRAD>(vol, V, dm, ix, iz0, ev, ev_stride, cnt, s);
#endif
    }
    zq_drain<GRAD>(vol, V, dm, ix, iz0, ev, ev_stride, cnt, s);
#pragma unroll
    for (int k = 0; k < ZQ_G; ++k)
        if (!((live >> k) & 1u)) {                          // rays that never touch the volume project to zero
            s.acc[k] = 0.f;
#pragma unroll
            for (int a = 0; a < 3; ++a) { s.s0[k][a] = 0.f; s.s1[k][a] = 0.f; }
        }
}
