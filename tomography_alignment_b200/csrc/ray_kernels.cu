// ray_kernels.cu -- ray-driven forward projector and projection + 6-DOF gradient, sm_100a.
//
// One thread owns one ray (view, ix, iz) and marches the samples p_j = p0 + j*D of
// utilities/ray_voxel_utilities.py:89-94, evaluating the zero-padded trilinear interpolant of
// src/ray_wt_grad.f90:20-91 (forward) and :121-222 (gradient).  Differences in *how* (not what):
//   * nothing is materialised: p0 comes from the affine view record (tomo_common.h), the ray is
//     clipped to the samples that can touch the volume, and the sample index j keeps the
//     reference's phase so positions are identical;
//   * the volume is read from a copy with a zero border of TOMO_PAD voxels, which realises the
//     per-corner bounds checks of the Fortran without branches;
//   * positions are carried as (integer cell, float32 fraction) with the fraction re-based from
//     float64 every REBASE samples, so the interpolation weights carry ~1e-6 absolute error at any
//     volume size instead of float32-at-coordinate-512 error;
//   * axes along which the ray runs backwards are mirrored so the per-step carry is one-sided;
//   * the gradient needs only S0 = sum_j G_j and S1 = sum_j j*G_j (G = spatial gradient of the
//     interpolant) per ray, because d p_j / d theta is affine in j (ray_wt_grad.f90:136-141).
// Lanes run along iz, the fastest detector axis, which maps to z, the fastest volume axis: corner
// loads of a warp are contiguous for small tilts.  Blocks are ordered z-tile-major so that the
// z-slab of the volume a tile row needs stays L2-resident while all views sweep over it.
#include <cuda_runtime.h>
#include "tomo_common.h"
#include "ray_core.h"
#include "zq_core.h"

namespace {

constexpr int TILE_Z = 32;     // lanes: iz
#ifndef RAY_TILE_X
#define RAY_TILE_X 8
#endif
constexpr int TILE_X = RAY_TILE_X;      // warps: ix
#ifndef RAY_BAND_TILES
#define RAY_BAND_TILES 16
#endif
constexpr int NRED   = 7;      // 6 gradient components + cost

struct RayArgs {
    const float*  volpad;
    const double* views;
    const float*  meas;      // nullable
    float*        proj;      // nullable
    float*        dproj;     // nullable
    double*       partial;   // nullable, [n_blocks][NRED]
    int nx, ny, nz, ndx, ndz, n_proj;
    int sxp, syp;            // padded strides (floats) of x and y; z stride is 1
    int nxt, nzt;            // detector tiles along x and z
    int xparts, xpp;         // launch order: bands of xpp x-tiles (see ray_kernel_body)
    int skip_separable;      // leave views with V_SEP == 1 to the separable kernels
    int skip_zq;             // leave views with V_ZQ == 1 to the z-quad kernels
};

template <bool GRAD>
__device__ __forceinline__ void ray_kernel_body(const RayArgs& A)
{
    // launch order: x-tile within a part fastest, then view, then part, then z-tile.  All views sweep over one z-slab
    // restricted to one band of detector columns before the next band starts, so the working set in L2 is a band of the
    // slab (which rotates slowly with the view angle) instead of the whole slab.
    const int pb   = blockIdx.x;
    const int xl   = pb % A.xpp;
    const int view = (pb / A.xpp) % A.n_proj;
    const int part = (pb / (A.xpp * A.n_proj)) % A.xparts;
    const int zt   = pb / (A.xpp * A.n_proj * A.xparts);
    const int xt   = part * A.xpp + xl;
    if (xt >= A.nxt) return;                                       // block-uniform (last part may be ragged)
    const int bid  = (zt * A.n_proj + view) * A.nxt + xt;          // logical tile id: layout of the block partials
    const int iz = zt * TILE_Z + threadIdx.x;
    const int ix = xt * TILE_X + threadIdx.y;
    const bool active = (ix < A.ndx) && (iz < A.ndz);

    const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
    if (A.skip_separable && V[V_SEP] != 0.0) return;               // untilted view (block-uniform): the separable kernels do it
    if (A.skip_zq && V[V_ZQ] != 0.0) return;                       // nearly untilted view: the z-quad kernels do it
    const size_t n_det = (size_t)A.ndx * A.ndz;
    const size_t ray = (size_t)ix * A.ndz + iz;

    RaySums s;
    s.acc = 0.f;
    if (active) {
        const RayDims dm = {A.nx, A.ny, A.nz, A.sxp, A.syp};
        ray_march<GRAD>(A.volpad, V, dm, ix, iz, s);
        if (A.proj) A.proj[(size_t)view * n_det + ray] = s.acc;
    }

    if (GRAD) {
        double red[NRED] = {0, 0, 0, 0, 0, 0, 0};
        if (active) {
            float dp[6];
            ray_gradient(V, ix, iz, s, dp);
            if (A.dproj) {
#pragma unroll
                for (int k = 0; k < 6; ++k) A.dproj[((size_t)view * 6 + k) * n_det + ray] = dp[k];
            }
            if (A.meas) {
                // residual and s.residual of utilities/alignment_functions.py:27-37,176-186
                const double res = (double)A.meas[(size_t)view * n_det + ray] - (double)s.acc;
#pragma unroll
                for (int k = 0; k < 6; ++k) red[k] = -(double)dp[k] * res;
                red[6] = 0.5 * res * res;
            }
        }
        if (A.partial) {       // fixed-order block reduction: shuffle tree, then warps in order
            __shared__ double sm[TILE_X][NRED];
#pragma unroll
            for (int k = 0; k < NRED; ++k) {
                double v = red[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if (threadIdx.x == 0) sm[threadIdx.y][k] = v;
            }
            __syncthreads();
            if (threadIdx.y == 0 && threadIdx.x < NRED) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < TILE_X; ++w) v += sm[w][threadIdx.x];
                A.partial[(size_t)bid * NRED + threadIdx.x] = v;
            }
        }
    }
}

// Measured on B200 (profiles/README.md): the forward march is fastest at 48 registers (5 blocks/SM) with the
// sample loop unrolled twice (6 blocks / 40 registers: -0.6 %, 4 blocks: +0.4 %).
#ifndef RAY_FWD_MINB
#define RAY_FWD_MINB (40 / TILE_X)
#endif
__global__ void __launch_bounds__(TILE_Z * TILE_X, RAY_FWD_MINB) ray_kernel_forward(const RayArgs A) { ray_kernel_body<false>(A); }
// gradient: 5 blocks per SM (48 registers, a few spills outside the sample loop) beat ptxas' own 52 registers / 4 blocks by 6.6 %
// once the compile-time strides had shortened the loop (long-scoreboard stalls dominate it): 48.4 -> 45.2 ms per 180 views
#ifndef RAY_GRAD_MINB
#define RAY_GRAD_MINB 5
#endif
__global__ void __launch_bounds__(TILE_Z * TILE_X, RAY_GRAD_MINB) ray_kernel_gradient(const RayArgs A) { ray_kernel_body<true>(A); }

// ---- z-quad kernels (zq_core.h): a thread owns four z-adjacent rays -------------------------------------------------
// Block = 8 warps; a warp covers 4 detector columns x 32 rows (8 lanes of 4 rays per column), a block 32 columns x 32 rows.
// Launch order as above (z-tile, band of columns, view, x-tile within the band) with the tile sizes below.
#ifndef ZQ_NTHREADS
#define ZQ_NTHREADS 256
#endif
constexpr int ZQ_THREADS = ZQ_NTHREADS, ZQ_TILE_X = ZQ_THREADS / 8, ZQ_TILE_Z = 32;
#ifndef ZQ_BAND_COLS
#define ZQ_BAND_COLS 128             // detector columns per band of the launch order
#endif
#define ZQ_BAND_TILES (ZQ_BAND_COLS / ZQ_TILE_X)

template <bool GRAD>
__device__ __forceinline__ void zq_kernel_body(const RayArgs& A)
{
    __shared__ unsigned short ev[ZQ_CAP + 2 * ZQ_G][ZQ_THREADS];     // per-thread event lists + clip ranges, interleaved (conflict-free)
    const int pb   = blockIdx.x;
    const int xl   = pb % A.xpp;
    const int view = (pb / A.xpp) % A.n_proj;
    const int part = (pb / (A.xpp * A.n_proj)) % A.xparts;
    const int zt   = pb / (A.xpp * A.n_proj * A.xparts);
    const int xt   = part * A.xpp + xl;
    if (xt >= A.nxt) return;                                       // block-uniform (last part may be ragged)
    const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
    if (V[V_ZQ] == 0.0) return;                                    // block-uniform: another kernel family owns this view
    const int bid  = (zt * A.n_proj + view) * A.nxt + xt;          // logical tile id: layout of the block partials
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ix = xt * ZQ_TILE_X + warp * 4 + (lane >> 3);
    const int iz0 = zt * ZQ_TILE_Z + (lane & 7) * ZQ_G;
    int nrays = 0;
    if (ix < A.ndx) nrays = min(ZQ_G, max(0, A.ndz - iz0));
    const size_t n_det = (size_t)A.ndx * A.ndz;
    const RayDims dm = {A.nx, A.ny, A.nz, A.sxp, A.syp};

    ZqSums s;
    zq_march<GRAD>(A.volpad, V, dm, ix, iz0, nrays, &ev[0][tid], ZQ_THREADS, s);

    const size_t ray0 = (size_t)ix * A.ndz + iz0;
    if (A.proj) {
        float* __restrict__ o = A.proj + (size_t)view * n_det + ray0;
        if (nrays == ZQ_G && ((A.ndz & 3) == 0)) *reinterpret_cast<float4*>(o) = make_float4(s.acc[0], s.acc[1], s.acc[2], s.acc[3]);
        else
#pragma unroll
            for (int k = 0; k < ZQ_G; ++k) if (k < nrays) o[k] = s.acc[k];
    }
    if (GRAD) {
        double red[NRED] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < ZQ_G; ++k) {
            if (k < nrays) {
                RaySums r;
                r.acc = s.acc[k];
#pragma unroll
                for (int a = 0; a < 3; ++a) { r.s0[a] = s.s0[k][a]; r.s1[a] = s.s1[k][a]; }
                float dp[6];
                ray_gradient(V, ix, iz0 + k, r, dp);
                if (A.dproj) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) A.dproj[((size_t)view * 6 + c) * n_det + ray0 + k] = dp[c];
                }
                if (A.meas) {
                    const double res = (double)A.meas[(size_t)view * n_det + ray0 + k] - (double)r.acc;
#pragma unroll
                    for (int c = 0; c < 6; ++c) red[c] += -(double)dp[c] * res;
                    red[6] += 0.5 * res * res;
                }
            }
        }
        if (A.partial) {       // fixed-order block reduction: the thread's four rays, shuffle tree, then warps in order
            __shared__ double sm[ZQ_THREADS / 32][NRED];
#pragma unroll
            for (int c = 0; c < NRED; ++c) {
                double v = red[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if (lane == 0) sm[warp][c] = v;
            }
            __syncthreads();
            if (tid < NRED) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < ZQ_THREADS / 32; ++w) v += sm[w][tid];
                A.partial[(size_t)bid * NRED + tid] = v;
            }
        }
    }
}

#ifndef ZQ_FWD_MINB
#define ZQ_FWD_MINB 2
#endif
#ifndef ZQ_GRAD_MINB
#define ZQ_GRAD_MINB 2
#endif
__global__ void __launch_bounds__(ZQ_THREADS, ZQ_FWD_MINB) zq_kernel_forward(const RayArgs A) { zq_kernel_body<false>(A); }
__global__ void __launch_bounds__(ZQ_THREADS, ZQ_GRAD_MINB) zq_kernel_gradient(const RayArgs A) { zq_kernel_body<true>(A); }

// Second pass of the deterministic reduction: one thread per (view, component) sums the block
// partials of that view in (zt, xt) order.
__global__ void grad_finalize_kernel(const double* __restrict__ partial, const double* __restrict__ views, int kind,
                                     int n_proj, int nxt, int nzt, double* __restrict__ grad6, double* __restrict__ cost)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_proj * NRED) return;
    const int view = t / NRED, k = t % NRED;
    // kind 0: views reduced by ray_kernel_gradient, 1: by sep_gradient_kernel, 2: by zq_kernel_gradient (each with its own tiling)
    const double* __restrict__ V = views + (size_t)view * TOMO_VIEW_STRIDE;
    const int mine = (V[V_SEP] != 0.0) ? 1 : (V[V_ZQ] != 0.0) ? 2 : 0;
    if (mine != kind) return;
    double v = 0.0;
    for (int zt = 0; zt < nzt; ++zt)
        for (int xt = 0; xt < nxt; ++xt)
            v += partial[((size_t)(zt * n_proj + view) * nxt + xt) * NRED + k];
    if (k < 6) { if (grad6) grad6[view * 6 + k] = v; }
    else       { if (cost)  cost[view] = v; }
}

__global__ void pad_volume_kernel(const float* __restrict__ vol, float* __restrict__ pad,
                                  int nx, int ny, int nz, int nyp, int nzp)
{
    // one thread per buffer element (TOMO_PAD_HEAD zeros, the padded volume z fastest, TOMO_PAD_TAIL zeros)
    const size_t body = (size_t)(nx + 2 * TOMO_PAD) * nyp * nzp, total = body + TOMO_PAD_HEAD + TOMO_PAD_TAIL;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += (size_t)gridDim.x * blockDim.x) {
        if (b < TOMO_PAD_HEAD || b >= body + TOMO_PAD_HEAD) { pad[b] = 0.f; continue; }
        const size_t i = b - TOMO_PAD_HEAD;
        const int zp = (int)(i % nzp);
        const size_t r = i / nzp;
        const int yp = (int)(r % nyp), xp = (int)(r / nyp);
        const int x = xp - TOMO_PAD, y = yp - TOMO_PAD, z = zp - TOMO_PAD;
        float v = 0.f;
        if (x >= 0 && x < nx && y >= 0 && y < ny && z >= 0 && z < nz) v = vol[((size_t)x * ny + y) * nz + z];
        pad[b] = v;
    }
}

}  // namespace

extern "C" void tomo_set_error(const char* msg);
int tomo_check_cuda(cudaError_t e, const char* what);
int tomo_forward_separable_launch(const TomoGeom* g, const void* views, int n_proj, const float* volpad, float* proj,
                                  void* stream);
void tomo_grad_separable_tiles(const TomoGeom* g, int* nxt, int* nchunk);
int tomo_grad_separable_launch(const TomoGeom* g, const void* views, int n_proj, const float* volpad, const float* meas,
                               float* proj, float* dproj, double* partial, void* stream);

static int check_sizes(const TomoGeom* g)
{
    int nyp, syp;
    tomo_pad_pitch(g->ny, g->nz, &nyp, &syp);
    const double padded = (double)(g->nx + 2 * TOMO_PAD) * nyp * syp;
    if (padded >= 2147483647.0) { tomo_set_error("padded volume exceeds 2^31 elements (32-bit kernel offsets)"); return TOMO_E_RANGE; }
    return 0;
}

extern "C" size_t tomo_padded_volume_bytes(const TomoGeom* g)
{
    if (!g) return 0;
    int nyp, syp;
    tomo_pad_pitch(g->ny, g->nz, &nyp, &syp);
    return sizeof(float) * ((size_t)(g->nx + 2 * TOMO_PAD) * nyp * syp + TOMO_PAD_HEAD + TOMO_PAD_TAIL);
}

extern "C" int tomo_pad_volume(const TomoGeom* g, const float* vol, float* pad, void* stream)
{
    if (!g || !vol || !pad) { tomo_set_error("tomo_pad_volume: null pointer"); return TOMO_E_ARG; }
    if (int e = check_sizes(g)) return e;
    const size_t total = tomo_padded_volume_bytes(g) / sizeof(float);
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads > 148 * 64 ? 148 * 64 : (total + threads - 1) / threads);
    int nyp, syp;
    tomo_pad_pitch(g->ny, g->nz, &nyp, &syp);
    pad_volume_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(vol, pad, g->nx, g->ny, g->nz, nyp, syp);
    return tomo_check_cuda(cudaGetLastError(), "pad_volume_kernel");
}

static int fill_args(const TomoGeom* g, const void* views, int n_proj, const float* volpad, RayArgs* A)
{
    if (!g || !views || !volpad || n_proj <= 0) { tomo_set_error("ray operator: null pointer or n_proj <= 0"); return TOMO_E_ARG; }
    if (int e = check_sizes(g)) return e;
    A->volpad = volpad + TOMO_PAD_HEAD; A->views = (const double*)views;      // the padded volume proper starts behind the head slack
    A->meas = nullptr; A->proj = nullptr; A->dproj = nullptr; A->partial = nullptr; A->skip_separable = 0; A->skip_zq = 0;
    A->nx = g->nx; A->ny = g->ny; A->nz = g->nz; A->ndx = g->ndx; A->ndz = g->ndz; A->n_proj = n_proj;
    int nyp;
    tomo_pad_pitch(g->ny, g->nz, &nyp, &A->syp);
    A->sxp = nyp * A->syp;
    A->nxt = (g->ndx + TILE_X - 1) / TILE_X;
    A->nzt = (g->ndz + TILE_Z - 1) / TILE_Z;
    // bands of about RAY_BAND_TILES x-tiles: 16 tiles = 128 detector columns keep a band of a 36-plane slab of a 512^2
    // cross-section at ~12 MB (measured best on B200, profiles/README.md)
    A->xparts = (A->nxt + RAY_BAND_TILES - 1) / RAY_BAND_TILES;
    A->xpp = (A->nxt + A->xparts - 1) / A->xparts;
    const double nblocks = (double)A->xpp * A->xparts * A->nzt * n_proj;
    if (nblocks >= 2147483647.0) { tomo_set_error("too many detector tiles for one launch"); return TOMO_E_RANGE; }
    return 0;
}

// tile counts and launch order of the z-quad kernels
static void zq_tiling(const TomoGeom* g, int n_proj, RayArgs* Z)
{
    (void)n_proj;
    Z->nxt = (g->ndx + ZQ_TILE_X - 1) / ZQ_TILE_X;
    Z->nzt = (g->ndz + ZQ_TILE_Z - 1) / ZQ_TILE_Z;
    Z->xparts = (Z->nxt + ZQ_BAND_TILES - 1) / ZQ_BAND_TILES;
    Z->xpp = (Z->nxt + Z->xparts - 1) / Z->xparts;
}

extern "C" int tomo_forward(const TomoGeom* g, const void* views, int n_proj,
                            const float* volpad, float* proj, void* stream)
{
    return tomo_forward_ex(g, views, n_proj, 0, volpad, proj, stream);
}

extern "C" int tomo_forward_ex(const TomoGeom* g, const void* views, int n_proj, int kinds,
                               const float* volpad, float* proj, void* stream)
{
    RayArgs A;
    if (int e = fill_args(g, views, n_proj, volpad, &A)) return e;
    if (!proj) { tomo_set_error("tomo_forward: proj_dev is NULL"); return TOMO_E_ARG; }
    A.proj = proj;
    A.skip_separable = 1;
    const bool known = (kinds & TOMO_KINDS_KNOWN) != 0;
    // nearly untilted views (V_ZQ): z-quad kernel, which needs 16-byte aligned plane rows (128-bit loads)
    const bool zq_ok = (((uintptr_t)A.volpad) & 15u) == 0;
    A.skip_zq = zq_ok ? 1 : 0;
    const dim3 block(TILE_Z, TILE_X);
    if (!known || (kinds & TOMO_KINDS_GENERIC) || (!zq_ok && (kinds & TOMO_KINDS_ZQUAD))) {
        ray_kernel_forward<<<A.xpp * A.xparts * A.nzt * n_proj, block, 0, (cudaStream_t)stream>>>(A);
        if (int e = tomo_check_cuda(cudaGetLastError(), "ray_kernel_forward")) return e;
    }
    if (zq_ok && (!known || (kinds & TOMO_KINDS_ZQUAD))) {
        RayArgs Z = A;
        zq_tiling(g, n_proj, &Z);
        zq_kernel_forward<<<Z.xpp * Z.xparts * Z.nzt * n_proj, ZQ_THREADS, 0, (cudaStream_t)stream>>>(Z);
        if (int e = tomo_check_cuda(cudaGetLastError(), "zq_kernel_forward")) return e;
    }
    // untilted views (alpha = beta = 0): separable kernel; both kernels return at once for views of the other kind
    if (!known || (kinds & TOMO_KINDS_SEPARABLE)) return tomo_forward_separable_launch(g, views, n_proj, volpad, proj, stream);
    return 0;
}

extern "C" size_t tomo_proj_grad_workspace_bytes(const TomoGeom* g, int n_proj)
{
    if (!g || n_proj <= 0) return 0;
    const size_t nxt = (g->ndx + TILE_X - 1) / TILE_X, nzt = (g->ndz + TILE_Z - 1) / TILE_Z;
    int sxt, sch;
    tomo_grad_separable_tiles(g, &sxt, &sch);
    const size_t zxt = (g->ndx + ZQ_TILE_X - 1) / ZQ_TILE_X, zzt = (g->ndz + ZQ_TILE_Z - 1) / ZQ_TILE_Z;
    // block partials of ray_kernel_gradient followed by those of sep_gradient_kernel and of zq_kernel_gradient
    return sizeof(double) * NRED * (nxt * nzt + (size_t)sxt * sch + zxt * zzt) * (size_t)n_proj;
}

extern "C" int tomo_proj_grad(const TomoGeom* g, const void* views, int n_proj,
                              const float* volpad, const float* meas,
                              float* proj, float* dproj, double* grad6, double* cost,
                              void* workspace, size_t workspace_bytes, void* stream)
{
    return tomo_proj_grad_ex(g, views, n_proj, 0, volpad, meas, proj, dproj, grad6, cost, workspace, workspace_bytes, stream);
}

extern "C" int tomo_proj_grad_ex(const TomoGeom* g, const void* views, int n_proj, int kinds,
                                 const float* volpad, const float* meas,
                                 float* proj, float* dproj, double* grad6, double* cost,
                                 void* workspace, size_t workspace_bytes, void* stream)
{
    RayArgs A;
    if (int e = fill_args(g, views, n_proj, volpad, &A)) return e;
    A.meas = meas; A.proj = proj; A.dproj = dproj;
    const bool reduce = (grad6 != nullptr) || (cost != nullptr);
    if (reduce) {
        if (!meas) { tomo_set_error("tomo_proj_grad: grad6/cost need meas_dev"); return TOMO_E_ARG; }
        if (!workspace || workspace_bytes < tomo_proj_grad_workspace_bytes(g, n_proj)) {
            tomo_set_error("tomo_proj_grad: workspace too small (see tomo_proj_grad_workspace_bytes)");
            return TOMO_E_WORKSPACE;
        }
        A.partial = (double*)workspace;
    }
    A.skip_separable = 1;
    const bool known = (kinds & TOMO_KINDS_KNOWN) != 0;
    const bool zq_ok = (((uintptr_t)A.volpad) & 15u) == 0;
    A.skip_zq = zq_ok ? 1 : 0;
    const bool want_gen = !known || (kinds & TOMO_KINDS_GENERIC) || (!zq_ok && (kinds & TOMO_KINDS_ZQUAD));
    const bool want_sep = !known || (kinds & TOMO_KINDS_SEPARABLE);
    const bool want_zq = zq_ok && (!known || (kinds & TOMO_KINDS_ZQUAD));
    const dim3 block(TILE_Z, TILE_X);
    if (want_gen) {
        ray_kernel_gradient<<<A.xpp * A.xparts * A.nzt * n_proj, block, 0, (cudaStream_t)stream>>>(A);
        if (int e = tomo_check_cuda(cudaGetLastError(), "ray_kernel_gradient")) return e;
    }
    // untilted views: separable kernel with its own block partials behind the generic ones; then the z-quad kernel's
    double* sep_partial = reduce ? A.partial + (size_t)NRED * A.nxt * A.nzt * n_proj : nullptr;
    int sxt0, sch0;
    tomo_grad_separable_tiles(g, &sxt0, &sch0);
    RayArgs Z = A;
    zq_tiling(g, n_proj, &Z);
    Z.partial = reduce ? sep_partial + (size_t)NRED * sxt0 * sch0 * n_proj : nullptr;
    if (want_zq) {
        zq_kernel_gradient<<<Z.xpp * Z.xparts * Z.nzt * n_proj, ZQ_THREADS, 0, (cudaStream_t)stream>>>(Z);
        if (int e = tomo_check_cuda(cudaGetLastError(), "zq_kernel_gradient")) return e;
    }
    if (want_sep)
        if (int e = tomo_grad_separable_launch(g, views, n_proj, volpad, meas, proj, dproj, sep_partial, stream)) return e;
    if (reduce) {
        const int n = n_proj * NRED;
        int sxt, sch;
        tomo_grad_separable_tiles(g, &sxt, &sch);
        if (want_gen)
            grad_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(A.partial, A.views, 0, n_proj, A.nxt, A.nzt, grad6, cost);
        if (want_sep)
            grad_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sep_partial, A.views, 1, n_proj, sxt, sch, grad6, cost);
        if (want_zq)
            grad_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(Z.partial, A.views, 2, n_proj, Z.nxt, Z.nzt, grad6, cost);
        return tomo_check_cuda(cudaGetLastError(), "grad_finalize_kernel");
    }
    return 0;
}
