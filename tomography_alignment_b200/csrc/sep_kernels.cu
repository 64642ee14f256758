// sep_kernels.cu -- separable forward projector for untilted views (V_SEP == 1), sm_100a.  See sep_core.h.
//
// Block = 8 warps = 8 adjacent ix; a warp handles one ix and one chunk of SEP_CHUNK padded z planes, four planes per
// lane with 128-bit loads (the 32 lanes of a warp read 512 contiguous bytes of each of the 4 (x, y) corner rows).  All
// lanes of a warp share the (x, y) march, so there is no divergence.  The warp then stages S in shared memory and
// applies the 2-tap z interpolation for the detector rows whose floor plane falls in its chunk.  Blocks are launched
// band by band (sep_item) so that the part of the volume a band of detector columns reads stays in L2 across views.
#include <cuda_runtime.h>
#include "tomo_common.h"
#include "sep_core.h"

namespace {

constexpr int SEP_WARPS = 8;
#ifndef SEP_BAND_TILES
#define SEP_BAND_TILES 16      // launch order: bands of 16 x-tiles (128 detector columns), see sep_item()
#endif

// Launch order of the ray-driven separable kernels: x-tile within a band fastest, then view, then band, then z chunk.
// All views sweep over one band of detector columns of one z chunk before the next band starts; the part of the chunk
// a band touches rotates slowly with the view angle, so it stays in L2 instead of being streamed from HBM once per view.
// Returns false for the padding blocks of a ragged last band.  `logical` is the (chunk, view, xt) id the partials use.
__device__ __forceinline__ bool sep_item(long long pb, int nxt, int n_proj, int& xt, int& view, int& chunk, long long& logical)
{
    const int parts = (nxt + SEP_BAND_TILES - 1) / SEP_BAND_TILES, xpp = (nxt + parts - 1) / parts;
    const int xl = (int)(pb % xpp);
    view = (int)((pb / xpp) % n_proj);
    const int part = (int)((pb / ((long long)xpp * n_proj)) % parts);
    chunk = (int)(pb / ((long long)xpp * n_proj * parts));
    xt = part * xpp + xl;
    logical = ((long long)chunk * n_proj + view) * nxt + xt;
    return xt < nxt;
}
__host__ inline double sep_grid_blocks(int nxt, int nchunk, int n_proj)
{
    const int parts = (nxt + SEP_BAND_TILES - 1) / SEP_BAND_TILES, xpp = (nxt + parts - 1) / parts;
    return (double)xpp * parts * nchunk * n_proj;
}

struct SepArgs {
    const float*  volpad;
    const double* views;
    float*        proj;
    int nx, ny, nz, ndx, ndz, n_proj, sxp, syp, nzp;
    int nxt, nchunk;
};

__global__ void __launch_bounds__(32 * SEP_WARPS)
sep_forward_kernel(const SepArgs A)
{
    __shared__ float Srow[SEP_WARPS][SEP_CHUNK];
    const int lane = threadIdx.x, warp = threadIdx.y;
    // one block per (chunk, view, x-tile) item (a persistent grid-stride loop measured 25 % slower); blocks of
    // tilted views return at once
    {
    int xt, view, chunk; long long item;
    if (!sep_item(blockIdx.x, A.nxt, A.n_proj, xt, view, chunk, item)) return;
    const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
    if (V[V_SEP] == 0.0) return;                                   // tilted view: the generic kernel does it
    const int ix = xt * SEP_WARPS + warp;
    if (ix >= A.ndx) return;                                       // whole warp
    const RayDims dm = {A.nx, A.ny, A.nz, A.sxp, A.syp};
    const int zp0 = chunk * SEP_OUT;
    const int zq = zp0 + 4 * lane;
    float S[4] = {0.f, 0.f, 0.f, 0.f};
    if (zq < A.nzp) {
        SepSetup r;
        sep_setup(V, dm, ix, r);
        sep_march_xy(A.volpad, V, dm, r, zq, S);
    }
    *reinterpret_cast<float4*>(&Srow[warp][4 * lane]) = make_float4(S[0], S[1], S[2], S[3]);
    __syncwarp();
    // detector rows whose floor plane lies in [zp0, zp0 + SEP_OUT); rows below / above the padded volume go to the
    // first / last chunk and read zeros
    const double z0 = V[V_P00 + 2], wzv = V[V_W + 2];
    const int last = (chunk == A.nchunk - 1);
    int iz_lo = (chunk == 0) ? 0 : (int)fmin(fmax(ceil(((double)(zp0 - TOMO_PAD) - z0) / wzv) - 1.0, 0.0), (double)A.ndz);
    int iz_hi = last ? A.ndz : (int)fmin(fmax(ceil(((double)(zp0 + SEP_OUT - TOMO_PAD) - z0) / wzv) + 1.0, 0.0), (double)A.ndz);
    float* __restrict__ out = A.proj + ((size_t)view * A.ndx + ix) * A.ndz;
    for (int iz = iz_lo + lane; iz < iz_hi; iz += 32) {
        int fzp; float wz;
        sep_zcell(V, iz, fzp, wz);
        const int fc = min(max(fzp, 0), A.nzp - 2);                // plane used to pick the owning chunk
        const int owner = min(fc / SEP_OUT, A.nchunk - 1);
        if (owner != chunk) continue;
        float v = 0.f;
        if (fzp >= 0 && fzp <= A.nzp - 2) {
            const int k = fzp - zp0;                               // 0 <= k, k + 1 < SEP_CHUNK (k <= SEP_OUT + 2 in the last chunk)
            v = fmaf(wz, Srow[warp][k + 1] - Srow[warp][k], Srow[warp][k]);
        }
        out[iz] = v;
    }
    }
}

// ---- separable projection + gradient -----------------------------------------------------------------------------
struct SepGradArgs {
    const float*  volpad;
    const double* views;
    const float*  meas;      // nullable
    float*        proj;      // nullable
    float*        dproj;     // nullable
    double*       partial;   // nullable, [nchunk][n_proj][nxt][7]
    int nx, ny, nz, ndx, ndz, n_proj, sxp, syp, nzp;
    int nxt, nchunk;
};

#ifndef SEP_GRAD_MINB
#define SEP_GRAD_MINB 4      // 64 registers (4 blocks/SM) measured best: 26.5 ms vs 45 ms at ptxas' own 102
#endif
__global__ void __launch_bounds__(32 * SEP_WARPS, SEP_GRAD_MINB)
sep_gradient_kernel(const SepGradArgs A)
{
    __shared__ float M[SEP_WARPS][6][SEP_CHUNK];
    __shared__ double red_sm[SEP_WARPS][7];
    int xt, view, chunk; long long item;
    if (!sep_item(blockIdx.x, A.nxt, A.n_proj, xt, view, chunk, item)) return;
    const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
    if (V[V_SEP] == 0.0) return;                                   // tilted view (block-uniform): ray_kernel_gradient does it
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int ix = xt * SEP_WARPS + warp;
    const bool row = ix < A.ndx;                                   // warp-uniform
    const RayDims dm = {A.nx, A.ny, A.nz, A.sxp, A.syp};
    const int zp0 = chunk * SEP_OUT;
    const int zq = zp0 + 4 * lane;
    SepMoments m;
#pragma unroll
    for (int k = 0; k < 4; ++k) { m.S[k] = m.T[k] = m.Gx[k] = m.Gy[k] = m.TGx[k] = m.TGy[k] = 0.f; }
    if (row && zq < A.nzp) {
        SepSetup r;
        sep_setup(V, dm, ix, r);
        sep_march_xy_grad(A.volpad, V, dm, r, zq, m);
    }
    *reinterpret_cast<float4*>(&M[warp][0][4 * lane]) = make_float4(m.S[0], m.S[1], m.S[2], m.S[3]);
    *reinterpret_cast<float4*>(&M[warp][1][4 * lane]) = make_float4(m.T[0], m.T[1], m.T[2], m.T[3]);
    *reinterpret_cast<float4*>(&M[warp][2][4 * lane]) = make_float4(m.Gx[0], m.Gx[1], m.Gx[2], m.Gx[3]);
    *reinterpret_cast<float4*>(&M[warp][3][4 * lane]) = make_float4(m.Gy[0], m.Gy[1], m.Gy[2], m.Gy[3]);
    *reinterpret_cast<float4*>(&M[warp][4][4 * lane]) = make_float4(m.TGx[0], m.TGx[1], m.TGx[2], m.TGx[3]);
    *reinterpret_cast<float4*>(&M[warp][5][4 * lane]) = make_float4(m.TGy[0], m.TGy[1], m.TGy[2], m.TGy[3]);
    __syncwarp();
    double red[7] = {0, 0, 0, 0, 0, 0, 0};
    if (row) {
        const double z0 = V[V_P00 + 2], wzv = V[V_W + 2];
        const int last = (chunk == A.nchunk - 1);
        const int iz_lo = (chunk == 0) ? 0 : (int)fmin(fmax(ceil(((double)(zp0 - TOMO_PAD) - z0) / wzv) - 1.0, 0.0), (double)A.ndz);
        const int iz_hi = last ? A.ndz : (int)fmin(fmax(ceil(((double)(zp0 + SEP_OUT - TOMO_PAD) - z0) / wzv) + 1.0, 0.0), (double)A.ndz);
        const size_t n_det = (size_t)A.ndx * A.ndz;
        for (int iz = iz_lo + lane; iz < iz_hi; iz += 32) {
            int fzp; float wz;
            sep_zcell(V, iz, fzp, wz);
            const int fc = min(max(fzp, 0), A.nzp - 2);
            if (min(fc / SEP_OUT, A.nchunk - 1) != chunk) continue;
            RaySums sm;
            sm.acc = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) { sm.s0[k] = 0.f; sm.s1[k] = 0.f; }
            if (fzp >= 0 && fzp <= A.nzp - 2)
                sep_ray_sums(M[warp][0], M[warp][1], M[warp][2], M[warp][3], M[warp][4], M[warp][5], fzp - zp0, wz, sm);
            const size_t ray = (size_t)ix * A.ndz + iz;
            if (A.proj) A.proj[(size_t)view * n_det + ray] = sm.acc;
            float dp[6];
            ray_gradient(V, ix, iz, sm, dp);
            if (A.dproj) {
#pragma unroll
                for (int k = 0; k < 6; ++k) A.dproj[((size_t)view * 6 + k) * n_det + ray] = dp[k];
            }
            if (A.meas) {
                const double res = (double)A.meas[(size_t)view * n_det + ray] - (double)sm.acc;
#pragma unroll
                for (int k = 0; k < 6; ++k) red[k] += -(double)dp[k] * res;
                red[6] += 0.5 * res * res;
            }
        }
    }
    if (A.partial) {        // fixed-order block reduction, as in ray_kernel_gradient
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            double v = red[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) red_sm[warp][k] = v;
        }
        __syncthreads();
        if (warp == 0 && lane < 7) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < SEP_WARPS; ++w) v += red_sm[w][lane];
            A.partial[(size_t)item * 7 + lane] = v;
        }
    }
}

// ---- separable adjoint ------------------------------------------------------------------------------------------
constexpr int SA_WARPS_X = 4, SA_WARPS_Y = 4;        // a block = 4 x 4 voxel columns, one warp each
constexpr int SA_QPL = 4;                            // z quads per lane: a warp covers 32 * 4 * 4 = 512 planes

struct SepBackArgs {
    const float*  proj;      // [n_proj][ndx][ndz]
    const double* views;
    float*        yz;        // workspace [n_proj][ndx][nzw]
    float*        vol;
    int nx, ny, nz, ndx, ndz, n_proj, nzw, accumulate;
    int x_begin, x_end;      // x-slab of the volume this launch writes
};

// Yz[view][ix][z] (z-transpose of the 2-tap interpolation) for the separable views of the table
__global__ void __launch_bounds__(256)
sep_zgather_kernel(const SepBackArgs A)
{
    if (A.views[V_NSEP] == 0.0) return;                               // no untilted view in the table
    const long long nrows = (long long)A.n_proj * A.ndx;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int view = (int)(row / A.ndx), ix = (int)(row % A.ndx);
        const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
        if (V[V_SEP] == 0.0) continue;
        const float* __restrict__ Prow = A.proj + ((size_t)view * A.ndx + ix) * A.ndz;
        for (int z = threadIdx.x; z < A.nzw; z += 256)
            A.yz[((size_t)view * A.ndx + ix) * A.nzw + z] = (z < A.nz) ? sep_zgather(Prow, V, A.ndz, z) : 0.f;
    }
}

__global__ void __launch_bounds__(32 * SA_WARPS_X * SA_WARPS_Y)
sep_adjoint_kernel(const SepBackArgs A)
{
    const int lane = threadIdx.x;
    if (A.views[V_NSEP] == 0.0) return;                               // no untilted view in the table
    const int x = A.x_begin + blockIdx.y * SA_WARPS_X + threadIdx.z, y = blockIdx.x * SA_WARPS_Y + threadIdx.y;
    if (x >= A.x_end || y >= A.ny) return;                            // whole warp
    const int zbase = blockIdx.z * (128 * SA_QPL);
    f2 acc[SA_QPL][2];
#pragma unroll
    for (int q = 0; q < SA_QPL; ++q) { acc[q][0] = f2_make(0.f, 0.f); acc[q][1] = f2_make(0.f, 0.f); }
    for (int view = 0; view < A.n_proj; ++view) {
        const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
        if (V[V_SEP] == 0.0) continue;
        SepColumn c;
        sep_column_setup(V, A.ndx, x, y, c);                           // warp-uniform
        // lane k evaluates K for candidate ix number k, then the values are broadcast one by one
        const int nmi = c.mihi - c.milo + 1;
        for (int m0 = 0; m0 < nmi; m0 += 32) {
            const int mine = c.milo + m0 + lane;
            const float Kl = (mine <= c.mihi) ? sep_column_weight(V, c, mine) : 0.f;
            const int cnt = min(32, nmi - m0);
            for (int k = 0; k < cnt; ++k) {
                const float K = __shfl_sync(0xffffffffu, Kl, k);
                if (K == 0.f) continue;                                // warp-uniform
                const int ix = c.n0i + c.milo + m0 + k;
                const float* __restrict__ row = A.yz + ((size_t)view * A.ndx + ix) * A.nzw + zbase + 4 * lane;
                const f2 K2 = f2_make(K, K);
#pragma unroll
                for (int q = 0; q < SA_QPL; ++q) {
                    if (zbase + 128 * q + 4 * lane < A.nzw) {
                        const f4 v = sep_ld4(row + 128 * q);
                        acc[q][0] = f2_fma(K2, f2_make(v.x, v.y), acc[q][0]);
                        acc[q][1] = f2_fma(K2, f2_make(v.z, v.w), acc[q][1]);
                    }
                }
            }
        }
    }
    float* __restrict__ out = A.vol + ((size_t)x * A.ny + y) * A.nz;
#pragma unroll
    for (int q = 0; q < SA_QPL; ++q) {
        const int z = zbase + 128 * q + 4 * lane;
        const float r[4] = {acc[q][0].x, acc[q][0].y, acc[q][1].x, acc[q][1].y};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (z + k < A.nz) out[z + k] = A.accumulate ? out[z + k] + r[k] : r[k];
    }
}

}  // namespace

extern "C" void tomo_set_error(const char* msg);
int tomo_check_cuda(cudaError_t e, const char* what);

// Launch the separable forward projector over all views of the table; views with V_SEP == 0 return at once.
int tomo_forward_separable_launch(const TomoGeom* g, const void* views, int n_proj, const float* volpad, float* proj,
                                  void* stream)
{
    SepArgs A;
    A.volpad = volpad + TOMO_PAD_HEAD; A.views = (const double*)views; A.proj = proj;
    A.nx = g->nx; A.ny = g->ny; A.nz = g->nz; A.ndx = g->ndx; A.ndz = g->ndz; A.n_proj = n_proj;
    A.nzp = tomo_nzp(g->nz);                       // planes to visit; the pitches may be those of the next cube (tomo_pad_pitch)
    int nyp;
    tomo_pad_pitch(g->ny, g->nz, &nyp, &A.syp);
    A.sxp = nyp * A.syp;
    A.nxt = (g->ndx + SEP_WARPS - 1) / SEP_WARPS;
    A.nchunk = (A.nzp - 1 + SEP_OUT - 1) / SEP_OUT;
    // the last chunk must be able to serve floor planes up to nzp - 2 from SEP_CHUNK staged planes
    while ((A.nchunk - 1) * SEP_OUT + SEP_CHUNK < A.nzp) ++A.nchunk;
    const double nblocks = sep_grid_blocks(A.nxt, A.nchunk, n_proj);
    if (nblocks >= 2147483647.0) { tomo_set_error("separable forward: too many blocks for one launch"); return TOMO_E_RANGE; }
    sep_forward_kernel<<<(unsigned)nblocks, dim3(32, SEP_WARPS), 0, (cudaStream_t)stream>>>(A);
    return tomo_check_cuda(cudaGetLastError(), "sep_forward_kernel");
}

// Workspace (floats) the separable adjoint needs: Yz for every view of the table
size_t tomo_back_separable_workspace_bytes(const TomoGeom* g, int n_proj)
{
    return sizeof(float) * (size_t)n_proj * g->ndx * (size_t)(((g->nz + 3) / 4) * 4);
}

// vol (+)= A^T y restricted to the separable views of the table (the caller has already handled the others)
int tomo_back_separable_launch(const TomoGeom* g, const void* views, int n_proj, const float* proj, float* vol,
                               int accumulate, void* workspace, int x_begin, int x_end, void* stream)
{
    SepBackArgs A;
    A.proj = proj; A.views = (const double*)views; A.yz = (float*)workspace; A.vol = vol;
    A.nx = g->nx; A.ny = g->ny; A.nz = g->nz; A.ndx = g->ndx; A.ndz = g->ndz; A.n_proj = n_proj;
    A.nzw = ((g->nz + 3) / 4) * 4;
    A.accumulate = accumulate;
    A.x_begin = x_begin; A.x_end = x_end;
    const double nrows = (double)g->ndx * n_proj;
    if (x_begin == 0) {      // Yz does not depend on the slab: the launch of the first slab fills it for the ones that follow
        sep_zgather_kernel<<<(unsigned)(nrows < 148.0 * 32 ? nrows : 148.0 * 32), 256, 0, (cudaStream_t)stream>>>(A);
        if (int e = tomo_check_cuda(cudaGetLastError(), "sep_zgather_kernel")) return e;
    }
    const dim3 grid((g->ny + SA_WARPS_Y - 1) / SA_WARPS_Y, (x_end - x_begin + SA_WARPS_X - 1) / SA_WARPS_X,
                    (A.nzw + 128 * SA_QPL - 1) / (128 * SA_QPL));
    if (grid.y > 65535u || grid.z > 65535u) { tomo_set_error("separable adjoint: volume too large for the launch grid"); return TOMO_E_RANGE; }
    sep_adjoint_kernel<<<grid, dim3(32, SA_WARPS_Y, SA_WARPS_X), 0, (cudaStream_t)stream>>>(A);
    return tomo_check_cuda(cudaGetLastError(), "sep_adjoint_kernel");
}

// Tiling of the separable gradient (for the workspace layout and the finalize pass)
void tomo_grad_separable_tiles(const TomoGeom* g, int* nxt, int* nchunk)
{
    const int nzp = tomo_nzp(g->nz);
    *nxt = (g->ndx + SEP_WARPS - 1) / SEP_WARPS;
    int nc = (nzp - 1 + SEP_OUT - 1) / SEP_OUT;
    while ((nc - 1) * SEP_OUT + SEP_CHUNK < nzp) ++nc;
    *nchunk = nc;
}

int tomo_grad_separable_launch(const TomoGeom* g, const void* views, int n_proj, const float* volpad, const float* meas,
                               float* proj, float* dproj, double* partial, void* stream)
{
    SepGradArgs A;
    A.volpad = volpad + TOMO_PAD_HEAD; A.views = (const double*)views; A.meas = meas; A.proj = proj; A.dproj = dproj; A.partial = partial;
    A.nx = g->nx; A.ny = g->ny; A.nz = g->nz; A.ndx = g->ndx; A.ndz = g->ndz; A.n_proj = n_proj;
    A.nzp = tomo_nzp(g->nz);                       // planes to visit; the pitches may be those of the next cube (tomo_pad_pitch)
    int nyp;
    tomo_pad_pitch(g->ny, g->nz, &nyp, &A.syp);
    A.sxp = nyp * A.syp;
    tomo_grad_separable_tiles(g, &A.nxt, &A.nchunk);
    const double nblocks = sep_grid_blocks(A.nxt, A.nchunk, n_proj);
    if (nblocks >= 2147483647.0) { tomo_set_error("separable gradient: too many blocks for one launch"); return TOMO_E_RANGE; }
    sep_gradient_kernel<<<(unsigned)nblocks, dim3(32, SEP_WARPS), 0, (cudaStream_t)stream>>>(A);
    return tomo_check_cuda(cudaGetLastError(), "sep_gradient_kernel");
}
