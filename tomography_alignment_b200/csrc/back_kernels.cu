// back_kernels.cu -- backprojectors, sm_100a.  No atomics anywhere: every voxel is owned by one
// thread which accumulates all views in a register and writes once (bitwise deterministic).
//
// (1) adjoint_gather_kernel: vol (+)= A^T y, the exact transpose of the ray-driven trilinear
//     forward projector (src/ray_wt_grad.f90:20-91); this is the backprojection the reference's
//     solvers apply through scipy (recon/sirt.py:61).  The trilinear weight a sample at p gives
//     voxel v is prod_axis tent(p_a - v_a), tent(d) = max(0, 1 - |d|), so a voxel can gather from
//     the regular sample lattice p(n) = P00 + L n, n = (ix, iz, j): map the voxel into lattice
//     coordinates q = Linv (v - P00) in float64, round to the nearest lattice point n0, and visit
//     the integer offsets m with |m_k - rho_k| <= sum_a |Linv[k][a]| (rho = q - n0); every lattice
//     point with a non-zero weight lies in that box because |p - v|_inf < 1 there.
//     Distances d = L (m - rho) are formed from small numbers, so weights are accurate to ~1e-7.
// (1b) adjoint_tile_kernel: the same operator, sample-driven.  A block owns a TX x TY x TZ voxel tile
//     whose accumulators live in shared memory for the whole launch (all views), so the volume is
//     written exactly once.  For every view the block walks the rays that cross its tile: a warp takes
//     one ray column (fixed ix, 32 consecutive iz on its lanes), marches j through the tile and adds
//     y * w to the 8 corner cells of every sample with plain shared-memory read-modify-writes.
//     Race freedom without atomics:
//       * rays are processed in colour classes ix mod C with a __syncthreads between classes; the host
//         picks C per view (views.cpp, V_NCOL) so that samples of two same-colour rays never share a
//         corner voxel;
//       * inside a warp instruction two lanes alias only if they sit in the same z cell (adjacent
//         lanes, W_z < 1); those lanes are deferred to a second pass;
//       * the four z-floor corners and the four z-ceil corners are separated by __syncwarp (lane l's
//         ceil plane is lane l+1's floor plane).
//     A sample is handled by every tile that owns one of its corner voxels; contributions that land
//     on the tile's ghost cells are dropped (the neighbouring tile adds them), so each voxel sums
//     exactly the reference's terms, in a fixed order: results are bitwise reproducible.
// (2) voxel_bilinear_kernel: the orphan voxel-driven backprojector of src/back_projection.f90 /
//     src/external_back_projection.f90 (inverse pose convention, 4 bilinear taps, y ignored).
//
// Lanes run along z, which is contiguous in the volume and (for small tilts) maps to iz, which is
// contiguous in the projections, so both sides are coalesced.
#include <cuda_runtime.h>
#include "tomo_common.h"
#include "back_core.h"

namespace {

constexpr int BZ = 32, BY = 4, BX = 2;     // voxel tile of a block

struct BackArgs {
    const float*  proj;      // [n_proj][ndx][ndz]
    const double* views;
    float*        vol;       // [nx][ny][nz]
    int nx, ny, nz, ndx, ndz, n_proj, accumulate;
    int only_uncoloured;     // gather kernel: visit only the views the tile kernel skipped (V_NCOL == 0)
    double origin[3];        // voxel_bilinear only: the Fortran's origin argument
    double vox0[3], vpix[3]; // voxel_bilinear only: physical voxel centres = vox0 + idx*vpix
};

__global__ void __launch_bounds__(BZ * BY * BX)
adjoint_gather_kernel(const BackArgs A)
{
    const int z = blockIdx.x * BZ + threadIdx.x;
    const int y = blockIdx.y * BY + threadIdx.y;
    const int x = blockIdx.z * BX + threadIdx.z;
    if (x >= A.nx || y >= A.ny || z >= A.nz) return;
    const size_t n_det = (size_t)A.ndx * A.ndz;
    if (A.only_uncoloured && A.views[V_NUNCOL] == 0.0) return;      // record 0 holds the count
    float acc = 0.f;
    for (int view = 0; view < A.n_proj; ++view) {
        const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
        if (A.only_uncoloured && V[V_NCOL] != 0.0) continue;
        acc += adjoint_gather_view(A.proj + (size_t)view * n_det, V, A.ndx, A.ndz, x, y, z);
    }
    const size_t vi = ((size_t)x * A.ny + y) * A.nz + z;
    A.vol[vi] = A.accumulate ? A.vol[vi] + acc : acc;
}

__global__ void __launch_bounds__(BZ * BY * BX)
voxel_bilinear_kernel(const BackArgs A)
{
    const int z = blockIdx.x * BZ + threadIdx.x;
    const int y = blockIdx.y * BY + threadIdx.y;
    const int x = blockIdx.z * BX + threadIdx.z;
    if (x >= A.nx || y >= A.ny || z >= A.nz) return;
    const size_t n_det = (size_t)A.ndx * A.ndz;
    const double cx = A.vox0[0] + x * A.vpix[0], cy = A.vox0[1] + y * A.vpix[1], cz = A.vox0[2] + z * A.vpix[2];
    float acc = 0.f;
    for (int view = 0; view < A.n_proj; ++view)
        acc += voxel_bilinear_view(A.proj + (size_t)view * n_det, A.views + (size_t)view * TOMO_VIEW_STRIDE,
                                   A.ndx, A.ndz, A.origin, cx, cy, cz);
    const size_t vi = ((size_t)x * A.ny + y) * A.nz + z;
    A.vol[vi] = A.accumulate ? A.vol[vi] + acc : acc;
}


constexpr int TNW = 8;     // warps per block of the tile kernel

template <int TX, int TY>
__global__ void __launch_bounds__(TNW * 32)
adjoint_tile_kernel(const BackArgs A, const int ntx, const int nty, const int ntz)
{
    constexpr int TZ = TOMO_BT_Z, SX = TX + 4, SY = TY + 4, SZ = TZ + 4;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ float acc[];                       // [SX][SY][SZ], 2 ghost cells per side
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int bb = blockIdx.x;
    const int tz = bb % ntz; bb /= ntz;
    const int ty = bb % nty;
    const int tx = bb / nty;
    const int org[3] = {tx * TX - 2, ty * TY - 2, tz * TZ - 2};   // voxel coordinate of smem index 0
    for (int i = threadIdx.x; i < SX * SY * SZ; i += TNW * 32) acc[i] = 0.f;
    __syncthreads();

    const size_t n_det = (size_t)A.ndx * A.ndz;
    const int ust[3] = {SY * SZ, SZ, 1};
    // a sample is ours iff its floor cell lies in [1, T+1] per axis, i.e. 1 <= p_s < T+2 (smem coordinates)
    const float hi[3] = {(float)(TX + 2), (float)(TY + 2), (float)(TZ + 2)};
    const double hw[3] = {0.5 * (TX + 1) + 1e-3, 0.5 * (TY + 1) + 1e-3, 0.5 * (TZ + 1) + 1e-3};

    for (int view = 0; view < A.n_proj; ++view) {
        const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
        const float* __restrict__ P = A.proj + (size_t)view * n_det;
        const int ncol = (int)V[V_NCOL];
        const int nsamp = (int)V[V_N];
        if (ncol == 0) continue;                         // outside the scatter envelope: the gather kernel adds it
        // lattice bounding box of the active region: centre +- sum |Linv| * half widths
        double B[3], cc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            B[a] = V[V_P00 + a] - (double)org[a];        // p_s = B + ix U + iz W + j D
            cc[a] = 1.0 + hw[a] - 1e-3 - B[a];           // box centre minus lattice origin
        }
        const double ixc = V[V_LINV + 0] * cc[0] + V[V_LINV + 1] * cc[1] + V[V_LINV + 2] * cc[2];
        const double izc = V[V_LINV + 3] * cc[0] + V[V_LINV + 4] * cc[1] + V[V_LINV + 5] * cc[2];
        const double rix = fabs(V[V_LINV + 0]) * hw[0] + fabs(V[V_LINV + 1]) * hw[1] + fabs(V[V_LINV + 2]) * hw[2];
        const double riz = fabs(V[V_LINV + 3]) * hw[0] + fabs(V[V_LINV + 4]) * hw[1] + fabs(V[V_LINV + 5]) * hw[2];
        const int ix_lo = max(0, (int)ceil(fmax(ixc - rix, -1.0))), ix_hi = min(A.ndx - 1, (int)floor(fmin(ixc + rix, 2.0e9)));
        const int iz_lo = max(0, (int)ceil(fmax(izc - riz, -1.0))), iz_hi = min(A.ndz - 1, (int)floor(fmin(izc + riz, 2.0e9)));
        if (ix_lo > ix_hi || iz_lo > iz_hi) continue;    // block-uniform: no ray of this view crosses the tile

        // mirrored-frame constants (as in ray_core.h): one-sided carries, signed smem strides
        float df[3], invd[3], dD[3];
        int sg[3], st[3], stepoff = 0, zstep;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double d = V[V_D + a], ad = fabs(d), ai = floor(ad);
            dD[a] = (float)d;
            invd[a] = (float)V[V_INVD + a];
            df[a] = (float)(ad - ai);
            sg[a] = (d < 0.0) ? -1 : 1;
            st[a] = sg[a] * ust[a];
            stepoff += (int)ai * st[a];
            if (a == 2) zstep = (int)ai * sg[a];
        }
        const int o01 = st[1], o10 = st[0], o11 = st[0] + st[1], oz = st[2];

        for (int c = 0; c < ncol; ++c) {
            const int first = ix_lo + (((c - ix_lo) % ncol) + ncol) % ncol;
            for (int ix = first + ncol * warp; ix <= ix_hi; ix += ncol * TNW) {
                for (int izb = iz_lo; izb <= iz_hi; izb += 32) {
                    const int iz = izb + lane;
                    const bool valid = iz <= iz_hi;
                    double pr[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a) pr[a] = B[a] + (double)ix * V[V_U + a] + (double)iz * V[V_W + a];
                    // per-lane sample range inside the tile's active box
                    float jlo = 0.f, jhi = (float)nsamp;
                    bool empty = !valid;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float p = (float)pr[a];
                        if (dD[a] > 0.f)      { jlo = fmaxf(jlo, (1.f - p) * invd[a]);   jhi = fminf(jhi, (hi[a] - p) * invd[a]); }
                        else if (dD[a] < 0.f) { jlo = fmaxf(jlo, (hi[a] - p) * invd[a]); jhi = fminf(jhi, (1.f - p) * invd[a]); }
                        else if (p < 1.f || p >= hi[a]) empty = true;
                    }
                    jlo = fminf(fmaxf(jlo, 0.f), (float)nsamp);
                    jhi = fminf(fmaxf(jhi, -1.f), (float)nsamp);
                    int j0 = (int)ceilf(jlo), j1 = min((int)floorf(jhi) + 1, nsamp);
                    if (empty || j1 <= j0) { j0 = 0x7fffffff; j1 = -0x7fffffff; }
                    const int jw0 = __reduce_min_sync(FULL, j0), jw1 = __reduce_max_sync(FULL, j1);
                    if (jw0 >= jw1) continue;                                   // warp-uniform
                    const float yv = (j1 > j0) ? __ldg(P + (size_t)ix * A.ndz + iz) : 0.f;

                    // state at sample jw0, from float64
                    float f[3];
                    int off = 0, zc = 0;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double q = (double)sg[a] * (pr[a] + (double)jw0 * V[V_D + a]);
                        const double qi = floor(q);
                        f[a] = (float)(q - qi);
                        int i = (int)fmin(fmax(qi, -1.0e6), 1.0e6);
                        if (f[a] >= 1.0f) { f[a] -= 1.0f; i += 1; }
                        off += sg[a] * i * ust[a];
                        if (a == 2) zc = sg[a] * i;
                    }
                    for (int j = jw0; j < jw1; ++j) {
                        const bool act = (j >= j0) && (j < j1);
                        // adjacent lanes in the same z cell would alias inside one instruction
                        const int zprev = __shfl_up_sync(FULL, zc, 1);
                        const bool aprev = __shfl_up_sync(FULL, (int)act, 1) != 0;
                        const bool dup = act && aprev && (lane > 0) && (zprev == zc);
                        const bool anydup = __any_sync(FULL, dup);
                        const float wx1 = f[0] * yv, wx0 = yv - wx1;
                        const float wy1 = f[1], wy0 = 1.f - wy1;
                        const float w00 = wx0 * wy0, w01 = wx0 * wy1, w10 = wx1 * wy0, w11 = wx1 * wy1;
                        const float wz1 = f[2], wz0 = 1.f - wz1;
                        float* __restrict__ s = acc + off;
                        bool go = act && !dup;
#pragma unroll 1
                        for (int pass = 0; pass < 2; ++pass) {
                            if (go) {
                                const float a0 = s[0], a1 = s[o01], a2 = s[o10], a3 = s[o11];
                                s[0]   = fmaf(w00, wz0, a0);
                                s[o01] = fmaf(w01, wz0, a1);
                                s[o10] = fmaf(w10, wz0, a2);
                                s[o11] = fmaf(w11, wz0, a3);
                            }
                            __syncwarp();
                            if (go) {
                                const float a0 = s[oz], a1 = s[oz + o01], a2 = s[oz + o10], a3 = s[oz + o11];
                                s[oz]       = fmaf(w00, wz1, a0);
                                s[oz + o01] = fmaf(w01, wz1, a1);
                                s[oz + o10] = fmaf(w10, wz1, a2);
                                s[oz + o11] = fmaf(w11, wz1, a3);
                            }
                            __syncwarp();
                            if (!anydup) break;
                            go = dup;
                        }
                        // advance one sample
                        f[0] += df[0]; f[1] += df[1]; f[2] += df[2];
                        off += stepoff; zc += zstep;
                        if (f[0] >= 1.0f) { f[0] -= 1.0f; off += st[0]; }
                        if (f[1] >= 1.0f) { f[1] -= 1.0f; off += st[1]; }
                        if (f[2] >= 1.0f) { f[2] -= 1.0f; off += st[2]; zc += sg[2]; }
                    }
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // write the interior of the tile once
    for (int i = threadIdx.x; i < TX * TY * 32; i += TNW * 32) {
        const int zz = i & 31, yy = (i >> 5) % TY, xx = (i >> 5) / TY;
        const int x = tx * TX + xx, y = ty * TY + yy, z = tz * TZ + zz;
        if (zz < TZ && x < A.nx && y < A.ny && z < A.nz) {
            const float v = acc[((xx + 2) * SY + (yy + 2)) * SZ + zz + 2];
            const size_t vi = ((size_t)x * A.ny + y) * A.nz + z;
            A.vol[vi] = A.accumulate ? A.vol[vi] + v : v;
        }
    }
}

}  // namespace

extern "C" void tomo_set_error(const char* msg);
int tomo_check_cuda(cudaError_t e, const char* what);

static int fill_back(const TomoGeom* g, const void* views, int n_proj, const float* proj, float* vol,
                     int accumulate, BackArgs* A, dim3* grid)
{
    if (!g || !views || !proj || !vol || n_proj <= 0) { tomo_set_error("backprojector: null pointer or n_proj <= 0"); return TOMO_E_ARG; }
    A->proj = proj; A->views = (const double*)views; A->vol = vol;
    A->nx = g->nx; A->ny = g->ny; A->nz = g->nz; A->ndx = g->ndx; A->ndz = g->ndz;
    A->n_proj = n_proj; A->accumulate = accumulate; A->only_uncoloured = 0;
    for (int a = 0; a < 3; ++a) { A->origin[a] = 0.0; A->vox0[a] = g->vox_origin[a]; A->vpix[a] = g->vox_pix[a]; }
    *grid = dim3((g->nz + BZ - 1) / BZ, (g->ny + BY - 1) / BY, (g->nx + BX - 1) / BX);
    if (grid->y > 65535u || grid->z > 65535u) { tomo_set_error("backprojector: volume too large for the launch grid"); return TOMO_E_RANGE; }
    return 0;
}

extern "C" int tomo_back_adjoint_gather(const TomoGeom* g, const void* views, int n_proj,
                                        const float* proj, float* vol, int accumulate, void* stream)
{
    BackArgs A; dim3 grid;
    if (int e = fill_back(g, views, n_proj, proj, vol, accumulate, &A, &grid)) return e;
    adjoint_gather_kernel<<<grid, dim3(BZ, BY, BX), 0, (cudaStream_t)stream>>>(A);
    return tomo_check_cuda(cudaGetLastError(), "adjoint_gather_kernel");
}

extern "C" int tomo_back_adjoint(const TomoGeom* g, const void* views, int n_proj,
                                 const float* proj, float* vol, int accumulate, void* stream)
{
    BackArgs A; dim3 grid;
    if (int e = fill_back(g, views, n_proj, proj, vol, accumulate, &A, &grid)) return e;
    constexpr int TX = TOMO_BT_X, TY = TOMO_BT_Y, TZ = TOMO_BT_Z;
    const int ntx = (g->nx + TX - 1) / TX, nty = (g->ny + TY - 1) / TY, ntz = (g->nz + TZ - 1) / TZ;
    const size_t smem = sizeof(float) * (TX + 4) * (TY + 4) * (TZ + 4);
    const double nblocks = (double)ntx * nty * ntz;
    if (nblocks >= 2147483647.0) { tomo_set_error("tomo_back_adjoint: too many tiles"); return TOMO_E_RANGE; }
    cudaError_t ce = cudaFuncSetAttribute(adjoint_tile_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (int e = tomo_check_cuda(ce, "cudaFuncSetAttribute(adjoint_tile_kernel)")) return e;
    adjoint_tile_kernel<TX, TY><<<(unsigned)(ntx * nty * ntz), TNW * 32, smem, (cudaStream_t)stream>>>(A, ntx, nty, ntz);
    if (int e = tomo_check_cuda(cudaGetLastError(), "adjoint_tile_kernel")) return e;
    // views outside the scatter envelope (rays nearly parallel to z; none for tomographic poses): the
    // gather kernel adds them; it returns at once when record 0 says there are none
    A.only_uncoloured = 1; A.accumulate = 1;
    adjoint_gather_kernel<<<grid, dim3(BZ, BY, BX), 0, (cudaStream_t)stream>>>(A);
    return tomo_check_cuda(cudaGetLastError(), "adjoint_gather_kernel(uncoloured)");
}

extern "C" int tomo_back_voxel_bilinear(const TomoGeom* g, const void* views, int n_proj,
                                        const double origin[3], const float* proj, float* vol,
                                        int accumulate, void* stream)
{
    BackArgs A; dim3 grid;
    if (!origin) { tomo_set_error("tomo_back_voxel_bilinear: origin is NULL"); return TOMO_E_ARG; }
    if (int e = fill_back(g, views, n_proj, proj, vol, accumulate, &A, &grid)) return e;
    for (int a = 0; a < 3; ++a) A.origin[a] = origin[a];
    voxel_bilinear_kernel<<<grid, dim3(BZ, BY, BX), 0, (cudaStream_t)stream>>>(A);
    return tomo_check_cuda(cudaGetLastError(), "voxel_bilinear_kernel");
}
