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
//     The inner loop is branch-free: a lane outside its sample range is redirected to dummy cells in its
//     own bank (guard / ghost area) instead of being predicated off, and the corner pairs go through
//     packed fp32x2 FMAs -- 51 SASS instructions per warp-sample (16 of them the LDS/STS of the 8 RMWs).
//     A sample is handled by every tile that owns one of its corner voxels; contributions that land
//     on the tile's ghost cells are dropped (the neighbouring tile adds them), so each voxel sums
//     exactly the reference's terms, in a fixed order: results are bitwise reproducible.
// (2) voxel_bilinear_kernel: the orphan voxel-driven backprojector of src/back_projection.f90 /
//     src/external_back_projection.f90 (inverse pose convention, 4 bilinear taps, y ignored).
// (2b) voxel_bilinear_tma_kernel: the same operator with the projection tile of every view staged in shared
//     memory by TMA.  A block owns a 16 x 16 x 32 voxel brick (32 accumulators per thread, lanes along z) for
//     the whole launch; one thread computes the brick's detector footprint per view and issues one
//     cp.async.bulk.tensor (3-D map {z', x', view}, box 44 x 32 x 1, z' start a multiple of 4) into a 4-stage ring of mbarriers, so the
//     loads of the next views overlap the interpolation of the current one.  Out-of-detector parts of the box
//     are zero-filled by the TMA unit, which IS the reference's per-tap bounds check.  No atomics, each voxel
//     written once, bitwise reproducible.
//
// Lanes run along z, which is contiguous in the volume and (for small tilts) maps to iz, which is
// contiguous in the projections, so both sides are coalesced.
#include <cuda.h>            // CUtensorMap types; the encoder is fetched through cudaGetDriverEntryPoint (no libcuda link)
#include <cuda_runtime.h>
#include <cmath>
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
    int skip_separable;      // leave views with V_SEP == 1 to the separable adjoint (workspace variant)
    int only_vbig;           // voxel_bilinear: visit only the views the TMA kernel skipped (V_VBOK == 0)
    int x_begin, x_end;      // adjoint kernels: the x-slab [x_begin, x_end) of the volume this launch writes
    double origin[3];        // voxel_bilinear only: the Fortran's origin argument
    double vox0[3], vpix[3]; // voxel_bilinear only: physical voxel centres = vox0 + idx*vpix
};

__global__ void __launch_bounds__(BZ * BY * BX)
adjoint_gather_kernel(const BackArgs A)
{
    const int z = blockIdx.x * BZ + threadIdx.x;
    const int y = blockIdx.y * BY + threadIdx.y;
    const int x = A.x_begin + blockIdx.z * BX + threadIdx.z;
    if (x >= A.x_end || y >= A.ny || z >= A.nz) return;
    const size_t n_det = (size_t)A.ndx * A.ndz;
    if (A.only_uncoloured && A.views[V_NUNCOL] == 0.0) return;      // every record holds the table's count
    float acc = 0.f;
    for (int view = 0; view < A.n_proj; ++view) {
        const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
        if (A.only_uncoloured && V[V_NCOL] != 0.0) continue;
        if (A.skip_separable && V[V_SEP] != 0.0) continue;
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
    if (A.only_vbig && A.views[V_NVBIG] == 0.0) return;             // every record holds the table's count
    const double cx = A.vox0[0] + x * A.vpix[0], cy = A.vox0[1] + y * A.vpix[1], cz = A.vox0[2] + z * A.vpix[2];
    float acc = 0.f;
    for (int view = 0; view < A.n_proj; ++view) {
        const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
        if (A.only_vbig && V[V_VBOK] != 0.0) continue;
        acc += voxel_bilinear_view(A.proj + (size_t)view * n_det, V, A.ndx, A.ndz, A.origin, cx, cy, cz);
    }
    const size_t vi = ((size_t)x * A.ny + y) * A.nz + z;
    A.vol[vi] = A.accumulate ? A.vol[vi] + acc : acc;
}


// ---------------------------------------------------------------------------------------------------
// (2b) TMA-staged voxel-driven backprojector
constexpr int VBX = TOMO_VB_X, VBY = TOMO_VB_Y, VBZ = TOMO_VB_Z;     // voxel brick of a block
constexpr int VTX = TOMO_VB_TX, VTZ = TOMO_VB_TZ;                    // staged detector box: VTX rows (x') of VTZ pixels (z')
constexpr int VB_WARPS = 8, VB_STAGES = 4;
constexpr int VB_YPW = VBY / VB_WARPS;                               // y rows of the brick per warp
constexpr unsigned VB_TILE_BYTES = VTX * VTZ * sizeof(float);
static_assert(VBZ == 32 && VBY % VB_WARPS == 0 && (VTZ * sizeof(float)) % 16 == 0, "brick / box shape");

struct VbStage { float ux0, uz0, a00, a01, a02, a20, a21, a22; };   // per (brick, view) constants written by the producer

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__global__ void __launch_bounds__(VB_WARPS * 32)
voxel_bilinear_tma_kernel(const BackArgs A, const __grid_constant__ CUtensorMap tmap)
{
    __shared__ __align__(128) float tile[VB_STAGES][VTX * VTZ];
    __shared__ VbStage hdr[VB_STAGES];
    __shared__ __align__(8) unsigned long long full[VB_STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.z * VBX, y0 = blockIdx.y * VBY, z0 = blockIdx.x * VBZ;
    const unsigned tile_b = (unsigned)__cvta_generic_to_shared(&tile[0][0]);
    const unsigned full_b = (unsigned)__cvta_generic_to_shared(&full[0]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < VB_STAGES; ++s) mbar_init(full_b + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- producer (thread 0): footprint of the brick under the next staged view, then one TMA load ----
    int next_view = 0;
    auto issue = [&](int stage) {
        while (next_view < A.n_proj && A.views[(size_t)next_view * TOMO_VIEW_STRIDE + V_VBOK] == 0.0) ++next_view;
        if (next_view >= A.n_proj) return;
        const double* __restrict__ V = A.views + (size_t)next_view * TOMO_VIEW_STRIDE;
        const double c0[3] = {A.vox0[0] + x0 * A.vpix[0], A.vox0[1] + y0 * A.vpix[1], A.vox0[2] + z0 * A.vpix[2]};
        const int ext[3] = {VBX - 1, VBY - 1, VBZ - 1};
        double u0[2], umin[2], a[2][3];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = 2 * r;                                           // x' (row 0) and z' (row 2) of Ry Rx Rz
            u0[r] = V[V_VROT + 3 * row] * c0[0] + V[V_VROT + 3 * row + 1] * c0[1] + V[V_VROT + 3 * row + 2] * c0[2]
                    + V[V_VTR + row] - A.origin[row];
            umin[r] = u0[r];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                a[r][k] = V[V_VROT + 3 * row + k] * A.vpix[k];
                umin[r] += fmin(0.0, a[r][k] * ext[k]);
            }
        }
        // box origin; clamped far outside the detector (an all-zero box) so the int conversion is defined.  The z' start is
        // rounded down to a multiple of 4: the TMA unit faults on a start that is not 16-byte aligned along the inner dimension
        const double ox = floor(fmin(fmax(umin[0] - 1e-3, -1.0e6), 1.0e6));
        const double oz = 4.0 * floor(0.25 * fmin(fmax(umin[1] - 1e-3, -1.0e6), 1.0e6));
        VbStage h;
        h.ux0 = (float)(fmin(fmax(u0[0] - ox, -1.0e6), 1.0e6) - 0.5); h.uz0 = (float)(fmin(fmax(u0[1] - oz, -1.0e6), 1.0e6) - 0.5);
        h.a00 = (float)a[0][0]; h.a01 = (float)a[0][1]; h.a02 = (float)a[0][2];
        h.a20 = (float)a[1][0]; h.a21 = (float)a[1][1]; h.a22 = (float)a[1][2];
        hdr[stage] = h;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // earlier generic reads of this stage vs the async write
        mbar_expect_tx(full_b + 8u * stage, VB_TILE_BYTES);
        tma_load_3d(tile_b + VB_TILE_BYTES * stage, &tmap, full_b + 8u * stage, (int)oz, (int)ox, next_view);
        ++next_view;
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < VB_STAGES; ++s) issue(s);

    float acc[VB_YPW * VBX];
#pragma unroll
    for (int i = 0; i < VB_YPW * VBX; ++i) acc[i] = 0.f;

    // floor by magic number: t = (u - 0.5) + 1.5 * 2^23 rounds u - 0.5 to the nearest integer n = floor(u) and leaves n
    // in the low mantissa bits; on an exact tie either neighbour cell gives the same bilinear value (weight 0 / 1).
    constexpr float MAGIC = 12582912.0f;
    const float fl = (float)lane;
    int cnt = 0;
    for (int view = 0; view < A.n_proj; ++view) {
        if (A.views[(size_t)view * TOMO_VIEW_STRIDE + V_VBOK] == 0.0) continue;      // block-uniform; the plain kernel adds it
        const int stage = cnt % VB_STAGES;
        mbar_wait(full_b + 8u * stage, (unsigned)((cnt / VB_STAGES) & 1));
        const VbStage h = hdr[stage];
        const unsigned basec = tile_b + VB_TILE_BYTES * stage;       // shared byte address of box cell (0, 0)
        const float bxz = fmaf(fl, h.a02, h.ux0), bzz = fmaf(fl, h.a22, h.uz0);
#pragma unroll
        for (int yy = 0; yy < VB_YPW; ++yy) {
            const float fy = (float)(warp * VB_YPW + yy);
            const float uxy = fmaf(fy, h.a01, bxz), uzy = fmaf(fy, h.a21, bzz);
#pragma unroll
            for (int xx = 0; xx < VBX; ++xx) {
                const float ux = fmaf((float)xx, h.a00, uxy), uz = fmaf((float)xx, h.a20, uzy);   // u - 0.5, box-local
                const float tx = ux + MAGIC, tz = uz + MAGIC;
                const float dx = ux - (tx - MAGIC), dz = uz - (tz - MAGIC);                       // frac - 0.5
                const float ax = 0.5f + dx, wx0 = 0.5f - dx, az = 0.5f + dz, wz0 = 0.5f - dz;
                const unsigned off = (__float_as_uint(tx) & 0xffu) * (unsigned)VTZ + (__float_as_uint(tz) & 0xffu);
                const float* __restrict__ t = (const float*)__cvta_shared_to_generic(basec + 4u * off);
                const float lo = fmaf(t[1], az, t[0] * wz0), hi = fmaf(t[VTZ + 1], az, t[VTZ] * wz0);
                acc[yy * VBX + xx] = fmaf(hi, ax, fmaf(lo, wx0, acc[yy * VBX + xx]));
            }
        }
        __syncthreads();                                   // every thread is done with this stage
        if (threadIdx.x == 0) issue(stage);
        ++cnt;
    }

    const int z = z0 + lane;
    if (z < A.nz) {
#pragma unroll
        for (int yy = 0; yy < VB_YPW; ++yy) {
            const int y = y0 + warp * VB_YPW + yy;
#pragma unroll
            for (int xx = 0; xx < VBX; ++xx) {
                const int x = x0 + xx;
                if (x < A.nx && y < A.ny) {
                    const size_t vi = ((size_t)x * A.ny + y) * A.nz + z;
                    A.vol[vi] = A.accumulate ? A.vol[vi] + acc[yy * VBX + xx] : acc[yy * VBX + xx];
                }
            }
        }
    }
}

constexpr int TNW = TOMO_BT_WARPS;     // warps per block of the tile kernel
// Shared-memory tile: one ghost cell per side, [TSX][TSY][TSZ] with TSZ = 32 so that lanes that straddle two
// rows or planes still hit 32 distinct banks.  A sample whose float32-marched cell lands one cell outside the
// active range (only possible within ~1e-5 of the boundary) writes one cell further out: along z and y that
// address wraps into a ghost cell of the neighbouring row / plane, along x into the guard planes allocated
// before and after the tile.  Ghost and guard cells are never read back, so such writes (and races on
// them) cannot reach an interior voxel.
constexpr int TSX = TOMO_BT_X + 2, TSY = TOMO_BT_Y + 2, TSZ = TOMO_BT_Z + 2;
constexpr int TGUARD = TSY * TSZ + TSZ + 1;                  // floats before and after the tile
constexpr int TSMEM_FLOATS = TSX * TSY * TSZ + 2 * TGUARD;
static_assert(TSZ == 32, "the tile kernel maps the z cells of a tile onto the 32 lanes / banks");

// Per-view constants of the tile kernel, float32 and tile-local: p_s = Bc + di*U + dk*W + dj*D with
// di = ix - ix_lo, dk = iz - iz_lo, dj = j - jc (all small), p_s in smem-cell coordinates.
struct TileView {
    float Bc[3], U[3], W[3], D[3], invD[3];
    float df[3];            // fractional part of |D|
    int   di_step;          // integer part of |D| folded into the address step (mirrored frame)
    int   ix_lo, ix_hi, iz_lo, iz_hi, jc, djmin, djmax, ncol;
    float dupthr;           // same-z-cell test threshold for adjacent lanes
    int   view;             // index of the view these constants belong to; -1: no view left for this tile
    int   sgn;              // bit 2/1/0 set: D_x / D_y / D_z negative
};

// Four read-modify-writes s[off_k] += w_k * wz on shared memory as two packed fp32x2 FMAs; the four loads
// are issued before the first store.  There is no predicate: a lane with nothing to add is pointed at its
// own dummy cells (see tile_march_view), one select instead of a divergent branch + reconvergence barrier.
template <int O1, int O2, int O3>
__device__ __forceinline__ void rmw4(float* s, float2 wA, float2 wB, float wz)
{
    const float2 aA = make_float2(s[0], s[O1]), aB = make_float2(s[O2], s[O3]);
    const float2 z2 = make_float2(wz, wz);
    const float2 rA = __ffma2_rn(wA, z2, aA), rB = __ffma2_rn(wB, z2, aB);
    s[0] = rA.x; s[O1] = rA.y; s[O2] = rB.x; s[O3] = rB.y;
}

// March the rays of one view through the tile.  SGX/SGY/SGZ = sign of D per axis: with the mirrored
// frame q = SG * p_s the carries are one-sided and the 7 corner offsets are compile-time immediates.
template <int SGX, int SGY, int SGZ>
__device__ __forceinline__ void tile_march_view(float* __restrict__ acc, const TileView& tv,
                                                const float* __restrict__ P, int ndz, int lane, int warp)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int STX = SGX * TSY * TSZ, STY = SGY * TSZ, STZ = SGZ;
    constexpr int O01 = STY, O10 = STX, O11 = STX + STY, OZ = STZ;
    const float hi[3] = {(float)(TOMO_BT_X + 1), (float)(TOMO_BT_Y + 1), (float)(TOMO_BT_Z + 1)};
    const unsigned stepoff4 = 4u * (unsigned)tv.di_step;        // byte step of the integer part of |D| (0 for steps below one voxel)
    // Dummy cells: a lane with nothing to add is redirected to (its own bank) + DUMMY, so it stays conflict-free
    // against the active lanes.  The eight corner offsets span [OMIN, OMIN + TGUARD - 1]; with DUMMY = -OMIN
    // rounded up to a multiple of 32 every dummy access falls into the front guard or the x = 0 ghost plane of the
    // tile (never read back).  smem_raw is 128-byte aligned, so "own bank" is bits 2..6 of the shared address.
    constexpr int OMIN = (STX < 0 ? STX : 0) + (STY < 0 ? STY : 0) + (STZ < 0 ? STZ : 0);
    constexpr int DUMMY = ((-OMIN + 31) / 32) * 32;
    static_assert(DUMMY + 31 + OMIN + TGUARD - 1 < TGUARD + TSY * TSZ, "dummy cells leave the ghost plane");
    const unsigned acc_b = (unsigned)__cvta_generic_to_shared(acc);
    unsigned dummy_b = acc_b - 4u * TGUARD + 4u * DUMMY;
    asm volatile("" : "+r"(dummy_b));                       // a register, not a per-sample recomputation
    const float df0 = tv.df[0], df1 = tv.df[1], df2 = tv.df[2], dupthr = tv.dupthr;
    const int ncol = tv.ncol, ix_lo = tv.ix_lo, ix_hi = tv.ix_hi;
    for (int c = 0; c < ncol; ++c) {
        const int first = ix_lo + (((c - ix_lo) % ncol) + ncol) % ncol;
        for (int ix = first + ncol * warp; ix <= ix_hi; ix += ncol * TNW) {
            for (int izb = tv.iz_lo; izb <= tv.iz_hi; izb += 32) {
                const int iz = izb + lane;
                const bool valid = iz <= tv.iz_hi;
                const float fdi = (float)(ix - tv.ix_lo), fdk = (float)(iz - tv.iz_lo);
                float pr[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) pr[a] = fmaf(fdk, tv.W[a], fmaf(fdi, tv.U[a], tv.Bc[a]));
                // per-lane sample range (relative to jc) inside the tile's active box 0 <= p_s < T+1
                float jlo = (float)tv.djmin, jhi = (float)tv.djmax;
                bool empty = !valid;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    if (tv.D[a] > 0.f)      { jlo = fmaxf(jlo, (0.f - pr[a]) * tv.invD[a]);   jhi = fminf(jhi, (hi[a] - pr[a]) * tv.invD[a]); }
                    else if (tv.D[a] < 0.f) { jlo = fmaxf(jlo, (hi[a] - pr[a]) * tv.invD[a]); jhi = fminf(jhi, (0.f - pr[a]) * tv.invD[a]); }
                    else if (pr[a] < 0.f || pr[a] >= hi[a]) empty = true;
                }
                jlo = fminf(fmaxf(jlo, -1.0e6f), 1.0e6f);
                jhi = fminf(fmaxf(jhi, -1.0e6f), 1.0e6f);
                int j0 = (int)ceilf(jlo), j1 = min((int)floorf(jhi) + 1, tv.djmax);
                if (empty || j1 <= j0) { j0 = 0x7fffffff; j1 = -0x7fffffff; }
                const int jw0 = __reduce_min_sync(FULL, j0), jw1 = __reduce_max_sync(FULL, j1);
                if (jw0 >= jw1) continue;                                   // warp-uniform
                const bool live = j1 > j0;
                const float yv = live ? __ldg(P + (size_t)ix * ndz + iz) : 0.f;
                unsigned span = live ? (unsigned)(j1 - j0) : 0u;
                unsigned k = (unsigned)(jw0 - (live ? j0 : 0));             // j - j0, wraps below zero
                asm volatile("" : "+r"(span));                              // keep it in a register (no per-sample recompute)

                // mirrored-frame state at sample jw0
                const float fj = (float)jw0;
                const float qx = (float)SGX * fmaf(fj, tv.D[0], pr[0]);
                const float qy = (float)SGY * fmaf(fj, tv.D[1], pr[1]);
                const float qz = (float)SGZ * fmaf(fj, tv.D[2], pr[2]);
                const float flx = floorf(qx), fly = floorf(qy), flz = floorf(qz);
                float f0 = qx - flx, f1 = qy - fly, f2 = qz - flz;
                int off = (int)flx * STX + (int)fly * STY + (int)flz * STZ;
                unsigned sb = acc_b + 4u * (unsigned)off;                  // shared byte address of the floor cell
                const float2 yv2 = make_float2(yv, yv);
                for (int j = jw0; j < jw1; ++j) {
                    const bool act = k < span;
                    // lane l+1 sits in lane l's z cell iff its own fraction says so (no shuffle needed):
                    // mirrored z decreases (SGZ < 0) or increases (SGZ > 0) by W_z per lane
                    const bool dup = act && ((SGZ > 0) ? (f2 >= dupthr) : (f2 < 1.f - dupthr));
                    // corner weights as the pairs (x0y0, x1y0) and (x0y1, x1y1) of the packed FMAs
                    const float wx1 = f0 * yv, wx0 = yv - wx1;
                    const float2 wx = make_float2(wx0, wx1);
                    const float2 wy1 = __fmul2_rn(wx, make_float2(f1, f1));           // (x0y1, x1y1)
                    const float2 wy0 = __ffma2_rn(wy1, make_float2(-1.f, -1.f), wx);  // (x0y0, x1y0)
                    const float wz0 = 1.f - f2;
                    if (__builtin_expect(!__any_sync(FULL, dup), 1)) {
                        float* const se = (float*)__cvta_shared_to_generic(act ? sb : ((sb & 0x7cu) | dummy_b));
                        rmw4<O10, O01, O11>(se, wy0, wy1, wz0);
                        __syncwarp();
#ifndef TOMO_PROBE_HALF_RMW          // timing probe only: drops the z-ceil plane
                        rmw4<O10, O01, O11>(se + OZ, wy0, wy1, f2);
                        __syncwarp();
#endif
                    } else {
                        // rare (W_z < 1 makes two adjacent lanes share a z cell ~ once per 1/(1-W_z) samples):
                        // lanes flagged dup go in a second pass
#pragma unroll 1
                        for (int pass = 0; pass < 2; ++pass) {
                            float* const se = (float*)__cvta_shared_to_generic(
                                (act && (dup == (pass == 1))) ? sb : ((sb & 0x7cu) | dummy_b));
                            rmw4<O10, O01, O11>(se, wy0, wy1, wz0);
                            __syncwarp();
                            rmw4<O10, O01, O11>(se + OZ, wy0, wy1, f2);
                            __syncwarp();
                        }
                    }
                    // advance one sample: one-sided carries, the strides folded into one address increment
                    ++k;
                    f0 += df0; f1 += df1; f2 += df2;
                    sb += stepoff4;
                    if (f0 >= 1.0f) { f0 -= 1.0f; sb += 4u * (unsigned)STX; }
                    if (f1 >= 1.0f) { f1 -= 1.0f; sb += 4u * (unsigned)STY; }
                    if (f2 >= 1.0f) { f2 -= 1.0f; sb += 4u * (unsigned)STZ; }
                }
            }
        }
        __syncthreads();
    }
}

// Constants of the next view (from `view` on) that has rays crossing the tile, or tv.view = -1.  One thread, float64.
__device__ __noinline__ void tile_next_view(const BackArgs& A, int view, const int org[3], TileView& tv)
{
    constexpr int TX = TOMO_BT_X, TY = TOMO_BT_Y, TZ = TOMO_BT_Z;
    // a sample is ours iff its floor cell lies in [0, T] per axis, i.e. 0 <= p_s < T+1 (smem coordinates)
    const double hw[3] = {0.5 * (TX + 1), 0.5 * (TY + 1), 0.5 * (TZ + 1)};
    for (; view < A.n_proj; ++view) {
        const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
        tv.ncol = (int)V[V_NCOL];
        if (tv.ncol == 0) continue;                      // outside the scatter envelope: the gather kernel adds it
        if (A.skip_separable && V[V_SEP] != 0.0) continue;   // untilted view: the separable adjoint adds it
        // lattice coordinates of the active box: centre +- sum |Linv| * half widths (exact for a linear map)
        double B[3], cc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            B[a] = V[V_P00 + a] - (double)org[a];        // p_s = B + ix U + iz W + j D
            cc[a] = hw[a] - B[a];
        }
        double lc[3], lr[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lc[k] = V[V_LINV + 3 * k] * cc[0] + V[V_LINV + 3 * k + 1] * cc[1] + V[V_LINV + 3 * k + 2] * cc[2];
            lr[k] = fabs(V[V_LINV + 3 * k]) * hw[0] + fabs(V[V_LINV + 3 * k + 1]) * hw[1] + fabs(V[V_LINV + 3 * k + 2]) * hw[2] + 1e-3;
        }
        tv.ix_lo = max(0, (int)ceil(fmax(lc[0] - lr[0], -1.0)));
        tv.ix_hi = min(A.ndx - 1, (int)floor(fmin(lc[0] + lr[0], 2.0e9)));
        tv.iz_lo = max(0, (int)ceil(fmax(lc[1] - lr[1], -1.0)));
        tv.iz_hi = min(A.ndz - 1, (int)floor(fmin(lc[1] + lr[1], 2.0e9)));
        const int nsamp = (int)V[V_N];
        const int j_lo = max(0, (int)ceil(fmax(lc[2] - lr[2], -1.0))), j_hi = min(nsamp - 1, (int)floor(fmin(lc[2] + lr[2], 2.0e9)));
        if (tv.ix_lo > tv.ix_hi || tv.iz_lo > tv.iz_hi || j_lo > j_hi) continue;   // nothing of this view crosses the tile
        tv.jc = (j_lo + j_hi) / 2;
        tv.djmin = j_lo - tv.jc;
        tv.djmax = j_hi - tv.jc + 1;
        tv.di_step = 0; tv.sgn = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double d = V[V_D + a], ad = fabs(d), ai = floor(ad);
            tv.Bc[a] = (float)(B[a] + (double)tv.ix_lo * V[V_U + a] + (double)tv.iz_lo * V[V_W + a] + (double)tv.jc * d);
            tv.U[a] = (float)V[V_U + a]; tv.W[a] = (float)V[V_W + a]; tv.D[a] = (float)d;
            tv.invD[a] = (float)V[V_INVD + a];
            tv.df[a] = (float)(ad - ai);
            const int sg = (d < 0.0) ? -1 : 1;
            if (sg < 0) tv.sgn |= 4 >> a;
            tv.di_step += (int)ai * sg * ((a == 0) ? TSY * TSZ : (a == 1) ? TSZ : 1);
        }
        // adjacent lanes share a z cell iff frac >= W_z (mirrored: frac < 1 - W_z); margin for float32 marching
        tv.dupthr = (float)fabs(V[V_W + 2]) - 2e-4f;
        tv.view = view;
        return;
    }
    tv.view = -1;
}

__global__ void __launch_bounds__(TNW * 32)
adjoint_tile_kernel(const BackArgs A, const int ntx, const int nty, const int ntz)
{
    constexpr int TX = TOMO_BT_X, TY = TOMO_BT_Y, TZ = TOMO_BT_Z;
    extern __shared__ __align__(128) float smem_raw[];   // guard | [TSX][TSY][TSZ] | guard
    float* const acc = smem_raw + TGUARD;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int bb = blockIdx.x;
    const int tz = bb % ntz; bb /= ntz;
    const int ty = bb % nty;
    const int tx = bb / nty + A.x_begin / TX;             // slab launches start at a tile boundary
    const int org[3] = {tx * TX - 1, ty * TY - 1, tz * TZ - 1};   // voxel coordinate of smem cell 0 (the low ghost cell)
    // Nothing to scatter (every view of this table is left to the separable or the gather kernel): initialise the tile and leave,
    // instead of walking the view table in every block
    {
        int need = 0;
        for (int v = threadIdx.x; v < A.n_proj; v += TNW * 32) {
            const double* __restrict__ V = A.views + (size_t)v * TOMO_VIEW_STRIDE;
            need |= (V[V_NCOL] != 0.0) && !(A.skip_separable && V[V_SEP] != 0.0);
        }
        if (!__syncthreads_or(need)) {
            if (!A.accumulate)
                for (int i = threadIdx.x; i < TX * TY * 32; i += TNW * 32) {
                    const int zz = i & 31, yy = (i >> 5) % TY, xx = (i >> 5) / TY;
                    const int x = tx * TX + xx, y = ty * TY + yy, z = tz * TZ + zz;
                    if (zz < TZ && x < A.x_end && y < A.ny && z < A.nz) A.vol[((size_t)x * A.ny + y) * A.nz + z] = 0.f;
                }
            return;
        }
    }
    for (int i = threadIdx.x; i < TSMEM_FLOATS; i += TNW * 32) smem_raw[i] = 0.f;
    __syncthreads();

    // Per-view constants: computed once per (block, view) by one thread -- for the NEXT view, while the block marches the current
    // one -- and handed over through shared memory (they used to be recomputed in float64 by every thread).
    __shared__ TileView tvs[2];
    if (threadIdx.x == 0) tile_next_view(A, 0, org, tvs[0]);
    __syncthreads();
    const size_t n_det = (size_t)A.ndx * A.ndz;
    for (int buf = 0; ; buf ^= 1) {
        const TileView& tv = tvs[buf];       // rewritten two views from now, after the barriers that end this view's classes
        if (tv.view < 0) break;              // block-uniform
        if (threadIdx.x == TNW * 32 - 1) tile_next_view(A, tv.view + 1, org, tvs[buf ^ 1]);
        __syncwarp();
        const float* __restrict__ P = A.proj + (size_t)tv.view * n_det;
        switch (tv.sgn) {                                                       // block-uniform
            case 0: tile_march_view< 1,  1,  1>(acc, tv, P, A.ndz, lane, warp); break;
            case 1: tile_march_view< 1,  1, -1>(acc, tv, P, A.ndz, lane, warp); break;
            case 2: tile_march_view< 1, -1,  1>(acc, tv, P, A.ndz, lane, warp); break;
            case 3: tile_march_view< 1, -1, -1>(acc, tv, P, A.ndz, lane, warp); break;
            case 4: tile_march_view<-1,  1,  1>(acc, tv, P, A.ndz, lane, warp); break;
            case 5: tile_march_view<-1,  1, -1>(acc, tv, P, A.ndz, lane, warp); break;
            case 6: tile_march_view<-1, -1,  1>(acc, tv, P, A.ndz, lane, warp); break;
            default: tile_march_view<-1, -1, -1>(acc, tv, P, A.ndz, lane, warp); break;
        }
    }
    __syncthreads();
    // write the interior of the tile once
    for (int i = threadIdx.x; i < TX * TY * 32; i += TNW * 32) {
        const int zz = i & 31, yy = (i >> 5) % TY, xx = (i >> 5) / TY;
        const int x = tx * TX + xx, y = ty * TY + yy, z = tz * TZ + zz;
        if (zz < TZ && x < A.x_end && y < A.ny && z < A.nz) {
            const float v = acc[((xx + 1) * TSY + (yy + 1)) * TSZ + zz + 1];
            const size_t vi = ((size_t)x * A.ny + y) * A.nz + z;
            A.vol[vi] = A.accumulate ? A.vol[vi] + v : v;
        }
    }
}

// Voxel-driven forward splat + gradient image (src/vox_wt_grad.f90:1-55).  One thread per (voxel, view); lanes along z.
// A scatter: FIXED = false adds float32 contributions with atomicAdd (summation order varies between runs); FIXED = true
// adds them as 64-bit fixed-point integers (atomicAdd on unsigned long long: integer addition is associative, so the sums
// are bitwise reproducible) that splat_fixed_finalize_kernel converts back.  blockIdx.x = view * nzb + z block, so the view
// count is not limited by the 65535 cap of grid.y.
struct SplatFixed {
    unsigned long long* det;      // [n_proj][n_det]
    unsigned long long* grad;     // [n_proj][6][n_det], nullable
    const unsigned* maxabs;       // float bits of max |vol| (splat_maxabs_kernel)
    double terms;                 // bound on the number of contributions one detector pixel can receive
    double ext;                   // bound on |voxel centre coordinate| summed over the axes (for the derivative magnitudes)
};

// scale of the fixed-point sums: |contribution| <= m, at most `terms` of them per pixel -> |sum| * scale <= 2^62
__device__ __forceinline__ double splat_scale(double m, double terms)
{
    return (m > 0.0) ? 4611686018427387904.0 / (m * terms) : 1.0;
}
// bound on |g0|, |g2| of a view (rows of derivative_rigid, voxel_utilities.py:23-48: rotation rows applied to a centre / to t)
__device__ __forceinline__ double splat_gmax(const double* __restrict__ V, double ext)
{
    return 2.0 * (ext + fabs(V[V_VTR + 0]) + fabs(V[V_VTR + 1]) + fabs(V[V_VTR + 2])) + 2.0;
}

__global__ void splat_maxabs_kernel(const float* __restrict__ vol, size_t n, unsigned* __restrict__ out)
{
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float a = fabsf(vol[i]);
        m = (a > m && a < 3.0e38f) ? a : m;               // Inf / NaN voxels do not set the scale
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));          // non-negative floats order like their bits
}

template <bool FIXED>
__global__ void __launch_bounds__(BZ * BY * BX)
voxel_splat_kernel(const BackArgs A, const float* __restrict__ vol, float* __restrict__ det, float* __restrict__ grad,
                   const SplatFixed F, int nzb)
{
    const int view = blockIdx.x / nzb;
    const int z = (blockIdx.x % nzb) * BZ + threadIdx.x;
    const int y = blockIdx.y * BY + threadIdx.y;
    const int x = blockIdx.z * BX + threadIdx.z;
    if (x >= A.nx || y >= A.ny || z >= A.nz) return;
    const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
    const float rec = vol[((size_t)x * A.ny + y) * A.nz + z];
    const double cx = A.vox0[0] + x * A.vpix[0], cy = A.vox0[1] + y * A.vpix[1], cz = A.vox0[2] + z * A.vpix[2];
    const double ux = V[V_VROT + 0] * cx + V[V_VROT + 1] * cy + V[V_VROT + 2] * cz + V[V_VTR + 0] - V[V_SORG + 0];
    const double uz = V[V_VROT + 6] * cx + V[V_VROT + 7] * cy + V[V_VROT + 8] * cz + V[V_VTR + 2] - V[V_SORG + 1];
    const double flx = floor(ux), flz = floor(uz);
    const float ax = (float)(ux - flx), az = (float)(uz - flz);       // alpha_x, alpha_z are float32 in the reference
    const int fx = (int)fmin(fmax(flx, -2.0), 1.0e9), fz = (int)fmin(fmax(flz, -2.0), 1.0e9);
    const size_t n_det = (size_t)A.ndx * A.ndz;
    const bool want_grad = FIXED ? (F.grad != nullptr) : (grad != nullptr);
    float g0[6], g2[6];
    if (want_grad) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double* q0 = V + V_SPL + (k * 2 + 0) * 4;
            const double* q2 = V + V_SPL + (k * 2 + 1) * 4;
            g0[k] = (float)(q0[0] * cx + q0[1] * cy + q0[2] * cz + q0[3]);
            g2[k] = (float)(q2[0] * cx + q2[1] * cy + q2[2] * cz + q2[3]);
        }
    }
    double sc = 0.0, scg = 0.0;
    if (FIXED) {
        const double m = (double)__uint_as_float(*F.maxabs);
        sc = splat_scale(m, F.terms);
        scg = splat_scale(m * splat_gmax(V, F.ext), F.terms);
    }
    // taps (fx,fz), (fx+1,fz), (fx,fz+1), (fx+1,fz+1): weights and d/dx', d/dz' factors of vox_wt_grad.f90:25-50
    const int   tx[4] = {fx, fx + 1, fx, fx + 1}, tzz[4] = {fz, fz, fz + 1, fz + 1};
    const float w[4]  = {(1.f - ax) * (1.f - az), ax * (1.f - az), (1.f - ax) * az, ax * az};
    const float G0[4] = {(1.f - az), -(1.f - az), az, -az};
    const float G2[4] = {(1.f - ax), ax, -(1.f - ax), -ax};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (tx[t] < 0 || tx[t] >= A.ndx || tzz[t] < 0 || tzz[t] >= A.ndz) continue;
        const size_t di = (size_t)tzz[t] * A.ndx + tx[t];
        const float c = rec * w[t];
        if (FIXED) atomicAdd(F.det + (size_t)view * n_det + di, (unsigned long long)__double2ll_rn((double)c * sc));
        else       atomicAdd(det + (size_t)view * n_det + di, c);
        if (want_grad) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const float cg = g0[k] * G0[t] * rec + g2[k] * G2[t] * rec;
                if (FIXED) atomicAdd(F.grad + ((size_t)view * 6 + k) * n_det + di, (unsigned long long)__double2ll_rn((double)cg * scg));
                else       atomicAdd(grad + ((size_t)view * 6 + k) * n_det + di, cg);
            }
        }
    }
}

__global__ void splat_fixed_finalize_kernel(const BackArgs A, const SplatFixed F, float* __restrict__ det, float* __restrict__ grad)
{
    const size_t n_det = (size_t)A.ndx * A.ndz, n_img = (size_t)A.n_proj * (F.grad ? 7 : 1);
    const double m = (double)__uint_as_float(*F.maxabs);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_img * n_det; i += (size_t)gridDim.x * blockDim.x) {
        const size_t img = i / n_det, px = i % n_det;
        if (img < (size_t)A.n_proj) {
            det[img * n_det + px] = (float)((double)(long long)F.det[img * n_det + px] / splat_scale(m, F.terms));
        } else {
            const size_t gi = img - A.n_proj, view = gi / 6;
            const double scg = splat_scale(m * splat_gmax(A.views + view * TOMO_VIEW_STRIDE, F.ext), F.terms);
            grad[gi * n_det + px] = (float)((double)(long long)F.grad[gi * n_det + px] / scg);
        }
    }
}

// Transpose of the splat: vol[v] (+)= sum_views sum_taps w(v, tap) * det[view][tap] with the splat's own weights (the matrix
// bilinear_sparse emits, src/vox_wt_grad.f90:58-112, applied transposed).  A gather: one thread per voxel, no atomics.
__global__ void __launch_bounds__(BZ * BY * BX)
voxel_splat_adjoint_kernel(const BackArgs A)
{
    const int z = blockIdx.x * BZ + threadIdx.x;
    const int y = blockIdx.y * BY + threadIdx.y;
    const int x = blockIdx.z * BX + threadIdx.z;
    if (x >= A.nx || y >= A.ny || z >= A.nz) return;
    const size_t n_det = (size_t)A.ndx * A.ndz;
    const double cx = A.vox0[0] + x * A.vpix[0], cy = A.vox0[1] + y * A.vpix[1], cz = A.vox0[2] + z * A.vpix[2];
    float acc = 0.f;
    for (int view = 0; view < A.n_proj; ++view) {
        const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
        const float* __restrict__ P = A.proj + (size_t)view * n_det;              // [ndz][ndx], x fastest
        const double ux = V[V_VROT + 0] * cx + V[V_VROT + 1] * cy + V[V_VROT + 2] * cz + V[V_VTR + 0] - V[V_SORG + 0];
        const double uz = V[V_VROT + 6] * cx + V[V_VROT + 7] * cy + V[V_VROT + 8] * cz + V[V_VTR + 2] - V[V_SORG + 1];
        const double flx = floor(ux), flz = floor(uz);
        const float ax = (float)(ux - flx), az = (float)(uz - flz);
        const int fx = (int)fmin(fmax(flx, -2.0), 1.0e9), fz = (int)fmin(fmax(flz, -2.0), 1.0e9);
        const bool x0 = (fx >= 0 && fx < A.ndx), x1 = (fx + 1 >= 0 && fx + 1 < A.ndx);
        const bool z0 = (fz >= 0 && fz < A.ndz), z1 = (fz + 1 >= 0 && fz + 1 < A.ndz);
        const float* __restrict__ c = P + (ptrdiff_t)fz * A.ndx + fx;
        float v = 0.f;
        if (x0 && z0) v = fmaf(__ldg(c), (1.f - ax) * (1.f - az), v);
        if (x1 && z0) v = fmaf(__ldg(c + 1), ax * (1.f - az), v);
        if (x0 && z1) v = fmaf(__ldg(c + A.ndx), (1.f - ax) * az, v);
        if (x1 && z1) v = fmaf(__ldg(c + A.ndx + 1), ax * az, v);
        acc += v;
    }
    const size_t vi = ((size_t)x * A.ny + y) * A.nz + z;
    A.vol[vi] = A.accumulate ? A.vol[vi] + acc : acc;
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
    A->n_proj = n_proj; A->accumulate = accumulate; A->only_uncoloured = 0; A->skip_separable = 0; A->only_vbig = 0;
    A->x_begin = 0; A->x_end = g->nx;
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

size_t tomo_back_separable_workspace_bytes(const TomoGeom* g, int n_proj);
int tomo_back_separable_launch(const TomoGeom* g, const void* views, int n_proj, const float* proj, float* vol,
                               int accumulate, void* workspace, int x_begin, int x_end, void* stream);

extern "C" size_t tomo_back_adjoint_workspace_bytes(const TomoGeom* g, int n_proj)
{
    if (!g || n_proj <= 0) return 0;
    return tomo_back_separable_workspace_bytes(g, n_proj);
}

// kinds: 0 = unknown (every kernel is launched and leaves at once when the table holds no view for it), otherwise the
// TOMO_KINDS_* mask of tomo_views_kinds() for this table (or any table it is a part of): kernels without views are not launched.
static int back_adjoint_impl(const TomoGeom* g, const void* views, int n_proj, int kinds, const float* proj, float* vol,
                             int accumulate, void* workspace, size_t workspace_bytes, int x_begin, int x_end, void* stream)
{
    BackArgs A; dim3 grid;
    if (int e = fill_back(g, views, n_proj, proj, vol, accumulate, &A, &grid)) return e;
    constexpr int TX = TOMO_BT_X, TY = TOMO_BT_Y, TZ = TOMO_BT_Z;
    if (x_begin < 0 || x_end > g->nx || x_begin >= x_end || x_begin % TX != 0 || (x_end % TX != 0 && x_end != g->nx)) {
        tomo_set_error("tomo_back_adjoint_slab: [x_begin, x_end) must be a non-empty range of whole tile rows "
                       "(multiples of tomo_back_adjoint_slab_granularity(); x_end may also be nx)");
        return TOMO_E_ARG;
    }
    const bool sep = workspace != nullptr;
    if (sep && workspace_bytes < tomo_back_separable_workspace_bytes(g, n_proj)) {
        tomo_set_error("tomo_back_adjoint_ws: workspace too small (see tomo_back_adjoint_workspace_bytes)");
        return TOMO_E_WORKSPACE;
    }
    const bool known = (kinds & TOMO_KINDS_KNOWN) != 0;
    A.skip_separable = sep ? 1 : 0;
    A.x_begin = x_begin; A.x_end = x_end;
    grid.z = (unsigned)((x_end - x_begin + BX - 1) / BX);
    const int ntx = (x_end - x_begin + TX - 1) / TX, nty = (g->ny + TY - 1) / TY, ntz = (g->nz + TZ - 1) / TZ;
    const size_t smem = sizeof(float) * TSMEM_FLOATS;
    const double nblocks = (double)ntx * nty * ntz;
    if (nblocks >= 2147483647.0) { tomo_set_error("tomo_back_adjoint: too many tiles"); return TOMO_E_RANGE; }
    // tilted views inside the scatter envelope (and, without a workspace, the untilted ones): the tile kernel.  It writes
    // every voxel of the slab, so the kernels after it accumulate; when it is not needed the first kernel launched
    // inherits the caller's accumulate flag.
    const bool want_tile = !known || (kinds & TOMO_KINDS_TILE) || (!sep && (kinds & TOMO_KINDS_SEPARABLE));
    const bool want_gather = !known || (kinds & TOMO_KINDS_UNCOLOURED);
    const bool want_sep = sep && (!known || (kinds & TOMO_KINDS_SEPARABLE));
    if (want_tile || !(want_gather || want_sep)) {
        cudaError_t ce = cudaFuncSetAttribute(adjoint_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (int e = tomo_check_cuda(ce, "cudaFuncSetAttribute(adjoint_tile_kernel)")) return e;
        adjoint_tile_kernel<<<(unsigned)(ntx * nty * ntz), TNW * 32, smem, (cudaStream_t)stream>>>(A, ntx, nty, ntz);
        if (int e = tomo_check_cuda(cudaGetLastError(), "adjoint_tile_kernel")) return e;
        accumulate = 1;
    }
    // views outside the scatter envelope (rays nearly parallel to z; none for tomographic poses): the gather kernel
    if (want_gather) {
        A.only_uncoloured = 1; A.accumulate = accumulate;
        adjoint_gather_kernel<<<grid, dim3(BZ, BY, BX), 0, (cudaStream_t)stream>>>(A);
        if (int e = tomo_check_cuda(cudaGetLastError(), "adjoint_gather_kernel(uncoloured)")) return e;
        accumulate = 1;
    }
    // untilted views: separable adjoint (needs the Yz workspace, filled by the launch of the slab that starts at x = 0)
    if (want_sep) return tomo_back_separable_launch(g, views, n_proj, proj, vol, accumulate, workspace, x_begin, x_end, stream);
    return 0;
}

extern "C" int tomo_back_adjoint_slab_granularity(void) { return TOMO_BT_X; }

extern "C" int tomo_back_adjoint_slab(const TomoGeom* g, const void* views, int n_proj, int kinds, const float* proj, float* vol,
                                      int accumulate, void* workspace, size_t workspace_bytes, int x_begin, int x_end, void* stream)
{
    return back_adjoint_impl(g, views, n_proj, kinds, proj, vol, accumulate, workspace, workspace_bytes, x_begin, x_end, stream);
}

extern "C" int tomo_back_adjoint(const TomoGeom* g, const void* views, int n_proj,
                                 const float* proj, float* vol, int accumulate, void* stream)
{
    return back_adjoint_impl(g, views, n_proj, 0, proj, vol, accumulate, nullptr, 0, 0, g ? g->nx : 0, stream);
}

extern "C" int tomo_back_adjoint_ws(const TomoGeom* g, const void* views, int n_proj, const float* proj, float* vol,
                                    int accumulate, void* workspace, size_t workspace_bytes, void* stream)
{
    if (!workspace) { tomo_set_error("tomo_back_adjoint_ws: workspace is NULL"); return TOMO_E_ARG; }
    return back_adjoint_impl(g, views, n_proj, 0, proj, vol, accumulate, workspace, workspace_bytes, 0, g ? g->nx : 0, stream);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point table (the library does not link libcuda).
typedef CUresult (*TomoEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TomoEncodeTiled tomo_encode_tiled()
{
    static TomoEncodeTiled fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) p = nullptr;
        return (TomoEncodeTiled)p;
    }();
    return fn;
}

// 3-D tensor map {z', x', view} over the projections with the box the TMA kernel stages; false when the layout
// does not meet the TMA constraints (16-byte aligned base and row pitch) or the detector is smaller than the box.
static bool voxback_tensor_map(const TomoGeom* g, int n_proj, const float* proj, CUtensorMap* map)
{
    if (g->ndz % 4 != 0 || ((uintptr_t)proj & 15u) != 0 || g->ndz < VTZ || g->ndx < VTX) return false;
    TomoEncodeTiled enc = tomo_encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)g->ndz, (cuuint64_t)g->ndx, (cuuint64_t)n_proj};
    const cuuint64_t strides[2] = {(cuuint64_t)g->ndz * sizeof(float), (cuuint64_t)g->ndz * g->ndx * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)VTZ, (cuuint32_t)VTX, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)proj, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int tomo_back_voxel_bilinear(const TomoGeom* g, const void* views, int n_proj,
                                        const double origin[3], const float* proj, float* vol,
                                        int accumulate, void* stream)
{
    BackArgs A; dim3 grid;
    if (!origin) { tomo_set_error("tomo_back_voxel_bilinear: origin is NULL"); return TOMO_E_ARG; }
    if (int e = fill_back(g, views, n_proj, proj, vol, accumulate, &A, &grid)) return e;
    for (int a = 0; a < 3; ++a) A.origin[a] = origin[a];
    CUtensorMap map;
    const dim3 bricks((g->nz + VBZ - 1) / VBZ, (g->ny + VBY - 1) / VBY, (g->nx + VBX - 1) / VBX);
    if (bricks.y <= 65535u && bricks.z <= 65535u && voxback_tensor_map(g, n_proj, proj, &map)) {
        // views whose brick footprint fits the staged box (V_VBOK), then the rest through the plain kernel
        voxel_bilinear_tma_kernel<<<bricks, VB_WARPS * 32, 0, (cudaStream_t)stream>>>(A, map);
        if (int e = tomo_check_cuda(cudaGetLastError(), "voxel_bilinear_tma_kernel")) return e;
        A.accumulate = 1; A.only_vbig = 1;
    }
    voxel_bilinear_kernel<<<grid, dim3(BZ, BY, BX), 0, (cudaStream_t)stream>>>(A);
    return tomo_check_cuda(cudaGetLastError(), "voxel_bilinear_kernel");
}

static int splat_launch(const TomoGeom* g, const void* views, int n_proj, const float* vol, float* det, float* grad,
                        void* workspace, size_t workspace_bytes, bool fixed, void* stream)
{
    BackArgs A; dim3 grid;
    if (!det) { tomo_set_error("tomo_voxel_splat: det_dev is NULL"); return TOMO_E_ARG; }
    if (int e = fill_back(g, views, n_proj, vol, det, 0, &A, &grid)) return e;
    const size_t n_det = (size_t)g->ndx * g->ndz, n_vox = (size_t)g->nx * g->ny * g->nz;
    const int nzb = (g->nz + BZ - 1) / BZ;
    if ((double)nzb * n_proj >= 2147483647.0) { tomo_set_error("tomo_voxel_splat: too many blocks for one launch"); return TOMO_E_RANGE; }
    grid.x = (unsigned)(nzb * n_proj);
    cudaStream_t st = (cudaStream_t)stream;
    SplatFixed F = {nullptr, nullptr, nullptr, 0.0, 0.0};
    if (!fixed) {
        cudaError_t ce = cudaMemsetAsync(det, 0, sizeof(float) * n_det * n_proj, st);
        if (ce == cudaSuccess && grad) ce = cudaMemsetAsync(grad, 0, sizeof(float) * 6 * n_det * n_proj, st);
        if (int e = tomo_check_cuda(ce, "tomo_voxel_splat: cudaMemsetAsync")) return e;
        voxel_splat_kernel<false><<<grid, dim3(BZ, BY, BX), 0, st>>>(A, vol, det, grad, F, nzb);
        return tomo_check_cuda(cudaGetLastError(), "voxel_splat_kernel");
    }
    const size_t need = tomo_voxel_splat_workspace_bytes(g, n_proj, grad != nullptr);
    if (!workspace || workspace_bytes < need) {
        tomo_set_error("tomo_voxel_splat_deterministic: workspace too small (see tomo_voxel_splat_workspace_bytes)");
        return TOMO_E_WORKSPACE;
    }
    cudaError_t ce = cudaMemsetAsync(workspace, 0, need, st);
    if (int e = tomo_check_cuda(ce, "tomo_voxel_splat_deterministic: cudaMemsetAsync")) return e;
    unsigned* maxabs = (unsigned*)workspace;                             // 16 bytes reserved, then the integer images
    F.det = (unsigned long long*)((char*)workspace + 16);
    F.grad = grad ? F.det + n_det * n_proj : nullptr;
    F.maxabs = maxabs;
    // a detector pixel collects the voxels of a tube of cross-section <= 2 x 2 pixels through the volume, 4 taps each
    F.terms = 16.0 * ((double)g->nx + g->ny + g->nz) + 16.0;
    F.ext = 0.0;
    for (int a = 0; a < 3; ++a) {
        const double n = (a == 0) ? g->nx : (a == 1) ? g->ny : g->nz;
        const double lo = std::fabs(g->vox_origin[a]), hi = std::fabs(g->vox_origin[a] + (n - 1.0) * g->vox_pix[a]);
        F.ext += lo > hi ? lo : hi;
    }
    splat_maxabs_kernel<<<148 * 4, 256, 0, st>>>(vol, n_vox, maxabs);
    voxel_splat_kernel<true><<<grid, dim3(BZ, BY, BX), 0, st>>>(A, vol, nullptr, nullptr, F, nzb);
    splat_fixed_finalize_kernel<<<148 * 8, 256, 0, st>>>(A, F, det, grad);
    return tomo_check_cuda(cudaGetLastError(), "voxel_splat_kernel<fixed>");
}

extern "C" size_t tomo_voxel_splat_workspace_bytes(const TomoGeom* g, int n_proj, int with_grad)
{
    if (!g || n_proj <= 0) return 0;
    return 16 + sizeof(unsigned long long) * (size_t)g->ndx * g->ndz * (size_t)n_proj * (with_grad ? 7 : 1);
}

extern "C" int tomo_voxel_splat(const TomoGeom* g, const void* views, int n_proj, const float* vol,
                                float* det, float* grad, void* stream)
{
    return splat_launch(g, views, n_proj, vol, det, grad, nullptr, 0, false, stream);
}

extern "C" int tomo_voxel_splat_deterministic(const TomoGeom* g, const void* views, int n_proj, const float* vol,
                                              float* det, float* grad, void* workspace, size_t workspace_bytes, void* stream)
{
    return splat_launch(g, views, n_proj, vol, det, grad, workspace, workspace_bytes, true, stream);
}

extern "C" int tomo_voxel_splat_adjoint(const TomoGeom* g, const void* views, int n_proj, const float* det,
                                        float* vol, int accumulate, void* stream)
{
    BackArgs A; dim3 grid;
    if (int e = fill_back(g, views, n_proj, det, vol, accumulate, &A, &grid)) return e;
    voxel_splat_adjoint_kernel<<<grid, dim3(BZ, BY, BX), 0, (cudaStream_t)stream>>>(A);
    return tomo_check_cuda(cudaGetLastError(), "voxel_splat_adjoint_kernel");
}
