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
    float acc = 0.f;
    for (int view = 0; view < A.n_proj; ++view)
        acc += adjoint_gather_view(A.proj + (size_t)view * n_det, A.views + (size_t)view * TOMO_VIEW_STRIDE,
                                   A.ndx, A.ndz, x, y, z);
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

}  // namespace

extern "C" void tomo_set_error(const char* msg);
int tomo_check_cuda(cudaError_t e, const char* what);

static int fill_back(const TomoGeom* g, const void* views, int n_proj, const float* proj, float* vol,
                     int accumulate, BackArgs* A, dim3* grid)
{
    if (!g || !views || !proj || !vol || n_proj <= 0) { tomo_set_error("backprojector: null pointer or n_proj <= 0"); return TOMO_E_ARG; }
    A->proj = proj; A->views = (const double*)views; A->vol = vol;
    A->nx = g->nx; A->ny = g->ny; A->nz = g->nz; A->ndx = g->ndx; A->ndz = g->ndz;
    A->n_proj = n_proj; A->accumulate = accumulate;
    for (int a = 0; a < 3; ++a) { A->origin[a] = 0.0; A->vox0[a] = g->vox_origin[a]; A->vpix[a] = g->vox_pix[a]; }
    *grid = dim3((g->nz + BZ - 1) / BZ, (g->ny + BY - 1) / BY, (g->nx + BX - 1) / BX);
    if (grid->y > 65535u || grid->z > 65535u) { tomo_set_error("backprojector: volume too large for the launch grid"); return TOMO_E_RANGE; }
    return 0;
}

extern "C" int tomo_back_adjoint(const TomoGeom* g, const void* views, int n_proj,
                                 const float* proj, float* vol, int accumulate, void* stream)
{
    BackArgs A; dim3 grid;
    if (int e = fill_back(g, views, n_proj, proj, vol, accumulate, &A, &grid)) return e;
    adjoint_gather_kernel<<<grid, dim3(BZ, BY, BX), 0, (cudaStream_t)stream>>>(A);
    return tomo_check_cuda(cudaGetLastError(), "adjoint_gather_kernel");
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
