// tv_kernels.cu -- total-variation proximal step (dual FISTA of utilities/tv_denoise.py:98-170), sm_100a.
//
// Two fused, HBM-bound stencil kernels per dual iteration instead of the reference's ~12 numpy passes:
//   tv_dual_error :  err = weight * div(p) - im                               (tv_denoise.py:147; div :20-31)
//   tv_dual_update:  g = gradient(err) / (factor * weight); aux += g;         (:148-150; gradient :34-59)
//                    tmp = aux / max(|aux|_2, 1);                             (_projector_on_dual :67-74)
//                    aux <- (1 + tf) * tmp - tf * gim;  gim <- tmp            (:153-154)
// Fields: p, aux, gim are [3][nx][ny][nz] float32 (component-major, z fastest); im, err are [nx][ny][nz].
// One thread per voxel, lanes along z: every access is a coalesced row read; the +-1 neighbours along x and y are
// other rows of the same kernel's working set and come from L2.
#include <cuda_runtime.h>
#include "tomo_common.h"

namespace {

struct TvArgs { int nx, ny, nz; };

__device__ __forceinline__ bool tv_index(const TvArgs& A, int& x, int& y, int& z)
{
    z = blockIdx.x * blockDim.x + threadIdx.x;
    y = blockIdx.y * blockDim.y + threadIdx.y;
    x = blockIdx.z;
    return z < A.nz && y < A.ny;
}

// div(p)[i] = sum_d ( [i_d < n_d - 1] p_d[i] - [i_d > 0] p_d[i - e_d] )      (tv_denoise.py:24-30)
__device__ __forceinline__ float tv_div(const float* __restrict__ p, const TvArgs& A, int x, int y, int z)
{
    const size_t sy = (size_t)A.nz, sx = (size_t)A.ny * A.nz, N = sx * A.nx;
    const size_t i = (size_t)x * sx + (size_t)y * sy + z;
    float r = 0.f;
    if (x < A.nx - 1) r += p[i];
    if (x > 0)        r -= p[i - sx];
    if (y < A.ny - 1) r += p[N + i];
    if (y > 0)        r -= p[N + i - sy];
    if (z < A.nz - 1) r += p[2 * N + i];
    if (z > 0)        r -= p[2 * N + i - 1];
    return r;
}

__global__ void __launch_bounds__(256)
tv_dual_error_kernel(const TvArgs A, const float weight, const float* __restrict__ p, const float* __restrict__ im,
                     float* __restrict__ err)
{
    int x, y, z;
    if (!tv_index(A, x, y, z)) return;
    const size_t i = ((size_t)x * A.ny + y) * A.nz + z;
    err[i] = weight * tv_div(p, A, x, y, z) - im[i];
}

__global__ void __launch_bounds__(256)
tv_dual_update_kernel(const TvArgs A, const float inv_fw, const float tf, const float* __restrict__ err,
                      float* __restrict__ aux, float* __restrict__ gim)
{
    int x, y, z;
    if (!tv_index(A, x, y, z)) return;
    const size_t sy = (size_t)A.nz, sx = (size_t)A.ny * A.nz, N = sx * A.nx;
    const size_t i = (size_t)x * sx + (size_t)y * sy + z;
    const float e = err[i];
    // forward differences, zero on the last index of each axis (tv_denoise.py:53-57)
    const float g0 = (x < A.nx - 1) ? err[i + sx] - e : 0.f;
    const float g1 = (y < A.ny - 1) ? err[i + sy] - e : 0.f;
    const float g2 = (z < A.nz - 1) ? err[i + 1] - e : 0.f;
    float a0 = aux[i] + g0 * inv_fw, a1 = aux[N + i] + g1 * inv_fw, a2 = aux[2 * N + i] + g2 * inv_fw;
    const float nrm = fmaxf(sqrtf(a0 * a0 + a1 * a1 + a2 * a2), 1.f);
    a0 /= nrm; a1 /= nrm; a2 /= nrm;
    const float o0 = gim[i], o1 = gim[N + i], o2 = gim[2 * N + i];
    aux[i]         = (1.f + tf) * a0 - tf * o0;
    aux[N + i]     = (1.f + tf) * a1 - tf * o1;
    aux[2 * N + i] = (1.f + tf) * a2 - tf * o2;
    gim[i] = a0; gim[N + i] = a1; gim[2 * N + i] = a2;
}

}  // namespace

extern "C" void tomo_set_error(const char* msg);
int tomo_check_cuda(cudaError_t e, const char* what);

static int tv_launch_dims(int nx, int ny, int nz, dim3* grid, dim3* block)
{
    if (nx <= 0 || ny <= 0 || nz <= 0) { tomo_set_error("tv: non-positive volume shape"); return TOMO_E_ARG; }
    *block = dim3(32, 8, 1);
    *grid = dim3((nz + 31) / 32, (ny + 7) / 8, nx);
    if (grid->y > 65535u || grid->z > 65535u) { tomo_set_error("tv: volume too large for the launch grid"); return TOMO_E_RANGE; }
    return 0;
}

extern "C" int tomo_tv_dual_error(int nx, int ny, int nz, float weight, const float* p, const float* im, float* err,
                                  void* stream)
{
    if (!p || !im || !err) { tomo_set_error("tomo_tv_dual_error: null pointer"); return TOMO_E_ARG; }
    dim3 grid, block;
    if (int e = tv_launch_dims(nx, ny, nz, &grid, &block)) return e;
    const TvArgs A = {nx, ny, nz};
    tv_dual_error_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(A, weight, p, im, err);
    return tomo_check_cuda(cudaGetLastError(), "tv_dual_error_kernel");
}

extern "C" int tomo_tv_dual_update(int nx, int ny, int nz, float inv_factor_weight, float t_factor, const float* err,
                                   float* aux, float* gim, void* stream)
{
    if (!err || !aux || !gim) { tomo_set_error("tomo_tv_dual_update: null pointer"); return TOMO_E_ARG; }
    dim3 grid, block;
    if (int e = tv_launch_dims(nx, ny, nz, &grid, &block)) return e;
    const TvArgs A = {nx, ny, nz};
    tv_dual_update_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(A, inv_factor_weight, t_factor, err, aux, gim);
    return tomo_check_cuda(cudaGetLastError(), "tv_dual_update_kernel");
}
