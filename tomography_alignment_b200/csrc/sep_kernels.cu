// sep_kernels.cu -- separable forward projector for untilted views (V_SEP == 1), sm_100a.  See sep_core.h.
//
// Block = 8 warps = 8 adjacent ix; a warp handles one ix and one chunk of SEP_CHUNK padded z planes, four planes per
// lane with 128-bit loads (the 32 lanes of a warp read 512 contiguous bytes of each of the 4 (x, y) corner rows).  All
// lanes of a warp share the (x, y) march, so there is no divergence.  The warp then stages S in shared memory and
// applies the 2-tap z interpolation for the detector rows whose floor plane falls in its chunk.  The volume is
// streamed once per view: this kernel is HBM/L2-bandwidth bound, not issue bound.
#include <cuda_runtime.h>
#include "tomo_common.h"
#include "sep_core.h"

namespace {

constexpr int SEP_WARPS = 8;

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
    int bid = blockIdx.x;
    const int xt = bid % A.nxt; bid /= A.nxt;
    const int view = bid % A.n_proj;
    const int chunk = bid / A.n_proj;
    const double* __restrict__ V = A.views + (size_t)view * TOMO_VIEW_STRIDE;
    if (V[V_SEP] == 0.0) return;                                   // tilted view: the generic kernel does it
    const int lane = threadIdx.x, warp = threadIdx.y;
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

}  // namespace

extern "C" void tomo_set_error(const char* msg);
int tomo_check_cuda(cudaError_t e, const char* what);

// Launch the separable forward projector over all views of the table; views with V_SEP == 0 return at once.
int tomo_forward_separable_launch(const TomoGeom* g, const void* views, int n_proj, const float* volpad, float* proj,
                                  void* stream)
{
    SepArgs A;
    A.volpad = volpad; A.views = (const double*)views; A.proj = proj;
    A.nx = g->nx; A.ny = g->ny; A.nz = g->nz; A.ndx = g->ndx; A.ndz = g->ndz; A.n_proj = n_proj;
    A.nzp = tomo_nzp(g->nz);
    A.syp = A.nzp;
    A.sxp = (g->ny + 2 * TOMO_PAD) * A.syp;
    A.nxt = (g->ndx + SEP_WARPS - 1) / SEP_WARPS;
    A.nchunk = (A.nzp - 1 + SEP_OUT - 1) / SEP_OUT;
    // the last chunk must be able to serve floor planes up to nzp - 2 from SEP_CHUNK staged planes
    while ((A.nchunk - 1) * SEP_OUT + SEP_CHUNK < A.nzp) ++A.nchunk;
    const double nblocks = (double)A.nxt * A.nchunk * n_proj;
    if (nblocks >= 2147483647.0) { tomo_set_error("separable forward: too many blocks for one launch"); return TOMO_E_RANGE; }
    sep_forward_kernel<<<(unsigned)nblocks, dim3(32, SEP_WARPS), 0, (cudaStream_t)stream>>>(A);
    return tomo_check_cuda(cudaGetLastError(), "sep_forward_kernel");
}
