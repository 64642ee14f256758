// emu.cpp -- TEST HARNESS ONLY: runs the __host__ __device__ cores of the CUDA kernels
// (tomography_alignment_b200/csrc/ray_core.h, back_core.h) on the CPU, one loop iteration per
// GPU thread, so the kernel arithmetic can be checked against the oracle in the CPU-only test tier.
// The product never loads this library.
#include <cstddef>
#include <vector>
#include "../../tomography_alignment_b200/csrc/ray_core.h"
#include "../../tomography_alignment_b200/csrc/back_core.h"
#include "../../tomography_alignment_b200/csrc/sep_core.h"
#include "../../tomography_alignment_b200/csrc/zq_core.h"

#define EMU_API extern "C" __attribute__((visibility("default")))

EMU_API void emu_pad(const TomoGeom* g, const float* vol, float* pad)
{
    int nyp, nzp;
    tomo_pad_pitch(g->ny, g->nz, &nyp, &nzp);
    const int nxp = g->nx + 2 * TOMO_PAD;
    for (size_t i = 0; i < (size_t)nxp * nyp * nzp + TOMO_PAD_HEAD + TOMO_PAD_TAIL; ++i) pad[i] = 0.f;
    pad += TOMO_PAD_HEAD;                                  // buffer layout of tomo_pad_volume: head slack, volume, tail slack
    for (int x = 0; x < g->nx; ++x) for (int y = 0; y < g->ny; ++y) for (int z = 0; z < g->nz; ++z)
        pad[((size_t)(x + TOMO_PAD) * nyp + y + TOMO_PAD) * nzp + z + TOMO_PAD] = vol[((size_t)x * g->ny + y) * g->nz + z];
}

// use_zq: take the z-quad core (zq_core.h) for the views that qualify (V_ZQ), as the CUDA path does
static int g_use_zq = 1;
EMU_API void emu_set_zq(int on) { g_use_zq = on; }

EMU_API void emu_proj_grad(const TomoGeom* g, const double* views, int n_proj, const float* volpad_buf,
                           const float* meas, float* proj, float* dproj, double* grad6, double* cost, int want_grad)
{
    const float* volpad = volpad_buf + TOMO_PAD_HEAD;
    int nyp_, syp_;
    tomo_pad_pitch(g->ny, g->nz, &nyp_, &syp_);
    const RayDims dm = {g->nx, g->ny, g->nz, nyp_ * syp_, syp_};
    const size_t n_det = (size_t)g->ndx * g->ndz;
    for (int v = 0; v < n_proj; ++v) {
        const double* V = views + (size_t)v * TOMO_VIEW_STRIDE;
        double red[7] = {0, 0, 0, 0, 0, 0, 0};
        if (!want_grad && V[V_SEP] != 0.0) {
            // untilted view: the separable cores (sep_core.h), S for every padded plane then the 2-tap z interpolation
            const int nzp = tomo_nzp(g->nz);
#pragma omp parallel for schedule(dynamic, 8)
            for (int ix = 0; ix < g->ndx; ++ix) {
                std::vector<float> S(nzp);
                SepSetup r;
                sep_setup(V, dm, ix, r);
                for (int zq = 0; zq < nzp; zq += 4) sep_march_xy(volpad, V, dm, r, zq, &S[zq]);
                for (int iz = 0; iz < g->ndz; ++iz) {
                    int fzp; float wz;
                    sep_zcell(V, iz, fzp, wz);
                    float val = 0.f;
                    if (fzp >= 0 && fzp <= nzp - 2) val = fmaf(wz, S[fzp + 1] - S[fzp], S[fzp]);
                    if (proj) proj[v * n_det + (size_t)ix * g->ndz + iz] = val;
                }
            }
            continue;
        }
        if (want_grad && V[V_SEP] != 0.0) {
            // untilted view: separable gradient cores (sep_core.h)
            const int nzp = tomo_nzp(g->nz);
#pragma omp parallel for schedule(dynamic, 8)
            for (int ix = 0; ix < g->ndx; ++ix) {
                std::vector<float> S(nzp), T(nzp), Gx(nzp), Gy(nzp), TGx(nzp), TGy(nzp);
                SepSetup r;
                sep_setup(V, dm, ix, r);
                for (int zq = 0; zq < nzp; zq += 4) {
                    SepMoments m;
                    sep_march_xy_grad(volpad, V, dm, r, zq, m);
                    for (int k = 0; k < 4; ++k) { S[zq + k] = m.S[k]; T[zq + k] = m.T[k]; Gx[zq + k] = m.Gx[k]; Gy[zq + k] = m.Gy[k]; TGx[zq + k] = m.TGx[k]; TGy[zq + k] = m.TGy[k]; }
                }
                double loc[7] = {0, 0, 0, 0, 0, 0, 0};
                for (int iz = 0; iz < g->ndz; ++iz) {
                    const size_t ray = (size_t)ix * g->ndz + iz;
                    int fzp; float wz;
                    sep_zcell(V, iz, fzp, wz);
                    RaySums sm; sm.acc = 0.f;
                    for (int k = 0; k < 3; ++k) { sm.s0[k] = 0.f; sm.s1[k] = 0.f; }
                    if (fzp >= 0 && fzp <= nzp - 2) sep_ray_sums(S.data(), T.data(), Gx.data(), Gy.data(), TGx.data(), TGy.data(), fzp, wz, sm);
                    if (proj) proj[v * n_det + ray] = sm.acc;
                    float dp[6];
                    ray_gradient(V, ix, iz, sm, dp);
                    if (dproj) for (int k = 0; k < 6; ++k) dproj[((size_t)v * 6 + k) * n_det + ray] = dp[k];
                    if (meas) {
                        const double res = (double)meas[v * n_det + ray] - (double)sm.acc;
                        for (int k = 0; k < 6; ++k) loc[k] += -(double)dp[k] * res;
                        loc[6] += 0.5 * res * res;
                    }
                }
#pragma omp critical
                for (int k = 0; k < 7; ++k) red[k] += loc[k];
            }
            if (grad6) for (int k = 0; k < 6; ++k) grad6[v * 6 + k] = red[k];
            if (cost) cost[v] = red[6];
            continue;
        }
        const bool zq = g_use_zq && V[V_ZQ] != 0.0;
#pragma omp parallel for schedule(dynamic, 8)
        for (int ix = 0; ix < g->ndx; ++ix) {
            double loc[7] = {0, 0, 0, 0, 0, 0, 0};
            ZqSums zs;
            unsigned short ev[ZQ_CAP + 2 * ZQ_G];
            for (int iz = 0; iz < g->ndz; ++iz) {
                const size_t ray = (size_t)ix * g->ndz + iz;
                RaySums s;
                if (zq) {                                   // one "thread" per group of four rays
                    if (iz % ZQ_G == 0) {
                        const int nr = g->ndz - iz < ZQ_G ? g->ndz - iz : ZQ_G;
                        if (want_grad) zq_march<true>(volpad, V, dm, ix, iz, nr, ev, 1, zs); else zq_march<false>(volpad, V, dm, ix, iz, nr, ev, 1, zs);
                    }
                    const int k = iz % ZQ_G;
                    s.acc = zs.acc[k];
                    for (int a = 0; a < 3; ++a) { s.s0[a] = zs.s0[k][a]; s.s1[a] = zs.s1[k][a]; }
                } else if (want_grad) ray_march<true>(volpad, V, dm, ix, iz, s); else ray_march<false>(volpad, V, dm, ix, iz, s);
                if (proj) proj[v * n_det + ray] = s.acc;
                if (want_grad) {
                    float dp[6];
                    ray_gradient(V, ix, iz, s, dp);
                    if (dproj) for (int k = 0; k < 6; ++k) dproj[((size_t)v * 6 + k) * n_det + ray] = dp[k];
                    if (meas) {
                        const double res = (double)meas[v * n_det + ray] - (double)s.acc;
                        for (int k = 0; k < 6; ++k) loc[k] += -(double)dp[k] * res;
                        loc[6] += 0.5 * res * res;
                    }
                }
            }
#pragma omp critical
            for (int k = 0; k < 7; ++k) red[k] += loc[k];
        }
        if (grad6) for (int k = 0; k < 6; ++k) grad6[v * 6 + k] = red[k];
        if (cost) cost[v] = red[6];
    }
}

EMU_API void emu_back(const TomoGeom* g, const double* views, int n_proj, const float* proj, float* vol,
                      int accumulate, int voxel_bilinear, const double* origin)
{
    const size_t n_det = (size_t)g->ndx * g->ndz;
#pragma omp parallel for schedule(dynamic, 1)
    for (int x = 0; x < g->nx; ++x) for (int y = 0; y < g->ny; ++y) for (int z = 0; z < g->nz; ++z) {
        float acc = 0.f;
        for (int v = 0; v < n_proj; ++v) {
            const double* V = views + (size_t)v * TOMO_VIEW_STRIDE;
            if (voxel_bilinear)
                acc += voxel_bilinear_view(proj + v * n_det, V, g->ndx, g->ndz, origin,
                                           g->vox_origin[0] + x * g->vox_pix[0], g->vox_origin[1] + y * g->vox_pix[1],
                                           g->vox_origin[2] + z * g->vox_pix[2]);
            else if (V[V_SEP] != 0.0) {
                // untilted view: separable cores (sep_core.h): K per candidate ix times the z-gathered projection row
                SepColumn c;
                sep_column_setup(V, g->ndx, x, y, c);
                for (int mi = c.milo; mi <= c.mihi; ++mi) {
                    const float K = sep_column_weight(V, c, mi);
                    if (K != 0.f)
                        acc = fmaf(K, sep_zgather(proj + v * n_det + (size_t)(c.n0i + mi) * g->ndz, V, g->ndz, z), acc);
                }
            } else
                acc += adjoint_gather_view(proj + v * n_det, V, g->ndx, g->ndz, x, y, z);
        }
        const size_t vi = ((size_t)x * g->ny + y) * g->nz + z;
        vol[vi] = accumulate ? vol[vi] + acc : acc;
    }
}
