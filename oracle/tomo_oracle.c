/*
 * tomo_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C, float64 restatement of the native loops of pandekan/tomography_alignment's
 * projection hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the shipped CUDA path never does.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or fixtures and
 * its Fortran cannot be compiled in this image (no gfortran).  The oracle is
 * therefore pinned against (a) the reference's own pure-numpy twins of these loops
 * (utilities/ray_voxel_utilities.py:173-345), imported from /root/reference with the f2py
 * modules stubbed -- see tests/golden/make_golden.py and tests/golden/*.npz -- and (b) the
 * known-answer identities of SURVEY.md section 4.  Everything here follows the cited lines.
 *
 * Conventions (all from the reference):
 *   volume linear index  (x*ny + y)*nz + z            src/ray_wt_grad.f90:38
 *   ray index            ix*ndz + iz                  utilities/geometry.py:90-94
 *   sample j of ray r    p = p0[:,r] + (j*step)*rhat[:,r]   utilities/ray_voxel_utilities.py:89-94
 *   corner order         fff ffc fcf fcc cff cfc ccf ccc (x,y,z)   src/ray_wt_grad.f90:35-89
 *   every corner is bounds-checked on its own (zero-padded volume)
 *
 * Array arguments named p0 / rhat are (3, n_rays) C-order float64, exactly the numpy arrays
 * the reference builds before it crosses the f2py boundary.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* One sample: floor, floor weights (w_floor = 1 - (p - floor)), as in
 * utilities/ray_voxel_utilities.py:96-99, then the Fortran's own ceil weights
 * wt_c = 1 - wt_f (src/ray_wt_grad.f90:29-34). */
typedef struct {
    long fx, fy, fz;          /* 0-based floor indices */
    double wfx, wfy, wfz;     /* floor weights */
    double wcx, wcy, wcz;     /* ceil weights */
} sample_t;

static inline void make_sample(const double *p0, const double *rhat, long n_rays, long r,
                               long j, double step_size, sample_t *s)
{
    const double js = (double)j * step_size;      /* j * step_size * r_hat: left to right */
    const double px = p0[0 * n_rays + r] + js * rhat[0 * n_rays + r];
    const double py = p0[1 * n_rays + r] + js * rhat[1 * n_rays + r];
    const double pz = p0[2 * n_rays + r] + js * rhat[2 * n_rays + r];
    const double flx = floor(px), fly = floor(py), flz = floor(pz);
    s->fx = (long)flx; s->fy = (long)fly; s->fz = (long)flz;
    s->wfx = 1.0 - (px - flx); s->wfy = 1.0 - (py - fly); s->wfz = 1.0 - (pz - flz);
    s->wcx = 1.0 - s->wfx; s->wcy = 1.0 - s->wfy; s->wcz = 1.0 - s->wfz;
}

#define IN(i, n) ((i) >= 0 && (i) < (n))

/* Number of threads the library will use (1 without OpenMP). */
int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* COO emitter: restates trilinear_ray_sparse, src/ray_wt_grad.f90:1-92.
 * dat/det/wts must hold 8*n_rays*n entries; returns the number written (n_inds).
 * Unused tail is set to -999 like the Fortran (ray_wt_grad.f90:15-17). */
long orc_ray_sparse(const double *p0, const double *rhat, long n_rays, long n, double step_size,
                    long nx, long ny, long nz, int32_t *dat, int32_t *det, double *wts)
{
    const long cap = 8 * n_rays * n;
    for (long i = 0; i < cap; ++i) { dat[i] = -999; det[i] = -999; wts[i] = -999.0; }
    long k = 0;
    for (long r = 0; r < n_rays; ++r) {
        for (long j = 0; j < n; ++j) {
            sample_t s; make_sample(p0, rhat, n_rays, r, j, step_size, &s);
            const long X[2] = {s.fx, s.fx + 1}, Y[2] = {s.fy, s.fy + 1}, Z[2] = {s.fz, s.fz + 1};
            const double WX[2] = {s.wfx, s.wcx}, WY[2] = {s.wfy, s.wcy}, WZ[2] = {s.wfz, s.wcz};
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) {
                if (IN(X[a], nx) && IN(Y[b], ny) && IN(Z[c], nz)) {
                    det[k] = (int32_t)r;
                    dat[k] = (int32_t)((X[a] * ny + Y[b]) * nz + Z[c]);
                    wts[k] = WX[a] * WY[b] * WZ[c];
                    ++k;
                }
            }
        }
    }
    return k;
}

/* y[r] = sum_j sum_corners w * vol[corner]  ==  row r of A applied to vol, float64 accumulate.
 * A is the matrix trilinear_ray_sparse emits (src/ray_wt_grad.f90:20-91) once duplicates are
 * summed (utilities/projection_operators.py:72-76).  If w32 != 0 every weight is first rounded
 * to float32 the way projection_operators.py:106 does (wts.astype(precision)). */
void orc_forward_view(const double *p0, const double *rhat, long n_rays, long n, double step_size,
                      long nx, long ny, long nz, const double *vol, int w32, double *proj)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (long r = 0; r < n_rays; ++r) {
        double acc = 0.0;
        for (long j = 0; j < n; ++j) {
            sample_t s; make_sample(p0, rhat, n_rays, r, j, step_size, &s);
            const long X[2] = {s.fx, s.fx + 1}, Y[2] = {s.fy, s.fy + 1}, Z[2] = {s.fz, s.fz + 1};
            const double WX[2] = {s.wfx, s.wcx}, WY[2] = {s.wfy, s.wcy}, WZ[2] = {s.wfz, s.wcz};
            for (int a = 0; a < 2; ++a) { if (!IN(X[a], nx)) continue;
                for (int b = 0; b < 2; ++b) { if (!IN(Y[b], ny)) continue;
                    for (int c = 0; c < 2; ++c) { if (!IN(Z[c], nz)) continue;
                        double w = WX[a] * WY[b] * WZ[c];
                        if (w32) w = (double)(float)w;
                        acc += w * vol[(X[a] * ny + Y[b]) * nz + Z[c]];
                    } } }
        }
        proj[r] = acc;
    }
}

/* vol += A^T y for one view: the exact transpose of orc_forward_view (the backprojection every
 * solver of the reference uses: recon/sirt.py:61, recon/cgls.py:54,72).  Serial scatter, float64. */
void orc_adjoint_view(const double *p0, const double *rhat, long n_rays, long n, double step_size,
                      long nx, long ny, long nz, const double *y, int w32, double *vol)
{
    for (long r = 0; r < n_rays; ++r) {
        const double yr = y[r];
        if (yr == 0.0) continue;
        for (long j = 0; j < n; ++j) {
            sample_t s; make_sample(p0, rhat, n_rays, r, j, step_size, &s);
            const long X[2] = {s.fx, s.fx + 1}, Y[2] = {s.fy, s.fy + 1}, Z[2] = {s.fz, s.fz + 1};
            const double WX[2] = {s.wfx, s.wcx}, WY[2] = {s.wfy, s.wcy}, WZ[2] = {s.wfz, s.wcz};
            for (int a = 0; a < 2; ++a) { if (!IN(X[a], nx)) continue;
                for (int b = 0; b < 2; ++b) { if (!IN(Y[b], ny)) continue;
                    for (int c = 0; c < 2; ++c) { if (!IN(Z[c], nz)) continue;
                        double w = WX[a] * WY[b] * WZ[c];
                        if (w32) w = (double)(float)w;
                        vol[(X[a] * ny + Y[b]) * nz + Z[c]] += w * yr;
                    } } }
        }
    }
}

/* Projection + gradient image for one view: restates trilinear_ray_interp,
 * src/ray_wt_grad.f90:95-223.
 *   der   (9, 3, n_rays) C-order, from derivative_ray_points (ray_voxel_utilities.py:15-50)
 *   step(r, j) = j*step_size/r_length0           (ray_voxel_utilities.py:148-151)
 *   g(0:3,:) = der(0:3,:,r);  g(3+k,:) = der(3+k,:,r) + step*der(6+k,:,r)   (ray_wt_grad.f90:136-141)
 *   per in-bounds corner: det_img += rec*w;  grad += rec*(sx*wy*wz*g[:,0] + sy*wx*wz*g[:,1] + sz*wx*wy*g[:,2])
 *   with s = -1 for a floor index and +1 for a ceil index on that axis (ray_wt_grad.f90:142-220).
 * det_img (n_rays), grad (6, n_rays) C-order == the Fortran's (6, n_rays) array as numpy sees it. */
void orc_interp_grad_view(const double *p0, const double *rhat, long n_rays, long n,
                          double step_size, double r_length0, long nx, long ny, long nz,
                          const double *vol, const double *der, double *det_img, double *grad)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (long r = 0; r < n_rays; ++r) {
        double acc = 0.0, ga[6] = {0, 0, 0, 0, 0, 0};
        double d[9][3];
        for (int k = 0; k < 9; ++k) for (int a = 0; a < 3; ++a)
            d[k][a] = der[((long)k * 3 + a) * n_rays + r];
        for (long j = 0; j < n; ++j) {
            sample_t s; make_sample(p0, rhat, n_rays, r, j, step_size, &s);
            const double st = (double)j * step_size / r_length0;
            double g[6][3];
            for (int a = 0; a < 3; ++a) {
                g[0][a] = d[0][a]; g[1][a] = d[1][a]; g[2][a] = d[2][a];
                g[3][a] = d[3][a] + st * d[6][a];
                g[4][a] = d[4][a] + st * d[7][a];
                g[5][a] = d[5][a] + st * d[8][a];
            }
            const long X[2] = {s.fx, s.fx + 1}, Y[2] = {s.fy, s.fy + 1}, Z[2] = {s.fz, s.fz + 1};
            const double WX[2] = {s.wfx, s.wcx}, WY[2] = {s.wfy, s.wcy}, WZ[2] = {s.wfz, s.wcz};
            const double SG[2] = {-1.0, 1.0};
            for (int a = 0; a < 2; ++a) { if (!IN(X[a], nx)) continue;
                for (int b = 0; b < 2; ++b) { if (!IN(Y[b], ny)) continue;
                    for (int c = 0; c < 2; ++c) { if (!IN(Z[c], nz)) continue;
                        const double v = vol[(X[a] * ny + Y[b]) * nz + Z[c]];
                        acc += v * (WX[a] * WY[b] * WZ[c]);
                        const double c1 = SG[a] * WY[b] * WZ[c] * v;
                        const double c2 = SG[b] * WX[a] * WZ[c] * v;
                        const double c3 = SG[c] * WX[a] * WY[b] * v;
                        for (int k = 0; k < 6; ++k)
                            ga[k] += c1 * g[k][0] + c2 * g[k][1] + c3 * g[k][2];
                    } } }
        }
        det_img[r] = acc;
        for (int k = 0; k < 6; ++k) grad[(long)k * n_rays + r] = ga[k];
    }
}

/* ---- large-volume variants (tests at the BASELINE.json sizes) ----------------------------------
 * Same loops as above with the volume / projections read as float32 (a float64 copy of a 1024^3
 * volume is 8 GiB) and the work restricted to a caller-chosen subset, so that the oracle finishes
 * in seconds at 512^3 / 1024^3.  Arithmetic is unchanged: float64 positions, weights, accumulators. */

/* Rows rays[0..n_sel) of A applied to vol (orc_forward_view for a subset of rays). */
void orc_forward_rays_f32(const double *p0, const double *rhat, long n_rays, long n, double step_size,
                          long nx, long ny, long nz, const float *vol, const int64_t *rays, long n_sel,
                          double *proj_sel)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (long q = 0; q < n_sel; ++q) {
        const long r = (long)rays[q];
        double acc = 0.0;
        for (long j = 0; j < n; ++j) {
            sample_t s; make_sample(p0, rhat, n_rays, r, j, step_size, &s);
            const long X[2] = {s.fx, s.fx + 1}, Y[2] = {s.fy, s.fy + 1}, Z[2] = {s.fz, s.fz + 1};
            const double WX[2] = {s.wfx, s.wcx}, WY[2] = {s.wfy, s.wcy}, WZ[2] = {s.wfz, s.wcz};
            for (int a = 0; a < 2; ++a) { if (!IN(X[a], nx)) continue;
                for (int b = 0; b < 2; ++b) { if (!IN(Y[b], ny)) continue;
                    for (int c = 0; c < 2; ++c) { if (!IN(Z[c], nz)) continue;
                        acc += (WX[a] * WY[b] * WZ[c]) * (double)vol[(X[a] * ny + Y[b]) * nz + Z[c]];
                    } } }
        }
        proj_sel[q] = acc;
    }
}

/* orc_interp_grad_view for a subset of rays; det_img_sel (n_sel), grad_sel (6, n_sel) C-order. */
void orc_interp_grad_rays_f32(const double *p0, const double *rhat, long n_rays, long n,
                              double step_size, double r_length0, long nx, long ny, long nz,
                              const float *vol, const double *der, const int64_t *rays, long n_sel,
                              double *det_img_sel, double *grad_sel)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (long q = 0; q < n_sel; ++q) {
        const long r = (long)rays[q];
        double acc = 0.0, ga[6] = {0, 0, 0, 0, 0, 0};
        double d[9][3];
        for (int k = 0; k < 9; ++k) for (int a = 0; a < 3; ++a)
            d[k][a] = der[((long)k * 3 + a) * n_rays + r];
        for (long j = 0; j < n; ++j) {
            sample_t s; make_sample(p0, rhat, n_rays, r, j, step_size, &s);
            const double st = (double)j * step_size / r_length0;
            double g[6][3];
            for (int a = 0; a < 3; ++a) {
                g[0][a] = d[0][a]; g[1][a] = d[1][a]; g[2][a] = d[2][a];
                g[3][a] = d[3][a] + st * d[6][a];
                g[4][a] = d[4][a] + st * d[7][a];
                g[5][a] = d[5][a] + st * d[8][a];
            }
            const long X[2] = {s.fx, s.fx + 1}, Y[2] = {s.fy, s.fy + 1}, Z[2] = {s.fz, s.fz + 1};
            const double WX[2] = {s.wfx, s.wcx}, WY[2] = {s.wfy, s.wcy}, WZ[2] = {s.wfz, s.wcz};
            const double SG[2] = {-1.0, 1.0};
            for (int a = 0; a < 2; ++a) { if (!IN(X[a], nx)) continue;
                for (int b = 0; b < 2; ++b) { if (!IN(Y[b], ny)) continue;
                    for (int c = 0; c < 2; ++c) { if (!IN(Z[c], nz)) continue;
                        const double v = (double)vol[(X[a] * ny + Y[b]) * nz + Z[c]];
                        acc += v * (WX[a] * WY[b] * WZ[c]);
                        const double c1 = SG[a] * WY[b] * WZ[c] * v;
                        const double c2 = SG[b] * WX[a] * WZ[c] * v;
                        const double c3 = SG[c] * WX[a] * WY[b] * v;
                        for (int k = 0; k < 6; ++k)
                            ga[k] += c1 * g[k][0] + c2 * g[k][1] + c3 * g[k][2];
                    } } }
        }
        det_img_sel[q] = acc;
        for (int k = 0; k < 6; ++k) grad_sel[(long)k * n_sel + q] = ga[k];
    }
}

/* orc_adjoint_view with the rays spread over the OpenMP threads (float64 atomic adds: the order of
 * the float64 sums varies at the 1e-16 level, nothing else does) and float32 projections. */
void orc_adjoint_view_par_f32(const double *p0, const double *rhat, long n_rays, long n, double step_size,
                              long nx, long ny, long nz, const float *y, double *vol)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (long r = 0; r < n_rays; ++r) {
        const double yr = (double)y[r];
        if (yr == 0.0) continue;
        for (long j = 0; j < n; ++j) {
            sample_t s; make_sample(p0, rhat, n_rays, r, j, step_size, &s);
            const long X[2] = {s.fx, s.fx + 1}, Y[2] = {s.fy, s.fy + 1}, Z[2] = {s.fz, s.fz + 1};
            const double WX[2] = {s.wfx, s.wcx}, WY[2] = {s.wfy, s.wcy}, WZ[2] = {s.wfz, s.wcz};
            for (int a = 0; a < 2; ++a) { if (!IN(X[a], nx)) continue;
                for (int b = 0; b < 2; ++b) { if (!IN(Y[b], ny)) continue;
                    for (int c = 0; c < 2; ++c) { if (!IN(Z[c], nz)) continue;
                        const double t = (WX[a] * WY[b] * WZ[c]) * yr;
#pragma omp atomic
                        vol[(X[a] * ny + Y[b]) * nz + Z[c]] += t;
                    } } }
        }
    }
}

/* Entries voxels[0..n_sel) of A^T y for one view, WITHOUT scattering the whole view: the samples that can
 * touch voxel v lie within one lattice step of the real-valued lattice coordinates (ix, iz, j) of v, where the
 * lattice p(ix, iz, j) = p0[:, ix*ndz + iz] + j*step*rhat is the one orc_adjoint_view walks.  Every candidate
 * sample in a window (the lattice-coordinate extent of a 2-voxel cube, plus 2) around the rounded coordinates is re-evaluated with make_sample (the same arithmetic
 * as the scatter) and contributes iff v is one of its eight corners -- so the terms summed are exactly those the
 * scatter adds to v.  tests/test_oracle_identities.py checks it against orc_adjoint_view entry for entry.
 * out_sel[q] += contribution (accumulates over views). */
void orc_adjoint_voxels_f32(const double *p0, const double *rhat, long ndx, long ndz, long n, double step_size,
                            long nx, long ny, long nz, const float *y, const int64_t *voxels, long n_sel,
                            double *out_sel)
{
    const long n_rays = ndx * ndz;
    (void)nx;
    /* lattice basis from the tabulated ray origins (affine in ix, iz) and ray 0's direction */
    double B[3][3], O0[3];
    for (int a = 0; a < 3; ++a) {
        O0[a] = p0[a * n_rays];
        B[a][0] = (ndx > 1) ? p0[a * n_rays + ndz] - O0[a] : (a == 0 ? 1.0 : 0.0);
        B[a][1] = (ndz > 1) ? p0[a * n_rays + 1] - O0[a] : (a == 2 ? 1.0 : 0.0);
        B[a][2] = step_size * rhat[a * n_rays];
    }
    const double det = B[0][0] * (B[1][1] * B[2][2] - B[1][2] * B[2][1])
                     - B[0][1] * (B[1][0] * B[2][2] - B[1][2] * B[2][0])
                     + B[0][2] * (B[1][0] * B[2][1] - B[1][1] * B[2][0]);
    double Bi[3][3];
    Bi[0][0] =  (B[1][1] * B[2][2] - B[1][2] * B[2][1]) / det;
    Bi[0][1] = -(B[0][1] * B[2][2] - B[0][2] * B[2][1]) / det;
    Bi[0][2] =  (B[0][1] * B[1][2] - B[0][2] * B[1][1]) / det;
    Bi[1][0] = -(B[1][0] * B[2][2] - B[1][2] * B[2][0]) / det;
    Bi[1][1] =  (B[0][0] * B[2][2] - B[0][2] * B[2][0]) / det;
    Bi[1][2] = -(B[0][0] * B[1][2] - B[0][2] * B[1][0]) / det;
    Bi[2][0] =  (B[1][0] * B[2][1] - B[1][1] * B[2][0]) / det;
    Bi[2][1] = -(B[0][0] * B[2][1] - B[0][1] * B[2][0]) / det;
    Bi[2][2] =  (B[0][0] * B[1][1] - B[0][1] * B[1][0]) / det;
#pragma omp parallel for schedule(dynamic, 16)
    for (long q = 0; q < n_sel; ++q) {
        const long v = (long)voxels[q];
        const long vz = v % nz, vy = (v / nz) % ny, vx = v / (nz * ny);
        const double d[3] = {(double)vx - O0[0], (double)vy - O0[1], (double)vz - O0[2]};
        long c[3], w[3];
        for (int k = 0; k < 3; ++k) {
            c[k] = (long)floor(Bi[k][0] * d[0] + Bi[k][1] * d[1] + Bi[k][2] * d[2] + 0.5);
            /* |p - v| < 1 per axis maps to at most sum_a |Bi[k][a]| lattice steps along k; + rounding + margin */
            w[k] = (long)ceil(fabs(Bi[k][0]) + fabs(Bi[k][1]) + fabs(Bi[k][2])) + 2;
        }
        double acc = 0.0;
        for (long ix = c[0] - w[0]; ix <= c[0] + w[0]; ++ix) { if (!IN(ix, ndx)) continue;
            for (long iz = c[1] - w[1]; iz <= c[1] + w[1]; ++iz) { if (!IN(iz, ndz)) continue;
                const long r = ix * ndz + iz;
                const double yr = (double)y[r];
                for (long j = c[2] - w[2]; j <= c[2] + w[2]; ++j) { if (!IN(j, n)) continue;
                    sample_t s; make_sample(p0, rhat, n_rays, r, j, step_size, &s);
                    const long dx = vx - s.fx, dy = vy - s.fy, dz = vz - s.fz;
                    if (dx < 0 || dx > 1 || dy < 0 || dy > 1 || dz < 0 || dz > 1) continue;
                    acc += ((dx ? s.wcx : s.wfx) * (dy ? s.wcy : s.wfy) * (dz ? s.wcz : s.wfz)) * yr;
                } } }
        out_sel[q] += acc;
    }
}

/* ---- orphan (matrix-free, never called from Python) voxel-driven semantics ----------------- */

/* vox += bilinear gather of one view's detector image at the rotated voxel centres.
 * Restates voxel_rigid_transformation + voxel_back_bilinear,
 * src/external_back_projection.f90:1-68, called per view by back_project,
 * src/back_projection.f90:25-32.  rot is the row-major 3x3 matrix Ry(b)*Rx(a)*Rz(p) and tr the
 * vector Ry(b)*t, so that x' = rot*x + tr == Ry(b)*(Rx(a)*Rz(p)*x + t)  (f90:17-25).
 * centres (3, n_vox) C-order, det_image (ndx, ndz) C-order (the Fortran's det_image(np,:,:)).
 * Four taps, each bounds-checked on its own; the y coordinate is ignored (f90:47-66). */
void orc_voxel_back_view(const double *rot, const double *tr, const double *centres, long n_vox,
                         const double *origin, const double *det_image, long ndx, long ndz,
                         double *vox)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n_vox; ++i) {
        const double x = centres[i], y = centres[n_vox + i], z = centres[2 * n_vox + i];
        const double xr = rot[0] * x + rot[1] * y + rot[2] * z + tr[0];
        const double zr = rot[6] * x + rot[7] * y + rot[8] * z + tr[2];
        const double ux = xr - origin[0], uz = zr - origin[2];
        const double flx = floor(ux), flz = floor(uz);
        const long fx = (long)flx, fz = (long)flz;
        const double ax = ux - flx, az = uz - flz;
        double acc = 0.0;
        if (IN(fx, ndx) && IN(fz, ndz))         acc += det_image[fx * ndz + fz] * (1.0 - ax) * (1.0 - az);
        if (IN(fx + 1, ndx) && IN(fz, ndz))     acc += det_image[(fx + 1) * ndz + fz] * ax * (1.0 - az);
        if (IN(fx, ndx) && IN(fz + 1, ndz))     acc += det_image[fx * ndz + fz + 1] * (1.0 - ax) * az;
        if (IN(fx + 1, ndx) && IN(fz + 1, ndz)) acc += det_image[(fx + 1) * ndz + fz + 1] * ax * az;
        vox[i] += acc;
    }
}

/* Voxel-driven forward splat + gradient image: restates bilinear_vox_interp,
 * src/vox_wt_grad.f90:1-55.  floor_x/floor_z/alpha_x/alpha_z/rec are (n_vox); der is
 * (6, 3, n_vox) C-order (numpy view of the Fortran der_points(:,:,i)); outputs use the
 * Fortran's detector layout det_img(fz, fx) flattened with x FASTEST, i.e. index fz + ndz*fx in
 * Fortran order == numpy det_img.ravel() of the (ndz, ndx) array returned by f2py, index
 * fz*ndx + fx.  grad is (6, ndz, ndx) C-order. */
void orc_voxel_splat_grad(long n_vox, const int32_t *floor_x, const int32_t *floor_z,
                          const double *alpha_x, const double *alpha_z, const double *rec,
                          long ndx, long ndz, const double *der, double *det_img, double *grad)
{
    memset(det_img, 0, sizeof(double) * (size_t)(ndx * ndz));
    memset(grad, 0, sizeof(double) * (size_t)(6 * ndx * ndz));
    for (long i = 0; i < n_vox; ++i) {
        const long fx = floor_x[i], fz = floor_z[i];
        const double ax = alpha_x[i], az = alpha_z[i], v = rec[i];
        /* weights and d(weight)/d(alpha) sign pattern, vox_wt_grad.f90:25-50 */
        const long TX[4] = {fx, fx + 1, fx, fx + 1};
        const long TZ[4] = {fz, fz, fz + 1, fz + 1};
        const double W[4]  = {(1.0 - ax) * (1.0 - az), ax * (1.0 - az), (1.0 - ax) * az, ax * az};
        const double G0[4] = {(1.0 - az), -(1.0 - az), az, -az};      /* factor on g(:,1) */
        const double G2[4] = {(1.0 - ax), ax, -(1.0 - ax), -ax};      /* factor on g(:,3) */
        for (int t = 0; t < 4; ++t) {
            if (!(IN(TX[t], ndx) && IN(TZ[t], ndz))) continue;
            const long di = TZ[t] * ndx + TX[t];
            det_img[di] += v * W[t];
            for (int k = 0; k < 6; ++k) {
                const double g1 = der[((long)k * 3 + 0) * n_vox + i];
                const double g3 = der[((long)k * 3 + 2) * n_vox + i];
                grad[(long)k * ndx * ndz + di] += g1 * G0[t] * v + g3 * G2[t] * v;
            }
        }
    }
}

/* COO emitter of the voxel-driven forward matrix: restates bilinear_sparse, src/vox_wt_grad.f90:58-112.
 * floor_x / floor_z int32, alpha_x / alpha_z float32 (the Fortran's real(4) arguments), 4 taps per voxel, each
 * bounds-checked on its own; det index = fx + ndim_x * fz (x fastest); weights are float32 products.
 * dat/det/wts hold 4*n_vox entries, unused tail = -999; returns n_inds. */
long orc_bilinear_sparse(long n_vox, const int32_t *floor_x, const int32_t *floor_z, const float *alpha_x,
                         const float *alpha_z, long ndim_x, long ndim_z, int32_t *dat, int32_t *det, float *wts)
{
    for (long i = 0; i < 4 * n_vox; ++i) { dat[i] = -999; det[i] = -999; wts[i] = -999.0f; }
    long k = 0;
    for (long i = 0; i < n_vox; ++i) {
        const long fx = floor_x[i], fz = floor_z[i];          /* 0-based */
        const float ax = alpha_x[i], az = alpha_z[i];
        if (IN(fx, ndim_x) && IN(fz, ndim_z))         { dat[k] = (int32_t)i; det[k] = (int32_t)(fx + ndim_x * fz);           wts[k] = (1.0f - ax) * (1.0f - az); ++k; }
        if (IN(fx + 1, ndim_x) && IN(fz, ndim_z))     { dat[k] = (int32_t)i; det[k] = (int32_t)(fx + 1 + ndim_x * fz);       wts[k] = ax * (1.0f - az); ++k; }
        if (IN(fx, ndim_x) && IN(fz + 1, ndim_z))     { dat[k] = (int32_t)i; det[k] = (int32_t)(fx + ndim_x * (fz + 1));     wts[k] = (1.0f - ax) * az; ++k; }
        if (IN(fx + 1, ndim_x) && IN(fz + 1, ndim_z)) { dat[k] = (int32_t)i; det[k] = (int32_t)(fx + 1 + ndim_x * (fz + 1)); wts[k] = ax * az; ++k; }
    }
    return k;
}

/* ---- orphan matrix-free forward projector, float32 throughout ---------------------------------------------------
 * Restates forward_project (src/forward_projection.f90:1-68) with rigid_transformation and ray_forward_trilinear
 * (src/external_forward_projection.f90:1-28, 73-160) and the rotation matrices of src/rotations_module.f90 in C `float`
 * arithmetic, operation for operation:
 *   p = Rz(phi) Rx(alpha) (Ry(beta) x + t) - origin  for the source and the detector points (f90:31-37)
 *   r_hat, r_length from RAY 1 ONLY (:41-43);  n_on_ray = NINT(r_length / step_size) (:44 -- the live path truncates)
 *   point j (1-based) = p0 + (j-1)*step_size*r_hat (:52-56);  the cor_shift ARGUMENT IS NEVER READ (:1,10 vs body)
 *   trilinear sum with per-corner bounds checks, w_floor = 1 - (p - real(floor(p))) (external:90-91)
 * source / detector (3, n_rays) C-order float32, xyz (n_proj, 3), ax (n_proj, n_rays) C-order. */
static void rotf(int axis, float a, float m[3][3])
{
    const float c = cosf(a), s = sinf(a);
    if (axis == 2)      { const float t[3][3] = {{c, -s, 0.f}, {s, c, 0.f}, {0.f, 0.f, 1.f}}; memcpy(m, t, sizeof(t)); }
    else if (axis == 0) { const float t[3][3] = {{1.f, 0.f, 0.f}, {0.f, c, -s}, {0.f, s, c}}; memcpy(m, t, sizeof(t)); }
    else                { const float t[3][3] = {{c, 0.f, s}, {0.f, 1.f, 0.f}, {-s, 0.f, c}}; memcpy(m, t, sizeof(t)); }
}

void orc_forward_project_orphan_f32(const float *alpha, const float *beta, const float *phi, const float *xyz,
                                    long n_proj, const float *source, const float *detector, long n_rays,
                                    const float *origin, float step_size, long nx, long ny, long nz,
                                    const float *recon, float *ax)
{
    for (long np = 0; np < n_proj; ++np) {
        float rz[3][3], rx[3][3], ry[3][3], rzx[3][3];
        rotf(2, phi[np], rz); rotf(0, alpha[np], rx); rotf(1, beta[np], ry);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
            rzx[i][j] = rz[i][0] * rx[0][j] + rz[i][1] * rx[1][j] + rz[i][2] * rx[2][j];
        const float *t = xyz + 3 * np;
        float rhat[3] = {0.f, 0.f, 0.f}, rlen = 0.f;
        long n_on_ray = 0;
        for (long r = 0; r < n_rays; ++r) {
            float p0[3], p1[3];
            for (int which = 0; which < 2; ++which) {
                const float *src = which ? detector : source;
                float x[3] = {src[r], src[n_rays + r], src[2 * n_rays + r]}, y[3], *out = which ? p1 : p0;
                for (int i = 0; i < 3; ++i) y[i] = (ry[i][0] * x[0] + ry[i][1] * x[1] + ry[i][2] * x[2]) + t[i];
                for (int i = 0; i < 3; ++i) out[i] = (rzx[i][0] * y[0] + rzx[i][1] * y[1] + rzx[i][2] * y[2]) - origin[i];
            }
            if (r == 0) {                                   /* direction and sample count from the first ray only */
                const float d[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
                rlen = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                for (int i = 0; i < 3; ++i) rhat[i] = d[i] / rlen;
                n_on_ray = lroundf(rlen / step_size);        /* NINT */
            }
            float acc = 0.f;
            for (long j = 0; j < n_on_ray; ++j) {
                float p[3], wf[3];
                long f[3];
                for (int i = 0; i < 3; ++i) {
                    p[i] = p0[i] + ((float)j * step_size) * rhat[i];
                    const float fl = floorf(p[i]);
                    f[i] = (long)fl;
                    wf[i] = 1.0f - (p[i] - fl);
                }
                const long X[2] = {f[0], f[0] + 1}, Y[2] = {f[1], f[1] + 1}, Z[2] = {f[2], f[2] + 1};
                const float WX[2] = {wf[0], 1.0f - wf[0]}, WY[2] = {wf[1], 1.0f - wf[1]}, WZ[2] = {wf[2], 1.0f - wf[2]};
                for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c)   /* fff ffc fcf fcc cff ... */
                    if (IN(X[a], nx) && IN(Y[b], ny) && IN(Z[c], nz))
                        acc = acc + recon[(X[a] * ny + Y[b]) * nz + Z[c]] * (WX[a] * WY[b] * WZ[c]);
            }
            ax[np * n_rays + r] = acc;
        }
    }
}
