// views.cpp -- host-side, float64 per-view constants (no CUDA).
//
// Restates the numpy setup the reference performs for every view before it calls its Fortran
// kernels, reduced to closed form:
//   rotations                  utilities/rotations.py:9-48
//   transform_points           utilities/ray_voxel_utilities.py:6-12     Rz(phi) Rx(alpha) (Ry(beta) x + t)
//   source/detector grids      utilities/geometry.py:90-100              s = (xd + cor_x, -sy, zd)
//   ray marching               utilities/ray_voxel_utilities.py:74-94    p_j = p0 + j*step*r_hat, n = int(|r|/step)
//   derivative_ray_points      utilities/ray_voxel_utilities.py:15-50    (9,3,n_rays) table
// Because the source grid is affine in the detector pixel (ix, iz), every per-ray quantity the
// reference tabulates is affine in (ix, iz); this file computes the affine coefficients.
#include <cmath>
#include <cstring>
#include "tomo_common.h"

namespace {

struct M3 { double m[3][3]; };
struct V3 { double v[3]; };

M3 rot_z(double a)  { return {{{std::cos(a), -std::sin(a), 0.}, {std::sin(a), std::cos(a), 0.}, {0., 0., 1.}}}; }
M3 drot_z(double a) { return {{{-std::sin(a), -std::cos(a), 0.}, {std::cos(a), -std::sin(a), 0.}, {0., 0., 0.}}}; }
M3 rot_x(double a)  { return {{{1., 0., 0.}, {0., std::cos(a), -std::sin(a)}, {0., std::sin(a), std::cos(a)}}}; }
M3 drot_x(double a) { return {{{0., 0., 0.}, {0., -std::sin(a), -std::cos(a)}, {0., std::cos(a), -std::sin(a)}}}; }
M3 rot_y(double a)  { return {{{std::cos(a), 0., std::sin(a)}, {0., 1., 0.}, {-std::sin(a), 0., std::cos(a)}}}; }
M3 drot_y(double a) { return {{{-std::sin(a), 0., std::cos(a)}, {0., 0., 0.}, {-std::cos(a), 0., -std::sin(a)}}}; }

M3 mul(const M3& a, const M3& b) {
    M3 c;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return c;
}
V3 mul(const M3& a, const V3& x) {
    V3 y;
    for (int i = 0; i < 3; ++i) y.v[i] = a.m[i][0] * x.v[0] + a.m[i][1] * x.v[1] + a.m[i][2] * x.v[2];
    return y;
}
V3 add(const V3& a, const V3& b) { return {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; }
V3 sub(const V3& a, const V3& b) { return {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; }
void put(double* dst, const V3& a) { dst[0] = a.v[0]; dst[1] = a.v[1]; dst[2] = a.v[2]; }

}  // namespace

extern "C" void tomo_set_error(const char* msg);

extern "C" int tomo_views_compute_host(const TomoGeom* g, const double* poses, int n_proj, double* out)
{
    if (!g || !poses || !out || n_proj <= 0) { tomo_set_error("tomo_views_compute_host: null pointer or n_proj <= 0"); return TOMO_E_ARG; }
    if (!(g->step_size > 0.0) || g->nx <= 0 || g->ny <= 0 || g->nz <= 0 || g->ndx <= 0 || g->ndz <= 0 ||
        !(g->det_y != g->src_y)) {
        tomo_set_error("tomo_views_compute_host: invalid geometry (sizes must be > 0, step_size > 0, det_y != src_y)");
        return TOMO_E_GEOM;
    }
    const V3 org = {{g->vox_origin[0], g->vox_origin[1], g->vox_origin[2]}};
    for (int v = 0; v < n_proj; ++v) {
        const double* ps = poses + (size_t)v * TOMO_POSE_STRIDE;
        double* o = out + (size_t)v * TOMO_VIEW_STRIDE;
        std::memset(o, 0, sizeof(double) * TOMO_VIEW_STRIDE);
        const double phi = ps[0], alpha = ps[1], beta = ps[2];
        const V3 t = {{ps[3], ps[4], ps[5]}};
        const double cor_x = ps[6];

        const M3 Rp = rot_z(phi), Ra = rot_x(alpha), Rb = rot_y(beta);
        const M3 dRp = drot_z(phi), dRa = drot_x(alpha), dRb = drot_y(beta);
        const M3 Rpa = mul(Rp, Ra);             // rot_pa, ray_voxel_utilities.py:8
        const M3 Rab = mul(Ra, Rb);             // R_ab,   ray_voxel_utilities.py:32

        // source / detector point of ray (0,0) with the centre-of-rotation shift applied to x
        // (ray_voxel_utilities.py:72-73), and the grid steps
        const V3 s00 = {{g->det_x0 + cor_x, g->src_y, g->det_z0}};
        const V3 d00 = {{g->det_x0 + cor_x, g->det_y, g->det_z0}};
        const V3 ex = {{g->det_dx, 0., 0.}}, ez = {{0., 0., g->det_dz}};

        const V3 Rb_s00_t = add(mul(Rb, s00), t);                  // Ry s + t
        const V3 p0 = sub(mul(Rpa, Rb_s00_t), org);                // :74
        const V3 p1 = sub(mul(Rpa, add(mul(Rb, d00), t)), org);    // :75
        const V3 U = mul(Rpa, mul(Rb, ex));
        const V3 W = mul(Rpa, mul(Rb, ez));
        const V3 r = sub(p1, p0);
        const double rlen = std::sqrt(r.v[0] * r.v[0] + r.v[1] * r.v[1] + r.v[2] * r.v[2]);   // :86
        const V3 rhat = {{r.v[0] / rlen, r.v[1] / rlen, r.v[2] / rlen}};
        // :88 (truncation).  The quotient sits on an integer (r_length = 2 sy up to rounding), so the caller passes the
        // count and length its own numpy evaluation produced (pose columns 9, 10); without them the formula is evaluated here.
        const int n = (ps[9] > 0.0) ? (int)ps[9] : (int)(rlen / g->step_size);
        const double rlen_ref = (ps[9] > 0.0 && ps[10] > 0.0) ? ps[10] : rlen;
        const V3 D = {{g->step_size * rhat.v[0], g->step_size * rhat.v[1], g->step_size * rhat.v[2]}};

        put(o + V_P00, p0); put(o + V_U, U); put(o + V_W, W); put(o + V_D, D);
        o[V_N] = (double)n; o[V_RLEN] = rlen_ref;
        for (int a = 0; a < 3; ++a) o[V_INVD + a] = (D.v[a] != 0.0) ? 1.0 / D.v[a] : 0.0;

        // translations: der[k] = (Rz Rx)[:, k]                    (:37-40)
        for (int k = 0; k < 3; ++k) for (int a = 0; a < 3; ++a) o[V_M + 3 * k + a] = Rpa.m[a][k];

        // angles, order phi, alpha, beta                          (:42-45)
        const M3 A3 = mul(dRp, Ra);     // dRz Rx   applied to (Ry s + t)
        const M3 A4 = mul(Rp, dRa);     // Rz dRx   applied to (Ry s + t)
        const M3 A5 = mul(Rpa, dRb);    // Rz Rx dRy applied to s
        put(o + V_E + 0, mul(A3, Rb_s00_t));
        put(o + V_F + 0, mul(A3, mul(Rb, ex)));
        put(o + V_H + 0, mul(A3, mul(Rb, ez)));
        put(o + V_E + 3, mul(A4, Rb_s00_t));
        put(o + V_F + 3, mul(A4, mul(Rb, ex)));
        put(o + V_H + 3, mul(A4, mul(Rb, ez)));
        put(o + V_E + 6, mul(A5, s00));
        put(o + V_F + 6, mul(A5, ex));
        put(o + V_H + 6, mul(A5, ez));

        // step-dependent parts on the untransformed ray vector d - s of ray 0   (:46-48), scaled by
        // d step / d j = step_size / r_length                                    (:148-151)
        const V3 rv = sub(d00, s00);
        const double sc = g->step_size / rlen_ref;
        const V3 k3 = mul(dRp, mul(Rab, rv));
        const V3 k4 = mul(Rp, mul(dRa, mul(Rb, rv)));
        const V3 k5 = mul(Rpa, mul(dRb, rv));
        for (int a = 0; a < 3; ++a) {
            o[V_K + 0 + a] = k3.v[a] * sc;
            o[V_K + 3 + a] = k4.v[a] * sc;
            o[V_K + 6 + a] = k5.v[a] * sc;
        }

        // inverse lattice map for the gather backprojector
        const double L[3][3] = {{U.v[0], W.v[0], D.v[0]}, {U.v[1], W.v[1], D.v[1]}, {U.v[2], W.v[2], D.v[2]}};
        const double det = L[0][0] * (L[1][1] * L[2][2] - L[1][2] * L[2][1])
                         - L[0][1] * (L[1][0] * L[2][2] - L[1][2] * L[2][0])
                         + L[0][2] * (L[1][0] * L[2][1] - L[1][1] * L[2][0]);
        if (!(std::fabs(det) > 0.0)) { tomo_set_error("tomo_views_compute_host: degenerate sample lattice"); return TOMO_E_GEOM; }
        const double id = 1.0 / det;
        double Li[3][3];
        Li[0][0] =  (L[1][1] * L[2][2] - L[1][2] * L[2][1]) * id;
        Li[0][1] = -(L[0][1] * L[2][2] - L[0][2] * L[2][1]) * id;
        Li[0][2] =  (L[0][1] * L[1][2] - L[0][2] * L[1][1]) * id;
        Li[1][0] = -(L[1][0] * L[2][2] - L[1][2] * L[2][0]) * id;
        Li[1][1] =  (L[0][0] * L[2][2] - L[0][2] * L[2][0]) * id;
        Li[1][2] = -(L[0][0] * L[1][2] - L[0][2] * L[1][0]) * id;
        Li[2][0] =  (L[1][0] * L[2][1] - L[1][1] * L[2][0]) * id;
        Li[2][1] = -(L[0][0] * L[2][1] - L[0][1] * L[2][0]) * id;
        Li[2][2] =  (L[0][0] * L[1][1] - L[0][1] * L[1][0]) * id;
        for (int k = 0; k < 3; ++k) {
            for (int a = 0; a < 3; ++a) o[V_LINV + 3 * k + a] = Li[k][a];
            o[V_RB + k] = std::fabs(Li[k][0]) + std::fabs(Li[k][1]) + std::fabs(Li[k][2]);
        }

        // Colour classes for the tile-scatter backprojector.  Rays ix and ix' are processed concurrently
        // by different warps of a block only if ix == ix' (mod C); their samples must then never share
        // a corner voxel.  Two samples differ by d = C k U + m D + n W (k != 0).  They can share a
        // z-corner only while |d_z| < 2, i.e. |n| <= n_over, where inside one tile |m|,|C k| are
        // bounded by the tile's lattice span; for those n the xy separation is at least
        // (C perp - n_over |W_xy|) / sqrt(2) in max-norm, perp = distance between neighbouring rays'
        // xy lines.  C is the smallest count that keeps this >= 2 (then the 2x2 corner cells differ).
        {
            const double dxy = std::sqrt(D.v[0] * D.v[0] + D.v[1] * D.v[1]);
            const double wz = W.v[2];        // lanes must advance towards +z (positive detector pitch)
            double ncol = 0.0;
            if (dxy > 1e-6 && wz >= 0.6) {
                const double perp = std::fabs(U.v[0] * D.v[1] - U.v[1] * D.v[0]) / dxy;
                const double wxy = std::sqrt(W.v[0] * W.v[0] + W.v[1] * W.v[1]);
                // lattice steps (in ix and in j) a tile can span: twice its xy extent plus ghost cells, per unit of xy advance
                // of one step -- for step sizes below 1 a tile holds 1/step as many samples, each drifting D_z in z
                const double uxy = std::sqrt(U.v[0] * U.v[0] + U.v[1] * U.v[1]);
                const double ext = 2.0 * (TOMO_BT_X + TOMO_BT_Y + 8);
                const double zdrift = ext * (std::fabs(U.v[2]) / std::fmax(uxy, 1e-6) + std::fabs(D.v[2]) / dxy);
                const double n_over = std::ceil((2.0 + zdrift) / wz) - 1.0;
                for (int C = 1; C <= 16; ++C)
                    if ((C * perp - n_over * wxy) / std::sqrt(2.0) >= 2.02) { ncol = C; break; }
            }
            o[V_NCOL] = ncol;
        }

        // untilted view: the z coordinate of a sample depends on iz only and (x, y) on (ix, j) only
        o[V_SEP] = (W.v[0] == 0.0 && W.v[1] == 0.0 && U.v[2] == 0.0 && D.v[2] == 0.0 && W.v[2] > 0.0) ? 1.0 : 0.0;

        // z-quad ray kernels (zq_core.h), opt-in per view (pose column 11; measured slower than the per-ray kernels on B200,
        // profiles/README.md): rays iz .. iz+3 of a detector column must almost always share an (x, y) cell (|3 W_xy| small)
        // and sit in consecutive z cells (|3 (W_z - 1)| small); the sample index must fit 13 bits
        o[V_ZQ] = (ps[11] != 0.0 && o[V_SEP] == 0.0 && std::fabs(W.v[2] - 1.0) <= 0.02 && std::fabs(W.v[0]) <= 0.06 && std::fabs(W.v[1]) <= 0.06 &&
                   n <= 8191 && n > 0) ? 1.0 : 0.0;

        // voxel-driven (inverse convention) transform  Ry (Rx Rz x + t)   (external_back_projection.f90:17-25)
        const M3 Vr = mul(Rb, mul(Ra, Rp));
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o[V_VROT + 3 * i + j] = Vr.m[i][j];
        put(o + V_VTR, mul(Rb, t));

        // footprint of a voxel brick on the detector under x' = Vr c + Ry t (detector index = x' - origin, pitch 1):
        // the span of x' (row 0) and z' (row 2) over the brick's voxel centres, plus the two taps and the floor
        {
            const int B[3] = {TOMO_VB_X, TOMO_VB_Y, TOMO_VB_Z};
            double sx = 0.0, sz = 0.0;
            for (int a = 0; a < 3; ++a) {
                sx += std::fabs(Vr.m[0][a] * g->vox_pix[a]) * (B[a] - 1);
                sz += std::fabs(Vr.m[2][a] * g->vox_pix[a]) * (B[a] - 1);
            }
            o[V_VBOK] = (sx + 3.01 <= TOMO_VB_TX && sz + 6.01 <= TOMO_VB_TZ) ? 1.0 : 0.0;     // + 3: box start rounded down to 4
        }

        // derivative_rigid (utilities/voxel_utilities.py:23-48), rows x' (0) and z' (2) only -- the Fortran never
        // reads the y' row (src/vox_wt_grad.f90:27-29).  Each is affine in the voxel centre c:
        //   k = 0..2: R_b[:, k]                      (constant)
        //   k = 3   : (R_b R_a dR_t) c               k = 4: (R_b dR_a R_t) c
        //   k = 5   : dR_b (R_a R_t c + t)
        {
            const M3 D3 = mul(mul(Rb, Ra), dRp), D4 = mul(Rb, mul(dRa, Rp)), D5 = mul(dRb, mul(Ra, Rp));
            const V3 d5t = mul(dRb, t);
            const int rowsel[2] = {0, 2};
            for (int c = 0; c < 2; ++c) {
                const int r_ = rowsel[c];
                for (int k = 0; k < 3; ++k) { double* q = o + V_SPL + (k * 2 + c) * 4; q[0] = q[1] = q[2] = 0.0; q[3] = Rb.m[r_][k]; }
                double* q3 = o + V_SPL + (3 * 2 + c) * 4; double* q4 = o + V_SPL + (4 * 2 + c) * 4; double* q5 = o + V_SPL + (5 * 2 + c) * 4;
                for (int a = 0; a < 3; ++a) { q3[a] = D3.m[r_][a]; q4[a] = D4.m[r_][a]; q5[a] = D5.m[r_][a]; }
                q3[3] = 0.0; q4[3] = 0.0; q5[3] = d5t.v[r_];
            }
            o[V_SORG + 0] = g->vox_origin[0] - ps[6];
            o[V_SORG + 1] = g->vox_origin[2] - ps[8];
        }
    }
    double n_uncoloured = 0.0;
    for (int v = 0; v < n_proj; ++v) if (out[(size_t)v * TOMO_VIEW_STRIDE + V_NCOL] == 0.0) n_uncoloured += 1.0;
    double n_sep = 0.0;
    for (int v = 0; v < n_proj; ++v) if (out[(size_t)v * TOMO_VIEW_STRIDE + V_SEP] != 0.0) n_sep += 1.0;
    double n_vbig = 0.0;
    for (int v = 0; v < n_proj; ++v) if (out[(size_t)v * TOMO_VIEW_STRIDE + V_VBOK] == 0.0) n_vbig += 1.0;
    for (int v = 0; v < n_proj; ++v) {
        out[(size_t)v * TOMO_VIEW_STRIDE + V_NUNCOL] = n_uncoloured;
        out[(size_t)v * TOMO_VIEW_STRIDE + V_NSEP] = n_sep;
        out[(size_t)v * TOMO_VIEW_STRIDE + V_NVBIG] = n_vbig;
    }
    return 0;
}
