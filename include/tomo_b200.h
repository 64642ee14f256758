/*
 * tomo_b200.h -- C ABI of libtomo_b200.so: B200 (sm_100a) projection operators for
 * pandekan/tomography_alignment.
 *
 * The reference has no C ABI; its native boundary is f2py around Fortran 90.  Each entry point
 * below names the reference interface it replaces (file:line under the reference tree).  All
 * pointers suffixed _dev are raw CUDA device pointers owned by the caller (e.g. torch
 * tensor.data_ptr()); the library allocates no device memory.  `stream` is a cudaStream_t passed
 * as void* (NULL = legacy default stream).  Calls are asynchronous with respect to the host unless
 * stated, re-entrant, and keep no global state except the thread-local last-error string.
 *
 * Return value: 0 on success, a negative TOMO_E_* code for argument errors, a positive
 * cudaError_t value for CUDA failures.  Nothing throws across the ABI.
 *
 * Layouts (all from the reference):
 *   volume        float32 [nx][ny][nz], z fastest              src/ray_wt_grad.f90:38
 *   projections   float32 [n_proj][ndx][ndz], iz fastest       utilities/geometry.py:90-94,
 *                                                              utilities/projection_operators.py:108
 *   poses         float64 [n_proj][TOMO_POSE_STRIDE = 12] = phi, alpha, beta, tx, ty, tz, cor_x, cor_y, cor_z,
 *                 n_samples, r_length0, flags
 *                 (angles, xyz_shift and Geometry.cor_shift rows of projection_operators.py:50-52,97-102;
 *                 only cor_x is used, ray_voxel_utilities.py:72-73).  n_samples / r_length0 are the
 *                 reference's n = int(r_length[0] / step_size) and r_length[0] (ray_voxel_utilities.py:85-88)
 *                 as ITS numpy expression evaluates them: the quotient sits on an integer, so the count
 *                 depends on last-bit rounding that only the caller's numpy can reproduce.  n_samples <= 0
 *                 asks the library to evaluate the formula itself in float64 (may differ by one trailing
 *                 sample, which lies sy - step beyond the rotation centre).  flags != 0 lets a nearly untilted view
 *                 take the z-quad ray kernels (four z-adjacent rays per thread, 128-bit window loads; same results to
 *                 rounding, measured slower than the per-ray kernels on B200 -- off by default).
 *   gradients     order [tx, ty, tz, phi, alpha, beta]         utilities/ray_voxel_utilities.py:39-46
 */
#ifndef TOMO_B200_H
#define TOMO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOMO_B200_VERSION 100            /* 0.1.0 */

#if defined(__GNUC__)
#define TOMO_API __attribute__((visibility("default")))
#else
#define TOMO_API
#endif

#define TOMO_E_ARG        (-1)           /* NULL pointer / non-positive size */
#define TOMO_E_GEOM       (-2)           /* geometry not representable (e.g. zero step) */
#define TOMO_E_WORKSPACE  (-3)           /* workspace too small */
#define TOMO_E_RANGE      (-4)           /* size exceeds 32-bit index budget of the kernels */

/* Doubles per view in the device-side view table written by tomo_views_*. */
#define TOMO_VIEW_STRIDE  160
#define TOMO_POSE_STRIDE  12
/* Zero border (voxels) of the padded volume on every side of every axis. */
#define TOMO_PAD          2

/* Parallel-beam geometry: the scalars utilities/geometry.py:14-105 derives its grids from. */
typedef struct TomoGeom {
    int32_t nx, ny, nz;        /* Geometry.vox_shape */
    int32_t ndx, ndz;          /* Geometry.det_shape */
    double  vox_origin[3];     /* Geometry.vox_origin = first voxel centre (geometry.py:87) */
    double  vox_pix[3];        /* Geometry.vox_pix (centre spacing, geometry.py:82-84) */
    double  det_x0, det_z0;    /* first detector-pixel centre (geometry.py:92-93) */
    double  det_dx, det_dz;    /* detector pitch */
    double  src_y, det_y;      /* source / detector plane y = -sy / +sy (geometry.py:95-100) */
    double  step_size;         /* Geometry.step_size */
} TomoGeom;

TOMO_API int         tomo_version(void);
TOMO_API const char* tomo_last_error(void);

/* ---- per-view constants --------------------------------------------------------------------- */
/* Host-only, float64: the per-view setup of forward_sparse / forward_proj_grad
 * (utilities/ray_voxel_utilities.py:66-94,124-159) and derivative_ray_points (:15-50) reduced to
 * the affine lattice  p(ix,iz,j) = P00 + ix*U + iz*W + j*D  and the affine pose-derivative tables.
 * out_host: [n_proj][TOMO_VIEW_STRIDE] doubles.  No CUDA call is made. */
TOMO_API int tomo_views_compute_host(const TomoGeom* geom, const double* poses, int n_proj, double* out_host);
TOMO_API size_t tomo_views_bytes(int n_proj);
/* Same, then cudaMemcpyAsync into views_dev (tomo_views_bytes(n_proj) bytes) on `stream`; the host
 * staging copy is synchronised before return so `poses` may be reused immediately. */
TOMO_API int tomo_views_upload(const TomoGeom* geom, const double* poses, int n_proj, void* views_dev, void* stream);

/* Which kernel families a view table needs (host-side, from the table tomo_views_compute_host wrote).  Passing the mask to
 * the *_ex / *_slab entry points lets them skip the launches that would find no view (each costs ~0.3 ms at 512^3: the
 * kernel is launched over the full grid and every block leaves at once).  A mask computed for a table is also valid for any
 * contiguous part of it.  0 = unknown: launch everything. */
#define TOMO_KINDS_KNOWN       1     /* the mask is valid */
#define TOMO_KINDS_GENERIC     2     /* tilted views: ray_kernel_forward / ray_kernel_gradient */
#define TOMO_KINDS_SEPARABLE   4     /* untilted views (alpha = beta = 0): separable kernels */
#define TOMO_KINDS_TILE        8     /* tilted views inside the scatter envelope: adjoint_tile_kernel */
#define TOMO_KINDS_UNCOLOURED 16     /* views outside it (rays nearly parallel to z): adjoint_gather_kernel */
#define TOMO_KINDS_ZQUAD      32     /* nearly untilted views (W ~ (0,0,1)): zq_kernel_forward / zq_kernel_gradient */
TOMO_API int tomo_views_kinds(const double* views_host, int n_proj);

/* ---- padded volume -------------------------------------------------------------------------- */
/* The ray-driven kernels read a zero-bordered copy of the volume (zero-padded-corner semantics of
 * src/ray_wt_grad.f90:35-89 without per-corner branches): [nx+2P][nyp][syp], data at offset (P,P,P), P = TOMO_PAD, with
 * nyp >= ny+2P rows per plane and syp >= nz+2P (a multiple of 32) floats per row -- the pitches are the library's choice
 * (those of the next cube of 64^3 ... 1024^3 when that costs at most twice the memory: the ray kernels have compile-time-stride
 * variants for them); callers only size the buffer with tomo_padded_volume_bytes() and fill it with tomo_pad_volume().  The
 * buffer holds 32 floats of zero slack before and after the volume (included in the byte count, written by tomo_pad_volume). */
TOMO_API size_t tomo_padded_volume_bytes(const TomoGeom* geom);
TOMO_API int tomo_pad_volume(const TomoGeom* geom, const float* vol_dev, float* volpad_dev, void* stream);

/* ---- operators ------------------------------------------------------------------------------ */
/* proj = A x for all views.  Replaces building A with trilinear_ray_sparse
 * (src/ray_wt_grad.f90:1-92 via utilities/projection_operators.py:54-76,95-110) and
 * sparse.csr_matrix.dot(A, x) (recon/sirt.py:59); also forward_project
 * (src/forward_projection.f90:1-68). */
TOMO_API int tomo_forward(const TomoGeom* geom, const void* views_dev, int n_proj,
                 const float* volpad_dev, float* proj_dev, void* stream);

/* tomo_forward with the table's TOMO_KINDS_* mask (see tomo_views_kinds). */
TOMO_API int tomo_forward_ex(const TomoGeom* geom, const void* views_dev, int n_proj, int kinds,
                    const float* volpad_dev, float* proj_dev, void* stream);

/* vol (+)= A^T y, the exact transpose of tomo_forward: replaces
 * sparse.csc_matrix.dot(sparse.csr_matrix.transpose(A), y) (recon/sirt.py:61, recon/cgls.py:54,72).
 * Tile-owned scatter in shared memory: no atomics, bitwise deterministic.  vol_dev is the UNPADDED
 * volume. */
TOMO_API int tomo_back_adjoint(const TomoGeom* geom, const void* views_dev, int n_proj,
                      const float* proj_dev, float* vol_dev, int accumulate, void* stream);

/* Same operator, with a caller-provided device workspace of tomo_back_adjoint_workspace_bytes() bytes: views
 * without tilt (alpha = beta = 0 exactly -- the default poses of projection_matrix) are then backprojected by the
 * separable adjoint (z-transposed projections Yz in the workspace, one warp per voxel column, register
 * accumulators), ~8x faster than the tile kernel; tilted views take the tile kernel as in tomo_back_adjoint. */
TOMO_API size_t tomo_back_adjoint_workspace_bytes(const TomoGeom* geom, int n_proj);
TOMO_API int tomo_back_adjoint_ws(const TomoGeom* geom, const void* views_dev, int n_proj,
                         const float* proj_dev, float* vol_dev, int accumulate,
                         void* workspace_dev, size_t workspace_bytes, void* stream);

/* tomo_back_adjoint_ws restricted to the x-slab [x_begin, x_end) of the volume (an x-slab of [nx][ny][nz] is contiguous):
 * only voxels of the slab are written.  x_begin and x_end must be multiples of tomo_back_adjoint_slab_granularity()
 * (x_end may also be nx).  Launching the slabs one after the other lets the caller start the all-reduce of slab k
 * (recon/sirt_mpi.py:103) while slab k+1 is still being backprojected.  workspace_dev may be NULL (then untilted views take
 * the tile kernel); with a workspace the slab that starts at x_begin = 0 must be launched first (it fills the z-transposed
 * projections the later slabs read).  kinds: TOMO_KINDS_* mask or 0. */
TOMO_API int tomo_back_adjoint_slab_granularity(void);
TOMO_API int tomo_back_adjoint_slab(const TomoGeom* geom, const void* views_dev, int n_proj, int kinds,
                           const float* proj_dev, float* vol_dev, int accumulate,
                           void* workspace_dev, size_t workspace_bytes, int x_begin, int x_end, void* stream);

/* Same operator and contract as tomo_back_adjoint, computed by the per-voxel gather kernel (an
 * independent formulation: ~8x slower, used as the cross-check of the tile-scatter kernel and for
 * poses outside its envelope, i.e. views whose rays are nearly parallel to z). */
TOMO_API int tomo_back_adjoint_gather(const TomoGeom* geom, const void* views_dev, int n_proj,
                             const float* proj_dev, float* vol_dev, int accumulate, void* stream);

/* vol (+)= voxel-driven bilinear backprojection, the orphan matrix-free back_project
 * (src/back_projection.f90:1-34, src/external_back_projection.f90:1-68): x' = Ry(b)(Rx(a)Rz(p)x + t),
 * 4 bilinear taps of the view's image at (x'_x - origin_x, x'_z - origin_z), y ignored.
 * NOT the transpose of tomo_forward (inverse pose convention).  origin[3] is the
 * Fortran's `origin` argument; det images use the [n_proj][ndx][ndz] layout above.
 * When ndz % 4 == 0, proj_dev is 16-byte aligned and the detector is at least 32 x 44 pixels the projection
 * tile of every view is staged in shared memory by TMA (one 3-D tensor map over proj_dev, encoded per call on
 * the host, no device allocation); other layouts and views tilted beyond the staged box take the plain gather
 * kernel.  Both write every voxel exactly once (no atomics, bitwise reproducible). */
TOMO_API int tomo_back_voxel_bilinear(const TomoGeom* geom, const void* views_dev, int n_proj,
                             const double origin[3], const float* proj_dev, float* vol_dev,
                             int accumulate, void* stream);

/* Voxel-driven forward projection ("splat") with optional gradient image, the orphan path of
 * utilities/voxel_utilities.py:82-108 -> bilinear_vox_interp (src/vox_wt_grad.f90:1-55): every voxel centre
 * is mapped with x' = Ry(b)(Rx(a)Rz(p)x + t), floor/fraction are taken relative to vox_origin - cor_shift
 * (x and z components) and rec * bilinear weight is added to the 4 detector taps, each bounds-checked on its
 * own; the gradient image adds rec * (g_x * dW/dx' + g_z * dW/dz') with g = derivative_rigid
 * (voxel_utilities.py:23-48), rows [sx, sy, sz, theta(phi), alpha, beta].
 *   det_dev   [n_proj][ndz][ndx]      float32  (x FASTEST: the Fortran's det_img(fz, fx), transposed w.r.t. the
 *                                               ray-driven projections)
 *   grad_dev  [n_proj][6][ndz][ndx]   float32, nullable
 * Outputs are zeroed by the call.  vol_dev is the UNPADDED volume.  The operator is a scatter:
 *   tomo_voxel_splat                float32 atomic adds: the summation ORDER (not the set of terms) varies between runs, so
 *                                   results agree to float32 rounding, not bitwise;
 *   tomo_voxel_splat_deterministic  every contribution is added as a 64-bit fixed-point integer (scale from max |vol| and the
 *                                   geometry, resolution ~1e-14 of the largest term) and converted at the end: integer addition
 *                                   is associative, so results are bitwise reproducible and more accurate than float32
 *                                   accumulation.  Needs tomo_voxel_splat_workspace_bytes() bytes of device workspace. */
TOMO_API int tomo_voxel_splat(const TomoGeom* geom, const void* views_dev, int n_proj, const float* vol_dev,
                     float* det_dev, float* grad_dev, void* stream);
TOMO_API size_t tomo_voxel_splat_workspace_bytes(const TomoGeom* geom, int n_proj, int with_grad);
TOMO_API int tomo_voxel_splat_deterministic(const TomoGeom* geom, const void* views_dev, int n_proj, const float* vol_dev,
                                   float* det_dev, float* grad_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* vol (+)= S^T det, the transpose of the splat matrix S that bilinear_sparse emits (src/vox_wt_grad.f90:58-112 through
 * utilities/voxel_utilities.py:50-79: rows = detector pixels fx + ndim_x * fz, columns = voxels, float32 weights, taps
 * bounds-checked one by one).  det_dev [n_proj][ndz][ndx] (x fastest, as tomo_voxel_splat writes it).  A per-voxel gather:
 * no atomics, bitwise reproducible. */
TOMO_API int tomo_voxel_splat_adjoint(const TomoGeom* geom, const void* views_dev, int n_proj, const float* det_dev,
                             float* vol_dev, int accumulate, void* stream);

/* Projection + 6-DOF rigid-body gradient for all views in one launch.  Replaces
 * ProjectionMatrix.projection_gradient (utilities/projection_operators.py:112-122) ->
 * forward_proj_grad (utilities/ray_voxel_utilities.py:113-170) -> trilinear_ray_interp
 * (src/ray_wt_grad.f90:95-223), and compute_gradient (src/projection_gradient.f90:1-79).
 *   proj_dev   [n_proj][n_det]      float32, nullable
 *   dproj_dev  [n_proj][6][n_det]   float32, nullable: d proj / d theta per ray (the `grad` the
 *              reference returns)
 *   meas_dev   [n_proj][n_det]      float32, nullable: measured projections b
 *   grad6_dev  [n_proj][6]          float64, nullable (needs meas): sum_rays (-dproj_k)*(b - proj),
 *              i.e. np.dot(s, residual) of utilities/alignment_functions.py:27-37,176-186
 *   cost_dev   [n_proj]             float64, nullable (needs meas): 0.5*||b - proj||^2
 *              (utilities/alignment_functions.py:163)
 * grad6/cost are reduced in float64 in a fixed order (bitwise deterministic); they need
 * tomo_proj_grad_workspace_bytes() bytes of workspace. */
TOMO_API size_t tomo_proj_grad_workspace_bytes(const TomoGeom* geom, int n_proj);
TOMO_API int tomo_proj_grad(const TomoGeom* geom, const void* views_dev, int n_proj,
                   const float* volpad_dev, const float* meas_dev,
                   float* proj_dev, float* dproj_dev, double* grad6_dev, double* cost_dev,
                   void* workspace_dev, size_t workspace_bytes, void* stream);

/* tomo_proj_grad with the table's TOMO_KINDS_* mask (see tomo_views_kinds). */
TOMO_API int tomo_proj_grad_ex(const TomoGeom* geom, const void* views_dev, int n_proj, int kinds,
                      const float* volpad_dev, const float* meas_dev,
                      float* proj_dev, float* dproj_dev, double* grad6_dev, double* cost_dev,
                      void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- TV proximal step ------------------------------------------------------ */
/* Dual FISTA iteration of utilities/tv_denoise.py:98-170 (denoise_fista), two fused stencil kernels.
 * Fields p / aux / gim: float32 [3][nx][ny][nz]; im / err: float32 [nx][ny][nz].
 *   tomo_tv_dual_error : err = weight * div(p) - im                       (tv_denoise.py:147, div :20-31)
 *   tomo_tv_dual_update: aux += gradient(err) * inv_factor_weight; tmp = aux / max(|aux|, 1);
 *                        aux = (1 + t_factor) * tmp - t_factor * gim; gim = tmp      (:148-154) */
TOMO_API int tomo_tv_dual_error(int nx, int ny, int nz, float weight, const float* p_dev, const float* im_dev,
                       float* err_dev, void* stream);
TOMO_API int tomo_tv_dual_update(int nx, int ny, int nz, float inv_factor_weight, float t_factor,
                        const float* err_dev, float* aux_dev, float* gim_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOMO_B200_H */
