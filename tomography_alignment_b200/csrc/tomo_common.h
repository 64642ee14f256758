// tomo_common.h -- view-table layout and small helpers shared by host and device code.
#pragma once
#include <stdint.h>
#include "../../include/tomo_b200.h"

// Offsets (in doubles) into one view record of TOMO_VIEW_STRIDE doubles.
// Sample lattice of a view (voxel-index coordinates, i.e. physical position minus vox_origin):
//     p(ix, iz, j) = P00 + ix*U + iz*W + j*D ,   j = 0 .. n-1
// which is p0 + j*step*r_hat of utilities/ray_voxel_utilities.py:74-94 with the affine dependence
// of the source point on the detector pixel made explicit.
enum {
    V_P00  = 0,    // 3   lattice origin  M (Ry s00 + t) - vox_origin
    V_U    = 3,    // 3   d p / d ix
    V_W    = 6,    // 3   d p / d iz
    V_D    = 9,    // 3   d p / d j = step_size * r_hat
    V_N    = 12,   // 1   n = int(r_length / step_size)                    (ray_voxel_utilities.py:88)
    V_RLEN = 13,   // 1   r_length of ray 0
    V_INVD = 14,   // 3   1/D per axis (0 when D == 0)
    V_M    = 17,   // 9   [k][a], k = tx,ty,tz: d p_a / d t_k = (Rz Rx)[a][k]  (ray_voxel_utilities.py:37-40)
    V_E    = 26,   // 9   [k][a], k = phi,alpha,beta: pose derivative at ix = iz = 0, j = 0  (:42-45)
    V_F    = 35,   // 9   [k][a]  its slope in ix
    V_H    = 44,   // 9   [k][a]  its slope in iz
    V_K    = 53,   // 9   [k][a]  its slope in j: der[6+k] * step_size / r_length   (:46-48, :148-151)
    V_LINV = 62,   // 9   row-major inverse of L = [U W D]: lattice coords (ix, iz, j) = Linv (p - P00)
    V_RB   = 71,   // 3   candidate half-widths in lattice coords: sum_a |Linv[k][a]|
    V_VROT = 74,   // 9   row-major Ry(b) Rx(a) Rz(p)      (src/external_back_projection.f90:17-25)
    V_VTR  = 83,   // 3   Ry(b) t
    V_NCOL = 86,   // 1   ray colour classes of the tile-scatter backprojector (0: not applicable)
    V_NUNCOL = 87, // 1   number of views of the whole table with V_NCOL == 0 (same in every record, so a
                   //     pointer to any record is a valid sub-table)
    V_SPL  = 96,   // 48  voxel-driven derivative_rigid (voxel_utilities.py:23-48): [k][c][4], k = sx,sy,sz,theta,alpha,beta,
                   //     c = 0 (x' row) / 1 (z' row): value = v[0]*cx + v[1]*cy + v[2]*cz + v[3] at voxel centre (cx,cy,cz)
    V_SORG = 144,  // 2   vox_origin - cor_shift, x and z components (voxel_utilities.py:61,90)
    V_SEP  = 146,  // 1   1.0 when the view has no tilt (alpha = beta = 0 exactly): W_x = W_y = U_z = D_z = 0, so z decouples
                   //     from (x, y) and the separable kernels apply
    V_NSEP = 147,  // 1   number of views of the whole table with V_SEP == 1 (same in every record)
    V_VBOK = 148,  // 1   1.0 when the detector footprint of a TOMO_VB_X x _Y x _Z voxel brick under the voxel-driven
                   //     transform fits the TOMO_VB_TX x TOMO_VB_TZ box the TMA-staged backprojector loads per view
    V_NVBIG = 149, // 1   number of views of the whole table with V_VBOK == 0 (same in every record)
    V_ZQ   = 150,  // 1   1.0 when the view qualifies for the z-quad ray kernels (zq_core.h): W ~ (0, 0, 1), i.e. four z-adjacent
                   //     rays share an (x, y) cell and sit in consecutive z cells for almost every sample
    V_END  = 151
};

static_assert(V_END <= TOMO_VIEW_STRIDE, "view record overflows TOMO_VIEW_STRIDE");

// Floats of slack before and after the padded volume inside its buffer (zeros): the z-quad kernels load aligned 8-plane
// windows around cells up to a few planes beside the zero border (zq_core.h).  128 bytes keep the 16-byte alignment.
#define TOMO_PAD_HEAD 32
#define TOMO_PAD_TAIL 32

// Number of z planes of the padded volume the kernels address (zero border included, rounded up to 32).
static inline
#ifdef __CUDACC__
__host__ __device__
#endif
int tomo_nzp(int nz) { return ((nz + 2 * TOMO_PAD + 31) / 32) * 32; }

// Pitches of the padded volume: *nyp rows per x plane, *syp floats per row (x stride = nyp * syp floats).  The natural pitches are
// (ny + 2 TOMO_PAD, tomo_nzp(nz)).  The ray kernels have variants with compile-time strides for the pitches of the cubes
// 64^3 ... 1024^3 (ray_core.h: all eight corner loads off one address register), so a volume whose (ny, nz) fits the next of those
// cubes takes that cube's pitches when it costs at most twice the memory of the natural layout; the rows / planes beyond the
// natural ones are never addressed.
static inline
#ifdef __CUDACC__
__host__ __device__
#endif
void tomo_pad_pitch(int ny, int nz, int* nyp, int* syp)
{
    const int ny0 = ny + 2 * TOMO_PAD, nz0 = tomo_nzp(nz);
    *nyp = ny0; *syp = nz0;
    for (int n = 64; n <= 1024; n *= 2) {
        const int cy = n + 2 * TOMO_PAD, cz = tomo_nzp(n);
        if (ny0 <= cy && nz0 <= cz) {
            if ((double)cy * cz <= 2.0 * (double)ny0 * nz0) { *nyp = cy; *syp = cz; }
            return;
        }
    }
}

// Tile of the scatter backprojector (back_kernels.cu); the host needs the xy extent to bound the
// z drift of a ray inside a tile when it counts colour classes.
#ifndef TOMO_BT_X
#define TOMO_BT_X 20     // 20 x 16 x 30 with 7 warps measured best on B200 (profiles/README.md)
#endif
#ifndef TOMO_BT_Y
#define TOMO_BT_Y 16
#endif
#ifndef TOMO_BT_Z
#define TOMO_BT_Z 30
#endif
#ifndef TOMO_BT_WARPS
#define TOMO_BT_WARPS 7      // warps per block of the tile kernel (7 measured 1.7 % faster than 8 and 2.7 % faster than 6)
#endif

// Voxel brick of the TMA-staged voxel-driven backprojector and the detector box (x' rows of z' pixels) it stages
// per view; the host marks the views whose footprint fits (V_VBOK).  The box starts at a z' index that is a multiple of 4
// (TMA needs a 16-byte aligned start along the innermost dimension), which costs up to 3 extra columns.
#define TOMO_VB_X 16
#define TOMO_VB_Y 16
#define TOMO_VB_Z 32
#define TOMO_VB_TX 32
#define TOMO_VB_TZ 44
