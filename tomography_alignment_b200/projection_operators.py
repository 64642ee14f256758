"""Drop-in for the reference's ``utilities/projection_operators.py`` (the L2 operator API).

``ProjectionMatrix`` keeps the reference's constructor and the signatures, defaults and side effects
of ``projection_matrix`` (utilities/projection_operators.py:22-76) and ``projection_gradient``
(:112-122).  Instead of a scipy CSR matrix (7.5*n_vox*n_proj non-zeros: 5.4 TiB at 512^3 x 720),
``projection_matrix`` returns a matrix-free ``ProjectionOperator`` that survives every way the
reference's solvers use the matrix:

    sparse.csr_matrix.dot(A, x)                                   recon/sirt.py:59, cgls.py:72
    sparse.csc_matrix.dot(sparse.csr_matrix.transpose(A), y)      recon/sirt.py:61, cgls.py:54
    A.shape

numpy in -> numpy out, CPU tensor in -> (pinned) CPU tensor out: host<->device copies happen inside the
call, overlapped with the kernels in view chunks (like a scipy matvec on host arrays, minus the wait);
CUDA tensor in -> CUDA tensor out, no copies (device-resident solvers, bench.py's kernel timing).  All arithmetic runs in libtomo_b200.so on the GPU; there is no CPU
path -- without CUDA, applying the operator raises.
"""
import numpy as np

try:
    import torch
except ImportError:  # pragma: no cover - torch is part of the image
    torch = None


def normalise_poses(geometry, alpha=None, beta=None, phi=None, xyz_shift=None):
    """Argument handling of ProjectionMatrix.projection_matrix, utilities/projection_operators.py:24-52.

    Returns (n_proj, angles (n_proj, 3) = [phi, alpha, beta], xyz_shift (n_proj, 3))."""
    if phi is None:
        n_proj = geometry.n_proj
        phi = np.linspace(0., np.pi, n_proj)
    else:
        n_proj = np.size(phi)
    if alpha is None:
        alpha = np.zeros_like(phi)
    if beta is None:
        beta = np.zeros_like(phi)
    if xyz_shift is None:
        xyz_shift = np.zeros((n_proj, 3))
    phi = np.squeeze(phi)
    alpha = np.squeeze(alpha)
    beta = np.squeeze(beta)
    xyz_shift = np.squeeze(xyz_shift)
    if n_proj == 1:
        phi = np.array([phi])
        alpha = np.array([alpha])
        beta = np.array([beta])
        xyz_shift = np.array([xyz_shift])
    angles = np.array([phi, alpha, beta]).T
    return n_proj, angles, xyz_shift


def pose_table(angles, xyz_shift, cor_shift):
    """(n_proj, 9) float64 rows [phi, alpha, beta, tx, ty, tz, cor_x, cor_y, cor_z] for the C ABI."""
    angles = np.asarray(angles, dtype=np.float64).reshape(-1, 3)
    n = angles.shape[0]
    poses = np.zeros((n, 9), dtype=np.float64)
    poses[:, 0:3] = angles
    poses[:, 3:6] = np.asarray(xyz_shift, dtype=np.float64).reshape(n, 3)
    poses[:, 6:9] = np.asarray(cor_shift, dtype=np.float64).reshape(n, 3)
    return poses


def _rot_z(a):
    return np.array([(np.cos(a), -np.sin(a), 0.), (np.sin(a), np.cos(a), 0.), (0., 0., 1.)])


def _rot_x(a):
    return np.array([(1., 0., 0.), (0., np.cos(a), -np.sin(a)), (0., np.sin(a), np.cos(a))])


def _rot_y(a):
    return np.array([(np.cos(a), 0., np.sin(a)), (0., 1., 0.), (-np.sin(a), 0., np.cos(a))])


def reference_sample_counts(geometry, poses):
    """Per view: (n, r_length[0]) with n = int(r_length[0] / step_size), the number of samples the reference
    marches along every ray (utilities/ray_voxel_utilities.py:85-88).

    In exact arithmetic r_length[0] = 2*sy, so for the usual step sizes r_length[0]/step_size sits ON an
    integer and int() returns 2*sy/step or one less depending on the last-bit rounding of the numpy
    expression.  To march exactly the samples the reference marches, the expression is evaluated here with
    the reference's own numpy calls (rotations.py:9-48, transform_points :6-12, forward_sparse :72-88) on the
    first columns of the source / detector grids -- np.dot on (3,3)x(3,K) gives column 0 the same bits for
    every K >= 2 (tests/test_views_host.py checks this against the full-width evaluation) -- and handed to
    the C ABI in the pose record instead of being recomputed in C++."""
    poses = np.asarray(poses, dtype=np.float64)
    poses = poses.reshape(-1, poses.shape[-1])
    k = min(int(geometry.n_det), 4)
    org = np.asarray(geometry.vox_origin, dtype=np.float64)[:, np.newaxis]
    out = np.zeros((poses.shape[0], 2))
    for i, ps in enumerate(poses):
        phi, alpha, beta, t, cor_x = ps[0], ps[1], ps[2], ps[3:6], ps[6]
        src = np.array(geometry.source_centers[:, :k], dtype=np.float64)
        det = np.array(geometry.det_centers[:, :k], dtype=np.float64)
        src[0, :] += cor_x
        det[0, :] += cor_x
        rot_pa = np.dot(_rot_z(phi), _rot_x(alpha))
        p0 = np.dot(rot_pa, np.dot(_rot_y(beta), src) + t[:, np.newaxis]) - org
        p1 = np.dot(rot_pa, np.dot(_rot_y(beta), det) + t[:, np.newaxis]) - org
        r_length = np.linalg.norm(p1 - p0, axis=0)
        out[i, 0] = int(r_length[0] / geometry.step_size)
        out[i, 1] = r_length[0]
    return out


def full_pose_table(geometry, poses, flags=0.0):
    """(n_proj, 9) rows of ``pose_table`` -> the (n_proj, TOMO_POSE_STRIDE = 12) records of the C ABI:
    columns 9, 10 = the reference's sample count and r_length[0] (``reference_sample_counts``), 11 = per-view flags
    (non-zero: the view may take the z-quad ray kernels)."""
    poses = np.asarray(poses, dtype=np.float64)
    poses = poses.reshape(-1, poses.shape[-1] if poses.ndim > 1 else 9)
    if poses.shape[1] == 12:
        return np.ascontiguousarray(poses)
    full = np.zeros((poses.shape[0], 12), dtype=np.float64)
    full[:, :9] = poses[:, :9]
    full[:, 9:11] = reference_sample_counts(geometry, poses)
    full[:, 11] = flags
    return full


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _numel(x):
    return x.numel() if _is_torch(x) else np.size(x)


class ProjectionOperator(object):
    """Matrix-free A (or A^T) of shape (n_proj*n_det, n_vox) ((n_vox, n_proj*n_det) transposed).

    Duck-types the parts of scipy's csr/csc matrices the reference's solvers touch through unbound
    methods: ``ndim``, ``shape``, ``data``/``indices``/``indptr`` (placeholders) and
    ``_csc_container`` for ``sparse.csr_matrix.transpose``; ``__matmul__`` for ``_spbase.dot``.
    """
    ndim = 2

    def __init__(self, backend, n_proj, n_det, n_vox, precision=np.float32, voxel_mask=None,
                 transposed=False, all_masked=False, poses=None, _state=None):
        self._backend = backend
        # The operator owns its poses (like the reference's CSR matrix owns its entries): a later
        # projection_matrix() call on the same ProjectionMatrix re-poses the shared backend, so every application
        # first re-binds this operator's own view table (a device tensor kept in _state, shared with the transposes).
        self._poses = None if poses is None else np.array(poses, dtype=np.float64, copy=True)
        self._state = {"views": None} if _state is None else _state
        self._n_proj, self._n_det, self._n_vox = int(n_proj), int(n_det), int(n_vox)
        self._transposed = bool(transposed)
        self._precision = precision
        self._mask = voxel_mask          # bool ndarray (n_vox,) or None
        self._mask_dev = None
        self._all_masked = all_masked
        rows, cols = self._n_proj * self._n_det, self._n_vox
        self.shape = (cols, rows) if transposed else (rows, cols)
        self.dtype = np.dtype(precision)
        # placeholders read by sparse.csr_matrix.transpose before it calls _csc_container
        self.data = self.indices = self.indptr = None

    # -- scipy unbound-method idioms -------------------------------------------------------------
    def _csc_container(self, arg1, shape=None, copy=False):
        return self.transpose()

    _csr_container = _csc_container

    def transpose(self, axes=None, copy=False):
        return ProjectionOperator(self._backend, self._n_proj, self._n_det, self._n_vox, self._precision,
                                  self._mask, not self._transposed, self._all_masked, self._poses, self._state)

    @property
    def T(self):
        return self.transpose()

    def dot(self, other):
        return self.__matmul__(other)

    def __matmul__(self, other):
        return self._rmatvec(other) if self._transposed else self._matvec(other)

    def matvec(self, x):
        return self.__matmul__(x)

    def rmatvec(self, y):
        return self.transpose().__matmul__(y)

    # -- application -----------------------------------------------------------------------------
    def _bind(self):
        """Make the backend apply THIS operator's poses (no-op while nobody else has re-posed it)."""
        b = self._backend
        if self._poses is None or getattr(b, "_bound_state", None) is self._state:
            return
        if self._state["views"] is not None and hasattr(b, "bind_views"):
            b.bind_views(self._state["views"], self._n_proj, self._state.get("kinds", 0))
        else:
            b.set_poses(self._poses)
            self._state["views"] = getattr(b, "views", None)
            self._state["kinds"] = getattr(b, "kinds", 0)
        b._bound_state = self._state

    def _mask_on(self, like):
        if self._mask_dev is None:
            self._mask_dev = torch.as_tensor(self._mask.astype(np.float32), device=like.device)
        return self._mask_dev

    def _matvec(self, x):
        if _numel(x) != self._n_vox:
            raise ValueError("dimension mismatch: operator has %d columns, vector has %d entries"
                             % (self._n_vox, _numel(x)))
        was_torch = _is_torch(x)
        self._bind()
        if not (was_torch and x.is_cuda) and self._mask is None and hasattr(self._backend, "forward_host"):
            # host buffers: copies overlapped with the kernels in view chunks (cuda_backend.forward_host)
            y = self._backend.forward_host(x).reshape(-1)
            return y if was_torch else y.numpy().astype(np.result_type(self._precision, np.asarray(x).dtype), copy=False)
        xd = self._backend._as_vol(x if was_torch else np.ascontiguousarray(np.asarray(x), dtype=np.float32))
        if self._mask is not None:
            # dropping the masked columns' entries (projection_operators.py:60-70) == zeroing x there
            xd = xd.reshape(-1) * (0.0 if self._all_masked else self._mask_on(xd))
        y = self._backend.forward(xd).reshape(-1)
        if was_torch:
            return y
        return y.cpu().numpy().astype(np.result_type(self._precision, np.asarray(x).dtype), copy=False)

    def _rmatvec(self, y):
        if _numel(y) != self._n_proj * self._n_det:
            raise ValueError("dimension mismatch: operator has %d rows, vector has %d entries"
                             % (self._n_proj * self._n_det, _numel(y)))
        was_torch = _is_torch(y)
        self._bind()
        if not (was_torch and y.is_cuda) and self._mask is None and hasattr(self._backend, "adjoint_host"):
            v = self._backend.adjoint_host(y).reshape(-1)
            return v if was_torch else v.numpy().astype(np.result_type(self._precision, np.asarray(y).dtype), copy=False)
        yd = self._backend._as_proj(y if was_torch else np.ascontiguousarray(np.asarray(y), dtype=np.float32))
        v = self._backend.adjoint(yd).reshape(-1)
        if self._mask is not None:
            v = v * (0.0 if self._all_masked else self._mask_on(v))
        if was_torch:
            return v
        return v.cpu().numpy().astype(np.result_type(self._precision, np.asarray(y).dtype), copy=False)


class ProjectionMatrix(object):
    """Same interface as the reference's class (utilities/projection_operators.py:11-122).

    ``backend`` is for tests that exercise the host logic without a GPU; by default the CUDA backend
    is created on first use and raises if no CUDA device is present."""

    def __init__(self, geometry, precision=np.float32, device=None, backend=None):
        self.geometry = geometry
        self.precision = precision
        self.n_proj = None
        self.angles = None
        self.xyz_shift = None
        self.voxel_mask = None
        self._device = device
        self._backend = backend
        self._grad_backend = None

    def _get_backend(self):
        if self._backend is None:
            from .cuda_backend import CudaBackend
            self._backend = CudaBackend(self.geometry, self._device)
        return self._backend

    def _aux_backend(self):
        """Backend of the single-call entry points (projection_gradient, forward_project, ...): a CudaBackend of its own, so
        that these calls do not disturb the view table of the operators projection_matrix() returned; an injected test
        backend (tests exercising the host logic without a GPU) is shared."""
        if self._grad_backend is None:
            if self._backend is not None and type(self._backend).__name__ != "CudaBackend":
                self._grad_backend = self._backend
            else:
                from .cuda_backend import CudaBackend
                self._grad_backend = CudaBackend(self.geometry, self._device)
        return self._grad_backend

    def projection_matrix(self, alpha=None, beta=None, phi=None, xyz_shift=None, voxel_mask=None):
        self.n_proj, self.angles, self.xyz_shift = normalise_poses(self.geometry, alpha, beta, phi, xyz_shift)
        self.voxel_mask = voxel_mask
        cor = np.asarray(self.geometry.cor_shift, dtype=np.float64)
        if cor.ndim == 1:      # a geometry whose cor_shift was replaced by one row (sirt_mpi.py:46-47 does this)
            cor = np.tile(cor, self.n_proj).reshape(self.n_proj, 3)
        poses = pose_table(self.angles, self.xyz_shift, cor[:self.n_proj])
        backend = self._get_backend()
        backend.set_poses(poses)
        state = {"views": getattr(backend, "views", None), "kinds": getattr(backend, "kinds", 0)}
        backend._bound_state = state
        mask, all_masked = None, False
        if voxel_mask is not None:
            mask = np.asarray(voxel_mask).ravel().astype(bool)
            if np.sum(mask) == 0:
                print('entire object is masked')     # projection_operators.py:63-65
                all_masked = True
        return ProjectionOperator(backend, self.n_proj, self.geometry.n_det, self.geometry.n_vox,
                                  self.precision, mask, False, all_masked, poses, state)

    def projection_gradient(self, rec, alpha, beta, phi, xyz_shift, cor_shift):
        """One view: (proj (n_det,), grad (6, n_det)), gradient rows [tx, ty, tz, phi, alpha, beta]
        (utilities/projection_operators.py:112-122, utilities/ray_voxel_utilities.py:39-46)."""
        out = self.projection_gradient_batch(rec, np.array([[phi, alpha, beta]], dtype=np.float64),
                                             np.asarray(xyz_shift, dtype=np.float64).reshape(1, 3),
                                             np.asarray(cor_shift, dtype=np.float64).reshape(-1)[:3].reshape(1, 3))
        proj, grad = out["proj"].reshape(-1), out["dproj"].reshape(6, -1)
        if _is_torch(rec):
            return proj, grad
        return (proj.cpu().numpy().astype(self.precision, copy=False),
                grad.cpu().numpy().astype(self.precision, copy=False))

    def forward_project(self, rec, alpha, beta, phi, xyz_shift, cor_shift=None):
        """The orphan matrix-free forward projector ``forward_project`` (src/forward_projection.f90:1-68), all views at once,
        with ITS semantics where they differ from the live path: the number of samples per ray is NINT(r_length / step_size)
        (:44; the live path truncates, ray_voxel_utilities.py:88) and ``cor_shift`` is accepted but never applied (:1,10 --
        the argument is ignored here as well).  Returns ax (n_proj, n_det) like the Fortran's ``ax(n_proj, n_rays)``.
        Arithmetic is the float64-setup / float32-interpolation of tomo_forward, i.e. at least as accurate as the
        all-float32 Fortran."""
        n_proj, angles, xyz = normalise_poses(self.geometry, alpha, beta, phi, xyz_shift)
        poses = np.zeros((n_proj, 12), dtype=np.float64)
        poses[:, 0:3] = angles
        poses[:, 3:6] = np.asarray(xyz, dtype=np.float64).reshape(n_proj, 3)
        g = self.geometry
        r_length = float(g._y_det) - float(g._y_source)          # |R (d - s)| = 2 sy for every pose
        poses[:, 9] = float(np.rint(r_length / float(g.step_size)))
        poses[:, 10] = r_length
        be = self._aux_backend()
        be.set_poses(poses)
        was_torch = _is_torch(rec)
        ax = be.forward(be._as_vol(rec if was_torch else np.ascontiguousarray(np.asarray(rec), dtype=np.float32)))
        ax = ax.reshape(n_proj, -1)
        return ax if was_torch else ax.cpu().numpy().astype(self.precision, copy=False)

    def voxel_projection_gradient(self, rec, alpha, beta, phi, xyz_shift, cor_shift):
        """The voxel-driven twin the reference keeps next to the ray-driven path
        (utilities/voxel_utilities.py:82-108, imported as vox_forward_proj_grad in projection_operators.py:8 but
        never called): (det_img.ravel() with x fastest, gradient (6, n_det)), rows [sx, sy, sz, theta, alpha, beta]."""
        self._aux_backend()
        cor = np.asarray(cor_shift, dtype=np.float64).reshape(-1)[:3].reshape(1, 3)
        self._grad_backend.set_poses(pose_table(np.array([[phi, alpha, beta]], dtype=np.float64),
                                                np.asarray(xyz_shift, dtype=np.float64).reshape(1, 3), cor))
        det, grad = self._grad_backend.voxel_splat(rec, want_grad=True)
        det, grad = det.reshape(-1), grad.reshape(6, -1)
        if _is_torch(rec):
            return det, grad
        return (det.cpu().numpy().astype(self.precision, copy=False), grad.cpu().numpy().astype(self.precision, copy=False))

    def projection_gradient_batch(self, rec, angles, xyz_shift, cor_shift=None, meas=None,
                                  want_dproj=True, want_proj=True):
        """All views at once on the device (what the reference does with n_proj separate calls,
        examples/align_rigid.py:40-49).  angles (n, 3) = [phi, alpha, beta].  With ``meas`` also
        returns grad6 (n, 6) = sum_rays (-dproj) * (meas - proj) and cost (n,) = 0.5*||meas - proj||^2,
        i.e. gradient_* / cost_* of utilities/alignment_functions.py before parameter masking."""
        self._aux_backend()
        angles = np.asarray(angles, dtype=np.float64).reshape(-1, 3)
        n = angles.shape[0]
        if cor_shift is None:
            cor_shift = np.asarray(self.geometry.cor_shift, dtype=np.float64).reshape(-1, 3)[:n]
        self._grad_backend.set_poses(pose_table(angles, xyz_shift, cor_shift))
        return self._grad_backend.proj_grad(rec, meas=meas, want_proj=want_proj, want_dproj=want_dproj)
