"""Device-resident reconstruction loops with the interfaces of the reference's ``recon/sirt.py`` and
``recon/cgls.py``.

The reference's solvers run unchanged on top of ``ProjectionMatrix``, but every
``A x`` / ``A^T y`` then crosses the host<->device boundary.  These classes keep the iteration on the GPU:
volume, projections, the SIRT normalisers W = 1/(A 1), V = 1/(A^T 1) and all norms live in device
memory; only the per-iteration error scalars come back.  Constructor arguments, option keys, return
values ``(rec.reshape(vox_shape), rms_error[:k])`` and the semi-convergence stop follow the reference
line by line (recon/sirt.py:9-107, recon/cgls.py:9-104).

With ``group`` (a torch.distributed process group; one rank per GPU) the views are sharded like
recon/sirt_mpi.py:36-72 / cgls_mpi.py:36-60 and the backprojections are all-reduced.
"""
import numpy as np
import torch
import torch.distributed as dist

from .projection_operators import ProjectionMatrix, normalise_poses, pose_table
from .sharding import adjoint_allreduce, check_world, shard_views


class _DeviceSolver(object):
    def __init__(self, geometry, projections, angles, xyz_shifts, options=None, group=None, device=None,
                 backend=None):
        options = {} if options is None else options
        self.geometry = geometry
        self.angles = np.asarray(angles, dtype=np.float64)
        self.xyz_shifts = np.asarray(xyz_shifts, dtype=np.float64)
        self.n_proj = self.angles.shape[0]
        self.precision = options['precision'] if 'precision' in options else np.float32
        self.voxel_mask = options['voxel_mask'] if 'voxel_mask' in options else None
        # views are sharded only when a process group is passed explicitly: group=None is an independent per-rank
        # reconstruction even inside a torchrun job
        self.group = group
        self.world = dist.get_world_size(group) if group is not None else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        check_world(self.n_proj, self.world)
        self.my_index = shard_views(self.n_proj, self.world, self.rank, options['shard'] if 'shard' in options else "contiguous")
        self.my_n_proj = int(len(self.my_index))
        if backend is None:
            from .cuda_backend import CudaBackend
            backend = CudaBackend(geometry, device)
        self.backend = backend
        self.device = getattr(backend, "device", torch.device("cpu"))
        cor = np.asarray(geometry.cor_shift, dtype=np.float64).reshape(-1, 3)
        self.backend.set_poses(pose_table(self.angles[self.my_index], self.xyz_shifts[self.my_index],
                                          cor[self.my_index]))
        self._dev = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a)).to(self.device, dt)
        proj = np.asarray(projections).reshape(self.n_proj, -1)
        self.projections = self._dev(proj[self.my_index])                 # this rank's measured views
        gt = options['ground_truth'] if 'ground_truth' in options else None
        self.ground_truth = None if gt is None else self._dev(np.asarray(gt).ravel())
        rec = options['rec'] if 'rec' in options else None
        self.rec = (torch.zeros(int(geometry.n_vox), dtype=torch.float32, device=self.device) if rec is None
                    else self._dev(np.asarray(rec).ravel()).clone())
        self.mask = None if self.voxel_mask is None else self._dev(np.asarray(self.voxel_mask).ravel().astype(np.float32))

    # A x on this rank's views, A^T y summed over all ranks
    def _A(self, x):
        if self.mask is not None:
            x = x * self.mask
        return self.backend.forward(x).reshape(self.my_n_proj, -1)

    def _At(self, y):
        v, works = adjoint_allreduce(self.backend, y, None, self.group, reduce=self.world > 1)
        for w in works:                   # slab all-reduces queued behind their kernels (sharding.adjoint_allreduce)
            w.wait()
        v = v.reshape(-1)
        if self.mask is not None:
            v = v * self.mask
        return v

    def _sum(self, t):
        """float64 sum over all ranks of a per-rank scalar tensor."""
        t = t.double().reshape(1)
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return t

    def _norm_factor(self):
        if self.ground_truth is not None:
            return float(torch.linalg.vector_norm(self.ground_truth.double()))
        return float(torch.sqrt(self._sum((self.projections.double() ** 2).sum())))

    def _result(self, rms_error, k):
        shape = tuple(int(v) for v in self.geometry.vox_shape)
        return self.rec.reshape(shape).cpu().numpy().astype(self.precision, copy=False), rms_error[:k]


class SIRT(_DeviceSolver):
    """recon/sirt.py:7-107 on the device.

    W = 1/(A 1), V = 1/(A^T 1) with zeros mapped to 0 (sirt.py:33-40; the *_mpi twin uses the threshold
    1e-8, sirt_mpi.py:69-72, which is what runs when ``group`` is given);
    x <- x + V * A^T (W * (b - A x)); optional positivity; stop when the RMS error rises (sirt.py:75-78)."""

    def __init__(self, geometry, projections, angles, xyz_shifts, options=None, group=None, device=None, backend=None):
        super().__init__(geometry, projections, angles, xyz_shifts, options, group, device, backend)
        self._initialize()

    def _initialize(self):
        ones_v = torch.ones(int(self.geometry.n_vox), dtype=torch.float32, device=self.device)
        W = self._A(ones_v)
        V = self._At(torch.ones_like(W))
        thr = 1.e-8 if self.world > 1 else 0.0
        self.W = torch.where(W > thr, 1.0 / W, torch.zeros_like(W)) if self.world > 1 else \
            torch.where(W == 0.0, torch.zeros_like(W), 1.0 / W)
        self.V = torch.where(V > thr, 1.0 / V, torch.zeros_like(V)) if self.world > 1 else \
            torch.where(V == 0.0, torch.zeros_like(V), 1.0 / V)

    def run_main_iteration(self, niter=100, make_plot=False, projections=None, positivity=False, debug=False):
        if projections is not None:
            self.projections = self._dev(np.asarray(projections).reshape(self.n_proj, -1)[self.my_index])
        norm_factor = self._norm_factor()
        rms_error = np.zeros((niter,))
        self.convergence = np.zeros((niter,))
        stop, k = 0, 0
        first_check = 1 if self.world > 1 else 0        # sirt.py:75 tests k > 0, sirt_mpi.py:118 tests k > 1
        while k < niter and not stop:
            res = self.projections - self._A(self.rec)
            back_proj = self._At(self.W * res)
            self.rec += back_proj * self.V
            if positivity:
                self.rec.clamp_(min=0.0)
            self.convergence[k] = float(torch.sqrt(self._sum((res.double() ** 2).sum())))
            if self.ground_truth is None:
                rms_error[k] = self.convergence[k] / norm_factor
            else:
                rms_error[k] = float(torch.linalg.vector_norm((self.ground_truth - self.rec).double())) / norm_factor
            if k > first_check and rms_error[k] > rms_error[k - 1]:
                stop = 1
                if self.rank == 0:
                    print('semi-convergence criterion reached: stopping at k %3d with RMSE = %4.5f' % (k, rms_error[k]))
            k += 1
        return self._result(rms_error, k)


class CGLS(_DeviceSolver):
    """recon/cgls.py:7-104 on the device (the reference file does not run as shipped: it imports a module that
    is not in the repository and reads an undefined attribute; the iteration itself is restated):
        r = b - A x, p = A^T r, gamma = |p|^2
        q = A p; alpha = gamma/|q|^2; x += alpha p; r -= alpha q; s = A^T r; beta = |s|^2/gamma; p = s + beta p
    with the reference's restart when the residual rises (cgls.py:59-67)."""

    def __init__(self, geometry, projections, angles, xyz_shift, options=None, group=None, device=None, backend=None):
        super().__init__(geometry, projections, angles, xyz_shift, options, group, device, backend)
        self.rms_error = None
        self._initialize()

    def _initialize(self):
        self._r = self.projections - self._A(self.rec)
        self._p = self._At(self._r)
        self._gamma = float((self._p.double() ** 2).sum())

    def run_main_iteration(self, make_plot=False, niter=100, debug=False):
        norm_factor = self._norm_factor()
        conv = np.zeros((niter,))
        self.rms_error = np.zeros((niter,))
        stop, k, reinit_iter = 0, 0, 0
        while not stop and k < niter:
            q = self._A(self._p)
            alpha = self._gamma / float(self._sum((q.double() ** 2).sum()))
            self.rec += alpha * self._p
            conv[k] = float(torch.sqrt(self._sum(((self.projections - self._A(self.rec)).double() ** 2).sum())))
            if k > 0 and conv[k] > conv[k - 1]:
                if self.rank == 0:
                    print('reinitializing at iteration %d' % k)
                if reinit_iter + 1 == k:
                    if self.rank == 0:
                        print('need to re-initialize at two consecutive iterations: quitting')
                    return self._result(self.rms_error, k)
                self.rec -= alpha * self._p
                self._initialize()
                reinit_iter = k
            self._r -= alpha * q
            s = self._At(self._r)
            gamma = float((s.double() ** 2).sum())
            beta = gamma / self._gamma
            self._gamma = gamma
            self._p = s + beta * self._p
            if self.ground_truth is None:
                self.rms_error[k] = float(torch.sqrt(self._sum((self._r.double() ** 2).sum()))) / norm_factor
            else:
                self.rms_error[k] = float(torch.linalg.vector_norm((self.rec - self.ground_truth).double())) / norm_factor
            k += 1
        return self._result(self.rms_error, k)


def soft_thresholding(x, _lambda):
    """recon/regularized.py:433-441: sgn(x) max(|x| - lambda, 0) (torch tensor in, torch tensor out)."""
    return torch.sign(x) * torch.clamp(x.abs() - _lambda, min=0.0)


def scalar_search_armijo(phi, phi0, derphi0, c1=1e-4, alpha0=1.0, amin=0.0):
    """The interpolating Armijo search behind scipy.optimize's ``line_search_armijo`` (Nocedal & Wright,
    Numerical Optimization, section 3.5: quadratic, then cubic interpolation), which recon/regularized.py:188
    calls.  Returns (alpha, phi(alpha)) or (None, phi) when no acceptable step larger than ``amin`` exists."""
    phi_a0 = phi(alpha0)
    if phi_a0 <= phi0 + c1 * alpha0 * derphi0:
        return alpha0, phi_a0
    alpha1 = -derphi0 * alpha0 ** 2 / 2.0 / (phi_a0 - phi0 - derphi0 * alpha0)
    phi_a1 = phi(alpha1)
    if phi_a1 <= phi0 + c1 * alpha1 * derphi0:
        return alpha1, phi_a1
    while alpha1 > amin:
        factor = alpha0 ** 2 * alpha1 ** 2 * (alpha1 - alpha0)
        a = alpha0 ** 2 * (phi_a1 - phi0 - derphi0 * alpha1) - alpha1 ** 2 * (phi_a0 - phi0 - derphi0 * alpha0)
        a = a / factor
        b = -alpha0 ** 3 * (phi_a1 - phi0 - derphi0 * alpha1) + alpha1 ** 3 * (phi_a0 - phi0 - derphi0 * alpha0)
        b = b / factor
        alpha2 = (-b + np.sqrt(abs(b ** 2 - 3 * a * derphi0))) / (3.0 * a)
        phi_a2 = phi(alpha2)
        if phi_a2 <= phi0 + c1 * alpha2 * derphi0:
            return alpha2, phi_a2
        if (alpha1 - alpha2) > alpha1 / 2.0 or (1 - alpha2 / alpha1) < 0.96:
            alpha2 = alpha1 / 2.0
        alpha0, alpha1, phi_a0, phi_a1 = alpha1, alpha2, phi_a1, phi_a2
    return None, phi_a1


class RegularizedRecon(_DeviceSolver):
    """recon/regularized.py:13-441 on the device: FISTA with a TV proximal step (``run_fista``), Tikhonov
    gradient descent (``run_tikhonov_gd``) and the two Lasso loops (``run_lasso_ista``, ``run_lasso_accelerated``).
    u_k = prox_{gamma g}(x_{k-1} + gamma A^T (b - A x_{k-1})); x_k = u_k + (t_{k-1} - 1)/t_k (u_k - u_{k-1}), gamma = 1/hyper;
    the prox is tv_denoise.denoise_fista(weight = gamma * beta_tv, niter = niter_tv) (regularized.py:84-103).
    With ``group`` the views are sharded and the backprojection all-reduced like recon/regularized_mpi.py:110-116;
    every rank then applies the (deterministic) prox redundantly, so no broadcast is needed (:118-137)."""

    def __init__(self, geometry, projections, angles, xyz_shifts, options=None, group=None, device=None, backend=None,
                 tv_ops=None):
        super().__init__(geometry, projections, angles, xyz_shifts, options, group, device, backend)
        self.tv_ops = tv_ops
        self.norm_factor = self._norm_factor()

    def run_fista(self, niter=100, make_plot=False, hyper=1.e4, beta_tv=1.0, niter_tv=20):
        from . import tv_denoise
        shape = tuple(int(v) for v in self.geometry.vox_shape)
        gamma = 1. / hyper
        t = 1.0
        rms_error = np.zeros(niter, )
        self.total_cost = np.zeros(niter, )
        self.data_fidelity_cost = np.zeros(niter, )
        u_old = self.rec.clone()
        k, stop = 0, 0
        while k < niter and not stop:
            res = self.projections - self._A(self.rec)
            back_proj = self._At(res)
            x_tmp = self.rec + gamma * back_proj
            u = tv_denoise.denoise_fista(x_tmp.reshape(shape), weight=gamma * beta_tv, niter=niter_tv, ops=self.tv_ops)
            t_old = t
            t = 0.5 * (1.0 + np.sqrt(1 + 4 * t_old ** 2))
            u = torch.as_tensor(u).to(self.rec.device).reshape(-1)
            self.rec = u + (t_old - 1) / t * (u - u_old)
            u_old = u
            self.data_fidelity_cost[k] = 0.5 * float(self._sum((res.double() ** 2).sum()))
            tv_value = beta_tv * tv_denoise.tv_norm_3d(self.rec.reshape(shape))
            self.total_cost[k] = self.data_fidelity_cost[k] + tv_value
            if self.ground_truth is None:
                rms_error[k] = np.sqrt(2 * self.data_fidelity_cost[k]) / self.norm_factor
            else:
                rms_error[k] = float(torch.linalg.vector_norm((self.ground_truth - self.rec).double())) / self.norm_factor
            if k > 0 and rms_error[k] > rms_error[k - 1]:
                stop = 1
                if self.rank == 0:
                    print('semi-convergence criterion reached: stopping at k %3d with RMSE = %4.5f' % (k, rms_error[k]))
            k += 1
        return self.rec.cpu().numpy().astype(self.precision, copy=False), rms_error[:k]

    # ---- shared bookkeeping of the three loops below (regularized.py:204-214, 290-299, 381-392) ----
    def _track(self, k, res_norm, rms_error, convergence, first_check=1):
        convergence[k] = res_norm
        if self.ground_truth is None:
            rms_error[k] = convergence[k] / self.norm_factor
        else:
            rms_error[k] = float(torch.linalg.vector_norm((self.ground_truth - self.rec).double())) / self.norm_factor
        if k > first_check and rms_error[k] > rms_error[k - 1]:
            if self.rank == 0:
                print('semi-convergence criterion reached: stopping at k %3d with RMSE = %4.5f' % (k, rms_error[k]))
            return 1
        return 0

    def _flat_result(self, rms_error, k, reshape):
        rec = self.rec.cpu().numpy().astype(self.precision, copy=False)
        if reshape:
            rec = rec.reshape(tuple(int(v) for v in self.geometry.vox_shape))
        return rec, rms_error[:k]

    def run_tikhonov_gd(self, niter=100, reg_param=1.0, positivity=False, make_plot=False):
        """regularized.py:156-237: x* = argmin 0.5|Ax - b|^2 + 0.5 lambda |x|^2 by gradient descent with the Armijo
        search of scipy (``line_search_armijo``, alpha0 = 1).

        The objective along the search direction is a quadratic in alpha, so it is evaluated from six inner
        products and ONE extra forward projection A g per iteration instead of one projection per trial step
        (the reference calls ``my_tikh_f`` per trial; that helper subtracts a 2-D array from a flat one and only
        runs for n_proj = 1, the intended flat residual is what is restated here).  Returns (rec flat, rms_error),
        like the reference."""
        rms_error, convergence = np.zeros(niter, ), np.zeros(niter, )
        stop, k = 0, 0
        while k < niter and not stop:
            res = self.projections - self._A(self.rec)                     # b - A x
            grad = -self._At(res) + reg_param * self.rec                   # A^T (A x - b) + lambda x
            Ag = self._A(grad)
            rr = float(self._sum((res.double() ** 2).sum()))
            rAg = float(self._sum((res.double() * Ag.double()).sum()))
            AgAg = float(self._sum((Ag.double() ** 2).sum()))
            xx = float((self.rec.double() ** 2).sum())
            xg = float((self.rec.double() * grad.double()).sum())
            gg = float((grad.double() ** 2).sum())
            # f(x - a g) = 0.5 |res + a A g|^2 + 0.5 lambda |x - a g|^2
            phi = lambda a: 0.5 * (rr + 2 * a * rAg + a * a * AgAg) + 0.5 * reg_param * (xx - 2 * a * xg + a * a * gg)
            cost = 0.5 * (rr + reg_param * xx)
            alpha, _ = scalar_search_armijo(phi, cost, -gg, alpha0=1.0)
            if alpha is None:
                print('line search failed at iteration %3d' % (k))
                break
            self.rec = self.rec - alpha * grad
            if positivity:
                self.rec.clamp_(min=0.0)
            stop = self._track(k, np.sqrt(rr), rms_error, convergence)
            k += 1
        return self._flat_result(rms_error, k, reshape=False)

    def _backtrack_lasso(self, t, beta, g0, dg0, _lambda):
        """regularized.py:317-332: shrink t until g(x+) <= g(x) - <grad g, G_t> + |G_t|^2 / (2 t), G_t = x - x+."""
        g0 = 0.5 * float(self._sum((g0.double() ** 2).sum()))
        xp = self.rec
        while t > 1.e-16:
            xp = soft_thresholding(self.rec - t * dg0, t * _lambda)
            Gt = (self.rec - xp).double()
            g = 0.5 * float(self._sum(((self._A(xp) - self.projections).double() ** 2).sum()))
            gp = g0 - float((dg0.double() * Gt).sum()) + (0.5 / t) * float((Gt ** 2).sum())
            if g <= gp:
                return xp, t, True
            t *= beta
        return xp, t, False

    def run_lasso_ista(self, niter=100, reg_param=1.0, alpha0=1.0, beta=0.5, make_plot=False):
        """regularized.py:239-315: proximal gradient descent for 0.5|Ax - b|^2 + lambda |x|_1 with backtracking;
        the step sizes are kept in ``self.step_size`` (the reference plots them)."""
        rms_error, convergence = np.zeros(niter, ), np.zeros(niter, )
        self.step_size = np.zeros(niter, )
        stop, k = 0, 0
        while k < niter and not stop:
            res = self._A(self.rec) - self.projections
            grad = self._At(res)
            _, alpha, success = self._backtrack_lasso(alpha0, beta, res, grad, reg_param)
            self.step_size[k] = alpha
            if not success:
                print('line search failed to converge')
                break
            self.rec = soft_thresholding(self.rec - alpha * grad, alpha * reg_param)
            stop = self._track(k, float(torch.sqrt(self._sum((res.double() ** 2).sum()))), rms_error, convergence)
            k += 1
        return self._flat_result(rms_error, k, reshape=True)

    def run_lasso_accelerated(self, niter=100, reg_param=1.0, alpha0=1.0, beta=0.5, make_plot=False):
        """regularized.py:334-413: the accelerated variant, v = x_1 + (k - 2)/(k + 1) (x_1 - x_0) with the step found by
        the same backtracking at the current iterate.  Returns (rec flat, rms_error), like the reference."""
        rms_error, convergence = np.zeros(niter, ), np.zeros(niter, )
        x_0, x_1 = torch.zeros_like(self.rec), torch.zeros_like(self.rec)
        stop, k = 0, 0
        while k < niter and not stop:
            res = self._A(self.rec) - self.projections
            grad = self._At(res)
            _, alpha, success = self._backtrack_lasso(alpha0, beta, res, grad, reg_param)
            if not success:
                print('line search failed to converge')
                break
            v = x_1 + (k - 2) / (k + 1) * (x_1 - x_0)
            self.rec = soft_thresholding(v - alpha * grad, alpha * reg_param)
            x_0, x_1 = x_1, self.rec.clone()
            stop = self._track(k, float(torch.sqrt(self._sum((res.double() ** 2).sum()))), rms_error, convergence)
            k += 1
        return self._flat_result(rms_error, k, reshape=False)
