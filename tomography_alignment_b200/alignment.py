"""Rigid-body alignment on top of the projection gradient.

Two layers:

* ``AlignmentUtilities`` and the ``cost_* / gradient_*`` closures keep the names, signatures and
  parameter conventions of the reference's ``utilities/alignment_functions.py`` (one view per call,
  numpy in / numpy out), so ``examples/align_rigid.py`` style drivers (``scipy.optimize.minimize`` per
  view, :40-52) run against the GPU operator unchanged.  The 9 mode pairs of the reference
  (xzpab, xzab, xz, x, z, ab, a, b, xzb; alignment_functions.py:113-485) are generated from one table.

* ``BatchedAlignment`` evaluates the cost and the masked gradient of **all views in one launch**
  (``tomo_proj_grad`` with the fused, deterministic float64 reduction) and runs a bounded, batched
  gradient descent with an Armijo backtracking line search per view -- the device-side replacement for
  the reference's serial loop of ``2 * n_proj`` single-view calls per optimiser step.

Parameter letters: x -> tx, z -> tz, p -> phi, a -> alpha, b -> beta (ty is never varied: motion along
the beam does not change the projection, examples/generate_data.py:20-23).  Gradient rows come from the
operator in the API order [tx, ty, tz, phi, alpha, beta] (utilities/ray_voxel_utilities.py:39-46).
"""
import numpy as np

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None

from .projection_operators import pose_table

# letter -> (row of the 6-row gradient, slot in (phi, alpha, beta) or xyz)
_ROW = {"x": 0, "z": 2, "p": 3, "a": 4, "b": 5}
MODES = ("xzpab", "xzab", "xz", "x", "z", "ab", "a", "b", "xzb")


def apply_parameters(mode, parameters, angles_in, xyz_in):
    """(angles [phi, alpha, beta], translations [tx, ty, tz]) for a parameter vector of ``mode``:
    offsets are ADDED to the inputs, exactly like cost_xzab etc. (alignment_functions.py:151-156)."""
    angles = np.array(angles_in, dtype=np.float64).copy()
    xyz = np.array(xyz_in, dtype=np.float64).copy()
    for k, letter in enumerate(mode):
        if letter == "x":
            xyz[..., 0] += parameters[..., k]
        elif letter == "z":
            xyz[..., 2] += parameters[..., k]
        else:
            angles[..., {"p": 0, "a": 1, "b": 2}[letter]] += parameters[..., k]
    return angles, xyz


def vary_mask(mode):
    """Boolean mask over the 6 gradient rows (the reference's ``vary_parameter`` arrays)."""
    m = np.zeros(6, dtype=bool)
    for letter in mode:
        m[_ROW[letter]] = True
    return m


class AlignmentUtilities(object):
    """alignment_functions.py:7-37: residual and (negated) gradient image of one view."""

    def __init__(self, proj, proj_obj, geometry):
        self.proj = proj
        self.proj_obj = proj_obj
        self.proj_mask = proj > 0
        self.geometry = geometry

    def cost(self, rec, angles, translations):
        phi, alpha, beta = angles
        this_proj, _ = self.proj_obj.projection_gradient(rec=rec, alpha=alpha, beta=beta, phi=phi,
                                                         xyz_shift=translations, cor_shift=self.geometry.cor_shift)
        return self.proj.ravel() - this_proj

    def gradient(self, rec, angles, translations):
        phi, alpha, beta = angles
        this_proj, this_grad = self.proj_obj.projection_gradient(rec=rec, alpha=alpha, beta=beta, phi=phi,
                                                                 xyz_shift=translations,
                                                                 cor_shift=self.geometry.cor_shift)
        residual = self.proj.ravel() - this_proj
        this_grad = this_grad * -1
        return residual, this_grad


def _make_pair(mode):
    rows = [_ROW[c] for c in mode]

    def cost(parameters, align_obj, rec, angles_in, xyz_in, scale_factor=None, return_vector=False):
        angles, translations = apply_parameters(mode, np.asarray(parameters, dtype=np.float64), angles_in, xyz_in)
        c = align_obj.cost(rec, angles, translations)
        if return_vector:
            return c
        return 0.5 * np.linalg.norm(c) ** 2

    def gradient(parameters, align_obj, rec, angles_in, xyz_in, scale_factor=None, return_vector=False):
        angles, translations = apply_parameters(mode, np.asarray(parameters, dtype=np.float64), angles_in, xyz_in)
        residual, s = align_obj.gradient(rec, angles, translations)
        s = s[rows]
        if scale_factor is None:
            scale_factor = np.ones(len(rows))
        s = s * np.asarray(scale_factor)[:, np.newaxis]
        if return_vector:
            return s.T
        return np.dot(s, residual)

    cost.__name__, gradient.__name__ = "cost_" + mode, "gradient_" + mode
    cost.__doc__ = "utilities/alignment_functions.py cost_%s: 0.5*||b - proj(parameters)||^2 of one view." % mode
    gradient.__doc__ = "utilities/alignment_functions.py gradient_%s: d cost / d parameters of one view." % mode
    return cost, gradient


for _m in MODES:
    globals()["cost_" + _m], globals()["gradient_" + _m] = _make_pair(_m)
del _m


class BatchedAlignment(object):
    """Cost and masked gradient of all views at once, and a batched bounded descent.

    ``projections`` (n_proj, n_det) are the measured data b; ``angles_in`` (n_proj, 3) = [phi, alpha, beta]
    and ``xyz_in`` (n_proj, 3) are the poses the parameter offsets are added to; ``cor_shift`` defaults to
    ``geometry.cor_shift``."""

    def __init__(self, geometry, projections, angles_in, xyz_in, mode="xzab", cor_shift=None, device=None,
                 backend=None):
        if mode not in MODES and any(c not in _ROW for c in mode):
            raise ValueError("unknown alignment mode %r" % (mode,))
        self.geometry = geometry
        self.mode = mode
        self.rows = [_ROW[c] for c in mode]
        self.angles_in = np.asarray(angles_in, dtype=np.float64).reshape(-1, 3)
        self.n_proj = self.angles_in.shape[0]
        self.xyz_in = np.asarray(xyz_in, dtype=np.float64).reshape(self.n_proj, 3)
        cor = geometry.cor_shift if cor_shift is None else cor_shift
        self.cor = np.asarray(cor, dtype=np.float64).reshape(-1, 3)[:self.n_proj]
        if backend is None:
            from .cuda_backend import CudaBackend
            backend = CudaBackend(geometry, device)
        self.backend = backend
        dev = getattr(backend, "device", "cpu")
        self.meas = torch.as_tensor(np.ascontiguousarray(np.asarray(projections, dtype=np.float32)
                                                         .reshape(self.n_proj, -1))).to(dev)
        self.evaluations = 0

    def cost_and_gradient(self, rec, parameters):
        """parameters (n_proj, P) -> (cost (n_proj,), grad (n_proj, P)) float64 numpy; one kernel launch."""
        parameters = np.asarray(parameters, dtype=np.float64).reshape(self.n_proj, len(self.mode))
        angles, xyz = apply_parameters(self.mode, parameters, self.angles_in, self.xyz_in)
        self.backend.set_poses(pose_table(angles, xyz, self.cor))
        out = self.backend.proj_grad(rec, meas=self.meas, want_proj=False, want_dproj=False)
        self.evaluations += 1
        g6 = out["grad6"].cpu().numpy()
        return out["cost"].cpu().numpy(), g6[:, self.rows]

    def minimize(self, rec, x0=None, bounds=None, maxiter=30, c1=1e-4, shrink=0.5, max_backtracks=12, eps=1e-6,
                 step0=None, verbose=False):
        """Projected gradient descent with a per-view Armijo backtracking line search, all views in lock-step.

        bounds: sequence of (lo, hi) per parameter (the reference bounds L-BFGS-B with +-3 px / +-0.02 rad,
        examples/align_rigid.py:48).  step0: initial step per parameter (defaults to a tenth of the bound
        width, or 1.0 px / 0.005 rad).  Returns (x (n_proj, P), cost (n_proj,), n_iter)."""
        P = len(self.mode)
        x = np.zeros((self.n_proj, P)) if x0 is None else np.array(x0, dtype=np.float64).reshape(self.n_proj, P)
        lo = np.full(P, -np.inf) if bounds is None else np.array([b[0] for b in bounds], dtype=np.float64)
        hi = np.full(P, np.inf) if bounds is None else np.array([b[1] for b in bounds], dtype=np.float64)
        if step0 is None:
            step0 = np.array([(0.1 * (h - l)) if np.isfinite(h - l) else (1.0 if c in "xz" else 5e-3)
                              for c, l, h in zip(self.mode, lo, hi)])
        step0 = np.asarray(step0, dtype=np.float64)
        x = np.clip(x, lo, hi)
        f, g = self.cost_and_gradient(rec, x)
        t = np.ones(self.n_proj)                     # per-view step multiplier, adapted between iterations
        it = 0
        for it in range(1, maxiter + 1):
            gn = np.abs(g).max(axis=1)
            # steepest descent scaled so that the first trial moves the largest component by step0
            d = -g / np.where(gn > 0, gn, 1.0)[:, None] * step0[None, :]
            active = gn > 0
            accepted = ~active
            x_new, f_new = x.copy(), f.copy()
            tt = t.copy()
            for _ in range(max_backtracks):
                trial = np.clip(x + tt[:, None] * d, lo, hi)
                ft, _g = self.cost_and_gradient(rec, np.where(accepted[:, None], x_new, trial))
                decrease = np.einsum("vp,vp->v", g, trial - x)
                ok = (~accepted) & (ft <= f + c1 * decrease) & (ft < f)
                x_new[ok], f_new[ok] = trial[ok], ft[ok]
                accepted |= ok
                if accepted.all():
                    break
                tt = np.where(accepted, tt, tt * shrink)
            moved = accepted & active
            rel = np.abs(f_new - f) / np.maximum(np.maximum(f_new, f), 1.0)
            # grow the step of views whose first trial was accepted, keep the backtracked one otherwise
            t = np.where(moved & (tt == t), np.minimum(t * 2.0, 4.0), np.maximum(tt, 1e-6))
            x, f = x_new, f_new
            # gradient at the accepted points (views that did not move keep theirs)
            _, g = self.cost_and_gradient(rec, x)
            if verbose:
                print("align iter %2d: cost %.6e, moved %d/%d" % (it, f.sum(), int(moved.sum()), self.n_proj))
            if (rel[active] <= eps).all() or not moved.any():
                break
        return x, f, it
