"""CPU ORACLE (test infrastructure) for the TV proximal step: numpy restatement of the reference's
utilities/tv_denoise.py (div :20-31, gradient :34-59, _projector_on_dual :67-74, dual_gap :77-95,
denoise_fista :98-170).  Pinned against the reference's own module, which is pure numpy and imports as is
(tests/golden/make_golden.py -> ref_numpy_cases.npz, keys tv/*)."""
import numpy as np


def div(grad):
    res = np.zeros(grad.shape[1:], dtype=grad.dtype)
    for d in range(grad.shape[0]):
        g = np.moveaxis(grad[d], d, 0)
        r = np.moveaxis(res, d, 0)
        r[:-1] += g[:-1]
        r[1:-1] -= g[:-2]
        r[-1] -= g[-2]
    return res


def gradient(img):
    out = np.zeros((img.ndim,) + img.shape, dtype=img.dtype)
    for d in range(img.ndim):
        sl = [slice(None)] * img.ndim
        sl[d] = slice(0, -1)
        out[d][tuple(sl)] = np.diff(img, axis=d)
    return out


def dual_gap(im, new, gap, weight):
    im_norm = (im ** 2).sum()
    g = gradient(new)
    tv_new = 2 * weight * np.sqrt((g ** 2).sum(axis=0)).sum()
    return 0.5 / im_norm * ((gap ** 2).sum() + tv_new - im_norm + (new ** 2).sum())


def tv_norm_3d(x):
    return np.linalg.norm(gradient(x))


def denoise_fista(im, weight=50, niter=200, eps=1.e-5, check_gap_frequency=3):
    factor = 12.0 if im.ndim == 3 else 8.0
    grad_im = np.zeros((im.ndim,) + im.shape, dtype=im.dtype)
    grad_aux = np.zeros_like(grad_im)
    t, i = 1., 0
    new = im.copy()
    while i < niter:
        error = weight * div(grad_aux) - im
        grad_aux = grad_aux + gradient(error) * (1 / (factor * weight))
        grad_tmp = grad_aux / np.maximum(np.sqrt(np.sum(grad_aux ** 2, 0)), 1.)
        t_new = 0.5 * (1 + np.sqrt(1 + 4 * t ** 2))
        t_factor = (t - 1) / t_new
        grad_aux = (1 + t_factor) * grad_tmp - t_factor * grad_im
        grad_im = grad_tmp
        t = t_new
        if (i % check_gap_frequency) == 0:
            gap = weight * div(grad_im)
            new = im - gap
            if dual_gap(im, new, gap, weight) < eps:
                break
        i += 1
    return new
