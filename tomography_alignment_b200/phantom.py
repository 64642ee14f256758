"""3-D Shepp-Logan phantom with the semantics of the reference's ``utilities/generate_phantom.py``
(tomopy-derived): the benchmark / example input generator (examples/generate_data.py:10).

Restated because the original (a) fails on numpy >= 2 (``np.lib.index_tricks``, generate_phantom.py:173)
and (b) materialises 3 float64 coordinate cubes (24*N^3 bytes: 3.2 GB at 512^3).  Here the cube is filled
slab by slab along x, on the CPU (numpy) or directly on the GPU (torch), with identical arithmetic:
    coords = linspace(-1, 1, N) per axis                         (generate_phantom.py:169-175)
    c' = (R(phi, theta, psi) c - (x0, y0, z0)) / (a, b, c)       (:147-166, :178-191)
    obj[sum(c'^2) <= 1] += A, for the 10 ellipsoids of          (:112-144, :194-209)
    result.clip(0, inf), float32                                 (:28-46)
"""
import numpy as np

# A, a, b, c, x0, y0, z0, phi, theta, psi   (generate_phantom.py:198-208)
SHEPP_ARRAY = [
    [1., .6900, .920, .810, 0., 0., 0., 90., 90., 90.],
    [-.8, .6624, .874, .780, 0., -.0184, 0., 90., 90., 90.],
    [-.2, .1100, .310, .220, .22, 0., 0., -108., 90., 100.],
    [-.2, .1600, .410, .280, -.22, 0., 0., 108., 90., 100.],
    [.1, .2100, .250, .410, 0., .35, -.15, 90., 90., 90.],
    [.1, .0460, .046, .050, 0., .1, .25, 90., 90., 90.],
    [.1, .0460, .046, .050, 0., -.1, .25, 90., 90., 90.],
    [.1, .0460, .023, .050, -.08, -.605, 0., 90., 90., 90.],
    [.1, .0230, .023, .020, 0., -.606, 0., 90., 90., 90.],
    [.1, .0230, .046, .020, .06, -.605, 0., 90., 90., 90.]]


def _rotation_matrix(phi, theta, psi):
    """Euler matrix of generate_phantom.py:147-166 (angles in degrees)."""
    cphi, sphi = np.cos(np.radians(phi)), np.sin(np.radians(phi))
    cth, sth = np.cos(np.radians(theta)), np.sin(np.radians(theta))
    cpsi, spsi = np.cos(np.radians(psi)), np.sin(np.radians(psi))
    return np.asarray([[cpsi * cphi - cth * sphi * spsi, cpsi * sphi + cth * cphi * spsi, spsi * sth],
                       [-spsi * cphi - cth * sphi * cpsi, -spsi * sphi + cth * cphi * cpsi, cpsi * sth],
                       [sth * sphi, -sth * cphi, cth]])


def phantom(size, params, dtype="float32", device=None, slab=16):
    """Cube filled with ellipsoids ``params`` (rows A, a, b, c, x0, y0, z0, phi, theta, psi)."""
    if not isinstance(size, tuple):
        size = (size, size, size)
    ax = [np.linspace(-1.0, 1.0, n) for n in size]
    if device is None:
        xp, obj = np, np.zeros(size, dtype=dtype)
        y, z = ax[1][None, :, None], ax[2][None, None, :]
    else:
        import torch
        xp = torch
        obj = torch.zeros(size, dtype=getattr(torch, dtype), device=device)
        ax = [torch.as_tensor(a, device=device) for a in ax]
        y, z = ax[1][None, :, None], ax[2][None, None, :]
    for x0 in range(0, size[0], slab):
        x = ax[0][x0:x0 + slab][:, None, None]
        view = obj[x0:x0 + slab]
        for p in params:
            A, a, b, c, cx, cy, cz, phi, theta, psi = (float(v) for v in p)
            R = _rotation_matrix(phi, theta, psi)
            s = 0.0
            for i, (m0, sc) in enumerate(((cx, a), (cy, b), (cz, c))):
                t = (R[i, 0] * x + R[i, 1] * y + R[i, 2] * z - m0) / sc
                s = s + t * t
            mask = s <= 1.0
            # the reference adds a float64 scalar into the float32 cube: each += is evaluated in float64
            # and rounded to float32 (generate_phantom.py:143)
            if device is None:
                view[mask] += np.float64(A)
            else:
                view[mask] = (view[mask].double() + A).to(view.dtype)
    return obj


def shepp3d(size=128, dtype="float32", device=None):
    """3-D Shepp-Logan phantom (generate_phantom.py:28-46): float32, clipped at 0."""
    obj = phantom(size, SHEPP_ARRAY, dtype, device)
    return obj.clip(0, np.inf) if device is None else obj.clamp_(min=0)


def benchmark_poses(n_proj, seed=20240229):
    """Poses of examples/generate_data.py:16-23 with a fixed seed:
    phi = linspace(0, pi), alpha/beta = +-1 degree in 0.01 degree steps, tx/tz = +-2 px in 0.01 px steps."""
    rng = np.random.default_rng(seed)
    phi = np.linspace(0.0, np.pi, n_proj)
    alpha = np.deg2rad(rng.integers(-100, 100, n_proj) / 100)
    beta = np.deg2rad(rng.integers(-100, 100, n_proj) / 100)
    xyz = np.zeros((n_proj, 3))
    xyz[:, 0] = rng.integers(-200, 200, n_proj) / 100
    xyz[:, 2] = rng.integers(-200, 200, n_proj) / 100
    return phi, alpha, beta, xyz
