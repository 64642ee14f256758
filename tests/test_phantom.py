"""phantom.shepp3d against the reference's generate_phantom.shepp3d (fixture from make_golden.py)."""
import os

import numpy as np

from tomography_alignment_b200.phantom import benchmark_poses, shepp3d

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_numpy_cases.npz"))


def test_shepp3d_bit_exact_numpy_and_torch():
    for n in (16, 24):
        ref = GOLD["shepp3d/%d" % n]
        a = shepp3d(n)
        assert a.dtype == np.float32 and np.array_equal(a, ref)
        assert np.array_equal(shepp3d(n, device="cpu").numpy(), ref)
    assert shepp3d(16).min() >= 0.0


def test_benchmark_poses_follow_generate_data():
    phi, alpha, beta, xyz = benchmark_poses(90)
    assert np.array_equal(phi, np.linspace(0.0, np.pi, 90))
    assert np.abs(alpha).max() <= np.deg2rad(1.0) and np.abs(beta).max() <= np.deg2rad(1.0)
    assert np.abs(xyz[:, [0, 2]]).max() <= 2.0 and np.all(xyz[:, 1] == 0)
    assert np.array_equal(benchmark_poses(90)[1], alpha)          # seeded
