"""Angle sharding + all-reduce over world_size-2 gloo on CPU (host logic of sharding.py; the per-rank
operators are oracle-backed here).  Mirrors what recon/sirt_mpi.py:36-72,97-110 does with mpi4py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from helpers import OracleBackend, make_geoms, random_poses, rel_l2
from tomography_alignment_b200.sharding import ShardedProjector, shard_views

N_PROJ, SHAPE, DSHAPE = 5, (8, 8, 8), (8, 8)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, shard="contiguous"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, og = make_geoms(SHAPE, DSHAPE, N_PROJ, cor=np.array([[0.1 * i, 0, 0] for i in range(N_PROJ)]))
        phi, alpha, beta, xyz = random_poses(N_PROJ, 21)
        sp = ShardedProjector(g, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz, backend_factory=OracleBackend, shard=shard)
        rng = np.random.default_rng(5)
        vol = rng.random(SHAPE).astype(np.float32)
        y = rng.random((N_PROJ, g.n_det)).astype(np.float32)
        meas = rng.random((N_PROJ, g.n_det)).astype(np.float32)
        mine = sp.my_index
        proj = sp.forward(vol)
        bp = sp.adjoint(torch.as_tensor(y[mine]))
        n2 = sp.residual_norm2(torch.as_tensor(y[mine]))
        pg = sp.proj_grad(vol, torch.as_tensor(meas[mine]))
        out[rank] = dict(index=mine, proj=proj.numpy(), bp=bp.numpy(), n2=float(n2), grad6=pg["grad6_all"].numpy(),
                         cost=pg["cost_all"].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shard", ["contiguous", "interleaved"])
def test_two_rank_sharding_matches_single_process(shard):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, shard), nprocs=world, join=True)
    g, og = make_geoms(SHAPE, DSHAPE, N_PROJ, cor=np.array([[0.1 * i, 0, 0] for i in range(N_PROJ)]))
    phi, alpha, beta, xyz = random_poses(N_PROJ, 21)
    rng = np.random.default_rng(5)
    vol = rng.random(SHAPE).astype(np.float32)
    y = rng.random((N_PROJ, g.n_det)).astype(np.float32)
    meas = rng.random((N_PROJ, g.n_det)).astype(np.float32)
    op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    ref_proj, ref_bp = op.forward(vol), op.adjoint(y)
    if shard == "contiguous":      # np.array_split (sirt_mpi.py:40): rank 0 gets 3 views, rank 1 gets 2
        assert list(out[0]["index"]) == [0, 1, 2] and list(out[1]["index"]) == [3, 4]
        assert np.array_equal(np.concatenate([shard_views(N_PROJ, world, r) for r in range(world)]), np.arange(N_PROJ))
    else:                          # view i on rank i mod world: every rank sees the whole angular range
        assert list(out[0]["index"]) == [0, 2, 4] and list(out[1]["index"]) == [1, 3]
    got_proj = np.zeros((N_PROJ, g.n_det))
    for r in range(world):
        got_proj[out[r]["index"]] = out[r]["proj"].reshape(len(out[r]["index"]), -1)
    assert rel_l2(got_proj, ref_proj) < 1e-6
    for r in range(world):                      # all-reduced quantities are replicated
        assert rel_l2(out[r]["bp"], ref_bp) < 1e-6
        assert abs(out[r]["n2"] - float((y.astype(np.float64) ** 2).sum())) < 1e-6 * float((y ** 2).sum())
    g6 = np.zeros((N_PROJ, 6))
    cost = np.zeros(N_PROJ)
    for i in range(N_PROJ):
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol)
        res = meas[i].astype(np.float64) - p
        g6[i], cost[i] = -gr @ res, 0.5 * res @ res
    for r in range(world):
        assert rel_l2(out[r]["grad6"], g6) < 1e-5 and rel_l2(out[r]["cost"], cost) < 1e-6
    assert np.array_equal(out[0]["grad6"], out[1]["grad6"])


def _solver_problem():
    n, n_proj = 8, 6
    g, og = make_geoms((n, n, n), (n, n), n_proj)
    phi, alpha, beta, xyz = random_poses(n_proj, 4, shift=0.5)
    c = np.arange(n) - (n - 1) / 2
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    truth = np.exp(-(X ** 2 + Y ** 2 + Z ** 2) / (0.1 * n * n)).astype(np.float32)
    b = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz).forward(truth).reshape(n_proj, -1)
    return g, truth, b.astype(np.float32), np.array([phi, alpha, beta]).T, xyz


def _run_solvers(g, truth, b, angles, xyz, group):
    from tomography_alignment_b200.recon import CGLS, RegularizedRecon
    opts = {"ground_truth": truth}
    out = {}
    out["cgls"] = CGLS(g, b, angles, xyz, options=opts, group=group, backend=OracleBackend(g)).run_main_iteration(niter=4)
    out["lasso"] = RegularizedRecon(g, b, angles, xyz, options=opts, group=group,
                                    backend=OracleBackend(g)).run_lasso_ista(niter=3, reg_param=0.02)
    out["tikh"] = RegularizedRecon(g, b, angles, xyz, options=opts, group=group,
                                   backend=OracleBackend(g)).run_tikhonov_gd(niter=3, reg_param=0.3)
    return {k: (np.asarray(v[0]), np.asarray(v[1])) for k, v in out.items()}


def _solver_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out[rank] = _run_solvers(*_solver_problem(), group=dist.group.WORLD)
    finally:
        dist.destroy_process_group()


def test_two_rank_solvers_match_single_process():
    """recon/cgls_mpi.py, recon/regularized_mpi.py: the sharded loops (views split, A^T y and the residual norms
    all-reduced) reproduce the single-process iterates."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_solver_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    single = _run_solvers(*_solver_problem(), group=None)
    for name, (rec, err) in single.items():
        for r in range(world):
            rec_r, err_r = out[r][name]
            assert rec_r.shape == rec.shape and len(err_r) == len(err), name
            assert rel_l2(rec_r.ravel(), rec.ravel()) < 2e-5, name
            np.testing.assert_allclose(err_r, err, rtol=1e-4, err_msg=name)
        assert np.array_equal(out[0][name][0], out[1][name][0]), name      # replicated state stays bitwise identical


def _worker_guards(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tomography_alignment_b200.recon import SIRT
        from tomography_alignment_b200.sharding import SharedHostBuffer
        g, og = make_geoms((6, 6, 6), (6, 6), 1)
        res = {}
        # more ranks than views: every rank raises the same error before any collective (nobody is left waiting)
        try:
            ShardedProjector(g, phi=np.array([0.3]), backend_factory=OracleBackend)
            res["sharded"] = "no error"
        except ValueError as e:
            res["sharded"] = str(e)
        try:
            SIRT(g, np.zeros((1, g.n_det), np.float32), np.zeros((1, 3)), np.zeros((1, 3)), group=dist.group.WORLD,
                 backend=OracleBackend(g))
            res["sirt"] = "no error"
        except ValueError as e:
            res["sirt"] = str(e)
        # group=None inside an initialised job = an independent reconstruction per rank: no sharding, no collective
        s = SIRT(g, np.ones((1, g.n_det), np.float32), np.zeros((1, 3)), np.zeros((1, 3)), group=None, backend=OracleBackend(g))
        res["independent_world"] = s.world
        # host buffer every rank sees: each rank writes its rows, rank 0 reads all of them
        buf = SharedHostBuffer("tomo_b200_test_%d" % port, (world, 5))
        buf.tensor[rank] = float(rank + 1)
        dist.barrier()
        res["shared"] = buf.tensor.clone().numpy()
        dist.barrier()
        buf.close()
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_empty_shards_are_refused_collectively_and_shared_host_buffer():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_guards, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        assert "every rank needs at least one view" in out[r]["sharded"] and "every rank needs at least one view" in out[r]["sirt"]
        assert out[r]["independent_world"] == 1
        assert np.array_equal(out[r]["shared"], np.array([[1.0] * 5, [2.0] * 5], np.float32))
