"""Data-parallel scheme on CPU with the gloo backend, world_size 2:
  * Dist.all_reduce_sum (the one collective the engines use) over gloo;
  * the scheme itself — per-rank BatchNorm statistic SUMS all-reduced before finalisation, loss terms scaled by the
    GLOBAL batch, gradients summed over ranks — applied to the oracle's VAE+DFC step on two batch shards reproduces
    the single-process step on the full batch (up to fp32 summation order)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        import torch.distributed.nn.functional as dfn
        from icsg3d_b200.engine import Dist
        from oracle import keras_ops as K, nets

        d = Dist()
        assert d.world == world and d.rank == rank
        t = torch.full((4,), float(rank + 1), dtype=torch.float64)
        d.all_reduce_sum(t)
        assert torch.equal(t, torch.full((4,), 3.0, dtype=torch.float64))

        # ---- full-batch reference (computed identically on every rank) ----
        g = torch.Generator().manual_seed(0)
        B, dd = 4, 16
        M = torch.rand(B, dd, dd, dd, 4, generator=g, dtype=torch.float64) * 2
        cond = torch.eye(10, dtype=torch.float64)[torch.randint(0, 10, (B,), generator=g)]
        eps = torch.randn(B, 256, generator=g, dtype=torch.float64)
        pv = nets.init_vae_params(1, dtype=torch.float64, d=dd)
        pu = nets.init_unet_params(2, dtype=torch.float64)
        names = nets.trainable_names(pv)

        def grads_of(Mx, cx, ex, scale):
            leaves = {k: pv[k].clone().requires_grad_(True) for k in names}
            p = dict(pv)
            p.update(leaves)
            (loss, pm, mse, kl), _ = nets.vae_dfc_step(p, pu, Mx, cx, ex, training=True)
            gr = torch.autograd.grad(loss * scale, [leaves[k] for k in names])
            return [float(loss), float(pm), float(mse), float(kl)], gr

        K.STAT_ALLREDUCE = None
        m_full, g_full = grads_of(M, cond, eps, 1.0)

        # ---- two shards with sync-BN statistics + summed gradients ----
        K.STAT_ALLREDUCE = lambda v: dfn.all_reduce(v, op=dist.ReduceOp.SUM)
        sl = slice(rank * B // world, (rank + 1) * B // world)
        m_loc, g_loc = grads_of(M[sl], cond[sl], eps[sl], 1.0 / world)  # local mean / world == share of the global mean
        K.STAT_ALLREDUCE = None
        mt = torch.tensor(m_loc, dtype=torch.float64)
        d.all_reduce_sum(mt)
        mt /= world
        worst = 0.0
        gmax = max(float(b.norm()) for b in g_full)
        for a, b in zip(g_loc, g_full):
            a = a.clone()
            d.all_reduce_sum(a)
            # conv biases in front of a BatchNorm have an analytically zero gradient (pure round-off): compare every
            # tensor on the scale of the largest gradient instead of its own norm
            worst = max(worst, float((a - b).norm() / (b.norm() + 1e-6 * gmax)))
        ret[rank] = (mt.tolist(), m_full, worst)
    finally:
        dist.destroy_process_group()


def test_two_rank_dp_equals_single_rank():
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert len(ret) == 2
    for rank in range(2):
        m_dp, m_full, worst = ret[rank]
        # loss / pm / mse / kld of the global batch
        for a, b in zip(m_dp, m_full):
            assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (m_dp, m_full)
        assert worst < 1e-8, worst
