"""CPU tests of the N > 1 host logic with the gloo backend, world size 2 (no GPU): utterance sharding, the Baum-Welch
statistics all-reduce and the replicated M-step.  The per-rank E-step statistics come from the oracle here (the CUDA
E-step is covered by the -m gpu tests); what is under test is that sharded-and-reduced == single process."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle
from pytorch_hmm_b200 import baum_welch as bw


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _problem():
    rng = np.random.default_rng(11)
    B, T, K, C, D = 6, 20, 3, 2, 4
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    means = rng.standard_normal((K, C, D)); var = np.exp(0.2 * rng.standard_normal((K, C, D)))
    w = rng.dirichlet(np.ones(C), size=K); P = rng.dirichlet(np.ones(K), size=K); p0 = rng.dirichlet(np.ones(K))
    return x, means, var, w, P, p0


def _oracle_stats(x, means, var, w, P, p0):
    K, C, D = means.shape
    comp = np.log(w)[None, None] - 0.5 * (((x[:, :, None, None, :] - means[None, None]) ** 2 / var[None, None]).sum(-1)
                                          + np.log(var).sum(-1)[None, None] + D * np.log(2 * np.pi))
    st = c_oracle.bw_stats_f64(x, comp, np.log(P), np.log(p0))
    vec = np.concatenate([st["gamma1"], st["xi"].ravel(), st["occ"].ravel(), st["sx"].ravel(), st["sxx"].ravel(),
                          [st["loglik"], x.shape[0] * x.shape[1], x.shape[0]]])
    return torch.from_numpy(vec)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, means, var, w, P, p0 = _problem()
    lo, hi = bw.shard_range(x.shape[0], rank, world)
    stats = _oracle_stats(x[lo:hi], means, var, w, P, p0)
    bw.all_reduce_stats(stats)
    p = bw.m_step_from_stats(stats, *means.shape)
    if rank == 0:
        torch.save({"stats": stats, "means": p.means, "trans": p.trans}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 64000):
        for world in (1, 2, 3, 8):
            spans = [bw.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_allreduce_equals_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    x, means, var, w, P, p0 = _problem()
    full = _oracle_stats(x, means, var, w, P, p0)
    assert torch.allclose(got["stats"], full, rtol=1e-12, atol=1e-12)
    p = bw.m_step_from_stats(full, *means.shape)
    assert torch.allclose(got["means"], p.means) and torch.allclose(got["trans"], p.trans)


def test_m_step_is_a_proper_update():
    x, means, var, w, P, p0 = _problem()
    K, C, D = means.shape
    p = bw.m_step_from_stats(_oracle_stats(x, means, var, w, P, p0), K, C, D, var_floor=1e-3)
    assert torch.allclose(p.trans.sum(1), torch.ones(K), atol=1e-6)
    assert torch.allclose(p.weights.sum(1), torch.ones(K), atol=1e-6)
    assert abs(p.init.sum().item() - 1.0) < 1e-6 and (p.vars >= 1e-3).all()
    assert bw.stats_size(12, 4, 80) == 12 + 144 + 48 + 2 * 3840 + 3
