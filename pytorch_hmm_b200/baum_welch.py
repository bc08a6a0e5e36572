"""Baum-Welch (EM) for a diagonal-GMM HMM on the sm_100a kernels, batch-sharded across GPUs.

The reference ships no Baum-Welch code: the formulas are docs/01_hmm_theory.md:196-227 (gamma :204, xi :209, pi-hat :216,
a-hat :221) plus the standard Gaussian-mixture M-step, and multi-GPU is a DDP snippet in docs/troubleshooting/faq.md:376-421.
This module is that functionality built on the path's kernels:

  E-step (device): GMM emission + per-component log-likelihoods -> forward-backward (true log-likelihood, EMIS_LOG) ->
                   `hmmb200_bw_accumulate_f32` adds the sufficient statistics in double.
  exchange       : ONE all-reduce (NCCL over NVLink when launched with torchrun) of the statistics vector per EM iteration
                   (K + K^2 + K*C + 2*K*C*D + 3 doubles; 7 888 values = 63 KB at K=12, C=4, D=80).  Utterances are sharded
                   contiguously across ranks; parameters are replicated; every rank then runs the identical M-step.
  M-step (host-side torch, O(K*C*D)): closed-form updates with a variance floor.

`m_step_from_stats`, `shard_range` and `all_reduce_stats` are pure torch, so the N > 1 logic is covered by world-size-2
gloo tests on CPU (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import _lib, ops

EXTRA = 3      # loglik_sum, n_frames, n_sequences appended to the kernel's statistics vector


def stats_slices(K: int, C: int, D: int) -> Dict[str, slice]:
    o, out = 0, {}
    for name, n in (("gamma1", K), ("xi", K * K), ("occ", K * C), ("sx", K * C * D), ("sxx", K * C * D), ("extra", EXTRA)):
        out[name] = slice(o, o + n)
        o += n
    return out


def stats_size(K: int, C: int, D: int) -> int:
    return K + K * K + K * C + 2 * K * C * D + EXTRA


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous partition of the utterance list: ranks [0, n % world) get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sum the statistics over all ranks (no-op without an initialised process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


@dataclass
class GMMHMMParams:
    trans: torch.Tensor      # [K,K] rows sum to 1
    init: torch.Tensor       # [K]
    weights: torch.Tensor    # [K,C]
    means: torch.Tensor      # [K,C,D]
    vars: torch.Tensor       # [K,C,D]

    def to(self, dev):
        return GMMHMMParams(*(t.to(dev) for t in (self.trans, self.init, self.weights, self.means, self.vars)))


def m_step_from_stats(stats: torch.Tensor, K: int, C: int, D: int, var_floor: float = 1e-3, eps: float = 1e-10) -> GMMHMMParams:
    """Closed-form M-step (docs/01_hmm_theory.md:216-227 + GMM extension) from the summed statistics (float64)."""
    sl = stats_slices(K, C, D)
    s = stats.double()
    g1, xi = s[sl["gamma1"]], s[sl["xi"]].view(K, K)
    occ, sx, sxx = s[sl["occ"]].view(K, C), s[sl["sx"]].view(K, C, D), s[sl["sxx"]].view(K, C, D)
    init = (g1 + eps) / (g1 + eps).sum()
    trans = (xi + eps) / (xi + eps).sum(1, keepdim=True)
    weights = (occ + eps) / (occ + eps).sum(1, keepdim=True)
    means = sx / (occ.unsqueeze(-1) + eps)
    var = (sxx / (occ.unsqueeze(-1) + eps) - means ** 2).clamp_min(var_floor)
    return GMMHMMParams(trans.float(), init.float(), weights.float(), means.float(), var.float())


class BaumWelch:
    """EM trainer.  `e_step` may be called on any number of batches per iteration (statistics accumulate on the device)."""

    def __init__(self, params: GMMHMMParams, var_floor: float = 1e-3, device=None):
        self.dev = ops.require_cuda(device)
        self.p = params.to(self.dev)
        self.K, self.C, self.D = self.p.means.shape
        self.var_floor = var_floor
        self.stats = torch.zeros(stats_size(self.K, self.C, self.D), dtype=torch.float64, device=self.dev)
        self._packed = None
        self._buf = None
        self._n_frames = self._n_seqs = 0

    @classmethod
    def from_layer(cls, layer, **kw) -> "BaumWelch":
        """Start from a MixtureGaussianHMMLayer's parameters (uniform prior, as the layer uses)."""
        K = layer.num_states
        with torch.no_grad():
            p = GMMHMMParams(layer.get_transition_matrix().detach().clone(), torch.full((K,), 1.0 / K),
                             torch.softmax(layer.mixture_weights_logits.detach(), -1), layer.means.detach().clone(),
                             torch.exp(layer._diag_log_vars().detach()).contiguous())
        return cls(p, **kw)

    def reset(self):
        self.stats.zero_()
        self._packed = None
        self._n_frames = self._n_seqs = 0

    def _pack(self):
        if self._packed is None:
            self._packed = ops.gmm_pack(self.p.means, torch.log(self.p.vars), 1.0, torch.log(self.p.weights))
            self._trans = self.p.trans.float().contiguous()
            self._init = self.p.init.float().contiguous()
        return self._packed

    def _buffers(self, B: int, T: int):
        """Per-shape scratch, allocated once: log b, per-component values, posteriors, log-likelihoods, sweep workspace."""
        if self._buf is None or self._buf[0] != (B, T):
            K, C, dev = self.K, self.C, self.dev
            self._buf = ((B, T), torch.empty(B, T, K, device=dev), torch.empty(B, T, K * C, device=dev),
                         {"gamma": torch.empty(B, T, K, device=dev), "loglik": torch.empty(B, device=dev)},
                         ops.fb_workspace(B, T, K, dev))
        return self._buf[1:]

    def e_step(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,T,D] (this rank's utterances).  Adds to self.stats; returns the batch's per-sequence log-likelihoods (a view
        of reused scratch: copy it if it must outlive the next call).  Everything is enqueued on the current stream; nothing
        is allocated after the first batch of a shape and no host scalar crosses to the device."""
        x = x.to(self.dev).float().contiguous()
        B, T, D = x.shape
        K, C = self.K, self.C
        lib = _lib.load()
        packed = self._pack()
        logb, comp, out, ws = self._buffers(B, T)
        with torch.cuda.device(self.dev):          # log b and the per-component values in one pass over x
            ops._check(lib.hmmb200_gmm_emission_components_f32(ops._p(x), ops._p(packed), B * T, K, C, D, ops._p(logb), ops._p(comp),
                                                               ops._stream(self.dev)), "hmmb200_gmm_emission_components_f32")
        # method="sweep": the accumulate kernel reads the scaled alpha / beta the sweeps leave in `ws`; the time-parallel scan
        # (which "auto" would pick for long utterances at small batch) keeps its own workspace and never writes them
        r = ops.forward_backward(logb, ops.EMIS_LOG, self._trans, self._init, want=("gamma",), out=out, workspace=ws, method="sweep")
        with torch.cuda.device(self.dev):
            ops._check(lib.hmmb200_bw_accumulate_f32(ops._p(x), ops._p(comp), ops._p(logb), ops._p(r["gamma"]), ops._p(logb),
                                                     ops.EMIS_LOG, 0.0, ops._p(self._trans), ops._p(ws), B, T, K, C, D,
                                                     ops._p(self.stats), ops._stream(self.dev)), "hmmb200_bw_accumulate_f32")
        ex = stats_slices(K, C, D)["extra"]
        self.stats[ex.start] += r["loglik"].double().sum()
        self._n_frames += B * T
        self._n_seqs += B
        return r["loglik"]

    def m_step(self) -> float:
        """All-reduce the statistics, update the parameters (identically on every rank); returns the mean log-likelihood
        per frame over ALL ranks' data for the parameters the E-step used."""
        ex = stats_slices(self.K, self.C, self.D)["extra"]
        self.stats[ex.start + 1:ex.stop] = torch.tensor([float(self._n_frames), float(self._n_seqs)], dtype=torch.float64)
        all_reduce_stats(self.stats)                 # one collective per EM iteration, enqueued behind the last statistics kernel
        exv = self.stats[ex]
        ll_per_frame = float(exv[0] / exv[1])
        self.p = m_step_from_stats(self.stats, self.K, self.C, self.D, self.var_floor).to(self.dev)
        self.reset()
        return ll_per_frame

    def fit(self, batches: Iterable[torch.Tensor], n_iter: int = 5):
        """EM over an iterable of [B,T,D] batches (this rank's shard); returns the per-iteration mean log-likelihood per frame."""
        batches = list(batches)
        history = []
        for _ in range(n_iter):
            for xb in batches:
                self.e_step(xb)
            history.append(self.m_step())
        return history
