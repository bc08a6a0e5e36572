"""HMM / HMMPyTorch -- drop-in for pytorch_hmm/hmm.py on top of the sm_100a kernels.

Public surface, argument meaning, return tuples, dtypes and error behaviour follow the reference
(pytorch_hmm/hmm.py:20-254); the per-time-step ATen loops (hmm.py:95-117, :162-178) are replaced by one
launch of the scaled-probability forward/backward sweeps and one launch of the Viterbi kernel.

`device` keeps its reference meaning (where parameters and RESULTS live).  Compute always runs on a CUDA device
(`compute_device`, default: the current one): CPU-resident inputs are copied to the GPU and results copied back,
which is the end-to-end path bench.py measures.  Without a CUDA device the compute methods raise; there is no
CPU implementation in this package.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import ops

EPS = 1e-8


class HMM:
    """Parameter holder: row-normalised P, p0 and their floored logs (reference hmm.py:20-55)."""

    def __init__(self, P: Union[np.ndarray, torch.Tensor], p0: Optional[Union[np.ndarray, torch.Tensor]] = None,
                 device: str = "cpu", compute_device: Optional[str] = None):
        if isinstance(P, np.ndarray):
            P = torch.from_numpy(P).float()
        P = P.to(device)
        self.K = P.shape[0]
        self.device = device
        self.compute_device = compute_device
        if P.dim() != 2:
            raise ValueError(f"P shape should have length 2. found {P.dim()}")
        if P.shape[0] != P.shape[1]:
            raise ValueError(f"P should be square, found {P.shape}")
        self.P = P / P.sum(dim=1, keepdim=True)
        self.log_P = torch.log(self.P + EPS)
        if p0 is None:
            self.p0 = torch.ones(self.K, device=device) / self.K
        else:
            if isinstance(p0, np.ndarray):
                p0 = torch.from_numpy(p0).float()
            p0 = p0.to(device)
            if len(p0) != self.K:
                raise ValueError(f"dimensions of p0 {p0.shape} must match P[0] {P.shape[0]}")
            self.p0 = p0 / p0.sum()
        self.log_p0 = torch.log(self.p0 + EPS)


class HMMPyTorch(HMM):
    """Forward-backward, Viterbi and likelihood on the B200 kernels (reference hmm.py:58-254)."""

    # -- helpers ---------------------------------------------------------------------------------------
    def _cuda(self) -> torch.device:
        dev = self.compute_device
        if dev is None and torch.device(self.device).type == "cuda":
            dev = self.device
        return ops.require_cuda(dev)

    def _effective_probs(self, dev):
        """exp(log_P), exp(log_p0) as the recursion's scaled-space operands: P + 1e-8 and p0 + 1e-8 (hmm.py:42,55).
        Derived from log_P / log_p0 so that callers who overwrite those (HMMLayer does) stay consistent."""
        return torch.exp(self.log_P.detach().to(dev).float()), torch.exp(self.log_p0.detach().to(dev).float())

    def _batched(self, observations: torch.Tensor):
        if observations.dim() == 2:
            return observations.unsqueeze(0), True
        return observations, False

    def _back(self, t: torch.Tensor) -> torch.Tensor:
        return t if t.device == torch.device(self.device) else t.to(self.device)

    # -- reference API ---------------------------------------------------------------------------------
    def forward_backward(self, observations: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """observations: probabilities (B,T,K) or (T,K).  Returns (posterior, forward, backward), each (B,T,K)
        fp32; forward/backward are exp(log alpha), exp(log beta) exactly as the reference returns them."""
        obs, _ = self._batched(observations)
        B, T, K = obs.shape
        assert K == self.K, f"Observation dim {K} must match model states {self.K}"
        dev = self._cuda()
        from . import autograd as ag
        if ag.needs_grad(observations, self.log_P, self.log_p0) and self.K <= 32:
            # training callers (HMMLayer.forward in training mode, the supervised loss of hmm_layer.py:161-167): the posterior
            # carries its gradient w.r.t. the observations and log_P / log_p0 (csrc/autograd_kernels.cu)
            g, f, b = ag.hmm_posteriors(obs, self.log_P, self.log_p0, ops.EMIS_PROB_FLOOR, EPS)
            return self._back(g), self._back(f), self._back(b)
        trans, init = self._effective_probs(dev)
        r = ops.forward_backward(obs.detach().to(dev), ops.EMIS_PROB_FLOOR, trans, init, eps=EPS)
        return self._back(r["gamma"]), self._back(r["fwd"]), self._back(r["bwd"])

    def viterbi_decode(self, observations: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (states int64 (B,T), log_delta (B,T,K)); both squeezed for 2-D input (hmm.py:180-182)."""
        obs, squeeze = self._batched(observations)
        B, T, K = obs.shape
        assert K == self.K, f"Observation dim {K} must match model states {self.K}"
        dev = self._cuda()
        r = ops.viterbi(obs.detach().to(dev), ops.EMIS_PROB_FLOOR, self.log_P.detach().to(dev), self.log_p0.detach().to(dev),
                        eps=EPS, want_delta=True, want_score=False)
        states, delta = self._back(r["states"]), self._back(r["delta"])
        if squeeze:
            states, delta = states.squeeze(0), delta.squeeze(0)
        return states, delta

    def compute_likelihood(self, observations: torch.Tensor) -> torch.Tensor:
        """The reference's value: logsumexp_k log(exp(log alpha_{T-1,k}) + 1e-8) (hmm.py:203-206), which saturates
        at log(K * 1e-8) for long sequences.  See log_likelihood() for the true quantity."""
        obs, squeeze = self._batched(observations)
        assert obs.shape[-1] == self.K
        dev = self._cuda()
        trans, init = self._effective_probs(dev)
        r = ops.forward_backward(obs.detach().to(dev), ops.EMIS_PROB_FLOOR, trans, init, eps=EPS, want=("fwd",))
        last = r["fwd"][:, -1]
        ll = torch.logsumexp(torch.log(last + EPS), dim=-1)
        from . import autograd as ag
        if ag.needs_grad(observations, self.log_P, self.log_p0):
            # Training callers (HMMLayer.compute_loss, hmm_layer.py:144-173).  The VALUE is the reference's saturating
            # formula; the gradient is that of the true log-likelihood damped by the saturation factor
            # sum_k alpha_k / sum_k (alpha_k + 1e-8) (equal to the reference's own gradient when no state is floored).
            true_ll = ag.hmm_log_likelihood(obs, self.log_P, self.log_p0, ops.EMIS_PROB_FLOOR, EPS).to(dev)
            sat = (last.sum(-1) / (last + EPS).sum(-1)).detach()
            ll = ll.detach() + sat * (true_ll - true_ll.detach())
        ll = self._back(ll)
        return ll.squeeze(0) if squeeze else ll

    def log_likelihood(self, observations: torch.Tensor) -> torch.Tensor:
        """True log p(o_1..T) = logsumexp_k log alpha_{T-1,k} (not available from the reference, SURVEY finding 4)."""
        obs, squeeze = self._batched(observations)
        dev = self._cuda()
        trans, init = self._effective_probs(dev)
        r = ops.forward_backward(obs.detach().to(dev), ops.EMIS_PROB_FLOOR, trans, init, eps=EPS, want=())
        ll = self._back(r["loglik"])
        return ll.squeeze(0) if squeeze else ll

    def sample(self, seq_length: int, batch_size: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
        """Ancestral sampling of a state path with one-hot 'observations' (reference hmm.py:213-245).  Host-side
        convenience, not on the hot path."""
        states = torch.zeros(batch_size, seq_length, dtype=torch.long, device=self.device)
        states[:, 0] = torch.multinomial(self.p0.detach().expand(batch_size, -1), 1).squeeze(1)
        for t in range(1, seq_length):
            states[:, t] = torch.multinomial(self.P.detach()[states[:, t - 1]], 1).squeeze(1)
        observations = torch.nn.functional.one_hot(states, self.K).to(torch.float32)
        return states, observations

    def to(self, device: str):
        self.device = device
        self.P, self.log_P = self.P.to(device), self.log_P.to(device)
        self.p0, self.log_p0 = self.p0.to(device), self.log_p0.to(device)
        return self
