"""NeuralHMMRecursion -- the inference core of pytorch_hmm/neural.py's NeuralHMM on the sm_100a kernels.

The reference's NeuralHMM is two torch networks (an observation model giving log-emissions [B,T,K] and a transition model giving
per-frame transition probabilities [B,T,K,K] from a context tensor) in front of the SAME recursions as HMMPyTorch, stepped through
Python loops (neural.py:403-511).  The networks stay whatever torch modules the caller uses; this class is the part behind them:
given `log_obs_probs`, `log_transition_probs` ([B,T,K,K] time-varying, or [K,K] static) and `log_initial_probs` it returns exactly what
NeuralHMM.forward / viterbi_decode / compute_likelihood return, from one launch per recursion.

    slice t of log_transition_probs carries frame t to frame t+1 (neural.py:424, :448, :490);
    log-emissions are used as they are (no 1e-8 floor, unlike HMMPyTorch);
    forward() returns (exp(log posterior), exp(log forward), exp(log backward))  -- the latter two underflow like the reference's.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops


class NeuralHMMRecursion:
    def __init__(self, num_states: int, compute_device: Optional[str] = None):
        self.num_states = num_states
        self.compute_device = compute_device

    def _dev(self, t: torch.Tensor) -> torch.device:
        return ops.require_cuda(self.compute_device if self.compute_device is not None else (t.device if t.is_cuda else None))

    def _check(self, log_obs_probs, log_transition_probs):
        B, T, K = log_obs_probs.shape
        if K != self.num_states:
            raise ValueError(f"log_obs_probs has {K} states, expected {self.num_states}")
        if log_transition_probs.dim() == 4 and tuple(log_transition_probs.shape) != (B, T, K, K):
            raise ValueError(f"log_transition_probs must be [B,T,K,K] or [K,K], got {tuple(log_transition_probs.shape)}")
        return log_transition_probs.dim() == 4

    def forward_backward(self, log_obs_probs: torch.Tensor, log_transition_probs: torch.Tensor,
                         log_initial_probs: torch.Tensor, want=("gamma", "fwd", "bwd")) -> dict:
        """dict(gamma, fwd, bwd, log_alpha, log_beta (as requested), loglik [B]) on the compute device."""
        dev = self._dev(log_obs_probs)
        tv = self._check(log_obs_probs, log_transition_probs)
        e = log_obs_probs.detach().to(dev)
        init = torch.exp(log_initial_probs.detach().to(dev).float())
        trans = torch.exp(log_transition_probs.detach().to(dev).float())         # host-side derivation of the recursion's operands
        if tv:
            return ops.tv_forward_backward(e, trans, init, want=want)
        return ops.forward_backward(e, ops.EMIS_LOG, trans, init, want=want)

    def forward(self, log_obs_probs, log_transition_probs, log_initial_probs) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(posteriors, forward, backward), each [B,T,K]: what NeuralHMM.forward returns (neural.py:355-401)."""
        r = self.forward_backward(log_obs_probs, log_transition_probs, log_initial_probs)
        back = (lambda t: t) if log_obs_probs.device == r["gamma"].device else (lambda t: t.to(log_obs_probs.device))
        return back(r["gamma"]), back(r["fwd"]), back(r["bwd"])

    def viterbi_decode(self, log_obs_probs, log_transition_probs, log_initial_probs) -> Tuple[torch.Tensor, torch.Tensor]:
        """(states int64 [B,T], log_delta [B,T,K]): NeuralHMM.viterbi_decode (neural.py:463-511), bit-identical on identical fp32 inputs."""
        dev = self._dev(log_obs_probs)
        tv = self._check(log_obs_probs, log_transition_probs)
        e = log_obs_probs.detach().to(dev)
        lt, li = log_transition_probs.detach().to(dev), log_initial_probs.detach().to(dev)
        r = ops.tv_viterbi(e, lt, li, want_score=False) if tv else ops.viterbi(e, ops.EMIS_LOG, lt, li, want_score=False)
        st, dl = r["states"], r["delta"]
        if log_obs_probs.device != st.device:
            st, dl = st.to(log_obs_probs.device), dl.to(log_obs_probs.device)
        return st, dl

    def compute_likelihood(self, log_obs_probs, log_transition_probs, log_initial_probs) -> torch.Tensor:
        """The reference's value logsumexp_k log(exp(log alpha_{T-1,k}) + 1e-8) (neural.py:513-519; saturates for long sequences)."""
        r = self.forward_backward(log_obs_probs, log_transition_probs, log_initial_probs, want=("fwd",))
        ll = torch.logsumexp(torch.log(r["fwd"][:, -1] + 1e-8), dim=-1)
        return ll if log_obs_probs.device == ll.device else ll.to(log_obs_probs.device)

    def log_likelihood(self, log_obs_probs, log_transition_probs, log_initial_probs) -> torch.Tensor:
        """True log p(o_1..T) (not available from the reference)."""
        r = self.forward_backward(log_obs_probs, log_transition_probs, log_initial_probs, want=())
        ll = r["loglik"]
        return ll if log_obs_probs.device == ll.device else ll.to(log_obs_probs.device)
