"""MixtureGaussianHMMLayer -- drop-in for pytorch_hmm/mixture_gaussian.py on the sm_100a kernels.

Parameter names (`transition_logits` / buffer `transition_matrix`, `mixture_weights_logits`, `means`, `log_vars`)
and initialisation follow the reference (mixture_gaussian.py:58-105).  forward() = GMM emission kernel + Viterbi
kernel on the RAW log-emissions with a uniform prior -log K and log(clamp(softmax(logits), 1e-8)) transitions
(mixture_gaussian.py:312, :357).  Covariance types: 'diag', 'tied' and 'spherical' run on the diagonal (tcgen05) kernel
(the latter two are broadcast special cases); 'full' (Cholesky parameters, mixture_gaussian.py:88-117) runs on the
triangular-contraction kernel csrc/emission_full.cu.
"""
from __future__ import annotations

import math
import warnings
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._cache import DerivedCache


class MixtureGaussianHMMLayer(nn.Module):
    def __init__(self, num_states: int, feature_dim: int, num_components: int = 3, covariance_type: str = "diag",
                 learnable_transitions: bool = True, max_sequence_length: int = 10000):
        super().__init__()
        self.num_states, self.feature_dim, self.num_components = num_states, feature_dim, num_components
        self.covariance_type = covariance_type
        self.learnable_transitions = learnable_transitions
        self.max_sequence_length = max_sequence_length
        self.eps = 1e-8
        self.log_eps = math.log(self.eps)
        self._derived = DerivedCache()          # kernel operands, re-derived only when a parameter changes
        S, Cn, D = num_states, num_components, feature_dim
        if learnable_transitions:
            self.transition_logits = nn.Parameter(torch.randn(S, S) * 0.1)
        else:
            self.register_buffer("transition_matrix", self._create_left_to_right_matrix())
        self.mixture_weights_logits = nn.Parameter(torch.randn(S, Cn) * 0.1)
        self.means = nn.Parameter(torch.randn(S, Cn, D) * math.sqrt(2.0 / D))
        if covariance_type == "diag":
            self.log_vars = nn.Parameter(torch.zeros(S, Cn, D))
        elif covariance_type == "tied":
            self.log_vars = nn.Parameter(torch.zeros(D))
        elif covariance_type == "spherical":
            self.log_vars = nn.Parameter(torch.zeros(S, Cn))
        elif covariance_type == "full":
            # Cholesky parameterisation of the reference (mixture_gaussian.py:88-117): the lower triangle row by row, the diagonal
            # entries are exponentiated; initialised to 0 with 0.1 on the diagonal
            self.cholesky_params = nn.Parameter(torch.zeros(S, Cn, D * (D + 1) // 2))
            with torch.no_grad():
                self.cholesky_params[:, :, [i * (i + 1) // 2 + i for i in range(D)]] = 0.1
        else:
            raise ValueError(f"Unknown covariance_type: {covariance_type}")

    def _create_left_to_right_matrix(self) -> torch.Tensor:
        S = self.num_states
        P = torch.zeros(S, S)
        i = torch.arange(S - 1)
        P[i, i], P[i, i + 1] = 0.8, 0.2
        P[S - 1, S - 1] = 1.0
        return P

    def get_transition_matrix(self) -> torch.Tensor:
        if self.learnable_transitions:
            return F.softmax(self.transition_logits, dim=-1)
        return self.transition_matrix

    def _safe_log(self, x: torch.Tensor) -> torch.Tensor:
        return torch.log(torch.clamp(x, min=self.eps))

    def _get_cholesky_factors(self) -> torch.Tensor:
        """[S,C,D,D] lower-triangular factors with exp() on the diagonal (mixture_gaussian.py:271-288)."""
        S, Cn, D = self.num_states, self.num_components, self.feature_dim
        L = torch.zeros(S * Cn, D, D, device=self.cholesky_params.device, dtype=self.cholesky_params.dtype)
        idx = torch.tril_indices(D, D)
        L[:, idx[0], idx[1]] = self.cholesky_params.view(S * Cn, -1)
        d = torch.arange(D)
        L[:, d, d] = torch.exp(L[:, d, d])
        return L.view(S, Cn, D, D)

    def _packed_full(self):
        def make():
            logw = self._safe_log(F.softmax(self.mixture_weights_logits, dim=-1))
            return ops.gmm_pack_full(self.means, self._get_cholesky_factors(), logw, self.eps)
        return self._derived.get("packed_full", (self.means, self.cholesky_params, self.mixture_weights_logits), make)

    def _emission(self, observations: torch.Tensor, dev) -> torch.Tensor:
        """log b on the emission kernels for the layer's covariance type (inference path: detached)."""
        if self.covariance_type == "full":
            return ops.gmm_emission_full(observations.detach().to(dev), self._packed_full(), self.num_states, self.num_components,
                                         self.feature_dim)
        packed, tc = self._packed_tc()
        return ops.gmm_emission(observations.detach().to(dev), packed, self.num_states, self.num_components, self.feature_dim,
                                tc_known=tc)

    def _diag_log_vars(self) -> torch.Tensor:
        S, Cn, D = self.num_states, self.num_components, self.feature_dim
        if self.covariance_type == "tied":
            return self.log_vars.view(1, 1, D).expand(S, Cn, D)
        if self.covariance_type == "spherical":
            return self.log_vars.unsqueeze(-1).expand(S, Cn, D)
        return self.log_vars

    def _packed(self) -> torch.Tensor:
        """Packed emission parameters (pack kernels + softmax / clamp / log run once per parameter update, not per call)."""
        def make():
            logw = self._safe_log(F.softmax(self.mixture_weights_logits, dim=-1))   # mixture_gaussian.py:178-179
            return ops.gmm_pack(self.means, self._diag_log_vars(), 1.0, logw)
        return self._derived.get("packed", (self.means, self.log_vars, self.mixture_weights_logits), make)

    def _packed_tc(self):
        """(packed, tc_known).  The setup-time query synchronises the stream once per parameter update, so it is only made
        for inference calls; training steps (parameters change every call) take the two-launch dispatch instead."""
        packed = self._packed()
        if torch.is_grad_enabled() and self.training:
            return packed, False
        tc = self._derived.get("tc_known", (packed,), lambda: ops.gmm_pack_on_tensor_cores(
            packed, self.num_states, self.num_components, self.feature_dim))
        return packed, tc

    def _log_transitions(self) -> torch.Tensor:
        src = self.transition_logits if self.learnable_transitions else self.transition_matrix
        return self._derived.get("log_trans", (src,), lambda: self._safe_log(self.get_transition_matrix()).contiguous())

    def _prior(self, dev) -> torch.Tensor:
        S = self.num_states
        return self._derived.get(f"prior@{dev}", (), lambda: torch.full((S,), -math.log(S), dtype=torch.float32, device=dev))

    def get_observation_log_probs(self, observations: torch.Tensor) -> torch.Tensor:
        """(B,T,D) -> (B,T,S) log-likelihood under each state's GMM (mixture_gaussian.py:157-198)."""
        if observations.shape[1] > self.max_sequence_length:
            warnings.warn(f"Sequence length {observations.shape[1]} exceeds recommended maximum "
                          f"{self.max_sequence_length}. Consider chunked processing.")
        dev = ops.require_cuda(self.means.device if self.means.is_cuda else None)
        from . import autograd as ag
        if self.covariance_type != "full" and ag.needs_grad(observations, self.means, self.log_vars, self.mixture_weights_logits):
            # training callers: differentiable w.r.t. means, log_vars, the mixture logits and x (autograd._GMMEmission)
            logw = self._safe_log(F.softmax(self.mixture_weights_logits, dim=-1))
            return ag.gmm_log_probs(observations, self.means, self._diag_log_vars(), logw, 1.0)
        out = self._emission(observations, dev)
        return out if observations.device == out.device else out.to(observations.device)

    def _viterbi_decode(self, obs_log_probs: torch.Tensor, log_transitions: torch.Tensor
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(B,T,S) raw log-emissions -> (states int64 (B,T), final_scores (B,))  (mixture_gaussian.py:290-338)."""
        dev = ops.require_cuda(obs_log_probs.device if obs_log_probs.is_cuda else None)
        S = obs_log_probs.shape[-1]
        prior = self._prior(dev) if S == self.num_states else torch.full((S,), -math.log(S), dtype=torch.float32, device=dev)
        r = ops.viterbi(obs_log_probs.detach().to(dev), ops.EMIS_LOG, log_transitions.detach().to(dev), prior,
                        want_delta=False, want_score=True)
        st, sc = r["states"], r["score"]
        if obs_log_probs.device != st.device:
            st, sc = st.to(obs_log_probs.device), sc.to(obs_log_probs.device)
        return st, sc

    def forward(self, observations: torch.Tensor, return_log_probs: bool = False
                ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        dev = ops.require_cuda(self.means.device if self.means.is_cuda else None)
        logb = self._emission(observations, dev)
        states, scores = self._viterbi_decode(logb, self._log_transitions())         # mixture_gaussian.py:357
        if return_log_probs and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # value: the kernel's score, bit for bit; gradient: along the decoded path
            log_trans = self._safe_log(self.get_transition_matrix())
            ps = self._path_score(observations.to(dev), states, log_trans)
            scores = scores.detach() + (ps - ps.detach())
        if observations.device != states.device:
            states, scores = states.to(observations.device), scores.to(observations.device)
        return (states, scores) if return_log_probs else (states, None)

    def _path_score(self, x: torch.Tensor, states: torch.Tensor, log_trans: torch.Tensor) -> torch.Tensor:
        """The Viterbi score as a differentiable function of the parameters along the decoded path (the sub-gradient the
        reference's autograd produces through torch.max, exercised by tests/test_mixture_gaussian.py:159-176):
        score = -log K + sum_t log b_t(s_t) + sum_t log A(s_{t-1}, s_t).  Plain torch ops on the path states only
        ([B,T,C,D] temporaries) -- this is the training caller, not the inference hot path."""
        S, Cn, D = self.num_states, self.num_components, self.feature_dim
        logw = self._safe_log(F.softmax(self.mixture_weights_logits, dim=-1))        # [S,C]
        mu, lv = self.means[states], self._diag_log_vars()[states]                    # [B,T,C,D]
        diff = x.unsqueeze(2) - mu
        comp = -0.5 * ((diff * diff / torch.exp(lv)).sum(-1) + lv.sum(-1) + D * math.log(2 * math.pi)) + logw[states]
        m = comp.max(-1, keepdim=True)[0]
        logb = (m + torch.log(torch.clamp(torch.exp(comp - m).sum(-1, keepdim=True), min=self.eps))).squeeze(-1)   # [B,T]
        score = logb.sum(1) - math.log(S)
        if states.shape[1] > 1:
            score = score + log_trans[states[:, :-1], states[:, 1:]].sum(1)
        return score

    def get_model_info(self) -> dict:
        total = sum(p.numel() for p in self.parameters())
        return {"num_states": self.num_states, "feature_dim": self.feature_dim, "num_components": self.num_components,
                "covariance_type": self.covariance_type, "learnable_transitions": self.learnable_transitions,
                "total_parameters": total,
                "trainable_parameters": sum(p.numel() for p in self.parameters() if p.requires_grad),
                "memory_efficient": True, "max_sequence_length": self.max_sequence_length}
