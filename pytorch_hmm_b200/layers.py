"""HMMLayer / GaussianHMMLayer -- drop-ins for pytorch_hmm/hmm_layer.py on the sm_100a kernels.

Parameter names (`log_transition_logits`, `log_initial_logits`, `means`, `log_scales`, buffer `transition_matrix`)
match the reference so its state_dicts load.  Semantics kept from the reference:
  * transitions = softmax(log_transition_logits, dim=1), prior = softmax(log_initial_logits)   hmm_layer.py:61-71
  * the HMM object is built on first use (row-normalising P) and afterwards UPDATED in place without
    re-normalisation                                                                            hmm_layer.py:73-89
  * training -> forward-backward posteriors; eval -> one-hot Viterbi (or posteriors)            hmm_layer.py:119-131
  * GaussianHMMLayer feeds exp(log N(x)) to the HMM (hmm_layer.py:336-337); with D = 80 this underflows and every
    emission collapses to the 1e-8 floor (SURVEY finding 3).  `normalize_emissions=True` (not in the reference)
    selects the per-frame max-normalised variant BASELINE.md section 3 uses for non-degenerate parity.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._cache import DerivedCache
from .core import HMMPyTorch, EPS
from .transitions import create_left_to_right_matrix, create_transition_matrix


class HMMLayer(nn.Module):
    def __init__(self, num_states: int, learnable_transitions: bool = True, transition_type: str = "left_to_right",
                 self_loop_prob: float = 0.7, viterbi_inference: bool = True, apply_sigmoid: bool = True):
        super().__init__()
        self.num_states = num_states
        self.viterbi_inference = viterbi_inference
        self.apply_sigmoid = apply_sigmoid
        if transition_type == "left_to_right":
            P_init = create_left_to_right_matrix(num_states, self_loop_prob)
        else:
            P_init = create_transition_matrix(num_states, transition_type, self_loop_prob)
        if learnable_transitions:
            self.log_transition_logits = nn.Parameter(torch.log(P_init + EPS))
        else:
            self.register_buffer("transition_matrix", P_init)
            self.log_transition_logits = None
        self.log_initial_logits = nn.Parameter(torch.log(torch.ones(num_states) / num_states + EPS))
        self._hmm: Optional[HMMPyTorch] = None

    # -- parameters -> probabilities --------------------------------------------------------------------
    def _get_transition_matrix(self) -> torch.Tensor:
        if self.log_transition_logits is not None:
            return F.softmax(self.log_transition_logits, dim=1)
        return self.transition_matrix

    def _get_initial_probabilities(self) -> torch.Tensor:
        return F.softmax(self.log_initial_logits, dim=0)

    def _get_hmm(self) -> HMMPyTorch:
        # (not detached: compute_loss differentiates the likelihood w.r.t. the transition / initial logits through
        # log_P / log_p0, like the reference, hmm_layer.py:73-89; the kernels always receive detached copies)
        P = self._get_transition_matrix()
        p0 = self._get_initial_probabilities()
        if self._hmm is None:
            self._hmm = HMMPyTorch(P, p0, device=str(P.device))
        else:
            h = self._hmm
            h.P, h.log_P = P, torch.log(P + EPS)
            h.p0, h.log_p0 = p0, torch.log(p0 + EPS)
            h.device = str(P.device)
        return self._hmm

    # -- reference API ------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, return_alignment: bool = False
                ) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        if self.apply_sigmoid:
            x = torch.sigmoid(x)
        if x.dim() == 2:
            x = x.unsqueeze(0)
        if x.shape[-1] != self.num_states:
            raise ValueError(f"Input feature dim {x.shape[-1]} must match num_states {self.num_states}")
        hmm = self._get_hmm()
        states = None
        if self.training or not self.viterbi_inference:
            posteriors, _, _ = hmm.forward_backward(x)
        else:
            states, _ = hmm.viterbi_decode(x)
            posteriors = F.one_hot(states, num_classes=self.num_states).float()
        if return_alignment and not self.training:
            alignment = states if states is not None else torch.argmax(posteriors, dim=-1)
            return posteriors, alignment
        return posteriors

    def compute_loss(self, observations: torch.Tensor, target_alignment: Optional[torch.Tensor] = None) -> torch.Tensor:
        hmm = self._get_hmm()
        if target_alignment is not None:
            posteriors = self.forward(observations)
            return F.cross_entropy(posteriors.reshape(-1, self.num_states), target_alignment.reshape(-1))
        if self.apply_sigmoid:
            observations = torch.sigmoid(observations)
        return -hmm.compute_likelihood(observations).mean()

    def align(self, observations: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        hmm = self._get_hmm()
        if self.apply_sigmoid:
            observations = torch.sigmoid(observations)
        return hmm.viterbi_decode(observations)

    def sample(self, seq_length: int, batch_size: int = 1):
        return self._get_hmm().sample(seq_length, batch_size)

    def get_transition_matrix(self) -> torch.Tensor:
        return self._get_transition_matrix()

    def get_initial_probabilities(self) -> torch.Tensor:
        return self._get_initial_probabilities()

    def extra_repr(self) -> str:
        return f"num_states={self.num_states}, viterbi_inference={self.viterbi_inference}"


class GaussianHMMLayer(nn.Module):
    def __init__(self, num_states: int, feature_dim: int, covariance_type: str = "diag",
                 learnable_transitions: bool = True, transition_type: str = "left_to_right",
                 normalize_emissions: bool = False):
        super().__init__()
        self.num_states, self.feature_dim, self.covariance_type = num_states, feature_dim, covariance_type
        self.normalize_emissions = normalize_emissions
        self.hmm_layer = HMMLayer(num_states, learnable_transitions=learnable_transitions,
                                  transition_type=transition_type, apply_sigmoid=False)
        self.means = nn.Parameter(torch.randn(num_states, feature_dim))
        if covariance_type == "full":
            self.log_scales = nn.Parameter(torch.zeros(num_states, feature_dim, feature_dim))
        elif covariance_type == "diag":
            self.log_scales = nn.Parameter(torch.zeros(num_states, feature_dim))
        elif covariance_type == "spherical":
            self.log_scales = nn.Parameter(torch.zeros(num_states, 1))
        else:
            raise ValueError(f"Unknown covariance_type: {covariance_type}")
        self._derived = DerivedCache()          # packed emission parameters, re-packed only when means / log_scales change

    def _packed(self) -> torch.Tensor:
        return self._derived.get("packed", (self.means, self.log_scales),
                                 lambda: ops.gmm_pack(self.means, self._diag_log_scales(), 2.0, None))

    def _diag_log_scales(self) -> torch.Tensor:
        """[K, D] log sigma for every covariance type: the reference's 'full' branch reads only the diagonal of its
        [K,D,D] parameter (hmm_layer.py:311-319) and 'spherical' broadcasts one value (:289-297)."""
        if self.covariance_type == "full":
            return torch.diagonal(self.log_scales, dim1=-2, dim2=-1)
        if self.covariance_type == "spherical":
            return self.log_scales.expand(self.num_states, self.feature_dim)
        return self.log_scales

    def _compute_gaussian_log_probs(self, observations: torch.Tensor) -> torch.Tensor:
        """(B,T,D) -> (B,T,K) log N(x | mu_k, diag(exp(2 log_scales_k)))  (emission kernel; differentiable w.r.t. means,
        log_scales and the observations when gradients are being recorded)."""
        dev = ops.require_cuda(self.means.device if self.means.is_cuda else None)
        from . import autograd as ag
        if ag.needs_grad(observations, self.means, self.log_scales):
            return ag.gmm_log_probs(observations, self.means, self._diag_log_scales(), None, 2.0)
        out = ops.gmm_emission(observations.detach().to(dev), self._packed(), self.num_states, 1, self.feature_dim)
        return out if observations.device == out.device else out.to(observations.device)

    def _posteriors(self, observations: torch.Tensor, want):
        dev = ops.require_cuda(self.means.device if self.means.is_cuda else None)
        logb = ops.gmm_emission(observations.detach().to(dev), self._packed(), self.num_states, 1, self.feature_dim)
        hmm = self.hmm_layer._get_hmm()
        trans, init = hmm._effective_probs(dev)
        mode = ops.EMIS_LOG_NORM_FLOOR if self.normalize_emissions else ops.EMIS_LOG_EXP_FLOOR
        return logb, mode, hmm, trans, init

    def forward(self, observations: torch.Tensor) -> torch.Tensor:
        from . import autograd as ag
        mode = ops.EMIS_LOG_NORM_FLOOR if self.normalize_emissions else ops.EMIS_LOG_EXP_FLOOR
        if (self.hmm_layer.training or not self.hmm_layer.viterbi_inference) and self.num_states <= 32 and \
                ag.needs_grad(observations, *self.parameters()):
            hmm = self.hmm_layer._get_hmm()                               # differentiable log_P / log_p0 (hmm_layer.py:73-89)
            post, _, _ = ag.hmm_posteriors(self._compute_gaussian_log_probs(observations), hmm.log_P, hmm.log_p0, mode, EPS)
            return post if observations.device == post.device else post.to(observations.device)
        logb, mode, hmm, trans, init = self._posteriors(observations, ("gamma",))
        if self.hmm_layer.training or not self.hmm_layer.viterbi_inference:
            r = ops.forward_backward(logb, mode, trans, init, eps=EPS, want=("gamma",))
            post = r["gamma"]
        else:
            r = ops.viterbi(logb, mode, hmm.log_P.to(logb.device), hmm.log_p0.to(logb.device), eps=EPS,
                            want_delta=False, want_score=False)
            post = F.one_hot(r["states"], num_classes=self.num_states).float()
        return post if observations.device == post.device else post.to(observations.device)

    def compute_loss(self, observations: torch.Tensor) -> torch.Tensor:
        """Negative mean log-likelihood (hmm_layer.py:342-359).  The VALUE is the reference's saturating formula
        logsumexp_k log(exp(log alpha_{T-1,k}) + 1e-8) (hmm.py:203-206); the GRADIENT (w.r.t. means, log_scales, the transition and
        initial logits) is that of the true log-likelihood damped by the saturation factor sum_k alpha_k / sum_k (alpha_k + 1e-8) --
        the reference's own gradient wherever no state is floored (it is identically zero once exp(alpha) has underflowed)."""
        from . import autograd as ag
        logb, mode, hmm, trans, init = self._posteriors(observations, ("fwd",))
        r = ops.forward_backward(logb, mode, trans, init, eps=EPS, want=("fwd",))
        last = r["fwd"][:, -1]
        ll = torch.logsumexp(torch.log(last + EPS), dim=-1)                   # hmm.py:206 via hmm_layer.py:358
        if ag.needs_grad(observations, *self.parameters()):
            logb_d = self._compute_gaussian_log_probs(observations)          # differentiable emission (means, log_scales, x)
            true_ll = ag.hmm_log_likelihood(logb_d, hmm.log_P, hmm.log_p0, mode, EPS).to(ll.device)
            sat = (last.sum(-1) / (last + EPS).sum(-1)).detach()
            ll = ll.detach() + sat * (true_ll - true_ll.detach())
        out = -ll.mean()
        return out if observations.device == out.device else out.to(observations.device)

    def extra_repr(self) -> str:
        return (f"num_states={self.num_states}, feature_dim={self.feature_dim}, "
                f"covariance_type={self.covariance_type}")
