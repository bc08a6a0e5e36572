"""HSMMLayer / SemiMarkovHMM -- drop-ins for pytorch_hmm/hsmm.py and pytorch_hmm/semi_markov.py on the sm_100a kernels.

HSMMLayer keeps the reference's parameters (`transition_logits`, `observation_means`, `observation_log_vars`,
`duration_shape/rate` | `duration_lambda` | `duration_scale/concentration`, buffer `duration_range`) and host-side
O(K*Dmax) tables (duration pdf hsmm.py:115-179, transition softmax with a -inf diagonal :108-113).  The emission runs on the
GMM kernel (C = 1) and the explicit-duration Viterbi on `hmmb200_hsmm_viterbi_f32`, which evaluates the reference's
recursion (hsmm.py:245-354) with the same fp32 operation order -- scores and paths are bit-identical given the same
log-emissions, at ~10^7 times the speed of the reference's five nested Python loops.

SemiMarkovHMM (semi_markov.py:193-570) differs in three documented ways: an initial-state prior, the Gaussian
constant counted once per segment, and log-duration tables taken without the 1e-8 floor.  Its unsupervised forward
(which crashes in the reference, semi_markov.py:353) runs on `hmmb200_hsmm_forward_f32`.
"""
from __future__ import annotations

import math
import warnings
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import autograd, ops
from ._cache import DerivedCache


class HSMMLayer(nn.Module):
    def __init__(self, num_states: int, feature_dim: int, duration_distribution: str = "gamma", max_duration: int = 50,
                 learnable_duration_params: bool = True, min_duration: int = 1):
        super().__init__()
        self.num_states, self.feature_dim = num_states, feature_dim
        self.duration_distribution = duration_distribution
        self.max_duration, self.min_duration = max_duration, min_duration
        self.learnable_duration_params = learnable_duration_params
        self.eps = 1e-8
        S = num_states
        self.transition_logits = nn.Parameter(torch.randn(S, S) * 0.1)
        self.observation_means = nn.Parameter(torch.randn(S, feature_dim) * 0.1)
        self.observation_log_vars = nn.Parameter(torch.zeros(S, feature_dim))
        if learnable_duration_params:
            if duration_distribution == "gamma":
                self.duration_shape = nn.Parameter(torch.ones(S) * 2.0)
                self.duration_rate = nn.Parameter(torch.ones(S) * 0.2)
            elif duration_distribution == "poisson":
                self.duration_lambda = nn.Parameter(torch.ones(S) * 10.0)
            elif duration_distribution == "weibull":
                self.duration_scale = nn.Parameter(torch.ones(S) * 10.0)
                self.duration_concentration = nn.Parameter(torch.ones(S) * 2.0)
            else:
                raise ValueError(f"Unknown duration distribution: {duration_distribution}")
        else:
            self.register_buffer("duration_means", torch.ones(S) * 10.0)
        self.register_buffer("duration_range", torch.arange(min_duration, max_duration + 1, dtype=torch.float))
        self._derived = DerivedCache()          # packed emission parameters and log tables, re-derived when a parameter changes

    def _duration_sources(self):
        names = ("duration_shape", "duration_rate", "duration_lambda", "duration_scale", "duration_concentration", "duration_means")
        return [getattr(self, n) for n in names if hasattr(self, n)]

    # -- host-side tables (O(K * Dmax), same formulas as the reference) -------------------------------------
    def get_transition_matrix(self) -> torch.Tensor:
        logits = self.transition_logits.clone()
        logits.fill_diagonal_(float("-inf"))
        return F.softmax(logits, dim=-1)

    def get_duration_probabilities(self) -> torch.Tensor:
        d = self.duration_range.unsqueeze(0)
        if self.duration_distribution == "gamma":
            shape = F.softplus(self.duration_shape).unsqueeze(1)
            rate = F.softplus(self.duration_rate).unsqueeze(1)
            logp = (shape - 1) * torch.log(d + self.eps) - rate * d - torch.lgamma(shape) + shape * torch.log(rate + self.eps)
        elif self.duration_distribution == "poisson":
            lam = F.softplus(self.duration_lambda).unsqueeze(1)
            logp = d * torch.log(lam + self.eps) - lam - torch.lgamma(d + 1)
        elif self.duration_distribution == "weibull":
            scale = F.softplus(self.duration_scale).unsqueeze(1)
            conc = F.softplus(self.duration_concentration).unsqueeze(1)
            logp = (torch.log(conc + self.eps) - conc * torch.log(scale + self.eps)
                    + (conc - 1) * torch.log(d + self.eps) - (d / scale) ** conc)
        else:
            raise ValueError(f"Unknown duration distribution: {self.duration_distribution}")
        logp = torch.where(d >= self.min_duration, logp, torch.full_like(logp, float("-inf")))
        return torch.exp(logp)

    # -- kernels ---------------------------------------------------------------------------------------------
    def _cuda(self) -> torch.device:
        return ops.require_cuda(self.observation_means.device if self.observation_means.is_cuda else None)

    def get_observation_log_probs(self, observations: torch.Tensor) -> torch.Tensor:
        """(B,T,D) -> (B,T,S) single diagonal Gaussian per state (hsmm.py:181-206) on the emission kernel."""
        dev = self._cuda()
        if autograd.needs_grad(observations, self.observation_means, self.observation_log_vars):
            return autograd.gmm_log_probs(observations, self.observation_means, self.observation_log_vars, None, 1.0)
        packed = self._derived.get("packed", (self.observation_means, self.observation_log_vars),
                                   lambda: ops.gmm_pack(self.observation_means, self.observation_log_vars, 1.0, None))
        out = ops.gmm_emission(observations.detach().to(dev), packed, self.num_states, 1, self.feature_dim)
        return out if observations.device == out.device else out.to(observations.device)

    def _tables(self, dev):
        def make():
            log_dur = torch.log(self.get_duration_probabilities().detach() + self.eps)       # hsmm.py:227
            log_trans = torch.log(self.get_transition_matrix().detach() + self.eps)          # hsmm.py:229
            return log_dur.to(dev).contiguous(), log_trans.to(dev).contiguous()
        return self._derived.get(f"tables@{dev}", [self.transition_logits] + self._duration_sources(), make)

    def viterbi_decode_hsmm(self, observations: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(B,T,D) -> (states int64 (B,T), scores (B,))  (hsmm.py:208-243)."""
        if observations.shape[1] > 1000:
            warnings.warn(f"Long sequence ({observations.shape[1]} frames) may cause memory issues in HSMM decoding.")
        dev = self._cuda()
        with torch.no_grad():                                        # decoding: the emission kernel alone, no autograd bookkeeping
            logb = self.get_observation_log_probs(observations.to(dev))
        states, scores = self._viterbi_from_log_probs(logb)
        if observations.device != states.device:
            states, scores = states.to(observations.device), scores.to(observations.device)
        return states, scores

    def _viterbi_from_log_probs(self, obs_log_probs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Explicit-duration Viterbi on given (B,T,S) log-emissions; bit-identical to hsmm.py:245-354 on the same inputs."""
        dev = self._cuda()
        log_dur, log_trans = self._tables(dev)
        return ops.hsmm_viterbi(obs_log_probs.to(dev), log_dur, log_trans, sum_order=0)

    def forward(self, observations: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.viterbi_decode_hsmm(observations)

    def forward_backward(self, observations: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Duration-augmented forward-backward (BASELINE config 4; new -- the reference's HSMMLayer is Viterbi only,
        hsmm.py:426-437).  Same model as viterbi_decode_hsmm (per-frame Gaussian log-density, log(pmf + 1e-8) durations,
        log(A + 1e-8) transitions without self loops, no prior on the first segment), sums instead of maxima.
        (B,T,D) -> (posterior (B,T,S) = P(state_t = s | x), log_likelihood (B,))."""
        dev = self._cuda()
        with torch.no_grad():
            logb = self.get_observation_log_probs(observations.to(dev))
        log_dur, log_trans = self._tables(dev)
        r = ops.hsmm_forward_backward(logb, log_dur, log_trans)
        g, ll = r["gamma"], r["total"]
        if observations.device != g.device:
            g, ll = g.to(observations.device), ll.to(observations.device)
        return g, ll

    def get_expected_durations(self) -> torch.Tensor:
        if self.duration_distribution == "gamma":
            return F.softplus(self.duration_shape) / F.softplus(self.duration_rate)
        if self.duration_distribution == "poisson":
            return F.softplus(self.duration_lambda)
        if self.duration_distribution == "weibull":
            conc = F.softplus(self.duration_concentration)
            return F.softplus(self.duration_scale) * torch.exp(torch.lgamma(1 + 1 / conc))
        return self.duration_means

    def generate_sequence(self, length: int, initial_state: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """Samples (states [length], observations [length, D]) from the model: a duration from the state's duration pmf, Gaussian
        frames for the whole segment at once, then a successor other than the current state (hsmm.py:356-423).  Host-side helper."""
        dev = self.observation_means.device
        states = torch.zeros(length, dtype=torch.long, device=dev)
        observations = torch.zeros(length, self.feature_dim, device=dev)
        with torch.no_grad():
            trans = self.get_transition_matrix()
            dur = self.get_duration_probabilities()
            std = torch.exp(0.5 * self.observation_log_vars)
            cur, t, it = int(initial_state), 0, 0
            while t < length and it < 2 * length:
                it += 1
                d = int(torch.multinomial(dur[cur], 1)) + self.min_duration
                end = min(t + d, length)
                states[t:end] = cur
                observations[t:end] = self.observation_means[cur] + std[cur] * torch.randn(end - t, self.feature_dim, device=dev)
                t = end
                if t < length:
                    w = trans[cur].clone()
                    w[cur] = 0
                    if w.sum() < 1e-8:
                        w = torch.ones_like(w)
                        w[cur] = 0
                        if w.sum() < 1e-8:
                            break
                    cur = int(torch.multinomial(w / w.sum(), 1))
            if it >= 2 * length and t < length:
                warnings.warn(f"generate_sequence reached maximum iterations ({2 * length}). Generated {t}/{length} timesteps.")
        return states, observations

    def get_model_info(self) -> dict:
        total = sum(p.numel() for p in self.parameters())
        return {"model_type": "HSMM", "num_states": self.num_states, "feature_dim": self.feature_dim,
                "duration_distribution": self.duration_distribution, "max_duration": self.max_duration,
                "min_duration": self.min_duration, "expected_durations": self.get_expected_durations().tolist(),
                "total_parameters": total,
                "trainable_parameters": sum(p.numel() for p in self.parameters() if p.requires_grad),
                "learnable_durations": self.learnable_duration_params}


class DurationModel(nn.Module):
    """Parametric duration log-densities of SemiMarkovHMM (semi_markov.py:9-153; 'neural' is out of scope)."""

    def __init__(self, num_states: int, max_duration: int = 50, distribution_type: str = "gamma", min_duration: int = 1):
        super().__init__()
        self.num_states, self.max_duration = num_states, max_duration
        self.distribution_type, self.min_duration = distribution_type, min_duration
        if distribution_type == "gamma":
            self.alpha_params = nn.Parameter(torch.ones(num_states))
            self.beta_params = nn.Parameter(torch.ones(num_states))
        elif distribution_type == "poisson":
            self.lambda_params = nn.Parameter(torch.ones(num_states) * 5)
        elif distribution_type == "gaussian":
            self.mean_params = nn.Parameter(torch.ones(num_states) * 10)
            self.std_params = nn.Parameter(torch.ones(num_states))
        else:
            raise ValueError(f"Unknown distribution_type: {distribution_type}")

    def _log_density(self, d: torch.Tensor, idx=None) -> torch.Tensor:
        """log p_s(d) by the per-state formulas of semi_markov.py:122-153 (no floor).  idx None: d is [1, N] and the result [K, N]
        (every state against every duration); idx [N]: d is [N] and the result [N] (state idx[i] against d[i])."""
        par = (lambda p: p.unsqueeze(1)) if idx is None else (lambda p: p[idx])
        if self.distribution_type == "gamma":
            a = par(F.softplus(self.alpha_params) + 1e-6)
            b = par(F.softplus(self.beta_params) + 1e-6)
            logp = (a - 1) * torch.log(d + 1e-8) - b * d
            logp = logp - (torch.lgamma(a) - a * torch.log(b))
        elif self.distribution_type == "poisson":
            lam = par(F.softplus(self.lambda_params) + 1e-6)
            logp = d * torch.log(lam + 1e-8) - lam
            logp = logp - torch.lgamma(d + 1)
        else:
            mean = par(F.softplus(self.mean_params) + self.min_duration)
            std = par(F.softplus(self.std_params) + 1e-6)
            logp = -0.5 * torch.log(2 * math.pi * std ** 2) - 0.5 * ((d - mean) / std) ** 2
        return torch.where(d >= self.min_duration, logp, torch.full_like(logp, float("-inf")))

    def log_table(self) -> torch.Tensor:
        """[K, Dmax] log p_s(d), d = 1..Dmax."""
        d = torch.arange(1, self.max_duration + 1, device=next(self.parameters()).device).float().unsqueeze(0)
        return self._log_density(d)

    def forward(self, state_indices: torch.Tensor, durations=None) -> torch.Tensor:
        """semi_markov.py:63-119: durations None -> [N, Dmax] rows of the table for the given states; else log p_{s_i}(d_i), [N]
        (the parametric formulas are evaluated at the given duration, also beyond max_duration, as the reference does)."""
        dev = next(self.parameters()).device
        idx = state_indices.to(dev).long()
        if durations is None:
            return self.log_table()[idx]
        return self._log_density(durations.to(dev).float(), idx)


class SemiMarkovHMM(nn.Module):
    def __init__(self, num_states: int, observation_dim: int, max_duration: int = 50, duration_distribution: str = "gamma",
                 observation_model: str = "gaussian", min_duration: int = 1):
        super().__init__()
        if observation_model != "gaussian":
            raise NotImplementedError("observation_model='neural' is outside the B200 hot path")
        self.num_states, self.observation_dim = num_states, observation_dim
        self.max_duration, self.min_duration = max_duration, min_duration
        self.duration_model = DurationModel(num_states, max_duration, duration_distribution, min_duration)
        self.transition_logits = nn.Parameter(torch.randn(num_states, num_states))
        self.initial_logits = nn.Parameter(torch.zeros(num_states))
        self.observation_means = nn.Parameter(torch.randn(num_states, observation_dim))
        self.observation_logvars = nn.Parameter(torch.zeros(num_states, observation_dim))
        self.observation_model_type = observation_model
        self._derived = DerivedCache()

    def _frame_terms(self, observations: torch.Tensor, dev):
        """Per-frame quadratic term q[t][s] = -0.5 sum_d (x-mu)^2/var and the per-SEGMENT constant
        c[s] = -0.5 sum log var - 0.5 D log 2pi (counted once per segment, semi_markov.py:422-424)."""
        def make():
            packed = ops.gmm_pack(self.observation_means, self.observation_logvars, 1.0, None)
            const = (-0.5 * self.observation_logvars.detach().sum(-1)
                     - 0.5 * self.observation_dim * math.log(2 * math.pi)).to(dev).float()
            return packed, const
        packed, const = self._derived.get(f"packed@{dev}", (self.observation_means, self.observation_logvars), make)
        logb = ops.gmm_emission(observations.detach().to(dev), packed, self.num_states, 1, self.observation_dim)
        return logb - const, const

    def _tables(self, dev):
        log_trans = torch.log(F.softmax(self.transition_logits.detach(), dim=1) + 1e-8)
        log_init = torch.log(F.softmax(self.initial_logits.detach(), dim=0) + 1e-8)
        return self.duration_model.log_table().detach().to(dev), log_trans.to(dev), log_init.to(dev)

    def forward(self, observations: torch.Tensor, state_sequence=None, duration_sequence=None) -> Dict[str, torch.Tensor]:
        if state_sequence is not None and duration_sequence is not None:
            return self._supervised_forward(observations, state_sequence, duration_sequence)
        return self._unsupervised_forward(observations)

    def _supervised_forward(self, observations: torch.Tensor, state_sequence: torch.Tensor,
                            duration_sequence: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Log-probability of a GIVEN segmentation (semi_markov.py:280-305): observation term per segment = the per-segment
        constant + the frame terms of the emission kernel summed over the segment (:385-411, :413-432; segments are taken in
        order until the first one that would end beyond T, the rest contribute no observation term), duration term (:100-119),
        transition term (:434-453).  The frame terms come from the CUDA emission kernel; the three sums are gathers over a
        handful of segments per sequence and stay on the device as tensor operations."""
        dev = ops.require_cuda(self.observation_means.device if self.observation_means.is_cuda else None)
        train = autograd.needs_grad(observations, *self.parameters())                # supervised training: keep the graph
        if train:
            const = (-0.5 * self.observation_logvars.sum(-1) - 0.5 * self.observation_dim * math.log(2 * math.pi)).to(dev)
            q = autograd.gmm_log_probs(observations.to(dev), self.observation_means, self.observation_logvars, None, 1.0) - const
        else:
            q, const = self._frame_terms(observations, dev)                          # [B, T, K], [K]
        B, T, K = q.shape
        st = state_sequence.to(dev).long()
        du = duration_sequence.to(dev).long()
        start = torch.cumsum(du, 1) - du
        fits = torch.cumprod((start + du <= T).long(), 1).bool()                     # the reference breaks at the first misfit
        csum = torch.cat([q.new_zeros(B, 1, K, dtype=torch.float64), torch.cumsum(q.double(), 1)], 1)    # [B, T+1, K]
        pick = lambda pos: csum.gather(1, pos.clamp(0, T).unsqueeze(-1).expand(-1, -1, K)).gather(2, st.unsqueeze(-1)).squeeze(-1)
        seg = (pick(start + du) - pick(start)).float() + const[st]
        log_obs = torch.where(fits, seg, torch.zeros_like(seg)).sum(1)
        log_dur = self.duration_model(st.flatten(), du.flatten()).to(dev).view(B, -1).sum(1)
        log_trans_m = torch.log(F.softmax(self.transition_logits, dim=1) + 1e-8).to(dev)
        if not train:
            log_dur, log_trans_m = log_dur.detach(), log_trans_m.detach()
        log_tr = log_trans_m[st[:, :-1], st[:, 1:]].sum(1) if st.shape[1] > 1 else q.new_zeros(B)
        back = (lambda t: t) if observations.device == dev else (lambda t: t.to(observations.device))
        return {"log_probability": back(log_obs + log_dur + log_tr), "log_observation": back(log_obs),
                "log_duration": back(log_dur), "log_transition": back(log_tr)}

    def _unsupervised_forward(self, observations: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Marginal log-probability over all segmentations (semi_markov.py:308-383).  Like the reference it reports batch
        element 0 under 'log_probability' / 'forward_variables'; all sequences are available under '*_batch'."""
        dev = ops.require_cuda(self.observation_means.device if self.observation_means.is_cuda else None)
        q, const = self._frame_terms(observations, dev)
        log_dur, log_trans, log_init = self._tables(dev)
        r = ops.hsmm_forward(q, log_dur, log_trans, seg_const=const, log_init=log_init, want_alpha=True)
        back = (lambda t: t) if observations.device == r["total"].device else (lambda t: t.to(observations.device))
        return {"log_probability": back(r["total"][0]), "forward_variables": back(r["alpha"][0]),
                "log_probability_batch": back(r["total"]), "forward_variables_batch": back(r["alpha"])}

    def viterbi_decode(self, observations: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """observations (T, D) -> (segment states, segment durations, log-probability)  (semi_markov.py:455-570)."""
        dev = ops.require_cuda(self.observation_means.device if self.observation_means.is_cuda else None)
        q, const = self._frame_terms(observations.unsqueeze(0), dev)
        log_dur, log_trans, log_init = self._tables(dev)
        states, score = ops.hsmm_viterbi(q, log_dur, log_trans, seg_const=const, log_init=log_init, sum_order=1)
        path = states[0]
        seg_states, seg_durs = torch.unique_consecutive(path, return_counts=True)    # no self transitions: runs = segments
        out_dev = observations.device
        return seg_states.to(out_dev), seg_durs.to(out_dev), score[0].to(out_dev)
