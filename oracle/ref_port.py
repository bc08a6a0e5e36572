"""oracle/ref_port.py -- torch-CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``pytorch_hmm_b200/`` imports this module.  It is used by ``tests/`` (as the checker), by
``__graft_entry__.smoke()`` and by ``bench.py`` (``cpu_baseline`` leg and ``--impl reference``).

The reference (crlotwhite/pytorch_hmm) is pure Python on top of ``torch``; its arithmetic lives in ATen
(``torch.logsumexp``, ``torch.max``, ``torch.exp/log``).  This port issues the *same ATen op sequence* per
time step, so on a given CPU build of torch it is bit-identical to the reference and it is dispatch-bound
in the same way (which is what makes it a fair CPU baseline: ``kind = "port"``).

Parity status: pinned.  ``tests/test_oracle_golden.py`` checks every function here against fixtures written
by ``oracle/make_golden.py``, which imports the real reference from /root/reference.

Citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

EPS = 1e-8


# ------------------------------------------------------------------------------------------------
# A1  HMM.__init__  (pytorch_hmm/hmm.py:20-55)
# ------------------------------------------------------------------------------------------------
def prepare_hmm(P: torch.Tensor, p0: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row-normalise P, floor with +1e-8 and take logs (hmm.py:39-55).  Returns (log_P, log_p0)."""
    P = P.float()
    if P.dim() != 2:
        raise ValueError(f"P shape should have length 2. found {P.dim()}")
    if P.shape[0] != P.shape[1]:
        raise ValueError(f"P should be square, found {tuple(P.shape)}")
    K = P.shape[0]
    P = P / P.sum(dim=1, keepdim=True)
    if p0 is None:
        p0 = torch.ones(K) / K
    else:
        p0 = p0.float()
        if len(p0) != K:
            raise ValueError("dimensions of p0 must match P")
        p0 = p0 / p0.sum()
    return torch.log(P + EPS), torch.log(p0 + EPS)


# ------------------------------------------------------------------------------------------------
# A2  forward_backward on *probabilities*  (hmm.py:66-130)
# ------------------------------------------------------------------------------------------------
def forward_backward_log(log_obs: torch.Tensor, log_P: torch.Tensor, log_p0: torch.Tensor):
    """Log-space alpha/beta for already-logged emissions.  Returns (log_alpha, log_beta), both [B,T,K]."""
    B, T, K = log_obs.shape
    la = torch.zeros(B, T, K)
    la[:, 0] = log_p0 + log_obs[:, 0]
    for t in range(1, T):                                           # hmm.py:95-101
        la[:, t] = torch.logsumexp(la[:, t - 1, :, None] + log_P[None, :, :], dim=1) + log_obs[:, t]
    lb = torch.zeros(B, T, K)
    for t in range(T - 2, -1, -1):                                  # hmm.py:110-117
        lb[:, t] = torch.logsumexp(log_P[None, :, :] + log_obs[:, t + 1, None, :] + lb[:, t + 1, None, :], dim=2)
    return la, lb


def forward_backward(obs: torch.Tensor, log_P: torch.Tensor, log_p0: torch.Tensor):
    """(posterior, forward, backward) exactly as HMMPyTorch.forward_backward returns them (hmm.py:120-130)."""
    if obs.dim() == 2:
        obs = obs.unsqueeze(0)
    assert obs.shape[-1] == log_P.shape[0]
    la, lb = forward_backward_log(torch.log(obs + EPS), log_P, log_p0)
    lp = la + lb
    lp = lp - torch.logsumexp(lp, dim=-1, keepdim=True)
    return torch.exp(lp), torch.exp(la), torch.exp(lb)


# ------------------------------------------------------------------------------------------------
# A3  viterbi_decode  (hmm.py:132-184) -- additionally returns psi, which the reference keeps internal
# ------------------------------------------------------------------------------------------------
def viterbi_log(log_obs: torch.Tensor, log_P: torch.Tensor, log_p0: torch.Tensor):
    """Viterbi on logged emissions.  Returns (states int64 [B,T], log_delta [B,T,K], psi int64 [B,T,K])."""
    B, T, K = log_obs.shape
    delta = torch.zeros(B, T, K)
    psi = torch.zeros(B, T, K, dtype=torch.long)
    delta[:, 0] = log_p0 + log_obs[:, 0]
    for t in range(1, T):                                           # hmm.py:162-168
        delta[:, t], psi[:, t] = torch.max(delta[:, t - 1, :, None] + log_P[None, :, :], dim=1)
        delta[:, t] += log_obs[:, t]
    states = torch.zeros(B, T, dtype=torch.long)
    states[:, -1] = torch.argmax(delta[:, -1], dim=1)               # hmm.py:174
    rows = torch.arange(B)
    for t in range(T - 2, -1, -1):                                  # hmm.py:177-178
        states[:, t] = psi[rows, t + 1, states[:, t + 1]]
    return states, delta, psi


def viterbi_decode(obs: torch.Tensor, log_P: torch.Tensor, log_p0: torch.Tensor):
    squeeze = obs.dim() == 2
    if squeeze:
        obs = obs.unsqueeze(0)
    states, delta, _ = viterbi_log(torch.log(obs + EPS), log_P, log_p0)
    if squeeze:
        states, delta = states.squeeze(0), delta.squeeze(0)
    return states, delta


# ------------------------------------------------------------------------------------------------
# A4  compute_likelihood  (hmm.py:186-211) -- the saturating value, and the true one
# ------------------------------------------------------------------------------------------------
def compute_likelihood(obs: torch.Tensor, log_P: torch.Tensor, log_p0: torch.Tensor) -> torch.Tensor:
    squeeze = obs.dim() == 2
    if squeeze:
        obs = obs.unsqueeze(0)
    _, fwd, _ = forward_backward(obs, log_P, log_p0)
    ll = torch.logsumexp(torch.log(fwd[:, -1] + EPS), dim=-1)       # hmm.py:206
    return ll.squeeze(0) if squeeze else ll


# ------------------------------------------------------------------------------------------------
# A6  diagonal Gaussian emission of GaussianHMMLayer  (hmm_layer.py:270-323, 'diag' branch)
# ------------------------------------------------------------------------------------------------
def gaussian_log_probs(x: torch.Tensor, means: torch.Tensor, log_scales: torch.Tensor) -> torch.Tensor:
    D = x.shape[-1]
    diff = x.unsqueeze(-2) - means[None, None]
    log_var = 2 * log_scales
    mahal = torch.sum(diff ** 2 / torch.exp(log_var)[None, None], dim=-1)
    log_norm = -0.5 * (D * math.log(2 * math.pi) + torch.sum(log_var, dim=-1))
    return log_norm[None, None] - 0.5 * mahal


# ------------------------------------------------------------------------------------------------
# A7  GMM emission of MixtureGaussianHMMLayer  (mixture_gaussian.py:137-214, 'diag' branch)
# ------------------------------------------------------------------------------------------------
def safe_log(v: torch.Tensor) -> torch.Tensor:
    return torch.log(torch.clamp(v, min=EPS))                       # mixture_gaussian.py:137-139


def own_logsumexp(v: torch.Tensor, dim: int) -> torch.Tensor:
    """The mixture layer's private logsumexp (mixture_gaussian.py:141-155)."""
    m = torch.max(v, dim=dim, keepdim=True)[0]
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    return safe_log(torch.sum(torch.exp(v - m), dim=dim)) + m.squeeze(dim)


def gmm_log_probs(x: torch.Tensor, means: torch.Tensor, log_vars: torch.Tensor,
                  mixture_logits: torch.Tensor) -> torch.Tensor:
    """x [B,T,D], means/log_vars [S,C,D], mixture_logits [S,C] -> [B,T,S]."""
    D = x.shape[-1]
    logw = safe_log(torch.softmax(mixture_logits, dim=-1))
    diff = x[:, :, None, None, :] - means[None, None]
    comp = -0.5 * (torch.sum(diff ** 2 / torch.exp(log_vars)[None, None], dim=-1)
                   + torch.sum(log_vars, dim=-1)[None, None] + D * math.log(2 * math.pi))
    return own_logsumexp(comp + logw[None, None], dim=-1)


# ------------------------------------------------------------------------------------------------
# A8  the mixture layer's private Viterbi  (mixture_gaussian.py:290-338): raw log-emissions, uniform prior
# ------------------------------------------------------------------------------------------------
def mixture_viterbi(logb: torch.Tensor, log_trans: torch.Tensor):
    """Returns (states [B,T] int64, final_scores [B], delta [B,T,S], psi [B,T,S])."""
    B, T, S = logb.shape
    delta = torch.full((B, T, S), float("-inf"), dtype=logb.dtype)
    psi = torch.zeros((B, T, S), dtype=torch.long)
    delta[:, 0, :] = logb[:, 0, :] - math.log(S)                    # :312
    for t in range(1, T):                                           # :315-324
        best, arg = torch.max(delta[:, t - 1, :].unsqueeze(-1) + log_trans.unsqueeze(0), dim=-2)
        delta[:, t, :] = best + logb[:, t, :]
        psi[:, t, :] = arg
    final_scores, last = torch.max(delta[:, -1, :], dim=-1)         # :327
    states = torch.zeros((B, T), dtype=torch.long)
    states[:, -1] = last
    rows = torch.arange(B)
    for t in range(T - 2, -1, -1):                                  # :334-336
        states[:, t] = psi[rows, t + 1, states[:, t + 1]]
    return states, final_scores, delta, psi


# ------------------------------------------------------------------------------------------------
# The headline pipeline of BASELINE.json config 2:
#   get_observation_log_probs -> HMMPyTorch.forward_backward on per-frame max-normalised probabilities
#   -> MixtureGaussianHMMLayer._viterbi_decode on the raw log-emissions.
# ------------------------------------------------------------------------------------------------
def headline_step(x, means, log_vars, mixture_logits, log_P_fb, log_p0_fb, log_trans_vit):
    logb = gmm_log_probs(x, means, log_vars, mixture_logits)
    obs = torch.exp(logb - logb.max(dim=-1, keepdim=True)[0])       # BASELINE.md section 3, last bullet
    post, fwd, bwd = forward_backward(obs, log_P_fb, log_p0_fb)
    states, scores, delta, _ = mixture_viterbi(logb, log_trans_vit)
    return {"logb": logb, "posterior": post, "forward": fwd, "backward": bwd,
            "states": states, "scores": scores, "log_delta": delta}
