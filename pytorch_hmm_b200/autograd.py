"""Gradients for the training callers of the hot path.

The reference gets its gradients from autograd through the per-time-step ATen ops (pytorch_hmm/hmm.py:95-117); here the
forward pass is one kernel launch, so the backward pass is written out: for L = log p(o_1..T),
    dL / d log b_t(k) = gamma_t(k),    dL / d log P(i,j) = sum_t xi_t(i,j),    dL / d log p0(k) = gamma_0(k)
(docs/01_hmm_theory.md:196-227).  gamma comes from the forward-backward kernels, the xi sums from `hmmb200_xi_sum_f32`
(weighted by the incoming gradient per sequence).  Small-K path only (K <= 32).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops


class _HMMLogLikelihood(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emis, log_P, log_p0, mode, eps):
        dev = ops.require_cuda(emis.device if emis.is_cuda else None)
        e = ops._f32c(emis.detach(), dev)
        B, T, K = e.shape
        if K > 32:
            raise NotImplementedError("gradients of the log-likelihood are implemented for K <= 32")
        trans = torch.exp(ops._f32c(log_P.detach(), dev))
        init = torch.exp(ops._f32c(log_p0.detach(), dev))
        ws = ops.fb_workspace(B, T, K, dev)
        r = ops.forward_backward(e, mode, trans, init, eps=eps, want=("gamma",), workspace=ws, method="sweep")
        ctx.save_for_backward(e, trans, r["gamma"], ws)
        ctx.mode, ctx.eps, ctx.devs = mode, eps, (emis.device, log_P.device, log_p0.device)
        return r["loglik"].to(emis.device)

    @staticmethod
    def backward(ctx, g):
        e, trans, gamma, ws = ctx.saved_tensors
        dev = e.device
        B, T, K = e.shape
        g = ops._f32c(g, dev)
        grad_e = grad_P = grad_p0 = None
        if ctx.needs_input_grad[0]:
            grad_e = gamma * g.view(B, 1, 1)                                   # d/d log b
            if ctx.mode == ops.EMIS_PROB_FLOOR:
                grad_e = grad_e / (e + ctx.eps)                                # d log(p + eps) / dp
            elif ctx.mode == ops.EMIS_LOG_EXP_FLOOR:
                p = torch.exp(e)
                grad_e = grad_e * p / (p + ctx.eps)
            elif ctx.mode == ops.EMIS_LOG_NORM_FLOOR:
                p = torch.exp(e - e.max(-1, keepdim=True)[0])
                grad_e = grad_e * p / (p + ctx.eps)
            grad_e = grad_e.to(ctx.devs[0])
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            xi = torch.zeros(K, K, dtype=torch.float64, device=dev)
            g1 = torch.zeros(K, dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                ops._check(_lib.load().hmmb200_xi_sum_f32(ops._p(e), int(ctx.mode), float(ctx.eps), ops._p(trans), ops._p(ws), ops._p(g),
                                                          B, T, K, ops._p(xi), ops._p(g1), ops._stream(dev)), "hmmb200_xi_sum_f32")
            grad_P = xi.float().to(ctx.devs[1]) if ctx.needs_input_grad[1] else None
            grad_p0 = g1.float().to(ctx.devs[2]) if ctx.needs_input_grad[2] else None
        return grad_e, grad_P, grad_p0, None, None


def hmm_log_likelihood(emis: torch.Tensor, log_P: torch.Tensor, log_p0: torch.Tensor, mode: int = ops.EMIS_LOG,
                       eps: float = ops.EPS) -> torch.Tensor:
    """log p(o) per sequence [B], differentiable w.r.t. the emissions (read according to `mode`), log_P and log_p0."""
    return _HMMLogLikelihood.apply(emis, log_P, log_p0, mode, eps)


def needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)
