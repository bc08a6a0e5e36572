"""Gradients for the training callers of the hot path.

The reference gets its gradients from autograd through the per-time-step ATen ops (pytorch_hmm/hmm.py:95-117); here the
forward pass is one kernel launch, so the backward pass is written out: for L = log p(o_1..T),
    dL / d log b_t(k) = gamma_t(k),    dL / d log P(i,j) = sum_t xi_t(i,j),    dL / d log p0(k) = gamma_0(k)
(docs/01_hmm_theory.md:196-227).  gamma comes from the forward-backward kernels, the xi sums from `hmmb200_xi_sum_f32`
(weighted by the incoming gradient per sequence); for K > 32 the same sums are one library GEMM per sequence over the scaled vectors
the large-K sweeps leave in the workspace.
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
            if K > 32:
                xi, g1 = _xi_sum_large_k(e, ctx.mode, ctx.eps, trans, ws, g)
            else:
                xi = torch.zeros(K, K, dtype=torch.float64, device=dev)
                g1 = torch.zeros(K, dtype=torch.float64, device=dev)
                with torch.cuda.device(dev):
                    ops._check(_lib.load().hmmb200_xi_sum_f32(ops._p(e), int(ctx.mode), float(ctx.eps), ops._p(trans), ops._p(ws), ops._p(g),
                                                              B, T, K, ops._p(xi), ops._p(g1), ops._stream(dev)), "hmmb200_xi_sum_f32")
            grad_P = xi.float().to(ctx.devs[1]) if ctx.needs_input_grad[1] else None
            grad_p0 = g1.float().to(ctx.devs[2]) if ctx.needs_input_grad[2] else None
        return grad_e, grad_P, grad_p0, None, None


def _xi_sum_large_k(e, mode, eps, trans, ws, g):
    """sum_b g_b sum_t xi_t(i,j) and sum_b g_b gamma_0 for K > 32, from the scaled alpha / beta vectors the large-K sweeps leave in the
    forward-backward workspace:  xi_t = a_t(i) P(i,j) u_{t+1}(j) / Z_t with u = b~ .* beta, i.e.  P .* (A^T U)  -- a [K,T] x [T,K]
    product per sequence, which is plain library GEMM work (torch.bmm), not a kernel of this package."""
    B, T, K = e.shape
    n = B * T
    stride = (n * K * 4 + 255) & ~255
    a = ws[: n * K * 4].view(torch.float32).view(B, T, K)
    b = ws[stride: stride + n * K * 4].view(torch.float32).view(B, T, K)
    if mode == ops.EMIS_PROB_FLOOR:
        bt = e + eps
    elif mode == ops.EMIS_LOG_EXP_FLOOR:
        bt = torch.exp(e) + eps
    else:
        bt = torch.exp(e - e.max(-1, keepdim=True)[0]) + (eps if mode == ops.EMIS_LOG_NORM_FLOOR else 0.0)
    g0 = a[:, 0] * b[:, 0]
    g1 = ((g0 / g0.sum(-1, keepdim=True)).double() * g.double().view(B, 1)).sum(0)
    if T < 2:
        return torch.zeros(K, K, dtype=torch.float64, device=e.device), g1
    U = bt[:, 1:] * b[:, 1:]                                               # [B,T-1,K]
    A = a[:, :-1]
    Z = (A * (U @ trans.t())).sum(-1, keepdim=True)                        # a_t . (P u_{t+1})
    An = A / Z * g.view(B, 1, 1)
    xi = trans.double() * torch.bmm(An.transpose(1, 2), U).double().sum(0)
    return xi, g1


def hmm_log_likelihood(emis: torch.Tensor, log_P: torch.Tensor, log_p0: torch.Tensor, mode: int = ops.EMIS_LOG,
                       eps: float = ops.EPS) -> torch.Tensor:
    """log p(o) per sequence [B], differentiable w.r.t. the emissions (read according to `mode`), log_P and log_p0."""
    return _HMMLogLikelihood.apply(emis, log_P, log_p0, mode, eps)


def needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


# ------------------------------------------------------------------------------------------------------------------
# posteriors: HMMLayer.forward in training mode, the supervised cross-entropy of HMMLayer.compute_loss (hmm_layer.py:161-167)
# ------------------------------------------------------------------------------------------------------------------
def _emis_chain(grad_logb, e, mode, eps):
    """d/d(emission tensor as passed) from d/d log b, for the emission modes of include/hmm_b200.h."""
    if mode == ops.EMIS_PROB_FLOOR:
        return grad_logb / (e + eps)                                     # d log(p + eps) / dp
    if mode == ops.EMIS_LOG_EXP_FLOOR:
        p = torch.exp(e)
        return grad_logb * p / (p + eps)
    if mode == ops.EMIS_LOG_NORM_FLOOR:
        p = torch.exp(e - e.max(-1, keepdim=True)[0])
        return grad_logb * p / (p + eps)
    return grad_logb


class _HMMPosteriors(torch.autograd.Function):
    """(gamma, exp(log alpha), exp(log beta)) with a backward pass for gamma (hmmb200_posterior_backward_f32: two more sweeps over
    time, see csrc/autograd_kernels.cu); the probability-space forward / backward tensors underflow and carry no gradient."""

    @staticmethod
    def forward(ctx, emis, log_P, log_p0, mode, eps):
        dev = ops.require_cuda(emis.device if emis.is_cuda else None)
        e = ops._f32c(emis.detach(), dev)
        B, T, K = e.shape
        if K > 32:
            raise NotImplementedError("gradients of the posteriors are implemented for K <= 32")
        trans = torch.exp(ops._f32c(log_P.detach(), dev))
        init = torch.exp(ops._f32c(log_p0.detach(), dev))
        ws = ops.fb_workspace(B, T, K, dev)
        r = ops.forward_backward(e, mode, trans, init, eps=eps, want=("gamma", "fwd", "bwd"), workspace=ws, method="sweep")
        ctx.save_for_backward(e, trans, r["gamma"], ws)
        ctx.mode, ctx.eps, ctx.devs = mode, eps, (emis.device, log_P.device, log_p0.device)
        out = tuple(r[k].to(emis.device) for k in ("gamma", "fwd", "bwd"))
        ctx.mark_non_differentiable(out[1], out[2])
        return out

    @staticmethod
    def backward(ctx, g_gamma, _g_fwd, _g_bwd):
        e, trans, gamma, ws = ctx.saved_tensors
        dev = e.device
        B, T, K = e.shape
        G = ops._f32c(g_gamma, dev)
        grad_logb = torch.empty(B, T, K, dtype=torch.float32, device=dev)
        gP = torch.zeros(K, K, dtype=torch.float64, device=dev)
        gp0 = torch.zeros(K, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            ops._check(_lib.load().hmmb200_posterior_backward_f32(ops._p(e), int(ctx.mode), float(ctx.eps), ops._p(trans), ops._p(ws),
                                                                  ops._p(gamma), ops._p(G), B, T, K, ops._p(grad_logb), ops._p(gP),
                                                                  ops._p(gp0), ops._stream(dev)), "hmmb200_posterior_backward_f32")
        grad_e = _emis_chain(grad_logb, e, ctx.mode, ctx.eps).to(ctx.devs[0]) if ctx.needs_input_grad[0] else None
        grad_P = gP.float().to(ctx.devs[1]) if ctx.needs_input_grad[1] else None
        grad_p0 = gp0.float().to(ctx.devs[2]) if ctx.needs_input_grad[2] else None
        return grad_e, grad_P, grad_p0, None, None


def hmm_posteriors(emis: torch.Tensor, log_P: torch.Tensor, log_p0: torch.Tensor, mode: int = ops.EMIS_LOG, eps: float = ops.EPS):
    """(posterior, forward, backward) as HMMPyTorch.forward_backward returns them; `posterior` is differentiable w.r.t. the
    emissions (read according to `mode`), log_P and log_p0."""
    return _HMMPosteriors.apply(emis, log_P, log_p0, mode, eps)


# ------------------------------------------------------------------------------------------------------------------
# emission: d log b / d (mu, log var, log w, x)
# ------------------------------------------------------------------------------------------------------------------
class _GMMEmission(torch.autograd.Function):
    """log b [.., K] = LSE_c(log w_kc + log N(x | mu_kc, exp(scale * log_vars_kc))) on the emission kernel, with the backward pass
    written out: with r the component responsibilities and g = dL/d log b, the weighted statistics
        occ = sum_n g r,   sx = sum_n g r x,   sxx = sum_n g r x^2          (hmmb200_gmm_stats_f32, accumulated in double)
    give  dL/dmu = (sx - occ mu) / var,  dL/dlog_vars = scale/2 ((sxx - 2 mu sx + mu^2 occ) / var - occ),  dL/dlog w = occ."""

    @staticmethod
    def forward(ctx, x, means, log_vars, log_weights, scale):
        dev = ops.require_cuda(means.device if means.is_cuda else None)
        squeeze = means.dim() == 2
        mu = ops._f32c(means.detach(), dev)
        lv = ops._f32c(log_vars.detach().expand_as(means), dev)
        if squeeze:
            mu, lv = mu.unsqueeze(1), lv.unsqueeze(1)
        K, Cn, D = mu.shape
        lw = None if log_weights is None else ops._f32c(log_weights.detach(), dev)
        xs = ops._f32c(x.detach(), dev)
        n = xs.numel() // D
        packed = ops.gmm_pack(mu, lv, float(scale), lw)
        logb = torch.empty(xs.shape[:-1] + (K,), dtype=torch.float32, device=dev)
        if Cn > 1:
            comp = torch.empty(xs.shape[:-1] + (K * Cn,), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                ops._check(_lib.load().hmmb200_gmm_emission_components_f32(ops._p(xs), ops._p(packed), n, K, Cn, D, ops._p(logb), ops._p(comp),
                                                                           ops._stream(dev)), "hmmb200_gmm_emission_components_f32")
        else:
            ops.gmm_emission(xs, packed, K, 1, D, out=logb)
            comp = logb
        ctx.save_for_backward(xs, mu, lv, logb, comp)
        ctx.scale, ctx.squeeze, ctx.has_w = float(scale), squeeze, log_weights is not None
        ctx.devs = (x.device, means.device, log_vars.device, None if log_weights is None else log_weights.device)
        ctx.lv_shape = tuple(log_vars.shape)
        return logb.to(x.device)

    @staticmethod
    def backward(ctx, g):
        xs, mu, lv, logb, comp = ctx.saved_tensors
        dev = xs.device
        K, Cn, D = mu.shape
        n = xs.numel() // D
        g = ops._f32c(g, dev)
        st = torch.zeros(K * Cn + 2 * K * Cn * D, dtype=torch.float64, device=dev)
        occ, sx, sxx = st[:K * Cn], st[K * Cn:K * Cn * (1 + D)], st[K * Cn * (1 + D):]
        with torch.cuda.device(dev):
            ops._check(_lib.load().hmmb200_gmm_stats_f32(ops._p(xs), ops._p(comp), ops._p(logb), ops._p(g), n, K, Cn, D,
                                                         ops._p(occ), ops._p(sx), ops._p(sxx), ops._stream(dev)), "hmmb200_gmm_stats_f32")
        occ = occ.view(K, Cn, 1)
        sx, sxx = sx.view(K, Cn, D), sxx.view(K, Cn, D)
        mud, var = mu.double(), torch.exp(ctx.scale * lv.double())
        g_mu = g_lv = g_lw = g_x = None
        if ctx.needs_input_grad[1]:
            g_mu = ((sx - occ * mud) / var).float()
            g_mu = (g_mu.squeeze(1) if ctx.squeeze else g_mu).to(ctx.devs[1])
        if ctx.needs_input_grad[2]:
            g_lv = (0.5 * ctx.scale * ((sxx - 2.0 * mud * sx + mud * mud * occ) / var - occ)).float()
            g_lv = g_lv.squeeze(1) if ctx.squeeze else g_lv
            # log_vars may have been a broadcast shape (tied / spherical / [K,1]): reduce back to it
            g_lv = g_lv.sum_to_size(ctx.lv_shape).to(ctx.devs[2])
        if ctx.has_w and ctx.needs_input_grad[3]:
            g_lw = occ.view(K, Cn).float().to(ctx.devs[3])
        if ctx.needs_input_grad[0]:
            W = (g.reshape(n, K, 1) * torch.exp(comp.reshape(n, K, Cn) - logb.reshape(n, K, 1))).reshape(n, K * Cn)
            iv = torch.exp(-ctx.scale * lv).reshape(K * Cn, D)
            g_x = (W @ (mu.reshape(K * Cn, D) * iv) - xs.reshape(n, D) * (W @ iv)).reshape(xs.shape).to(ctx.devs[0])
        return g_x, g_mu, g_lv, g_lw, None


def gmm_log_probs(x: torch.Tensor, means: torch.Tensor, log_vars: torch.Tensor, log_weights, scale: float = 1.0) -> torch.Tensor:
    """Differentiable emission log-likelihoods: means [K,C,D] (or [K,D]), log_vars broadcastable to means, log_weights [K,C] or None."""
    return _GMMEmission.apply(x, means, log_vars, log_weights, scale)
