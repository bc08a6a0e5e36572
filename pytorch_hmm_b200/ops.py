"""Tensor-level wrappers over the C ABI (include/hmm_b200.h).  torch is used for device memory and streams only.

Every function takes CUDA tensors and enqueues work on torch's current stream of the tensor's device.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch

from . import _lib

EMIS_LOG = 0
EMIS_PROB_FLOOR = 1
EMIS_LOG_NORM_FLOOR = 2
EMIS_LOG_EXP_FLOOR = 3

EPS = 1e-8


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev: torch.device):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {_lib.last_error()}")


def require_cuda(device=None) -> torch.device:
    """The engine is CUDA-only.  Resolves the compute device or raises -- there is no CPU fallback."""
    if not torch.cuda.is_available():
        raise RuntimeError("pytorch_hmm_b200 needs a CUDA device (B200 / sm_100a); it has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"compute device must be CUDA, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _f32c(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    if t.device != dev:
        t = t.to(dev, non_blocking=True)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------------------------------------------
# emission
# ------------------------------------------------------------------------------------------------------
def gmm_pack(means: torch.Tensor, log_vars: torch.Tensor, log_var_scale: float,
             log_weights: Optional[torch.Tensor]) -> torch.Tensor:
    """means/log_vars [K,C,D] (or [K,D]), log_weights [K,C] or None -> packed parameter buffer (device)."""
    dev = require_cuda(means.device if means.is_cuda else None)
    if means.dim() == 2:
        means, log_vars = means.unsqueeze(1), log_vars.unsqueeze(1)
    K, Cn, D = means.shape
    means, log_vars = _f32c(means.detach(), dev), _f32c(log_vars.detach(), dev)
    lw = None if log_weights is None else _f32c(log_weights.detach(), dev)
    lib = _lib.load()
    n = lib.hmmb200_gmm_packed_floats(K, Cn, D)
    packed = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _check(lib.hmmb200_gmm_pack_f32(_p(means), _p(log_vars), float(log_var_scale), _p(lw), K, Cn, D,
                                        _p(packed), _stream(dev)), "hmmb200_gmm_pack_f32")
    return packed


def gmm_pack_on_tensor_cores(packed: torch.Tensor, K: int, Cn: int, D: int) -> bool:
    """Setup-time query (synchronises the stream once): do these packed parameters run on the tcgen05 emission kernel?"""
    dev = packed.device
    with torch.cuda.device(dev):
        rc = _lib.load().hmmb200_gmm_pack_on_tensor_cores(_p(packed), K, Cn, D, _stream(dev))
    if rc < 0:
        _check(rc, "hmmb200_gmm_pack_on_tensor_cores")
    return rc == 1


def gmm_emission(x: torch.Tensor, packed: torch.Tensor, K: int, Cn: int, D: int,
                 out: Optional[torch.Tensor] = None, tc_known: bool = False) -> torch.Tensor:
    """x [..., D] (CUDA) -> log b [..., K].  tc_known: the caller has seen gmm_pack_on_tensor_cores(packed) == True."""
    dev = packed.device
    x = _f32c(x, dev)
    if x.shape[-1] != D:
        raise ValueError(f"feature dim {x.shape[-1]} != {D}")
    n = x.numel() // D
    if out is None:
        out = torch.empty(x.shape[:-1] + (K,), dtype=torch.float32, device=dev)
    lib = _lib.load()
    fn = lib.hmmb200_gmm_emission_tc_f32 if tc_known else lib.hmmb200_gmm_emission_f32
    with torch.cuda.device(dev):
        _check(fn(_p(x), _p(packed), n, K, Cn, D, _p(out), _stream(dev)), "hmmb200_gmm_emission_f32")
    return out


def gmm_pack_full(means: torch.Tensor, chol: torch.Tensor, log_weights: Optional[torch.Tensor], eps: float = 1e-8):
    """Full-covariance operands from means [K,C,D], Cholesky factors L [K,C,D,D] (lower triangular, positive diagonal) and log mixture
    weights [K,C] (or None): (W [K*C, D, DP], cvec [K*C, D], cst [K*C]) for gmm_emission_full.  Host-side derivation, O(K C D^3):
    W = L^-1 by a triangular solve; log det = 2 sum log(diag L + eps) as the reference computes it (mixture_gaussian.py:232-233)."""
    dev = require_cuda(means.device if means.is_cuda else None)
    mu = _f32c(means.detach(), dev)
    L = _f32c(chol.detach(), dev)
    K, Cn, D = mu.shape
    eye = torch.eye(D, device=dev, dtype=torch.float64).expand(K, Cn, D, D)
    W = torch.linalg.solve_triangular(L.double(), eye, upper=False)                        # [K,C,D,D], lower triangular
    cvec = -(W @ mu.double().unsqueeze(-1)).squeeze(-1)
    log_det = 2.0 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1).double() + eps).sum(-1)
    cst = -0.5 * (log_det + D * math.log(2.0 * math.pi))
    if log_weights is not None:
        cst = cst + _f32c(log_weights.detach(), dev).double()
    DP = (D + 3) & ~3
    Wp = torch.zeros(K * Cn, D, DP, dtype=torch.float32, device=dev)
    Wp[:, :, :D] = torch.tril(W).reshape(K * Cn, D, D).float()
    return Wp.contiguous(), cvec.reshape(K * Cn, D).float().contiguous(), cst.reshape(K * Cn).float().contiguous()


def gmm_emission_full(x: torch.Tensor, packed_full, K: int, Cn: int, D: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [..., D] (CUDA) -> log b [..., K] under full-covariance mixtures (operands from gmm_pack_full)."""
    W, cvec, cst = packed_full
    dev = W.device
    x = _f32c(x, dev)
    if x.shape[-1] != D:
        raise ValueError(f"feature dim {x.shape[-1]} != {D}")
    n = x.numel() // D
    if out is None:
        out = torch.empty(x.shape[:-1] + (K,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _check(_lib.load().hmmb200_gmm_emission_full_f32(_p(x), _p(W), _p(cvec), _p(cst), n, K, Cn, D, _p(out), None, _stream(dev)),
               "hmmb200_gmm_emission_full_f32")
    return out


# ------------------------------------------------------------------------------------------------------
# recursions
# ------------------------------------------------------------------------------------------------------
def fb_workspace(B: int, T: int, K: int, dev) -> torch.Tensor:
    """Scratch buffer for forward_backward (scaled alpha/beta + log-scales); reusable across calls on one stream."""
    n = _lib.load().hmmb200_fb_workspace_bytes(B, T, K)
    return torch.empty(max(n, 1), dtype=torch.uint8, device=dev)


def viterbi_workspace(B: int, T: int, K: int, dev) -> torch.Tensor:
    n = _lib.load().hmmb200_viterbi_workspace_bytes(B, T, K)
    return torch.empty(max(n, 1), dtype=torch.uint8, device=dev)


def use_time_parallel_scan(B: int, T: int, K: int) -> bool:
    """Schedule choice for forward-backward: the sequential sweeps take ~65 ns per frame of the LONGEST sequence whatever
    the batch; the scan does K x the arithmetic but in parallel over time.  Measured crossover (tools/bench_scan.py)."""
    return K <= 32 and T >= 8192 and B <= 64          # 2.3x at (B=1, T=16k), 17.7x at (B=1, T=1M), 1.5x at (B=64, T=16k)


def forward_backward(emis: torch.Tensor, mode: int, trans_prob: torch.Tensor, init_prob: torch.Tensor,
                     eps: float = EPS, add_rowmax: bool = False, want=("gamma", "fwd", "bwd"),
                     out: Optional[dict] = None, workspace: Optional[torch.Tensor] = None, method: str = "auto") -> dict:
    """emis [B,T,K] CUDA fp32.  Returns a dict with the requested tensors among
    gamma, fwd, bwd, log_alpha, log_beta (each [B,T,K]) and always 'loglik' [B].
    method: "sweep" (sequential in time), "scan" (time-parallel, K <= 32) or "auto"."""
    dev = require_cuda(emis.device)
    emis = _f32c(emis, dev)
    B, T, K = emis.shape
    if method == "scan" or (method == "auto" and use_time_parallel_scan(B, T, K)):
        return _forward_backward_scan(emis, mode, trans_prob, init_prob, eps, add_rowmax, want, out, dev)
    trans_prob, init_prob = _f32c(trans_prob, dev), _f32c(init_prob, dev)
    lib = _lib.load()
    res = {} if out is None else out
    for name in ("gamma", "fwd", "bwd", "log_alpha", "log_beta"):
        if name in want and name not in res:
            res[name] = torch.empty(B, T, K, dtype=torch.float32, device=dev)
    if "loglik" not in res:
        res["loglik"] = torch.empty(B, dtype=torch.float32, device=dev)
    ws_bytes = lib.hmmb200_fb_workspace_bytes(B, T, K)
    ws = workspace if workspace is not None and workspace.numel() >= ws_bytes else torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _check(lib.hmmb200_forward_backward_f32(
            _p(emis), int(mode), float(eps), int(bool(add_rowmax)), _p(trans_prob), _p(init_prob), B, T, K,
            _p(res.get("gamma")), _p(res.get("fwd")), _p(res.get("bwd")), _p(res.get("log_alpha")),
            _p(res.get("log_beta")), _p(res["loglik"]), _p(ws), ws_bytes, _stream(dev)),
            "hmmb200_forward_backward_f32")
    return res


def _forward_backward_scan(emis, mode, trans_prob, init_prob, eps, add_rowmax, want, out, dev) -> dict:
    B, T, K = emis.shape
    trans_prob, init_prob = _f32c(trans_prob, dev), _f32c(init_prob, dev)
    lib = _lib.load()
    res = {} if out is None else out
    for name in ("gamma", "fwd", "bwd", "log_alpha", "log_beta"):
        if name in want and name not in res:
            res[name] = torch.empty(B, T, K, dtype=torch.float32, device=dev)
    if "loglik" not in res:
        res["loglik"] = torch.empty(B, dtype=torch.float32, device=dev)
    n = lib.hmmb200_fb_scan_workspace_bytes(B, T, K)
    ws = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _check(lib.hmmb200_forward_backward_scan_f32(
            _p(emis), int(mode), float(eps), int(bool(add_rowmax)), _p(trans_prob), _p(init_prob), B, T, K,
            _p(res.get("gamma")), _p(res.get("fwd")), _p(res.get("bwd")), _p(res.get("log_alpha")),
            _p(res.get("log_beta")), _p(res["loglik"]), _p(ws), n, _stream(dev)),
            "hmmb200_forward_backward_scan_f32")
    return res


def viterbi(emis: torch.Tensor, mode: int, log_trans: torch.Tensor, log_init: torch.Tensor, eps: float = EPS,
            want_delta: bool = True, want_psi: bool = False, want_score: bool = True,
            out: Optional[dict] = None, workspace: Optional[torch.Tensor] = None) -> dict:
    """emis [B,T,K] CUDA fp32 -> dict(states int64 [B,T], delta [B,T,K], psi uint8 [B,T,K], score [B])."""
    dev = require_cuda(emis.device)
    emis = _f32c(emis, dev)
    B, T, K = emis.shape
    log_trans, log_init = _f32c(log_trans, dev), _f32c(log_init, dev)
    lib = _lib.load()
    res = {} if out is None else out
    if "states" not in res:
        res["states"] = torch.empty(B, T, dtype=torch.int64, device=dev)
    if want_delta and "delta" not in res:
        res["delta"] = torch.empty(B, T, K, dtype=torch.float32, device=dev)
    if want_psi and "psi" not in res:
        res["psi"] = torch.empty(B, T, K, dtype=torch.uint8 if K <= 256 else torch.int16, device=dev)   # packed: 1 or 2 bytes
    if want_score and "score" not in res:
        res["score"] = torch.empty(B, dtype=torch.float32, device=dev)
    ws_bytes = lib.hmmb200_viterbi_workspace_bytes(B, T, K)
    ws = workspace if workspace is not None and workspace.numel() >= ws_bytes else torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _check(lib.hmmb200_viterbi_f32(_p(emis), int(mode), float(eps), _p(log_trans), _p(log_init), B, T, K,
                                       _p(res.get("delta")), _p(res.get("psi")), _p(res["states"]),
                                       _p(res.get("score")), _p(ws), ws_bytes, _stream(dev)),
               "hmmb200_viterbi_f32")
    return res


def tv_forward_backward(log_emis: torch.Tensor, trans_prob: torch.Tensor, init_prob: torch.Tensor,
                        want=("gamma", "fwd", "bwd"), out: Optional[dict] = None, workspace: Optional[torch.Tensor] = None) -> dict:
    """Forward-backward with time-varying transitions (NeuralHMM form): log_emis [B,T,K] log-emissions, trans_prob [B,T,K,K]
    probabilities (slice t: frame t -> t+1), init_prob [K].  Same result dict as forward_backward()."""
    dev = require_cuda(log_emis.device)
    e = _f32c(log_emis, dev)
    B, T, K = e.shape
    tp, ip = _f32c(trans_prob, dev), _f32c(init_prob, dev)
    if tuple(tp.shape) != (B, T, K, K):
        raise ValueError(f"trans_prob must be [B,T,K,K] = {(B, T, K, K)}, got {tuple(tp.shape)}")
    lib = _lib.load()
    res = {} if out is None else out
    for name in ("gamma", "fwd", "bwd", "log_alpha", "log_beta"):
        if name in want and name not in res:
            res[name] = torch.empty(B, T, K, dtype=torch.float32, device=dev)
    if "loglik" not in res:
        res["loglik"] = torch.empty(B, dtype=torch.float32, device=dev)
    ws_bytes = lib.hmmb200_fb_workspace_bytes(B, T, K)
    ws = workspace if workspace is not None and workspace.numel() >= ws_bytes else torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _check(lib.hmmb200_tv_forward_backward_f32(_p(e), _p(tp), _p(ip), B, T, K, _p(res.get("gamma")), _p(res.get("fwd")),
                                                   _p(res.get("bwd")), _p(res.get("log_alpha")), _p(res.get("log_beta")),
                                                   _p(res["loglik"]), _p(ws), ws_bytes, _stream(dev)), "hmmb200_tv_forward_backward_f32")
    return res


def tv_viterbi(log_emis: torch.Tensor, log_trans: torch.Tensor, log_init: torch.Tensor, want_delta: bool = True,
               want_psi: bool = False, want_score: bool = True) -> dict:
    """Viterbi with time-varying transitions: log_trans [B,T,K,K] (slice t: frame t -> t+1) -> states / delta / psi / score."""
    dev = require_cuda(log_emis.device)
    e = _f32c(log_emis, dev)
    B, T, K = e.shape
    lt, li = _f32c(log_trans, dev), _f32c(log_init, dev)
    if tuple(lt.shape) != (B, T, K, K):
        raise ValueError(f"log_trans must be [B,T,K,K] = {(B, T, K, K)}, got {tuple(lt.shape)}")
    lib = _lib.load()
    res = {"states": torch.empty(B, T, dtype=torch.int64, device=dev)}
    if want_delta:
        res["delta"] = torch.empty(B, T, K, dtype=torch.float32, device=dev)
    if want_psi:
        res["psi"] = torch.empty(B, T, K, dtype=torch.uint8, device=dev)
    if want_score:
        res["score"] = torch.empty(B, dtype=torch.float32, device=dev)
    n = lib.hmmb200_tv_viterbi_workspace_bytes(B, T, K)
    ws = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _check(lib.hmmb200_tv_viterbi_f32(_p(e), _p(lt), _p(li), B, T, K, _p(res.get("delta")), _p(res.get("psi")), _p(res["states"]),
                                          _p(res.get("score")), _p(ws), n, _stream(dev)), "hmmb200_tv_viterbi_f32")
    return res


FUSED_PDL = 1
FUSED_BF16_OUT = 2


def fb_viterbi_workspace(B: int, T: int, K: int, dev) -> torch.Tensor:
    n = _lib.load().hmmb200_fb_viterbi_workspace_bytes(B, T, K)
    return torch.empty(max(n, 1), dtype=torch.uint8, device=dev)


def forward_backward_viterbi(emis: torch.Tensor, fb_mode: int, vit_mode: int, trans_prob: torch.Tensor, init_prob: torch.Tensor,
                             log_trans: torch.Tensor, log_init: torch.Tensor, eps: float = EPS, add_rowmax: bool = False,
                             want=("gamma", "fwd", "bwd"), want_delta: bool = True, want_psi: bool = False, want_score: bool = True,
                             out: Optional[dict] = None, workspace: Optional[torch.Tensor] = None, pdl: bool = False,
                             out_dtype: torch.dtype = torch.float32) -> dict:
    """Forward-backward and Viterbi on the same emissions in one pass (hmmb200_fb_viterbi_f32: one launch for K <= 32).
    Returns the union of forward_backward()'s and viterbi()'s dicts.  pdl: the kernel may overlap the tail of the previous
    kernel on the stream (the emission kernel that writes `emis`); only pass True when that kernel writes none of the tables."""
    dev = require_cuda(emis.device)
    emis = _f32c(emis, dev)
    B, T, K = emis.shape
    trans_prob, init_prob = _f32c(trans_prob, dev), _f32c(init_prob, dev)
    log_trans, log_init = _f32c(log_trans, dev), _f32c(log_init, dev)
    lib = _lib.load()
    res = {} if out is None else out
    bf16 = out_dtype == torch.bfloat16                           # posteriors / forward / backward as bfloat16 (K <= 32)
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("out_dtype must be torch.float32 or torch.bfloat16")
    for name in ("gamma", "fwd", "bwd", "log_alpha", "log_beta"):
        if name in want and name not in res:
            res[name] = torch.empty(B, T, K, dtype=out_dtype if name in ("gamma", "fwd", "bwd") else torch.float32, device=dev)
    if bf16 and any(res[n].dtype != torch.bfloat16 for n in ("gamma", "fwd", "bwd") if n in res):
        raise ValueError("out_dtype=bfloat16 needs bfloat16 output tensors")
    if "loglik" not in res:
        res["loglik"] = torch.empty(B, dtype=torch.float32, device=dev)
    if "states" not in res:
        res["states"] = torch.empty(B, T, dtype=torch.int64, device=dev)
    if want_delta and "delta" not in res:
        res["delta"] = torch.empty(B, T, K, dtype=torch.float32, device=dev)
    if want_psi and "psi" not in res:
        res["psi"] = torch.empty(B, T, K, dtype=torch.uint8 if K <= 256 else torch.int16, device=dev)
    if want_score and "score" not in res:
        res["score"] = torch.empty(B, dtype=torch.float32, device=dev)
    ws_bytes = lib.hmmb200_fb_viterbi_workspace_bytes(B, T, K)
    ws = workspace if workspace is not None and workspace.numel() >= ws_bytes else torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _check(lib.hmmb200_fb_viterbi_f32(
            _p(emis), int(fb_mode), int(vit_mode), float(eps), int(bool(add_rowmax)), _p(trans_prob), _p(init_prob),
            _p(log_trans), _p(log_init), B, T, K,
            _p(res.get("gamma")), _p(res.get("fwd")), _p(res.get("bwd")), _p(res.get("log_alpha")), _p(res.get("log_beta")),
            _p(res["loglik"]), _p(res.get("delta")), _p(res.get("psi")), _p(res["states"]), _p(res.get("score")),
            _p(ws), ws_bytes, (FUSED_PDL if pdl else 0) | (FUSED_BF16_OUT if bf16 else 0), _stream(dev)), "hmmb200_fb_viterbi_f32")
    return res


# ------------------------------------------------------------------------------------------------------
# explicit-duration (semi-Markov) recursions
# ------------------------------------------------------------------------------------------------------
def hsmm_viterbi(frame_logp: torch.Tensor, log_dur: torch.Tensor, log_trans: torch.Tensor,
                 seg_const: Optional[torch.Tensor] = None, log_init: Optional[torch.Tensor] = None,
                 sum_order: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """frame_logp [B,T,K] -> (states int64 [B,T], score [B]); see include/hmm_b200.h for the recursion."""
    dev = require_cuda(frame_logp.device)
    f = _f32c(frame_logp, dev)
    B, T, K = f.shape
    log_dur, log_trans = _f32c(log_dur, dev), _f32c(log_trans, dev)
    Dm = log_dur.shape[1]
    sc = None if seg_const is None else _f32c(seg_const, dev)
    li = None if log_init is None else _f32c(log_init, dev)
    lib = _lib.load()
    states = torch.empty(B, T, dtype=torch.int64, device=dev)
    score = torch.empty(B, dtype=torch.float32, device=dev)
    n = lib.hmmb200_hsmm_viterbi_workspace_bytes(B, T, K, Dm)
    ws = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _check(lib.hmmb200_hsmm_viterbi_f32(_p(f), _p(sc), _p(log_dur), _p(log_trans), _p(li), B, T, K, Dm, int(sum_order),
                                            _p(states), _p(score), _p(ws), n, _stream(dev)), "hmmb200_hsmm_viterbi_f32")
    return states, score


def hsmm_forward(frame_logp: torch.Tensor, log_dur: torch.Tensor, log_trans: torch.Tensor,
                 seg_const: Optional[torch.Tensor] = None, log_init: Optional[torch.Tensor] = None,
                 want_alpha: bool = True) -> dict:
    """frame_logp [B,T,K] -> dict(total [B], alpha [B,T,K,Dm] if want_alpha, end [B,T,K])."""
    dev = require_cuda(frame_logp.device)
    f = _f32c(frame_logp, dev)
    B, T, K = f.shape
    log_dur, log_trans = _f32c(log_dur, dev), _f32c(log_trans, dev)
    Dm = log_dur.shape[1]
    sc = None if seg_const is None else _f32c(seg_const, dev)
    li = None if log_init is None else _f32c(log_init, dev)
    lib = _lib.load()
    res = {"total": torch.empty(B, dtype=torch.float32, device=dev),
           "end": torch.empty(B, T, K, dtype=torch.float32, device=dev)}
    if want_alpha:
        res["alpha"] = torch.empty(B, T, K, Dm, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _check(lib.hmmb200_hsmm_forward_f32(_p(f), _p(sc), _p(log_dur), _p(log_trans), _p(li), B, T, K, Dm,
                                            _p(res.get("alpha")), _p(res["end"]), _p(res["total"]), _stream(dev)),
               "hmmb200_hsmm_forward_f32")
    return res


def hsmm_forward_backward(frame_logp: torch.Tensor, log_dur: torch.Tensor, log_trans: torch.Tensor,
                          seg_const: Optional[torch.Tensor] = None, log_init: Optional[torch.Tensor] = None,
                          want_beta: bool = False) -> dict:
    """frame_logp [B,T,K] -> dict(gamma [B,T,K] state-occupancy posterior, total [B], beta_begin/beta_end [B,T,K] if want_beta)."""
    dev = require_cuda(frame_logp.device)
    f = _f32c(frame_logp, dev)
    B, T, K = f.shape
    log_dur, log_trans = _f32c(log_dur, dev), _f32c(log_trans, dev)
    Dm = log_dur.shape[1]
    sc = None if seg_const is None else _f32c(seg_const, dev)
    li = None if log_init is None else _f32c(log_init, dev)
    lib = _lib.load()
    res = {"gamma": torch.empty(B, T, K, dtype=torch.float32, device=dev), "total": torch.empty(B, dtype=torch.float32, device=dev)}
    if want_beta:
        res["beta_begin"] = torch.empty(B, T, K, dtype=torch.float32, device=dev)
        res["beta_end"] = torch.empty(B, T, K, dtype=torch.float32, device=dev)
    n = lib.hmmb200_hsmm_fb_workspace_bytes(B, T, K)
    ws = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _check(lib.hmmb200_hsmm_forward_backward_f32(_p(f), _p(sc), _p(log_dur), _p(log_trans), _p(li), B, T, K, Dm,
                                                     _p(res["gamma"]), _p(res["total"]), _p(res.get("beta_begin")),
                                                     _p(res.get("beta_end")), _p(ws), n, _stream(dev)),
               "hmmb200_hsmm_forward_backward_f32")
    return res


# ------------------------------------------------------------------------------------------------------
# streaming (state carried between chunks)
# ------------------------------------------------------------------------------------------------------
def greedy_decode(logb: torch.Tensor, log_trans: torch.Tensor, state_io: torch.Tensor, want_scores: bool = True):
    """logb [B,T,K]; state_io int32 [B] (-1 before the first chunk; updated in place) -> (states int64 [B,T], scores [B,T])."""
    dev = require_cuda(logb.device)
    logb, log_trans = _f32c(logb, dev), _f32c(log_trans, dev)
    B, T, K = logb.shape
    states = torch.empty(B, T, dtype=torch.int64, device=dev)
    scores = torch.empty(B, T, dtype=torch.float32, device=dev) if want_scores else None
    with torch.cuda.device(dev):
        _check(_lib.load().hmmb200_greedy_decode_f32(_p(logb), _p(log_trans), B, T, K, _p(state_io), _p(states), _p(scores),
                                                     _stream(dev)), "hmmb200_greedy_decode_f32")
    return states, scores


def forward_chunk(emis: torch.Tensor, mode: int, trans_prob: torch.Tensor, init_prob: torch.Tensor, state: dict,
                  eps: float = EPS, want_filtered: bool = True) -> Optional[torch.Tensor]:
    """One chunk of the forward recursion.  `state` = dict(alpha [B,K] f32, loglik [B] f64, started [B] i32), updated in place
    (make one with new_forward_state).  Returns the filtered posteriors [B,T,K] (or None)."""
    dev = require_cuda(emis.device)
    emis = _f32c(emis, dev)
    B, T, K = emis.shape
    filt = torch.empty(B, T, K, dtype=torch.float32, device=dev) if want_filtered else None
    with torch.cuda.device(dev):
        _check(_lib.load().hmmb200_forward_chunk_f32(_p(emis), int(mode), float(eps), _p(_f32c(trans_prob, dev)),
                                                     _p(_f32c(init_prob, dev)), B, T, K, _p(state["alpha"]), _p(state["loglik"]),
                                                     _p(state["started"]), _p(filt), _stream(dev)), "hmmb200_forward_chunk_f32")
    return filt


def new_forward_state(B: int, K: int, dev) -> dict:
    return {"alpha": torch.zeros(B, K, dtype=torch.float32, device=dev),
            "loglik": torch.zeros(B, dtype=torch.float64, device=dev),
            "started": torch.zeros(B, dtype=torch.int32, device=dev)}


# ------------------------------------------------------------------------------------------------------
# alignment utilities (csrc/alignment.cu)
# ------------------------------------------------------------------------------------------------------
def ctc_trellis(direction: int, log_probs: torch.Tensor, targets: torch.Tensor, input_lengths: torch.Tensor,
                target_lengths: torch.Tensor, blank: int, want_table: bool = True, want_loglik: bool = True):
    """log_probs [T,B,C] (CUDA), targets [B,L], lengths [B] -> (table [B,T,2L+1] or None, loglik [B] or None).
    direction 0: log alpha and the log-likelihood (ctc.py:32-121); 1: log beta (ctc.py:124-199)."""
    dev = log_probs.device
    lp = _f32c(log_probs, dev)
    T, B, Cn = lp.shape
    tg = targets.to(dev, torch.int64).contiguous()
    L = tg.shape[1]
    il = input_lengths.to(dev, torch.int64).contiguous()
    tl = target_lengths.to(dev, torch.int64).contiguous()
    table = torch.empty(B, T, 2 * L + 1, dtype=torch.float32, device=dev) if want_table else None
    ll = torch.empty(B, dtype=torch.float32, device=dev) if (want_loglik and direction == 0) else None
    with torch.cuda.device(dev):
        _check(_lib.load().hmmb200_ctc_trellis_f32(direction, _p(lp), _p(tg), _p(il), _p(tl), int(blank), T, B, Cn, L, _p(table), _p(ll),
                                                   _stream(dev)), "hmmb200_ctc_trellis_f32")
    return table, ll


def dtw(dist: torch.Tensor, step_pattern: int):
    """dist [P,N,M] (CUDA) -> (cost [P,N,M], path_i [P,N+M-1], path_j [P,N+M-1], path_len [P] int32)."""
    dev = dist.device
    d = _f32c(dist, dev)
    P, N, M = d.shape
    cost = torch.empty_like(d)
    dirs = torch.empty(P * N * M, dtype=torch.uint8, device=dev)
    pi = torch.zeros(P, N + M - 1, dtype=torch.int64, device=dev)
    pj = torch.zeros(P, N + M - 1, dtype=torch.int64, device=dev)
    plen = torch.zeros(P, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _check(_lib.load().hmmb200_dtw_f32(_p(d), P, N, M, int(step_pattern), _p(cost), _p(dirs), _p(pi), _p(pj), _p(plen), _stream(dev)),
               "hmmb200_dtw_f32")
    return cost, pi, pj, plen
