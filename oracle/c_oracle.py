"""ctypes front-end of oracle/liboracle.so (built from oracle/hmm_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "hmm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def viterbi_f32(logb, logP, logp0):
    logb, logP, logp0 = _f32(logb), _f32(logP), _f32(logp0)
    B, T, K = logb.shape
    delta = np.empty((B, T, K), np.float32); psi = np.empty((B, T, K), np.int32)
    states = np.empty((B, T), np.int64); score = np.empty((B,), np.float32)
    lib().orc_viterbi_f32(_p(logb, C.c_float), _p(logP, C.c_float), _p(logp0, C.c_float), B, T, K,
                          _p(delta, C.c_float), _p(psi, C.c_int32), _p(states, C.c_int64), _p(score, C.c_float))
    return states, delta, psi, score


def forward_backward_f64(logb, logP, logp0):
    logb, logP, logp0 = _f64(logb), _f64(logP), _f64(logp0)
    B, T, K = logb.shape
    la = np.empty((B, T, K)); lb = np.empty((B, T, K)); g = np.empty((B, T, K)); ll = np.empty((B,))
    lib().orc_forward_backward_f64(_p(logb, C.c_double), _p(logP, C.c_double), _p(logp0, C.c_double), B, T, K,
                                   _p(la, C.c_double), _p(lb, C.c_double), _p(g, C.c_double), _p(ll, C.c_double))
    return la, lb, g, ll


def forward_backward_f32(logb, logP, logp0):
    logb, logP, logp0 = _f32(logb), _f32(logP), _f32(logp0)
    B, T, K = logb.shape
    la = np.empty((B, T, K), np.float32); lb = np.empty((B, T, K), np.float32); g = np.empty((B, T, K), np.float32)
    lib().orc_forward_backward_f32(_p(logb, C.c_float), _p(logP, C.c_float), _p(logp0, C.c_float), B, T, K,
                                   _p(la, C.c_float), _p(lb, C.c_float), _p(g, C.c_float))
    return la, lb, g


def gmm_emission_f64(x, means, log_vars, log_var_scale, logw):
    """x [..., D]; means/log_vars [K, C, D] (or [K, D] for a single Gaussian); logw [K, C] or None."""
    x = _f32(x); means = _f32(means); log_vars = _f32(log_vars)
    if means.ndim == 2:
        means = means[:, None, :]; log_vars = log_vars[:, None, :]
    K, Cc, D = means.shape
    N = x.size // D
    out = np.empty(x.shape[:-1] + (K,), np.float64)
    lw = None if logw is None else _f32(logw)
    lib().orc_gmm_emission_f64(_p(x, C.c_float), _p(means, C.c_float), _p(log_vars, C.c_float),
                               C.c_double(log_var_scale), _p(lw, C.c_float), C.c_int64(N), K, Cc, D,
                               _p(out, C.c_double))
    return out


def hsmm_viterbi_f32(logb, logdur, logA):
    logb, logdur, logA = _f32(logb), _f32(logdur), _f32(logA)
    B, T, K = logb.shape
    Dm = logdur.shape[1]
    states = np.empty((B, T), np.int64); score = np.empty((B,), np.float32)
    lib().orc_hsmm_viterbi_f32(_p(logb, C.c_float), _p(logdur, C.c_float), _p(logA, C.c_float), B, T, K, Dm,
                               _p(states, C.c_int64), _p(score, C.c_float))
    return states, score


def hsmm_forward_f64(seg, logdur, logA, logpi, want_beta=True):
    seg, logdur, logA, logpi = _f64(seg), _f64(logdur), _f64(logA), _f64(logpi)
    T, K, Dm = seg.shape
    alpha = np.empty((T, K, Dm)); beta = np.empty((T, K)) if want_beta else None
    tot = C.c_double(0.0)
    lib().orc_hsmm_forward_f64(_p(seg, C.c_double), _p(logdur, C.c_double), _p(logA, C.c_double),
                               _p(logpi, C.c_double), T, K, Dm, _p(alpha, C.c_double), _p(beta, C.c_double),
                               C.byref(tot))
    return alpha, beta, tot.value


def bw_stats_f64(x, comp, logP, logp0):
    x = _f32(x); comp = _f64(comp); logP = _f64(logP); logp0 = _f64(logp0)
    B, T, D = x.shape
    K, Cc = comp.shape[2], comp.shape[3]
    g1 = np.zeros((K,)); xi = np.zeros((K, K)); occ = np.zeros((K, Cc))
    sx = np.zeros((K, Cc, D)); sxx = np.zeros((K, Cc, D)); ll = np.zeros((1,))
    lib().orc_bw_stats_f64(_p(x, C.c_float), _p(comp, C.c_double), _p(logP, C.c_double), _p(logp0, C.c_double),
                           B, T, K, Cc, D, _p(g1, C.c_double), _p(xi, C.c_double), _p(occ, C.c_double),
                           _p(sx, C.c_double), _p(sxx, C.c_double), _p(ll, C.c_double))
    return {"gamma1": g1, "xi": xi, "occ": occ, "sx": sx, "sxx": sxx, "loglik": float(ll[0])}


def greedy_decode_f32(logb, logA, prev_state=-1):
    logb, logA = _f32(logb), _f32(logA)
    T, K = logb.shape
    states = np.empty((T,), np.int64); scores = np.empty((T,), np.float32)
    lib().orc_greedy_decode_f32(_p(logb, C.c_float), _p(logA, C.c_float), T, K, int(prev_state),
                                _p(states, C.c_int64), _p(scores, C.c_float))
    return states, scores


def tv_viterbi_f32(logb, logT, logp0):
    """NeuralHMM Viterbi with time-varying transitions (neural.py:463-511): logT [B,T,K,K], slice t carries frame t -> t+1."""
    logb, logT, logp0 = _f32(logb), _f32(logT), _f32(logp0)
    B, T, K = logb.shape
    delta = np.empty((B, T, K), np.float32); psi = np.empty((B, T, K), np.int32); states = np.empty((B, T), np.int64)
    lib().orc_tv_viterbi_f32(_p(logb, C.c_float), _p(logT, C.c_float), _p(logp0, C.c_float), B, T, K,
                             _p(delta, C.c_float), _p(psi, C.c_int32), _p(states, C.c_int64))
    return states, delta, psi


def tv_forward_backward_f64(logb, logT, logp0):
    logb, logT, logp0 = _f64(logb), _f64(logT), _f64(logp0)
    B, T, K = logb.shape
    la = np.empty((B, T, K)); lb = np.empty((B, T, K)); g = np.empty((B, T, K)); ll = np.empty((B,))
    lib().orc_tv_forward_backward_f64(_p(logb, C.c_double), _p(logT, C.c_double), _p(logp0, C.c_double), B, T, K,
                                      _p(la, C.c_double), _p(lb, C.c_double), _p(g, C.c_double), _p(ll, C.c_double))
    return la, lb, g, ll
