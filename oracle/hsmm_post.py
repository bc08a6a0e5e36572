"""Float64 posteriors of the explicit-duration HMM on top of oracle/hmm_oracle.c.  TEST INFRASTRUCTURE ONLY.

The reference has no HSMM backward pass (SURVEY finding 5; the only forward recursion is semi_markov.py:308-383), so this
row is "parity unpinned" by the reference: the float64 alpha/beta of the C restatement are combined here, and
tests/test_host_cpu.py checks the result against brute-force enumeration of every segmentation for tiny T."""
from __future__ import annotations

import itertools

import numpy as np

from . import c_oracle


def seg_table(f, segc, Dm):
    """f [T,K] per-frame log terms, segc [K] per-segment constant -> seg[t,s,d-1] = segc[s] + sum f[t-d+1..t, s]."""
    T, K = f.shape
    seg = np.full((T, K, Dm), -np.inf)
    for t in range(T):
        for d in range(1, min(Dm, t + 1) + 1):
            seg[t, :, d - 1] = segc + f[t - d + 1:t + 1].astype(np.float64).sum(0)
    return seg


def posteriors_f64(f, segc, logdur, logA, logpi):
    """-> (gamma [T,K], total).  gamma_t(s) = P(state_t = s | o) over all segmentations."""
    T, K = f.shape
    Dm = logdur.shape[1]
    alpha, beta, tot = c_oracle.hsmm_forward_f64(seg_table(f, segc, Dm), logdur, logA, logpi)
    gamma = np.zeros((T, K))
    for t in range(T):                      # segment (s, d) ending at t covers frames t-d+1..t
        for d in range(1, min(Dm, t + 1) + 1):
            w = np.exp(alpha[t, :, d - 1] + beta[t] - tot)
            gamma[t - d + 1:t + 1] += w[None, :]
    return gamma, tot


def brute_force(f, segc, logdur, logA, logpi):
    """Enumerates every segmentation (tiny T only) -> (gamma [T,K], total)."""
    T, K = f.shape
    Dm = logdur.shape[1]
    f = f.astype(np.float64)
    paths = []

    def comps(n):
        if n == 0:
            yield ()
            return
        for d in range(1, min(Dm, n) + 1):
            for rest in comps(n - d):
                yield (d,) + rest

    logs, occ = [], []
    for durs in comps(T):
        n = len(durs)
        for states in itertools.product(range(K), repeat=n):
            if any(states[i] == states[i + 1] for i in range(n - 1)):
                continue
            lp = logpi[states[0]]
            t = 0
            o = np.zeros((T, K))
            for i, (s, d) in enumerate(zip(states, durs)):
                if i > 0:
                    lp += logA[states[i - 1], s]
                lp += segc[s] + f[t:t + d, s].sum() + logdur[s, d - 1]
                o[t:t + d, s] = 1.0
                t += d
            logs.append(lp); occ.append(o)
    logs = np.array(logs, np.float64)
    m = logs.max()
    tot = m + np.log(np.exp(logs - m).sum())
    w = np.exp(logs - tot)
    gamma = np.tensordot(w, np.array(occ), axes=(0, 0))
    return gamma, tot
