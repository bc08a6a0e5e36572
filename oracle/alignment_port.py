"""oracle/alignment_port.py -- numpy restatement of the reference's CTC trellises and DTW.  TEST INFRASTRUCTURE ONLY.

Nothing under ``pytorch_hmm_b200/`` imports this module (tests only).  Follows pytorch_hmm/alignment/ctc.py:32-199 and
pytorch_hmm/alignment/dtw.py:47-153 cell for cell (fp32 arithmetic; the log-sum-exp is torch.logsumexp's max / sum-exp / log / add
in float32).  Parity status: pinned -- tests/test_oracle_golden.py checks it against tests/golden/alignment.npz, written by
oracle/make_golden.py from the real reference.
"""
from __future__ import annotations

import numpy as np

NEG = np.float32(-np.inf)


def expand_targets(targets: np.ndarray, blank: int) -> np.ndarray:
    """ctc.py:8-29"""
    B, L = targets.shape
    out = np.full((B, 2 * L + 1), blank, dtype=np.int64)
    out[:, 1::2] = targets
    return out


def _lse(c):
    c = np.asarray(c, dtype=np.float32)
    m = c.max()
    if not np.isfinite(m):
        return np.float32(m)
    return np.float32(np.log(np.exp(c - m, dtype=np.float32).sum(dtype=np.float32), dtype=np.float32) + m)


def ctc_forward(lp: np.ndarray, targets: np.ndarray, in_len: np.ndarray, tg_len: np.ndarray, blank: int = 0):
    """ctc.py:32-121 -> (log_alpha [B,T,2L+1], log_likelihood [B])"""
    T, B, _ = lp.shape
    ext = expand_targets(targets, blank)
    S = ext.shape[1]
    la = np.full((B, T, S), NEG, dtype=np.float32)
    for b in range(B):
        la[b, 0, 0] = lp[0, b, blank]
        if tg_len[b] > 0:
            la[b, 0, 1] = lp[0, b, ext[b, 1]]
    for t in range(1, T):
        for b in range(B):
            if t >= in_len[b]:
                continue
            E = 2 * int(tg_len[b]) + 1
            for s in range(min(E, S)):
                c = []
                if la[b, t - 1, s] > NEG:
                    c.append(la[b, t - 1, s])
                if s > 0 and la[b, t - 1, s - 1] > NEG:
                    c.append(la[b, t - 1, s - 1])
                if s > 1 and la[b, t - 1, s - 2] > NEG and ext[b, s] != ext[b, s - 2]:
                    c.append(la[b, t - 1, s - 2])
                if c:
                    la[b, t, s] = np.float32(lp[t, b, ext[b, s]] + _lse(c))
    ll = np.full((B,), NEG, dtype=np.float32)
    for b in range(B):
        ti = int(in_len[b]) - 1
        E = 2 * int(tg_len[b]) + 1
        c = []
        if E >= 1:
            c.append(la[b, ti, E - 1])
        if E >= 2:
            c.append(la[b, ti, E - 2])
        if c:
            ll[b] = _lse(c)
    return la, ll


def ctc_backward(lp: np.ndarray, targets: np.ndarray, in_len: np.ndarray, tg_len: np.ndarray, blank: int = 0):
    """ctc.py:124-199 -> log_beta [B,T,2L+1]"""
    T, B, _ = lp.shape
    ext = expand_targets(targets, blank)
    S = ext.shape[1]
    lb = np.full((B, T, S), NEG, dtype=np.float32)
    for b in range(B):
        ti = int(in_len[b]) - 1
        E = 2 * int(tg_len[b]) + 1
        if E >= 1:
            lb[b, ti, E - 1] = 0.0
        if E >= 2:
            lb[b, ti, E - 2] = 0.0
    for t in range(T - 2, -1, -1):
        for b in range(B):
            if t >= in_len[b]:
                continue
            E = 2 * int(tg_len[b]) + 1
            for s in range(min(E, S)):
                c = []
                if lb[b, t + 1, s] > NEG:
                    c.append(np.float32(lb[b, t + 1, s] + lp[t + 1, b, ext[b, s]]))
                if s + 1 < E and s + 1 < S and lb[b, t + 1, s + 1] > NEG:
                    c.append(np.float32(lb[b, t + 1, s + 1] + lp[t + 1, b, ext[b, s + 1]]))
                if s + 2 < E and s + 2 < S and lb[b, t + 1, s + 2] > NEG and ext[b, s] != ext[b, s + 2]:
                    c.append(np.float32(lb[b, t + 1, s + 2] + lp[t + 1, b, ext[b, s + 2]]))
                if c:
                    lb[b, t, s] = _lse(c)
    return lb


def dtw(dist: np.ndarray, pattern: str = "symmetric"):
    """dtw.py:47-153 -> (path_i, path_j, cost [N,M]); anti-diagonal order gives the same numbers as the reference's row-major loops."""
    dist = dist.astype(np.float32)
    N, M = dist.shape
    cost = np.full((N, M), np.inf, dtype=np.float32)
    cost[0, 0] = dist[0, 0]
    for i in range(N):
        for j in range(M):
            if i == 0 and j == 0:
                continue
            d = dist[i, j]
            c = []
            if i > 0 and j > 0:
                c.append(np.float32(cost[i - 1, j - 1] + (np.float32(2) * d if pattern == "rabiner_juang" else d)) if pattern != "symmetric" else cost[i - 1, j - 1])
            if i > 0:
                c.append(np.float32(cost[i - 1, j] + d) if pattern != "symmetric" else cost[i - 1, j])
            if j > 0:
                c.append(np.float32(cost[i, j - 1] + d) if pattern != "symmetric" else cost[i, j - 1])
            cost[i, j] = np.float32(d + min(c)) if pattern == "symmetric" else min(c)
    pi, pj = [], []
    i, j = N - 1, M - 1
    while i > 0 or j > 0:
        pi.append(i)
        pj.append(j)
        c = []
        if i > 0 and j > 0:
            c.append((cost[i - 1, j - 1], i - 1, j - 1))
        if i > 0:
            c.append((cost[i - 1, j], i - 1, j))
        if j > 0:
            c.append((cost[i, j - 1], i, j - 1))
        _, i, j = min(c)
    pi.append(0)
    pj.append(0)
    return np.array(pi[::-1], dtype=np.int64), np.array(pj[::-1], dtype=np.int64), cost
