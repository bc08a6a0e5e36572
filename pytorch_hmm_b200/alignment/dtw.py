"""Dynamic time warping on the engine: drop-in for pytorch_hmm/alignment/dtw.py (hard DTW; the reference's soft variant returns an
approximate linear path and is not part of the hot path).

`compute_dtw_path` keeps the reference's signature and returns (path_i, path_j, cost_matrix) (dtw.py:47-153); costs and paths are
bit-identical to the reference's (fp32 adds and minima, the reference's tie order).  The wavefront kernel needs 3*N floats of shared
memory, so the FIRST sequence may have at most ~17 000 frames.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

STEP_PATTERNS = {"symmetric": 0, "asymmetric": 1, "rabiner_juang": 2}


def compute_distance_matrix(x: torch.Tensor, y: torch.Tensor, distance_fn: str = "euclidean") -> torch.Tensor:
    """[N,D], [M,D] -> [N,M] with the reference's formulas (dtw.py:8-44)."""
    if distance_fn == "euclidean":
        return torch.norm(x.unsqueeze(1) - y.unsqueeze(0), dim=2)
    if distance_fn == "cosine":
        return 1 - torch.mm(F.normalize(x, p=2, dim=1), F.normalize(y, p=2, dim=1).t())
    if distance_fn == "manhattan":
        return torch.sum(torch.abs(x.unsqueeze(1) - y.unsqueeze(0)), dim=2)
    raise ValueError(f"Unknown distance function: {distance_fn}")


def compute_dtw_path(distance_matrix: torch.Tensor, step_pattern: str = "symmetric") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    if step_pattern not in STEP_PATTERNS:
        raise ValueError(f"Unknown step pattern: {step_pattern}")
    dev = ops.require_cuda(distance_matrix.device if distance_matrix.is_cuda else None)
    cost, pi, pj, plen = ops.dtw(distance_matrix.detach().to(dev).unsqueeze(0), STEP_PATTERNS[step_pattern])
    n = int(plen[0])                                                      # (the path length is data dependent: one host read)
    out_dev = distance_matrix.device
    return pi[0, :n].to(out_dev), pj[0, :n].to(out_dev), cost[0].to(out_dev)


def dtw_distance(x, y, distance_fn: str = "euclidean", step_pattern: str = "symmetric") -> torch.Tensor:
    d = compute_distance_matrix(x, y, distance_fn)
    _, _, cost = compute_dtw_path(d, step_pattern)
    return cost[-1, -1]


def dtw_alignment(x, y, distance_fn: str = "euclidean", step_pattern: str = "symmetric"):
    d = compute_distance_matrix(x, y, distance_fn)
    pi, pj, cost = compute_dtw_path(d, step_pattern)
    return pi, pj, cost[-1, -1]


class DTWAligner(nn.Module):
    """dtw.py:205-269 (hard DTW).  A batch [B,N,D] x [B,M,D] is one kernel launch (one CTA per pair)."""

    def __init__(self, distance_fn: str = "euclidean", step_pattern: str = "symmetric", bandwidth=None, soft_dtw: bool = False,
                 gamma: float = 0.1):
        super().__init__()
        if soft_dtw:
            raise NotImplementedError("soft DTW is outside the B200 hot path (the reference returns a linear approximate path for it)")
        self.distance_fn, self.step_pattern, self.bandwidth, self.soft_dtw, self.gamma = distance_fn, step_pattern, bandwidth, soft_dtw, gamma

    def forward(self, x: torch.Tensor, y: torch.Tensor):
        if x.dim() == 3:
            dev = ops.require_cuda(x.device if x.is_cuda else None)
            d = torch.stack([compute_distance_matrix(x[b], y[b], self.distance_fn) for b in range(x.shape[0])]).to(dev)
            cost, pi, pj, plen = ops.dtw(d, STEP_PATTERNS[self.step_pattern])
            lens = plen.tolist()
            return ([pi[b, :n].to(x.device) for b, n in enumerate(lens)], [pj[b, :n].to(x.device) for b, n in enumerate(lens)],
                    cost[:, -1, -1].to(x.device))
        return dtw_alignment(x, y, self.distance_fn, self.step_pattern)


def extract_phoneme_durations(alignment: torch.Tensor, num_phonemes: int) -> torch.Tensor:
    """frames per phoneme index (dtw.py:387-403)."""
    return torch.bincount(alignment.long().clamp(min=0), minlength=num_phonemes)[:num_phonemes].to(torch.long)
