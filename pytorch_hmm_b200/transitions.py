"""Transition-matrix builders (host-side, O(K^2), run once).

Same names, arguments and values as the reference's pytorch_hmm/utils.py:9-103 (`create_transition_matrix`,
`create_left_to_right_matrix`); they are the input generators of BASELINE configs 1 and 5.
"""
from __future__ import annotations

import torch


def create_transition_matrix(num_states: int, transition_type: str = "ergodic", self_loop_prob: float = 0.5,
                             forward_prob: float = 0.4, skip_prob: float = 0.1, device: str = "cpu") -> torch.Tensor:
    K = num_states
    idx = torch.arange(K, device=device)
    P = torch.zeros(K, K, device=device)
    if transition_type == "ergodic":
        P = torch.ones(K, K, device=device) + torch.eye(K, device=device) * self_loop_prob * K
    elif transition_type in ("left_to_right", "left_to_right_skip"):
        P[idx, idx] = self_loop_prob
        P[idx[:-1], idx[:-1] + 1] = forward_prob
        if transition_type == "left_to_right_skip" and K > 2:
            P[idx[:-2], idx[:-2] + 2] = skip_prob
        P[K - 1, K - 1] = 1.0
    elif transition_type == "circular":
        P[idx, idx] = self_loop_prob
        P[idx, (idx + 1) % K] = forward_prob
    else:
        raise ValueError(f"Unknown transition_type: {transition_type}")
    return P / P.sum(dim=1, keepdim=True)


def create_left_to_right_matrix(num_states: int, self_loop_prob: float = 0.7, device: str = "cpu") -> torch.Tensor:
    return create_transition_matrix(num_states, "left_to_right", self_loop_prob=self_loop_prob,
                                    forward_prob=1.0 - self_loop_prob, device=device)


def compute_state_durations(state_sequence: torch.Tensor) -> torch.Tensor:
    """Run lengths of a state path [T] -> [n_runs] (utils.py:447-476), as one vectorised pass instead of a Python loop over frames."""
    n = len(state_sequence)
    if n == 0:
        return torch.tensor([])
    s = state_sequence.reshape(-1)
    change = torch.ones(n, dtype=torch.bool, device=s.device)
    change[1:] = s[1:] != s[:-1]
    starts = change.nonzero().flatten()
    ends = torch.cat([starts[1:], torch.tensor([n], device=s.device)])
    return (ends - starts).to(torch.long)


def validate_transition_matrix(P: torch.Tensor, tolerance: float = 1e-6) -> dict:
    """Checks of a transition matrix with the reference's result keys (utils.py: validate_transition_matrix)."""
    row_sums = P.sum(dim=1)
    return {"row_sums_valid": bool(torch.allclose(row_sums, torch.ones_like(row_sums), atol=tolerance)),
            "non_negative": bool((P >= 0).all()), "finite": bool(torch.isfinite(P).all()),
            "shape_valid": P.dim() == 2 and P.shape[0] == P.shape[1],
            "min_value": float(P.min()), "max_value": float(P.max()),
            "max_row_sum_error": float((row_sums - 1).abs().max())}
