"""CTC trellises on the engine: drop-in for pytorch_hmm/alignment/ctc.py.

`ctc_forward_algorithm` / `ctc_backward_algorithm` keep the reference's signatures, shapes and conventions (ctc.py:32-199): time-major
log-probabilities [T,B,C], padded targets [B,L], -inf outside an utterance's length and outside its expanded target.  One kernel launch
per call (csrc/alignment.cu) instead of T x B x (2L+1) Python iterations.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from .. import ops


def expand_targets_with_blank(targets: torch.Tensor, blank_id: int) -> torch.Tensor:
    """[B,L] -> [B,2L+1] with blanks at the even positions (ctc.py:8-29)."""
    B, L = targets.shape
    out = torch.full((B, 2 * L + 1), blank_id, device=targets.device)
    out[:, 1::2] = targets
    return out


def _trellis(direction, log_probs, targets, input_lengths, target_lengths, blank_id, want_table, want_loglik):
    dev = ops.require_cuda(log_probs.device if log_probs.is_cuda else None)
    return ops.ctc_trellis(direction, log_probs.detach().to(dev), targets.to(dev), input_lengths.to(dev), target_lengths.to(dev),
                           blank_id, want_table, want_loglik)


def ctc_forward_algorithm(log_probs: torch.Tensor, targets: torch.Tensor, input_lengths: torch.Tensor,
                          target_lengths: torch.Tensor, blank_id: int = 0) -> torch.Tensor:
    """log-likelihood [B] of the targets under the frame posteriors (ctc.py:32-121)."""
    _, ll = _trellis(0, log_probs, targets, input_lengths, target_lengths, blank_id, False, True)
    return ll if ll.device == log_probs.device else ll.to(log_probs.device)


def ctc_forward_trellis(log_probs, targets, input_lengths, target_lengths, blank_id: int = 0):
    """(log alpha [B,T,2L+1], log-likelihood [B]): the table the reference builds internally."""
    return _trellis(0, log_probs, targets, input_lengths, target_lengths, blank_id, True, True)


def ctc_backward_algorithm(log_probs: torch.Tensor, targets: torch.Tensor, input_lengths: torch.Tensor,
                           target_lengths: torch.Tensor, blank_id: int = 0) -> torch.Tensor:
    """log beta [B,T,2L+1] (ctc.py:124-199)."""
    tab, _ = _trellis(1, log_probs, targets, input_lengths, target_lengths, blank_id, True, False)
    return tab if tab.device == log_probs.device else tab.to(log_probs.device)


def ctc_alignment_path(log_probs, targets, input_lengths, target_lengths, blank_id: int = 0) -> List[torch.Tensor]:
    """The reference's ctc_alignment_path (ctc.py:202-256), result for result: it combines log beta with a log alpha table that it
    never fills (all -inf), so every frame's best position is 0 and the returned token is the blank-expanded target's first entry.
    `ctc_posterior_alignment` below is the alignment that routine describes."""
    dev = log_probs.device
    expanded = expand_targets_with_blank(targets, blank_id)
    return [expanded[b, 0].repeat(int(input_lengths[b])).to(dev) for b in range(log_probs.shape[1])]


def ctc_posterior_alignment(log_probs, targets, input_lengths, target_lengths, blank_id: int = 0) -> List[torch.Tensor]:
    """Per frame, the token of the expanded-target position with the largest log alpha + log beta (what ctc.py:202-256 sets out to do)."""
    la, _ = _trellis(0, log_probs, targets, input_lengths, target_lengths, blank_id, True, False)
    lb, _ = _trellis(1, log_probs, targets, input_lengths, target_lengths, blank_id, True, False)
    expanded = expand_targets_with_blank(targets, blank_id).to(la.device)
    pos = (la + lb).argmax(-1)                                            # [B,T]
    tok = expanded.gather(1, pos)
    return [tok[b, : int(input_lengths[b])].to(log_probs.device) for b in range(tok.shape[0])]


def remove_ctc_blanks(sequence: torch.Tensor, blank_id: int = 0) -> torch.Tensor:
    return sequence[sequence != blank_id]


def collapse_repeated_tokens(sequence: torch.Tensor) -> torch.Tensor:
    if len(sequence) == 0:
        return sequence
    keep = torch.ones_like(sequence, dtype=torch.bool)
    keep[1:] = sequence[1:] != sequence[:-1]
    return sequence[keep]


def ctc_decode_sequence(sequence: torch.Tensor, blank_id: int = 0) -> torch.Tensor:
    return remove_ctc_blanks(collapse_repeated_tokens(sequence), blank_id)


class CTCAligner(nn.Module):
    """ctc.py:259-381: loss (torch's CTC loss, as in the reference), greedy decoding, forced alignment."""

    def __init__(self, num_classes: int, blank_id: int = 0, reduction: str = "mean"):
        super().__init__()
        self.num_classes, self.blank_id, self.reduction = num_classes, blank_id, reduction
        self.ctc_loss = nn.CTCLoss(blank=blank_id, reduction=reduction, zero_infinity=True)

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        flat = torch.cat([targets[b, : target_lengths[b]] for b in range(targets.shape[0])])
        return self.ctc_loss(log_probs, flat, input_lengths, target_lengths)

    def log_likelihood(self, log_probs, targets, input_lengths, target_lengths):
        return ctc_forward_algorithm(log_probs, targets, input_lengths, target_lengths, self.blank_id)

    def decode(self, log_probs, input_lengths, beam_width: int = 1):
        best = log_probs.argmax(-1)                                       # [T,B]  (beam search falls back to greedy, ctc.py:355-362)
        return [ctc_decode_sequence(best[: int(input_lengths[b]), b], self.blank_id) for b in range(log_probs.shape[1])]

    def align(self, log_probs, targets, input_lengths, target_lengths):
        return ctc_alignment_path(log_probs, targets, input_lengths, target_lengths, self.blank_id)
