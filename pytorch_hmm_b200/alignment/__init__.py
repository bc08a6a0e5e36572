"""Alignment utilities on the B200 engine: the reference's `pytorch_hmm.alignment` names for the CTC trellises and
dynamic time warping (pytorch_hmm/alignment/__init__.py:13-33).  The attention-based aligners of the reference are outside the
hot path and are not provided."""
from .ctc import (CTCAligner, collapse_repeated_tokens, ctc_alignment_path, ctc_backward_algorithm, ctc_decode_sequence,
                  ctc_forward_algorithm, ctc_posterior_alignment, expand_targets_with_blank, remove_ctc_blanks)
from .dtw import DTWAligner, compute_distance_matrix, compute_dtw_path, dtw_alignment, dtw_distance, extract_phoneme_durations

__all__ = [
    "DTWAligner", "dtw_alignment", "compute_dtw_path", "dtw_distance", "compute_distance_matrix", "extract_phoneme_durations",
    "CTCAligner", "ctc_alignment_path", "ctc_posterior_alignment", "expand_targets_with_blank", "ctc_forward_algorithm",
    "ctc_backward_algorithm", "remove_ctc_blanks", "collapse_repeated_tokens", "ctc_decode_sequence",
]
