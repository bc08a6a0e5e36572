"""Runs the REFERENCE's own test files, unchanged, against this package.  TEST INFRASTRUCTURE ONLY.

`install()` makes `import pytorch_hmm` (and the three submodules the reference tests import from) resolve to pytorch_hmm_b200.  Classes
of the reference that are outside the hot path (the duration-penalty toy model DurationConstrainedHMM, the
AdaptiveLatencyController of the streaming front end) get stand-ins that SKIP the tests that construct them, so that the rest of
their test files still collect and run.
"""
import sys
import types

import pytest


def _out_of_scope(name):
    class _Skip:
        def __init__(self, *a, **k):
            pytest.skip(f"{name} is outside the B200 hot path (SURVEY section 8); not provided by pytorch_hmm_b200")
    _Skip.__name__ = name
    return _Skip


def install():
    import pytorch_hmm_b200 as pkg
    from pytorch_hmm_b200 import gmm, hsmm_layer, layers, stream, transitions
    sys.modules["pytorch_hmm"] = pkg
    mg = types.ModuleType("pytorch_hmm.mixture_gaussian")
    mg.MixtureGaussianHMMLayer = gmm.MixtureGaussianHMMLayer
    hs = types.ModuleType("pytorch_hmm.hsmm")
    hs.HSMMLayer = hsmm_layer.HSMMLayer
    hs.DurationConstrainedHMM = _out_of_scope("DurationConstrainedHMM")
    st = types.ModuleType("pytorch_hmm.streaming")
    st.StreamingHMMProcessor = stream.StreamingHMMProcessor
    st.StreamingResult = stream.StreamingResult
    st.AdaptiveLatencyController = _out_of_scope("AdaptiveLatencyController")
    hl = types.ModuleType("pytorch_hmm.hmm_layer")
    hl.HMMLayer, hl.GaussianHMMLayer = layers.HMMLayer, layers.GaussianHMMLayer
    ut = types.ModuleType("pytorch_hmm.utils")
    for name in ("create_transition_matrix", "create_left_to_right_matrix", "compute_state_durations", "validate_transition_matrix"):
        setattr(ut, name, getattr(transitions, name))
    for m in (mg, hs, st, hl, ut):
        sys.modules[m.__name__] = m
