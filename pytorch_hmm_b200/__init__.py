"""pytorch_hmm_b200 -- B200-native (sm_100a) HMM inference engine, drop-in for pytorch_hmm's hot path.

The classes mirror crlotwhite/pytorch_hmm's public API; the work is done by hand-written CUDA kernels behind the
C ABI in include/hmm_b200.h (libhmm_b200.so, loaded with ctypes).  There is no CPU fallback.
"""
from .core import HMM, HMMPyTorch
from .layers import HMMLayer, GaussianHMMLayer
from .gmm import MixtureGaussianHMMLayer
from .transitions import create_transition_matrix, create_left_to_right_matrix
from . import ops

__version__ = "0.1.0"

__all__ = ["HMM", "HMMPyTorch", "HMMLayer", "GaussianHMMLayer", "MixtureGaussianHMMLayer",
           "create_transition_matrix", "create_left_to_right_matrix", "ops"]
