"""pytorch_hmm_b200 -- B200-native (sm_100a) HMM inference engine, drop-in for pytorch_hmm's hot path.

The classes mirror crlotwhite/pytorch_hmm's public API; the work is done by hand-written CUDA kernels behind the
C ABI in include/hmm_b200.h (libhmm_b200.so, loaded with ctypes).  There is no CPU fallback.
"""
from .core import HMM, HMMPyTorch
from .layers import HMMLayer, GaussianHMMLayer
from .gmm import MixtureGaussianHMMLayer
from .hsmm_layer import HSMMLayer, SemiMarkovHMM, DurationModel
from .stream import StreamingHMMProcessor, StreamingResult
from .neural import NeuralHMMRecursion
from .transitions import create_transition_matrix, create_left_to_right_matrix, compute_state_durations
from . import ops

__version__ = "0.1.0"


def create_speech_hmm(num_states: int, feature_dim: int, model_type: str = "mixture_gaussian", **kwargs):
    """Factory of the reference (pytorch_hmm/__init__.py:229-274).  The reference passes some keywords twice and raises
    TypeError for its own README call (SURVEY finding 7); here the README spellings work: `num_mixtures` and
    `num_components` are synonyms, `max_duration` / `chunk_size` / ... are forwarded once."""
    kw = dict(kwargs)
    if model_type == "mixture_gaussian":
        n = kw.pop("num_components", kw.pop("num_mixtures", 3))
        kw.pop("num_mixtures", None)
        return MixtureGaussianHMMLayer(num_states=num_states, feature_dim=feature_dim, num_components=n,
                                       covariance_type=kw.pop("covariance_type", "diag"), **kw)
    if model_type == "hsmm":
        return HSMMLayer(num_states=num_states, feature_dim=feature_dim,
                         duration_distribution=kw.pop("duration_distribution", "gamma"),
                         max_duration=kw.pop("max_duration", 50), **kw)
    if model_type == "streaming":
        return StreamingHMMProcessor(num_states=num_states, feature_dim=feature_dim, chunk_size=kw.pop("chunk_size", 160),
                                     use_beam_search=kw.pop("use_beam_search", True), **kw)
    raise ValueError(f"Unknown model_type: {model_type}. Choose from: 'mixture_gaussian', 'hsmm', 'streaming'")


class ModelFactory:
    """Common configurations (pytorch_hmm/__init__.py:342-376)."""

    @staticmethod
    def create_asr_model(vocabulary_size: int, acoustic_dim: int = 80):
        return MixtureGaussianHMMLayer(num_states=vocabulary_size, feature_dim=acoustic_dim, num_components=4,
                                       covariance_type="diag", learnable_transitions=True)

    @staticmethod
    def create_tts_model(num_phonemes: int, mel_dim: int = 80):
        return HSMMLayer(num_states=num_phonemes, feature_dim=mel_dim, duration_distribution="gamma", max_duration=30,
                         learnable_duration_params=True)

    @staticmethod
    def create_realtime_model(num_states: int, feature_dim: int = 80):
        return StreamingHMMProcessor(num_states=num_states, feature_dim=feature_dim, chunk_size=160,
                                     use_beam_search=False, lookahead_frames=3)


__all__ = ["HMM", "HMMPyTorch", "HMMLayer", "GaussianHMMLayer", "MixtureGaussianHMMLayer", "HSMMLayer", "SemiMarkovHMM",
           "DurationModel", "StreamingHMMProcessor", "StreamingResult", "NeuralHMMRecursion", "create_speech_hmm", "ModelFactory",
           "create_transition_matrix", "create_left_to_right_matrix", "compute_state_durations", "ops"]
