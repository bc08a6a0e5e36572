"""ctypes binding of libhmm_b200.so -- the only door from Python into the CUDA kernels.

The signatures below are exactly those declared in include/hmm_b200.h.  There is no fallback: if the library
cannot be loaded (or built with nvcc), importing the compute layer raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

c_f32p = C.c_void_p     # device pointers travel as raw addresses
c_ptr = C.c_void_p

# name -> (restype, argtypes); mirrors include/hmm_b200.h one to one
SIGNATURES = {
    "hmmb200_abi_version": (C.c_int, []),
    "hmmb200_last_error": (C.c_char_p, []),
    "hmmb200_device_check": (C.c_int, [C.c_int]),
    "hmmb200_gmm_packed_floats": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hmmb200_gmm_pack_f32": (C.c_int, [c_ptr, c_ptr, C.c_float, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr]),
    "hmmb200_gmm_emission_f32": (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr]),
    "hmmb200_gmm_pack_on_tensor_cores": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, c_ptr]),
    "hmmb200_gmm_emission_tc_f32": (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr]),
    "hmmb200_gmm_emission_full_f32": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_int64, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr]),
    "hmmb200_ctc_trellis_f32": (C.c_int, [C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr]),
    "hmmb200_dtw_f32": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "hmmb200_fb_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hmmb200_forward_backward_f32": (C.c_int, [c_ptr, C.c_int, C.c_float, C.c_int, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int,
                                               c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "hmmb200_fb_scan_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hmmb200_forward_backward_scan_f32": (C.c_int, [c_ptr, C.c_int, C.c_float, C.c_int, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int,
                                                    c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "hmmb200_viterbi_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hmmb200_viterbi_f32": (C.c_int, [c_ptr, C.c_int, C.c_float, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int,
                                      c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "hmmb200_fb_viterbi_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hmmb200_fb_viterbi_f32": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_float, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int,
                                         c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                         c_ptr, C.c_size_t, C.c_int, c_ptr]),
    "hmmb200_tv_forward_backward_f32": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                                  c_ptr, C.c_size_t, c_ptr]),
    "hmmb200_tv_viterbi_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hmmb200_tv_viterbi_f32": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                         c_ptr, C.c_size_t, c_ptr]),
    "hmmb200_posterior_backward_f32": (C.c_int, [c_ptr, C.c_int, C.c_float, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int,
                                                 c_ptr, c_ptr, c_ptr, c_ptr]),
    "hmmb200_gmm_stats_f32": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_int64, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "hmmb200_hsmm_viterbi_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "hmmb200_hsmm_viterbi_f32": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "hmmb200_hsmm_forward_f32": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int,
                                           c_ptr, c_ptr, c_ptr, c_ptr]),
    "hmmb200_hsmm_fb_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hmmb200_hsmm_forward_backward_f32": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int,
                                                    c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "hmmb200_greedy_decode_f32": (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "hmmb200_forward_chunk_f32": (C.c_int, [c_ptr, C.c_int, C.c_float, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int,
                                            c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "hmmb200_gmm_components_f32": (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr]),
    "hmmb200_gmm_emission_components_f32": (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr]),
    "hmmb200_bw_stats_doubles": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hmmb200_xi_sum_f32": (C.c_int, [c_ptr, C.c_int, C.c_float, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr]),
    "hmmb200_bw_accumulate_f32": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_float, c_ptr, c_ptr,
                                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr]),
}


def lib_path() -> str:
    # HMMB200_LIB_PATH: load another build of the same ABI (A/B timing of kernel variants); default = the in-tree build
    return os.environ.get("HMMB200_LIB_PATH") or _build.lib_path()


def load(build_if_missing: bool = True):
    """Returns the loaded CDLL.  The in-tree library is (re)built first when it is missing or was built from other sources
    than the ones in the tree (content hash recorded at build time) and nvcc is available; a stale library that cannot be
    rebuilt is an error -- argument-layout drift between header and binary must never be loaded silently."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        in_tree = path == _build.lib_path()
        force = os.environ.get("HMMB200_REBUILD") == "1"
        if in_tree and build_if_missing and (force or _build.is_stale(path)):
            import shutil
            if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
                path = _build.build_library(force=True)
            elif os.path.exists(path):
                raise RuntimeError(f"{path} was built from different sources than the tree holds and nvcc is not available "
                                   "to rebuild it")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `python -m pytorch_hmm_b200.build` "
                               "(pytorch_hmm_b200 has no CPU fallback)")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        if lib.hmmb200_abi_version() != 1:
            raise RuntimeError("libhmm_b200.so ABI version mismatch")
        _lib = lib
        return _lib


def last_error() -> str:
    return load().hmmb200_last_error().decode("utf-8", "replace")
