"""Per-layer cache of kernel operands derived from nn.Parameters (packed GMM buffer, floored log tables).

The drop-in layers keep the reference's call pattern -- parameters -> softmax / clamp / log -> kernel operands on EVERY call
(pytorch_hmm/hmm_layer.py:73-89, mixture_gaussian.py:178-179,357) -- but the derived operands only change when a parameter does.
torch bumps `tensor._version` on every in-place write (optimizer steps, copy_, load_state_dict), so (data_ptr, _version, device)
of the source tensors is an exact key: a hit costs no kernel launch, a miss re-derives and re-packs.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Tuple

import torch


def _key(tensors: Iterable[torch.Tensor]) -> Tuple:
    return tuple((t.data_ptr(), t._version, t.device, tuple(t.shape)) if t is not None else None for t in tensors)


class DerivedCache:
    """name -> (key, value).  `get(name, sources, make)` returns the cached value while the source tensors are unchanged."""

    def __init__(self):
        self._slots: Dict[str, Tuple[Tuple, object]] = {}

    def get(self, name: str, sources: Iterable[torch.Tensor], make: Callable[[], object]):
        sources = list(sources)
        k = _key(sources)
        hit = self._slots.get(name)
        if hit is not None and hit[0] == k:
            return hit[1]
        with torch.no_grad():
            v = make()
        self._slots[name] = (k, v)
        return v

    def clear(self):
        self._slots.clear()
