"""HMMInferenceEngine -- the batch-throughput front end of the hot path (emission -> forward-backward + Viterbi).

The reference runs `layer.get_observation_log_probs(x)`, `hmm.forward_backward(...)` and `viterbi_decode(...)` one
after the other on a whole batch (examples/benchmark.py:120-196).  On a B200 one such pass is bound by the LATENCY of
the T dependent recursion steps, not by bandwidth, and a host-resident batch is bound by the PCIe copy of x.  The
engine therefore cuts the batch of utterances into shards (utterances are independent: hmm.py:98 broadcasts one
[K,K] matrix against [B,K]) and keeps `n_streams` shards in flight on their own CUDA streams, each with its own
scratch: the host->device copy of shard i+1, the kernels of shard i and the device->host copy of shard i-1 overlap,
and so do the latency-bound recursions of different shards.  Results are written straight into caller-visible
[B,T,K] tensors at the shard's offset, so the outputs are exactly those of the un-sharded calls.

Outputs per pass (all fp32 except states):
    posterior, forward, backward  [B,T,K]   HMMPyTorch.forward_backward on per-frame max-normalised probabilities
                                            (hmm.py:66-130; BASELINE.md section 3 -- raw exp() underflows, SURVEY finding 3)
    log_delta [B,T,K], states int64 [B,T], score [B]
                                            MixtureGaussianHMMLayer._viterbi_decode on the raw log-emissions
                                            (mixture_gaussian.py:290-338)
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops
from .core import HMMPyTorch

OUT_NAMES = ("posterior", "forward", "backward", "log_delta", "states")


class _Slot:
    """One in-flight shard: a stream plus the scratch that must not be shared between concurrent shards."""

    def __init__(self, dev, bs: int, T: int, K: int, D: int, staged: bool):
        self.stream = torch.cuda.Stream(dev)
        self.aux = torch.cuda.Stream(dev)          # Viterbi runs beside forward-backward: they only share log b
        self.ev_emis, self.ev_vit = torch.cuda.Event(), torch.cuda.Event()
        self.logb = torch.empty(bs, T, K, device=dev)
        self.fb_ws = ops.fb_workspace(bs, T, K, dev)
        self.vit_ws = ops.viterbi_workspace(bs, T, K, dev)
        self.fused_ws = ops.fb_viterbi_workspace(bs, T, K, dev)
        self.x = torch.empty(bs, T, D, device=dev) if staged else None
        self.done = torch.cuda.Event()


class HMMInferenceEngine:
    def __init__(self, layer, batch: int, seq_len: int, shard: int = 64, n_streams: int = 4,
                 device: Optional[torch.device] = None, host_io: bool = False, fused: bool = True, pdl: bool = True):
        """layer: a MixtureGaussianHMMLayer (its parameters are packed once here; call refresh() after an update).
        batch/seq_len: the shape of one pass.  shard: utterances per in-flight shard.  host_io: also allocate the
        per-slot device staging for x so that run_host() can take pinned host tensors.
        fused: forward sweep, backward sweep and Viterbi of a shard in ONE launch (hmmb200_fb_viterbi_f32) instead of two
        kernels on two streams.  pdl: that launch may overlap the tail of the emission kernel (programmatic dependent launch)."""
        self.fused, self.pdl = bool(fused), bool(pdl)
        self.dev = ops.require_cuda(device if device is not None else (layer.means.device if layer.means.is_cuda else None))
        self.layer = layer
        self.K, self.C, self.D = layer.num_states, layer.num_components, layer.feature_dim
        self.B, self.T = int(batch), int(seq_len)
        self.shard = max(1, min(int(shard), self.B))
        self.n_shards = (self.B + self.shard - 1) // self.shard
        self.n_streams = max(1, min(int(n_streams), self.n_shards))
        self.slots = [_Slot(self.dev, self.shard, self.T, self.K, self.D, host_io) for _ in range(self.n_streams)]
        self.host_io = host_io
        self.refresh()
        n = (self.B, self.T, self.K)
        self.out: Dict[str, torch.Tensor] = {
            "posterior": torch.empty(n, device=self.dev), "forward": torch.empty(n, device=self.dev),
            "backward": torch.empty(n, device=self.dev), "log_delta": torch.empty(n, device=self.dev),
            "states": torch.empty(self.B, self.T, dtype=torch.int64, device=self.dev),
            "score": torch.empty(self.B, device=self.dev), "loglik": torch.empty(self.B, device=self.dev)}
        self._fork = torch.cuda.Event()
        # gmm_emission_tc (+ the fp32 kernel when unknown), then fb_viterbi + fb_combine (fused) or fb_sweep, fb_combine, viterbi
        self.kernels_per_shard = (1 if self.tc_known else 2) + (2 if self.fused else 3)

    def refresh(self):
        """Re-derive the kernel operands from the layer's parameters (O(K^2 + K*C*D), host side: SURVEY H4)."""
        layer, dev = self.layer, self.dev
        with torch.no_grad():
            P = layer.get_transition_matrix().detach().to(dev)
            self.hmm = HMMPyTorch(P, None, device=str(dev))
            self.trans, self.init = self.hmm._effective_probs(dev)                 # exp(log(P + 1e-8)), hmm.py:42,55
            self.log_trans = layer._safe_log(P).contiguous()                       # mixture_gaussian.py:357
            self.prior = torch.full((self.K,), -math.log(self.K), dtype=torch.float32, device=dev)   # :312
            self.packed = layer._packed()
            # setup-time only (one stream synchronisation): lets every pass launch the tcgen05 kernel alone
            self.tc_known = ops.gmm_pack_on_tensor_cores(self.packed, self.K, self.C, self.D)

    # ------------------------------------------------------------------------------------------------
    def _shard_kernels(self, slot: _Slot, x_sh: torch.Tensor, lo: int, hi: int):
        n = hi - lo
        logb = slot.logb[:n]
        o = self.out
        ops.gmm_emission(x_sh, self.packed, self.K, self.C, self.D, out=logb, tc_known=self.tc_known)
        if self.fused:
            ops.forward_backward_viterbi(
                logb, ops.EMIS_LOG_NORM_FLOOR, ops.EMIS_LOG, self.trans, self.init, self.log_trans, self.prior,
                want=("gamma", "fwd", "bwd"),
                out={"gamma": o["posterior"][lo:hi], "fwd": o["forward"][lo:hi], "bwd": o["backward"][lo:hi],
                     "loglik": o["loglik"][lo:hi], "states": o["states"][lo:hi], "delta": o["log_delta"][lo:hi],
                     "score": o["score"][lo:hi]},
                workspace=slot.fused_ws, pdl=self.pdl and self.tc_known)
            return
        slot.ev_emis.record(slot.stream)
        with torch.cuda.stream(slot.aux):
            slot.aux.wait_event(slot.ev_emis)
            ops.viterbi(logb, ops.EMIS_LOG, self.log_trans, self.prior,
                        out={"states": o["states"][lo:hi], "delta": o["log_delta"][lo:hi], "score": o["score"][lo:hi]},
                        workspace=slot.vit_ws)
            slot.ev_vit.record(slot.aux)
        ops.forward_backward(logb, ops.EMIS_LOG_NORM_FLOOR, self.trans, self.init, want=("gamma", "fwd", "bwd"),
                             out={"gamma": o["posterior"][lo:hi], "fwd": o["forward"][lo:hi], "bwd": o["backward"][lo:hi],
                                  "loglik": o["loglik"][lo:hi]}, workspace=slot.fb_ws)
        slot.stream.wait_event(slot.ev_vit)

    def _fan_out(self):
        main = torch.cuda.current_stream(self.dev)
        self._fork.record(main)
        for s in self.slots:
            s.stream.wait_event(self._fork)
        return main

    def _fan_in(self, main):
        for s in self.slots:
            s.done.record(s.stream)
            main.wait_event(s.done)

    def join(self):
        """Make the current stream wait for every shard enqueued so far."""
        self._fan_in(torch.cuda.current_stream(self.dev))

    def run_device(self, x: torch.Tensor, join: bool = True) -> Dict[str, torch.Tensor]:
        """x [B,T,D] resident on the device.  Enqueues one pass and returns the (device) output dict.  With join=True
        the caller's current stream waits for every shard, so ordinary stream semantics apply to the results; with
        join=False consecutive passes pipeline (shard i of the next pass queues behind shard i of this one on the same
        stream and scratch) and the caller calls join() before using the results."""
        assert x.shape == (self.B, self.T, self.D) and x.is_cuda
        main = self._fan_out()
        for i in range(self.n_shards):
            slot = self.slots[i % self.n_streams]
            lo, hi = i * self.shard, min(self.B, (i + 1) * self.shard)
            with torch.cuda.stream(slot.stream):
                self._shard_kernels(slot, x[lo:hi], lo, hi)
        if join:
            self._fan_in(main)
        return self.out

    def run_host(self, x_host: torch.Tensor, out_host: Dict[str, torch.Tensor], join: bool = True) -> None:
        """x_host [B,T,D] PINNED host memory in, every API-visible output to the pinned host tensors of `out_host`
        (keys among posterior, forward, backward, log_delta, states, score, loglik).  Asynchronous: returns once the
        work is enqueued; synchronise the current stream (or the device) before reading out_host."""
        assert self.host_io, "construct the engine with host_io=True"
        assert x_host.shape == (self.B, self.T, self.D) and x_host.is_pinned()
        main = self._fan_out()
        for i in range(self.n_shards):
            slot = self.slots[i % self.n_streams]
            lo, hi = i * self.shard, min(self.B, (i + 1) * self.shard)
            with torch.cuda.stream(slot.stream):
                xs = slot.x[:hi - lo]
                xs.copy_(x_host[lo:hi], non_blocking=True)
                self._shard_kernels(slot, xs, lo, hi)
                for name, dst in out_host.items():
                    dst[lo:hi].copy_(self.out[name][lo:hi], non_blocking=True)
        if join:
            self._fan_in(main)

    def capture_device(self, x: torch.Tensor, passes: int = 1) -> "torch.cuda.CUDAGraph":
        """Captures `passes` consecutive run_device(x) passes (all shards, all streams) into a CUDA graph: one launch per
        replay instead of 4-5 per shard and pass.  x and the output tensors are baked in by address; refill x in place
        and replay()."""
        self.run_device(x)                                   # warm up outside the capture (lazy inits, smem opt-ins)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(max(1, int(passes))):
                self.run_device(x, join=True)
        return g
