"""StreamingHMMProcessor -- API shell of pytorch_hmm/streaming.py on the sm_100a kernels.

Kept from the reference: constructor arguments and parameters (`transition_logits`, `emission_net`), the bounded frame
buffer and the buffering / look-ahead rules of process_chunk (streaming.py:183-265), `StreamingResult`, `flush_buffer`,
`reset_streaming_state`, `get_performance_stats`.  The per-frame Python decode loop (streaming.py:292-308) runs as one
launch of `hmmb200_greedy_decode_f32`, continuing from the previous chunk's last state.

New (north star): `forward_chunk()` runs the forward recursion over a chunk with the filtered state vector carried between
calls (`hmmb200_forward_chunk_f32`); chunked calls equal one unchunked pass.

The asynchronous front end (streaming.py:123-181: start / stop, add_audio_chunk_async, get_result_async) is one worker thread
feeding process_chunk from a bounded queue, chunks decoded in arrival order.

Out of scope: beam search, the adaptive latency controller.  With
`use_beam_search=True` (the reference's default) decoding uses the greedy kernel and says so in the result metadata.
"""
from __future__ import annotations

import queue
import threading
import time
import warnings
from collections import deque
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


@dataclass
class StreamingResult:
    decoded_states: Optional[torch.Tensor]
    confidence: float
    processing_time_ms: float
    buffer_size: int
    chunk_id: int
    status: str
    metadata: Dict[str, Any]


class StreamingHMMProcessor(nn.Module):
    def __init__(self, num_states: int, feature_dim: int, chunk_size: int = 160, overlap_size: int = 80,
                 lookahead_frames: int = 5, max_delay_frames: int = 50, use_beam_search: bool = True, beam_width: int = 8,
                 buffer_size: int = 1000):
        super().__init__()
        self.num_states, self.feature_dim = num_states, feature_dim
        self.chunk_size, self.overlap_size = chunk_size, overlap_size
        self.lookahead_frames, self.max_delay_frames = lookahead_frames, max_delay_frames
        self.use_beam_search, self.beam_width, self.buffer_size = use_beam_search, beam_width, buffer_size
        self.transition_logits = nn.Parameter(torch.randn(num_states, num_states) * 0.1)
        self.emission_net = nn.Sequential(nn.Linear(feature_dim, 128), nn.ReLU(), nn.Dropout(0.1),
                                          nn.Linear(128, num_states), nn.LogSoftmax(dim=-1))
        self.processing_times = deque(maxlen=1000)
        self.processing_queue: "queue.Queue" = queue.Queue(maxsize=buffer_size)
        self.result_queue: "queue.Queue" = queue.Queue(maxsize=buffer_size)
        self.is_processing, self.processing_thread = False, None
        self.reset_streaming_state()

    def reset_streaming_state(self):
        self.feature_buffer = deque(maxlen=self.max_delay_frames + self.lookahead_frames)
        self.viterbi_states, self.viterbi_scores = [], []
        self.last_output_frame = -1
        self.chunk_counter = 0
        self.total_frames_processed = 0
        self._greedy_state = None          # int32 [1] on the compute device: last decoded state (-1 = none yet)
        self._fwd_state = None             # carried forward-recursion state

    def get_transition_matrix(self) -> torch.Tensor:
        return F.softmax(self.transition_logits, dim=-1)

    def _cuda(self) -> torch.device:
        return ops.require_cuda(self.transition_logits.device if self.transition_logits.is_cuda else None)

    # -- asynchronous front end (streaming.py:123-181) ----------------------------------------------------------
    def start_async_processing(self):
        if self.is_processing:
            return
        self.is_processing = True
        self.processing_thread = threading.Thread(target=self._worker, name="hmm-stream", daemon=True)
        self.processing_thread.start()

    def stop_async_processing(self):
        self.is_processing = False
        t, self.processing_thread = self.processing_thread, None
        if t is not None:
            t.join(timeout=1.0)

    def _worker(self):
        while self.is_processing:
            try:
                chunk = self.processing_queue.get(timeout=0.1)
            except queue.Empty:
                continue
            try:
                with torch.no_grad():
                    result = self.process_chunk(chunk)
                if not self.result_queue.full():                 # a full result queue drops the newest result, as the reference does
                    self.result_queue.put(result)
            except Exception as exc:                            # the worker must survive a bad chunk
                warnings.warn(f"Error in async processing: {exc}")
            finally:
                self.processing_queue.task_done()

    def add_audio_chunk_async(self, audio_chunk: torch.Tensor) -> bool:
        """False when the input queue is full (the chunk is not taken)."""
        try:
            self.processing_queue.put_nowait(audio_chunk)
            return True
        except queue.Full:
            return False

    def get_result_async(self) -> Optional[StreamingResult]:
        try:
            return self.result_queue.get_nowait()
        except queue.Empty:
            return None

    # -- reference API -------------------------------------------------------------------------------------
    def process_chunk(self, audio_chunk: torch.Tensor) -> StreamingResult:
        t0 = time.time()
        for frame in audio_chunk:
            self.feature_buffer.append(frame)
        available = len(self.feature_buffer)
        required = self.chunk_size + self.lookahead_frames
        if available < required:
            return StreamingResult(None, 0.0, (time.time() - t0) * 1000, available, self.chunk_counter, "buffering",
                                   {"frames_needed": required - available})
        start = max(0, self.last_output_frame + 1)
        end = available - self.lookahead_frames
        if end <= start:
            return StreamingResult(None, 0.0, (time.time() - t0) * 1000, available, self.chunk_counter,
                                   "waiting_for_lookahead", {})
        features = torch.stack(list(self.feature_buffer)[start:end])
        states, confidence = self._greedy_decode(features)
        self.last_output_frame = end - 1
        self.total_frames_processed += len(features)
        dt = (time.time() - t0) * 1000
        self.processing_times.append(dt)
        self.chunk_counter += 1
        frame_ms = len(features) * 1000 / 100
        return StreamingResult(states, confidence.mean().item(), dt, available, self.chunk_counter, "decoded",
                               {"frames_processed": len(features), "real_time_factor": frame_ms / dt if dt > 0 else float("inf"),
                                "buffer_utilization": available / self.feature_buffer.maxlen,
                                "decoder": "greedy (sm_100a kernel)" + ("; beam search is out of scope" if self.use_beam_search else "")})

    def _greedy_decode(self, features: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(T, D) features -> (states (T,), confidence (T,) = exp(score))  (streaming.py:267-320)."""
        dev = self._cuda()
        with torch.no_grad():
            logb = self.emission_net(features.to(dev))
            log_trans = torch.log(self.get_transition_matrix() + 1e-8)
            if self._greedy_state is None:
                self._greedy_state = torch.full((1,), -1, dtype=torch.int32, device=dev)
            states, scores = ops.greedy_decode(logb.unsqueeze(0), log_trans, self._greedy_state)
        states, scores = states[0], scores[0]
        self.viterbi_states.extend(states.tolist())
        self.viterbi_scores.extend(scores.tolist())
        if len(self.viterbi_states) > self.max_delay_frames:
            cut = len(self.viterbi_states) - self.max_delay_frames
            self.viterbi_states, self.viterbi_scores = self.viterbi_states[cut:], self.viterbi_scores[cut:]
        out_dev = features.device
        return states.to(out_dev), torch.exp(scores).to(out_dev)

    def forward_chunk(self, features: torch.Tensor) -> Dict[str, torch.Tensor]:
        """(T, D) or (B, T, D) features -> filtered posteriors p(state_t | o_1..t) and the running log-likelihood, with the
        forward vector carried across calls (uniform prior, transitions = softmax(transition_logits))."""
        dev = self._cuda()
        x = features if features.dim() == 3 else features.unsqueeze(0)
        with torch.no_grad():
            logb = self.emission_net(x.to(dev))
            trans = self.get_transition_matrix()
            init = torch.full((self.num_states,), 1.0 / self.num_states, device=dev)
            if self._fwd_state is None or self._fwd_state["alpha"].shape[0] != x.shape[0]:
                self._fwd_state = ops.new_forward_state(x.shape[0], self.num_states, dev)
            filt = ops.forward_chunk(logb, ops.EMIS_LOG, trans, init, self._fwd_state)
        ll = self._fwd_state["loglik"].clone()
        if features.dim() == 2:
            filt, ll = filt[0], ll[0]
        return {"filtered": filt.to(features.device), "log_likelihood": ll.to(features.device)}

    def flush_buffer(self) -> Optional[StreamingResult]:
        if not self.feature_buffer:
            return None
        start = max(0, self.last_output_frame + 1)
        frames = list(self.feature_buffer)[start:]
        if not frames:
            return None
        t0 = time.time()
        states, confidence = self._greedy_decode(torch.stack(frames))
        self.last_output_frame = len(self.feature_buffer) - 1
        self.chunk_counter += 1                                           # (streaming.py:397)
        return StreamingResult(states, confidence.mean().item(), (time.time() - t0) * 1000, len(self.feature_buffer),
                               self.chunk_counter, "flushed", {"final_chunk": True, "frames_processed": len(frames)})

    def optimize_for_latency(self, target_latency_ms: float = 50.0):
        """Parameter adjustment of the reference (streaming.py:444-488): trade beam width / chunk size against the measured latency.
        Host-side bookkeeping only; decoding here is the greedy kernel whatever `use_beam_search` says."""
        stats = self.get_performance_stats()
        if "avg_processing_time_ms" not in stats:
            warnings.warn("No performance data available for optimization")
            return
        cur = stats["avg_processing_time_ms"]
        if cur > target_latency_ms:
            if self.use_beam_search and self.beam_width > 2:
                self.beam_width = max(2, self.beam_width - 1)
            elif self.use_beam_search:
                self.use_beam_search = False
            elif self.chunk_size > 80:
                self.chunk_size = max(80, int(self.chunk_size * 0.8))
        elif cur < target_latency_ms * 0.5:
            if not self.use_beam_search:
                self.use_beam_search = True
                self.beam_width = 2
            elif self.beam_width < 8:
                self.beam_width += 1

    def get_performance_stats(self) -> Dict[str, float]:
        if not self.processing_times:
            return {}
        t = torch.tensor(list(self.processing_times))
        return {"avg_processing_time_ms": t.mean().item(), "max_processing_time_ms": t.max().item(),
                "min_processing_time_ms": t.min().item(), "total_chunks": self.chunk_counter,
                "total_frames": self.total_frames_processed}
