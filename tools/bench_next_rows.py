"""Measurements for the SURVEY section 8(f) rows (the callers either side of the hot path): time-varying-transition recursions (NeuralHMM
form), full-covariance emission, CTC trellises, DTW -- CUDA-event times with resident inputs, algorithmic bytes / flops against the
measured peaks, and the oracle port of the reference's algorithm on the host (a bounded sample, single thread) beside them."""
import json, math, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_hmm_b200 as hm
from oracle import alignment_port as ap

dev = torch.device("cuda", 0)
torch.manual_seed(7)
HBM = 6452.8e9
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:
    pass


def ms(fn, it=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / it


out = {}
# ---- (f2) NeuralHMM recursion: per-frame [K,K] transitions, K = 12, B = 256, T = 2000 (the headline shape) -------------------------
B, T, K = 256, 2000, 12
le = torch.log_softmax(torch.randn(B, T, K, device=dev), -1)
tp = torch.softmax(torch.randn(B, T, K, K, device=dev), -1)
p0 = torch.full((K,), 1.0 / K, device=dev)
ltp, lp0 = torch.log(tp), torch.log(p0)
t_fb = ms(lambda: hm.ops.tv_forward_backward(le, tp, p0))
t_v = ms(lambda: hm.ops.tv_viterbi(le, ltp, lp0))
bytes_fb = B * T * (K * K * 4 + K * 4 + 3 * K * 4)           # transitions + emissions in, three tables out
bytes_v = B * T * (K * K * 4 + K * 4 + K * 4 + 8)
out["neural_hmm_recursion"] = {"shape": f"B={B} T={T} K={K}, transitions [B,T,K,K]", "forward_backward_ms": t_fb, "viterbi_ms": t_v,
                               "frames_per_s_fb_plus_viterbi": B * T / ((t_fb + t_v) * 1e-3),
                               "hbm_frac_forward_backward": bytes_fb / (t_fb * 1e-3) / HBM, "hbm_frac_viterbi": bytes_v / (t_v * 1e-3) / HBM}

# ---- (f3) full-covariance emission: K = 12, C = 4, D = 80, 512 000 frames ----------------------------------------------------------
K, C, D, N = 12, 4, 80, 256 * 2000
means = torch.randn(K, C, D, device=dev) * 0.2
A = torch.randn(K, C, D, D, device=dev) * 0.05
chol = torch.linalg.cholesky(A @ A.transpose(-1, -2) + torch.eye(D, device=dev))
packed = hm.ops.gmm_pack_full(means, chol, torch.log_softmax(torch.randn(K, C, device=dev), -1))
x = torch.randn(N, D, device=dev)
ob = torch.empty(N, K, device=dev)
t_full = ms(lambda: hm.ops.gmm_emission_full(x, packed, K, C, D, out=ob), it=5, warm=2)
flops = 2.0 * N * K * C * (D * (D + 1) / 2 + D)
out["full_covariance_emission"] = {"shape": f"N={N} frames, K={K}, C={C}, D={D}", "ms": t_full, "frames_per_s": N / (t_full * 1e-3),
                                   "tflops_fp32_algorithmic": flops / (t_full * 1e-3) / 1e12,
                                   "hbm_frac": N * (D + K) * 4 / (t_full * 1e-3) / HBM}

# ---- (f4) CTC trellises: 64 utterances, T = 1000 frames, targets of 100 labels, 64 classes ------------------------------------------
Tn, Bc, Cc, L = 1000, 64, 64, 100
lp = torch.log_softmax(torch.randn(Tn, Bc, Cc, device=dev), -1)
tg = torch.randint(1, Cc, (Bc, L), device=dev)
il = torch.full((Bc,), Tn, dtype=torch.int64, device=dev); tl = torch.full((Bc,), L, dtype=torch.int64, device=dev)
t_cf = ms(lambda: hm.ops.ctc_trellis(0, lp, tg, il, tl, 0))
t_cb = ms(lambda: hm.ops.ctc_trellis(1, lp, tg, il, tl, 0))
t0 = time.perf_counter()
ap.ctc_forward(lp[:, :2].cpu().numpy(), tg[:2].cpu().numpy(), il[:2].cpu().numpy(), tl[:2].cpu().numpy(), 0)
cpu_ctc = (time.perf_counter() - t0) / 2
out["ctc"] = {"shape": f"B={Bc} T={Tn} L={L} classes={Cc}", "forward_ms": t_cf, "backward_ms": t_cb,
              "cells_per_s_forward": Bc * Tn * (2 * L + 1) / (t_cf * 1e-3),
              "cpu_port_ms_per_utterance_forward": cpu_ctc * 1e3, "cpu_sample": "2 utterances, numpy port of alignment/ctc.py, 1 thread",
              "speedup_vs_cpu_port_per_utterance": cpu_ctc * 1e3 / (t_cf / Bc)}

# ---- (f4) DTW: 64 pairs of 1000 x 1000 --------------------------------------------------------------------------------------------
P, Nn, Mm = 64, 1000, 1000
dist = torch.rand(P, Nn, Mm, device=dev)
t_d = ms(lambda: hm.ops.dtw(dist, 0), it=5, warm=2)
d1 = dist[0, :300, :300].cpu().numpy()
t0 = time.perf_counter(); ap.dtw(d1, "symmetric"); cpu_dtw = time.perf_counter() - t0
out["dtw"] = {"shape": f"P={P} pairs of {Nn} x {Mm}", "ms": t_d, "cells_per_s": P * Nn * Mm / (t_d * 1e-3),
              "cpu_port_cells_per_s": 300 * 300 / cpu_dtw, "cpu_sample": "one 300 x 300 pair, python port of alignment/dtw.py, 1 thread"}
print(json.dumps(out))
