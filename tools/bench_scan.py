"""Time-parallel scan vs sequential sweeps for forward-backward at small batch / long T (K=12).  CUDA-event times."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_hmm_b200 as hm

K = 12
dev = torch.device("cuda", 0)
torch.manual_seed(1)
P = torch.softmax(0.5 * torch.randn(K, K), -1).to(dev) + 1e-8
p0 = torch.full((K,), 1.0 / K, device=dev) + 1e-8


def ms(fn, it=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / it


rows = []
for B, T in ((1, 4096), (1, 16384), (1, 100000), (1, 1000000), (4, 100000), (8, 16384), (16, 16384), (32, 16384), (64, 16384)):
    x = torch.randn(B, T, K, device=dev) * 5 - 60
    out = {k: torch.empty(B, T, K, device=dev) for k in ("gamma", "fwd", "bwd")}
    out["loglik"] = torch.empty(B, device=dev)
    t_scan = ms(lambda: hm.ops.forward_backward(x, hm.ops.EMIS_LOG_NORM_FLOOR, P, p0, out=out, method="scan"))
    t_sweep = ms(lambda: hm.ops.forward_backward(x, hm.ops.EMIS_LOG_NORM_FLOOR, P, p0, out=out, method="sweep"))
    rows.append({"B": B, "T": T, "scan_ms": round(t_scan, 4), "sweep_ms": round(t_sweep, 4), "speedup": round(t_sweep / t_scan, 2),
                 "scan_frames_per_s": B * T / (t_scan * 1e-3)})
    print(json.dumps(rows[-1]), flush=True)
