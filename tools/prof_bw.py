"""Per-kernel times (torch profiler / CUPTI) of one Baum-Welch E-step batch at the headline shape (K=12, C=4, D=80, B=256, T=2000)."""
import os, sys, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pytorch_hmm_b200 as hm
from pytorch_hmm_b200 import baum_welch as bwm
dev = torch.device("cuda", 0)
model = bench.make_model()
x = bench.make_frames(model, bench.BATCH, bench.SEQ, 2001).to(dev)
K, C, D = bench.K_STATES, bench.N_MIX, bench.FEAT
start = bwm.GMMHMMParams(torch.softmax(model["transition_logits"], -1), torch.full((K,), 1.0 / K),
                         torch.softmax(model["mixture_weights_logits"], -1), model["means"].clone(), torch.ones(K, C, D))
bw = bwm.BaumWelch(start, device=dev)
for _ in range(3):
    bw.e_step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        bw.e_step(x)
    torch.cuda.synchronize()
tot = 0.0
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]:
    print(f"   {ev.device_time_total / 4e3:9.4f} ms/batch  x{ev.count // 4:<3d} {ev.key[:120]}")
    tot += ev.device_time_total / 4e3
print("sum of kernel times per batch (ms):", round(tot, 4))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    bw.e_step(x)
e.record(); e.synchronize()
print("wall per batch (ms):", round(s.elapsed_time(e) / 20, 4))
