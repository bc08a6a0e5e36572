import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
print(bench.extra_config1(dev))
import pytorch_hmm_b200 as hm
K, D, B, T = 10, 80, 32, 1000
g = hm.GaussianHMMLayer(K, D, normalize_emissions=True).to(dev).eval()
x = torch.randn(B, T, D, device=dev)
hmm = hm.HMMPyTorch(hm.create_left_to_right_matrix(K, 0.7), None, device="cuda")
trans, init = hmm._effective_probs(dev); logP, logp0 = hmm.log_P.to(dev), hmm.log_p0.to(dev)
def step():
    logb = g._compute_gaussian_log_probs(x)
    hm.ops.forward_backward_viterbi(logb, hm.ops.EMIS_LOG_NORM_FLOOR, hm.ops.EMIS_LOG_NORM_FLOOR, trans, init, logP, logp0)
step(); step(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:10]:
    print(f"   {ev.device_time_total / 1e3:9.4f} ms  x{ev.count:<3d} {ev.key[:100]}")
