"""Runs the full-covariance emission kernel a few times at K=12, C=4, D=80 on 512 000 frames (ncu / timing target)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_hmm_b200 as hm
dev = torch.device("cuda", 0)
torch.manual_seed(7)
K, C, D, N = 12, 4, 80, 256 * 2000
means = torch.randn(K, C, D, device=dev) * 0.2
A = torch.randn(K, C, D, D, device=dev) * 0.05
chol = torch.linalg.cholesky(A @ A.transpose(-1, -2) + torch.eye(D, device=dev))
packed = hm.ops.gmm_pack_full(means, chol, torch.log_softmax(torch.randn(K, C, device=dev), -1))
x = torch.randn(N, D, device=dev)
ob = torch.empty(N, K, device=dev)
for _ in range(3):
    hm.ops.gmm_emission_full(x, packed, K, C, D, out=ob)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    hm.ops.gmm_emission_full(x, packed, K, C, D, out=ob)
e.record(); e.synchronize()
print("ms", s.elapsed_time(e) / 5)
