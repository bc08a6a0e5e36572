import sys, torch
sys.path.insert(0, "/root/repo")
import pytorch_hmm_b200 as hm
dev = torch.device("cuda", 0)
B, T, K = 256, 2000, 12
le = torch.log_softmax(torch.randn(B, T, K, device=dev), -1)
tp = torch.softmax(torch.randn(B, T, K, K, device=dev), -1)
p0 = torch.full((K,), 1.0 / K, device=dev)
for _ in range(3):
    hm.ops.tv_forward_backward(le, tp, p0)
    hm.ops.tv_viterbi(le, torch.log(tp), torch.log(p0))
torch.cuda.synchronize()
