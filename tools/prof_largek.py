"""Per-kernel times (torch profiler / CUPTI) of the large-K calls at BASELINE config 5 (K=512, B=64, T=4000)."""
import os, sys, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_hmm_b200 as hm
K, B, T = int(os.environ.get("K", 512)), int(os.environ.get("B", 64)), int(os.environ.get("T", 4000))
dev = torch.device("cuda", 0)
torch.manual_seed(5001)
hmm = hm.HMMPyTorch(hm.create_transition_matrix(K, "ergodic"), None, device="cuda")
obs = torch.softmax(torch.randn(B, T, K, device=dev), dim=-1)
trans, init = hmm._effective_probs(dev)
logP, logp0 = hmm.log_P.to(dev), hmm.log_p0.to(dev)
calls = {
    "forward_backward": lambda: hm.ops.forward_backward(obs, hm.ops.EMIS_PROB_FLOOR, trans, init),
    "viterbi": lambda: hm.ops.viterbi(obs, hm.ops.EMIS_PROB_FLOOR, logP, logp0),
    "forward_backward_viterbi": lambda: hm.ops.forward_backward_viterbi(obs, hm.ops.EMIS_PROB_FLOOR, hm.ops.EMIS_PROB_FLOOR, trans, init, logP, logp0),
}
for name, fn in calls.items():
    fn(); fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    print(f"== {name}")
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
        print(f"   {ev.device_time_total / 1e3:9.3f} ms  x{ev.count:<3d} {ev.key[:110]}")
