"""BASELINE config 1: HMMPyTorch left-to-right K=10, D=80 diag-Gaussian, forward_backward + viterbi_decode, B=32, T=1000.
GPU times (CUDA events, inputs resident) and, with --cpu, the op-for-op CPU port of the reference on the same inputs."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
import pytorch_hmm_b200 as hm
torch.set_grad_enabled(False)

K, D, B, T = 10, 80, 32, 1000
ap = argparse.ArgumentParser(); ap.add_argument("--cpu", action="store_true"); a = ap.parse_args()
torch.manual_seed(1001)
g = hm.GaussianHMMLayer(K, D, normalize_emissions=True).cuda().eval()
P = hm.create_left_to_right_matrix(K, 0.7)
path = (torch.arange(T) * K // T).expand(B, T)
x = (g.means.detach().cpu()[path] + torch.randn(B, T, D)).cuda()
hmm = hm.HMMPyTorch(P, None, device="cuda")
trans, init = hmm._effective_probs(torch.device("cuda", 0))
logP, logp0 = hmm.log_P.cuda(), hmm.log_p0.cuda()


def step():
    logb = g._compute_gaussian_log_probs(x)
    hm.ops.forward_backward(logb, hm.ops.EMIS_LOG_NORM_FLOOR, trans, init, want=("gamma", "fwd", "bwd"))
    hm.ops.viterbi(logb, hm.ops.EMIS_LOG_NORM_FLOOR, logP, logp0)


for _ in range(5):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(50):
    step()
e.record(); e.synchronize()
ms = s.elapsed_time(e) / 50
res = {"config": "1: K=10 left-to-right, D=80 diag-Gaussian, B=32, T=1000", "gpu_ms_per_step": ms, "gpu_frames_per_s": B * T / (ms * 1e-3)}
if a.cpu:
    from oracle import ref_port
    logb = g._compute_gaussian_log_probs(x).cpu()
    obs = torch.exp(logb - logb.max(-1, keepdim=True)[0])
    lP, lp0 = ref_port.prepare_hmm(P, None)
    t0 = time.perf_counter()
    for _ in range(3):
        ref_port.forward_backward(obs, lP, lp0); ref_port.viterbi_decode(obs, lP, lp0)
    dt = (time.perf_counter() - t0) / 3
    res.update({"cpu_port_s_per_step": dt, "cpu_port_frames_per_s": B * T / dt, "cpu_threads": torch.get_num_threads()})
print(json.dumps(res))
