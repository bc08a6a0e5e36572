"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck), one tool per run."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_hmm_b200 as hm
torch.set_grad_enabled(False)

dev = torch.device("cuda", 0)
torch.manual_seed(0)
which = sys.argv[1:] or ["smallk", "fused", "largek", "scan", "hsmm", "emission", "bw", "alignment"]
if "emission" in which or "smallk" in which:
    m = hm.MixtureGaussianHMMLayer(12, 80, num_components=4).to(dev).eval()
    x = torch.randn(3, 200, 80, device=dev)
    logb = m.get_observation_log_probs(x)
    with torch.no_grad():
        m(x, return_log_probs=True)
    P = m.get_transition_matrix().detach()
    h = hm.HMMPyTorch(P, None, device="cuda")
    tr, ini = h._effective_probs(dev)
    hm.ops.forward_backward(logb, hm.ops.EMIS_LOG_NORM_FLOOR, tr, ini, want=("gamma", "fwd", "bwd"), method="sweep")
if "fused" in which:
    K = 12
    P = torch.softmax(torch.randn(K, K), -1).to(dev) + 1e-8
    p0 = torch.full((K,), 1.0 / K, device=dev)
    lb = torch.randn(5, 300, K, device=dev) - 20
    hm.ops.forward_backward_viterbi(lb, hm.ops.EMIS_LOG_NORM_FLOOR, hm.ops.EMIS_LOG, P, p0, torch.log(P), torch.log(p0))
if "largek" in which:
    K = 96
    h = hm.HMMPyTorch(hm.create_transition_matrix(K, "ergodic"), None, device="cuda")
    obs = torch.softmax(torch.randn(5, 24, K, device=dev), -1)
    h.forward_backward(obs); h.viterbi_decode(obs)
if "scan" in which:
    K = 12
    P = torch.softmax(torch.randn(K, K), -1).to(dev) + 1e-8
    p0 = torch.full((K,), 1.0 / K, device=dev)
    hm.ops.forward_backward(torch.randn(2, 700, K, device=dev) - 30, hm.ops.EMIS_LOG, P, p0, method="scan")
if "hsmm" in which:
    s = hm.HSMMLayer(5, 12, max_duration=7).to(dev).eval()
    x = torch.randn(2, 40, 12, device=dev)
    s(x); s.forward_backward(x)
if "hsmm" in which:
    s2 = hm.HSMMLayer(10, 16, max_duration=20).to(dev).eval()       # the specialised K = 10, Dmax = 20 kernels
    x2 = torch.randn(2, 90, 16, device=dev)
    s2(x2); s2.forward_backward(x2)
if "alignment" in which:
    import pytorch_hmm_b200.alignment as al
    lp = torch.log_softmax(torch.randn(30, 3, 9, device=dev), -1)
    tg = torch.randint(1, 9, (3, 6), device=dev)
    il, tl = torch.tensor([30, 22, 30], device=dev), torch.tensor([6, 4, 0], device=dev)
    al.ctc_forward_algorithm(lp, tg, il, tl); al.ctc_backward_algorithm(lp, tg, il, tl)
    al.compute_dtw_path(torch.rand(40, 55, device=dev), "rabiner_juang")
if "bw" in which:
    from pytorch_hmm_b200.baum_welch import BaumWelch
    m = hm.MixtureGaussianHMMLayer(4, 8, num_components=2).to(dev)
    bw = BaumWelch.from_layer(m)
    bw.e_step(torch.randn(3, 50, 8, device=dev)); bw.m_step()
torch.cuda.synchronize()
print("sanitize_small ok:", which)
