"""BASELINE config 4: HSMM K=10, max_duration=20, D=80, B=128, T=2000 -- emission + duration-augmented forward-backward + Viterbi."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_hmm_b200 as hm
torch.set_grad_enabled(False)

K, Dm, D, B, T = 10, 20, 80, 128, 2000
torch.manual_seed(4001)
m = hm.HSMMLayer(K, D, duration_distribution="gamma", max_duration=Dm).cuda()
with torch.no_grad():
    m.observation_means.mul_(10.0)
x = torch.randn(B, T, D, device="cuda") + m.observation_means.detach()[torch.randint(0, K, (B, T), device="cuda")]


def ms(fn, it=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / it


res = {"config": "HSMM K=10 Dmax=20 D=80 B=128 T=2000",
       "emission_ms": ms(lambda: m.get_observation_log_probs(x)),
       "forward_backward_ms": ms(lambda: m.forward_backward(x)),
       "viterbi_ms": ms(lambda: m(x))}
logb = m.get_observation_log_probs(x)
log_dur, log_trans = m._tables(x.device)
res["kernel_only"] = {"hsmm_viterbi_ms": ms(lambda: hm.ops.hsmm_viterbi(logb, log_dur, log_trans, sum_order=0)),
                      "hsmm_forward_backward_ms": ms(lambda: hm.ops.hsmm_forward_backward(logb, log_dur, log_trans))}
res["frames_per_s_fb_plus_viterbi"] = B * T / ((res["forward_backward_ms"] + res["viterbi_ms"]) * 1e-3)
print(json.dumps(res))
