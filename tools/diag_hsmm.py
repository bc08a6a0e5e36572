"""Diagnostic: HSMM Viterbi kernel vs the C oracle at growing T (which sequences / frames / scores differ)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_hmm_b200 as hm
torch.set_grad_enabled(False)
from oracle import c_oracle

K, D, Dm = 10, 80, 20
torch.manual_seed(4001)
m = hm.HSMMLayer(K, D, duration_distribution="gamma", max_duration=Dm).cuda().eval()
for B, T in ((4, 300), (4, 600), (4, 1000), (12, 2000)):
    seg = torch.randint(0, K, (B, T // 10 + 1)).repeat_interleave(10, 1)[:, :T]
    x = (m.observation_means.detach().cpu()[seg] + torch.randn(B, T, D)).cuda()
    logb = m.get_observation_log_probs(x)
    states, scores = m._viterbi_from_log_probs(logb)
    log_dur, log_trans = m._tables(torch.device("cuda", 0))
    st, sc = c_oracle.hsmm_viterbi_f32(logb.cpu().numpy(), log_dur.cpu().numpy(), log_trans.cpu().numpy())
    s = states.cpu().numpy()
    bad = [(b, int((s[b] != st[b]).sum()), int(np.argmax(s[b] != st[b])), float(scores[b]), float(sc[b])) for b in range(B) if not np.array_equal(s[b], st[b])]
    print(f"B={B} T={T}: score bit-equal {np.array_equal(scores.cpu().numpy(), sc)}; mismatching sequences (b, n_frames, first, gpu_score, oracle_score): {bad}")
    for b, n, first, *_ in bad[:2]:
        lo, hi = max(0, first - 3), min(T, first + 25)
        print("  gpu   ", s[b, lo:hi].tolist())
        print("  oracle", st[b, lo:hi].tolist())

# the exact data of tests/test_gpu_fullsize.py::test_config4_full_shape_hsmm
K, D, Dm, B, T = 10, 80, 20, 128, 2000
torch.manual_seed(4001)
m = hm.HSMMLayer(K, D, duration_distribution="gamma", max_duration=Dm).cuda().eval()
seg = torch.randint(0, K, (B, T // 10 + 1)).repeat_interleave(10, 1)[:, :T]
x = (m.observation_means.detach().cpu()[seg] + torch.randn(B, T, D)).cuda()
logb = m.get_observation_log_probs(x)
states, scores = m._viterbi_from_log_probs(logb)
log_dur, log_trans = m._tables(torch.device("cuda", 0))
pick = list(range(0, B, 11))
st, sc = c_oracle.hsmm_viterbi_f32(logb[pick].cpu().numpy(), log_dur.cpu().numpy(), log_trans.cpu().numpy())
s = states[pick].cpu().numpy(); g = scores[pick].cpu().numpy()
for i, b in enumerate(pick):
    if not np.array_equal(s[i], st[i]) or g[i] != sc[i]:
        d = np.nonzero(s[i] != st[i])[0]
        print(f"seq {b}: score gpu {g[i]!r} oracle {sc[i]!r} equal {g[i] == sc[i]}; {d.size} frames differ, first {d[:1]}, last {d[-1:]}")
        lo = max(0, int(d[0]) - 3); hi = min(T, int(d[-1]) + 4)
        if hi - lo < 80:
            print("  gpu   ", s[i, lo:hi].tolist()); print("  oracle", st[i, lo:hi].tolist())
# alone (B = the picked sequences only): does the batch size matter?
st2, sc2 = m._viterbi_from_log_probs(logb[pick].contiguous())
print("picked sequences run as their own batch: states equal oracle", np.array_equal(st2.cpu().numpy(), st), "scores equal", np.array_equal(sc2.cpu().numpy(), sc))
