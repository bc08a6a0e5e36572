"""Diagnostic: HSMM Viterbi kernel vs the C oracle at growing T (which sequences / frames / scores differ)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_hmm_b200 as hm
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
