"""Race hunt for the large-K cluster sweeps: repeat Viterbi / forward-backward on fixed inputs and count runs whose
output differs bit-wise from the first run (the kernels are deterministic) and, for K = 64, from the reference's golden delta.
Prints the first differing positions.  Debug aid."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_hmm_b200 as hm

N = int(os.environ.get("N", 1000))
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "largek.npz"))
dev = "cuda"


def run_case(name, logb, logP, logp0, P, p0, golden_delta=None):
    first_d = first_g = None
    bad_d = bad_g = bad_gold = 0
    for i in range(N):
        r = hm.ops.viterbi(logb, hm.ops.EMIS_LOG, logP, logp0, want_delta=True)
        f = hm.ops.forward_backward(logb, hm.ops.EMIS_LOG, P, p0, want=("gamma",))
        d, gm = r["delta"], f["gamma"]
        if first_d is None:
            first_d, first_g = d.clone(), gm.clone()
        if not torch.equal(d, first_d):
            bad_d += 1
            if bad_d <= 2:
                idx = (d != first_d).nonzero()
                print(name, "run", i, "delta differs at", idx[:6].tolist(), "n =", idx.shape[0],
                      [(float(d[tuple(j)]), float(first_d[tuple(j)])) for j in idx[:3]])
        if not torch.equal(gm, first_g):
            bad_g += 1
            if bad_g <= 2:
                idx = (gm != first_g).nonzero()
                print(name, "run", i, "gamma differs at", idx[:6].tolist(), "n =", idx.shape[0])
        if golden_delta is not None and not torch.equal(d, golden_delta):
            bad_gold += 1
    torch.cuda.synchronize()
    print(f"{name}: runs={N} delta_nondeterministic={bad_d} gamma_nondeterministic={bad_g} golden_mismatch={bad_gold}")


tag = "k64"
log_obs = torch.log(torch.from_numpy(g[f"{tag}_obs"]) + 1e-8).to(dev)
P = torch.from_numpy(g[f"{tag}_P"]).to(dev) + 1e-8
p0 = torch.full((P.shape[0],), 1.0 / P.shape[0], device=dev)
run_case("k64 B=3 T=40", log_obs, torch.from_numpy(g[f"{tag}_log_P"]).to(dev), torch.from_numpy(g[f"{tag}_log_p0"]).to(dev), P, p0,
         torch.from_numpy(g[f"{tag}_log_delta"]).to(dev))
torch.manual_seed(1)
for K, B, T in ((512, 13, 64), (200, 7, 50), (96, 20, 33)):
    logb = torch.randn(B, T, K, device=dev) * 3 - 20
    Pm = torch.rand(K, K, device=dev) ** 3 + 0.01
    Pm = Pm / Pm.sum(1, keepdim=True)
    p0 = torch.full((K,), 1.0 / K, device=dev)
    run_case(f"K={K} B={B} T={T}", logb, torch.log(Pm), torch.log(p0), Pm, p0)
