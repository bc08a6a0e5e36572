"""Parity report at the headline shape (SURVEY finding 9 asks for three numbers on gamma: engine vs float64, an fp32 log-space
recursion like the reference's vs float64, engine vs that reference-like recursion), plus log-likelihood and Viterbi
bit-exactness counts.  Uses oracle/ (checker) on the GPU box; prints JSON."""
import json, math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
import pytorch_hmm_b200 as hm
from oracle import c_oracle, ref_port
import bench
torch.set_grad_enabled(False)

K, C, D, B, T = 12, 4, 80, 16, 2000
model = bench.make_model()
x = bench.make_frames(model, B, T, 2001)
layer = hm.MixtureGaussianHMMLayer(K, D, num_components=C).cuda().eval()
layer.load_state_dict({k: v.cuda() for k, v in model.items()})
logb = layer.get_observation_log_probs(x.cuda())
# emission vs float64
logw = ref_port.safe_log(torch.softmax(model["mixture_weights_logits"], -1)).numpy()
ref_logb = c_oracle.gmm_emission_f64(x.numpy(), model["means"].numpy(), model["log_vars"].numpy(), 1.0, logw)
emis_rel = float(np.max(np.abs(logb.cpu().numpy() - ref_logb) / np.abs(ref_logb)))
# forward-backward: HMMPyTorch semantics on per-frame max-normalised probabilities
P = layer.get_transition_matrix().detach()
hmm = hm.HMMPyTorch(P, None, device="cuda")
trans, init = hmm._effective_probs(torch.device("cuda", 0))
r = hm.ops.forward_backward(logb, hm.ops.EMIS_LOG_NORM_FLOOR, trans, init, want=("gamma",))
lb = logb.cpu()
obs = torch.exp(lb - lb.max(-1, keepdim=True)[0])
log_obs32 = torch.log(obs + 1e-8).numpy()
logP32, logp032 = hmm.log_P.cpu().numpy(), hmm.log_p0.cpu().numpy()
_, _, gam64, ll64 = c_oracle.forward_backward_f64(log_obs32.astype(np.float64), logP32.astype(np.float64), logp032.astype(np.float64))
_, _, gam32 = c_oracle.forward_backward_f32(log_obs32, logP32, logp032)
ours = r["gamma"].cpu().numpy()
mask = gam64 > 1e-6
rel = lambda a, b: float(np.max(np.abs(a - b)[mask] / b[mask]))
# Viterbi: mixture semantics on the raw log-emissions
log_trans = layer._safe_log(P).cpu().numpy()
prior = np.full((K,), -math.log(K), np.float32)
st, dl, psi, sc = c_oracle.viterbi_f32(lb.numpy(), log_trans, prior)
v = hm.ops.viterbi(logb, hm.ops.EMIS_LOG, torch.from_numpy(log_trans).cuda(), torch.from_numpy(prior).cuda(), want_delta=True, want_psi=True)
print(json.dumps({
    "shape": {"K": K, "C": C, "D": D, "B": B, "T": T},
    "emission_max_rel_err_vs_float64": emis_rel,
    "gamma_max_rel_err": {"engine_vs_float64": rel(ours, gam64), "fp32_logspace_reference_like_vs_float64": rel(gam32, gam64),
                          "engine_vs_reference_like": float(np.max(np.abs(ours - gam32)[mask] / gam32[mask]))},
    "loglik_max_rel_err_engine_vs_float64": float(np.max(np.abs(r["loglik"].cpu().numpy() - ll64) / np.abs(ll64))),
    "viterbi_same_fp32_inputs": {"frames": B * T, "state_mismatches": int((v["states"].cpu().numpy() != st).sum()),
                                 "delta_bit_mismatches": int((v["delta"].cpu().numpy() != dl).sum()),
                                 "psi_mismatches": int((v["psi"].cpu().numpy().astype(np.int32) != psi).sum()),
                                 "score_bit_mismatches": int((v["score"].cpu().numpy() != sc).sum())}}, indent=1))
