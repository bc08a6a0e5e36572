"""Per-phase CUDA-event times of one Baum-Welch E-step batch (config 3 shape: B=256, T=2000, K=12, C=4, D=80)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pytorch_hmm_b200 import baum_welch as bw, ops, _lib
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
model = bench.make_model(3001)
K, C, D = bench.K_STATES, bench.N_MIX, bench.FEAT
xb = bench.bw_utterances(model, 0, 256, bench.SEQ, dev)
start = bw.GMMHMMParams(torch.softmax(model["transition_logits"], -1), torch.full((K,), 1.0 / K),
                        torch.softmax(model["mixture_weights_logits"], -1), model["means"], torch.ones(K, C, D))
tr = bw.BaumWelch(start, device=dev)
for _ in range(3):
    tr.e_step(xb)
torch.cuda.synchronize()
print("e_step", round(bench.event_ms(lambda: tr.e_step(xb), 10), 4), "ms")
if len(sys.argv) > 1 and sys.argv[1] == "phases":
    lib = _lib.load()
    B, T = 256, bench.SEQ
    packed = tr._pack()
    logb, comp, out, ws = tr._buffers(B, T)
    S = ops._stream(dev)
    def em():
        ops._check(lib.hmmb200_gmm_emission_components_f32(ops._p(xb), ops._p(packed), B * T, K, C, D, ops._p(logb), ops._p(comp), S), "em")
    def fb():
        return ops.forward_backward(logb, ops.EMIS_LOG, tr._trans, tr._init, want=("gamma",), out=out, workspace=ws, method="sweep")
    r = fb()
    def acc():
        ops._check(lib.hmmb200_bw_accumulate_f32(ops._p(xb), ops._p(comp), ops._p(logb), ops._p(r["gamma"]), ops._p(logb), ops.EMIS_LOG, 0.0,
                                                 ops._p(tr._trans), ops._p(ws), B, T, K, C, D, ops._p(tr.stats), S), "acc")
    for name, fn in (("emission+components", em), ("forward_backward(gamma)", fb), ("bw_accumulate", acc)):
        print(name, round(bench.event_ms(fn, 10), 4), "ms")
    ws_b = ws
    def xi():
        ops._check(lib.hmmb200_xi_sum_f32(ops._p(logb), ops.EMIS_LOG, 0.0, ops._p(tr._trans), ops._p(ws), None, B, T, K,
                                          ops._p(tr.stats[K:]), ops._p(tr.stats), S), "xi")
    sl = bw.stats_slices(K, C, D)
    occ, sx, sxx = tr.stats[sl["occ"]], tr.stats[sl["sx"]], tr.stats[sl["sxx"]]
    def gs():
        ops._check(lib.hmmb200_gmm_stats_f32(ops._p(xb), ops._p(comp), ops._p(logb), ops._p(r["gamma"]), B * T, K, C, D,
                                             ops._p(occ), ops._p(sx), ops._p(sxx), S), "gs")
    for name, fn in (("xi_sum", xi), ("gmm_stats", gs)):
        print(name, round(bench.event_ms(fn, 10), 4), "ms")
