"""BASELINE config 5: K=512 ergodic, B=64, T=4000 -- forward_backward + viterbi_decode on softmax(randn) observations.
Prints CUDA-event times per call (inputs resident in HBM)."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_hmm_b200 as hm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=512); ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--T", type=int, default=4000); ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(5001)
    P = hm.create_transition_matrix(a.K, "ergodic")
    hmm = hm.HMMPyTorch(P, None, device="cuda")
    obs = torch.softmax(torch.randn(a.B, a.T, a.K, device=dev), dim=-1)
    trans, init = hmm._effective_probs(dev)
    n = (a.B, a.T, a.K)
    fb_out = {k: torch.empty(n, device=dev) for k in ("gamma", "fwd", "bwd")}
    fb_out["loglik"] = torch.empty(a.B, device=dev)
    v_out = {"states": torch.empty(a.B, a.T, dtype=torch.int64, device=dev), "delta": torch.empty(n, device=dev),
             "score": torch.empty(a.B, device=dev)}
    fb_ws = hm.ops.fb_workspace(a.B, a.T, a.K, dev); v_ws = hm.ops.viterbi_workspace(a.B, a.T, a.K, dev)
    logP, logp0 = hmm.log_P.to(dev), hmm.log_p0.to(dev)

    def fb(want=("gamma", "fwd", "bwd")):
        hm.ops.forward_backward(obs, hm.ops.EMIS_PROB_FLOOR, trans, init, want=want,
                                out=fb_out if want else {"loglik": fb_out["loglik"]}, workspace=fb_ws)

    def vit():
        hm.ops.viterbi(obs, hm.ops.EMIS_PROB_FLOOR, logP, logp0, out=v_out, workspace=v_ws)

    def ms(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.iters):
            fn()
        e.record(); e.synchronize()
        return s.elapsed_time(e) / a.iters

    fused_out = dict(fb_out); fused_out.update(v_out)
    fused_ws = torch.empty(hm._lib.load().hmmb200_fb_viterbi_workspace_bytes(a.B, a.T, a.K), dtype=torch.uint8, device=dev)

    def fused():                                                  # the three sweeps in one launch (csrc/recursion_largek.cu)
        hm.ops.forward_backward_viterbi(obs, hm.ops.EMIS_PROB_FLOOR, hm.ops.EMIS_PROB_FLOOR, trans, init, logP, logp0, out=fused_out,
                                        workspace=fused_ws)

    res = {"K": a.K, "B": a.B, "T": a.T, "forward_only_ms": ms(lambda: fb(())), "forward_backward_ms": ms(fb), "viterbi_ms": ms(vit),
           "fused_fb_viterbi_ms": ms(fused)}
    fb(); vit(); torch.cuda.synchronize()
    ref = {k: fb_out[k].clone() for k in ("gamma", "loglik")}; ref["states"] = v_out["states"].clone(); ref["delta"] = v_out["delta"].clone()
    fused(); torch.cuda.synchronize()
    res["fused_matches_separate"] = bool(all(torch.equal(ref[k], fused_out[k]) for k in ref))
    res["frames_per_s_fb_plus_viterbi"] = a.B * a.T / (min(res["forward_backward_ms"] + res["viterbi_ms"], res["fused_fb_viterbi_ms"]) * 1e-3)
    res["exchange_ok"] = bool(int(fb_ws[-256:].view(torch.int32)[0]) == 0 and int(v_ws[-256:].view(torch.int32)[0]) == 0)
    import subprocess
    try:
        res["sm_mhz_now"] = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader", "-i", "0"], text=True).strip()
    except Exception:
        pass
    print(json.dumps(res))


if __name__ == "__main__":
    main()
