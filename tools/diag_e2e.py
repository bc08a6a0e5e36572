"""Diagnostic: does the sharded multi-stream host path reproduce the single-pass device path bit for bit (fused / PDL on or off)?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pytorch_hmm_b200.engine import HMMInferenceEngine
import pytorch_hmm_b200 as hm

dev = torch.device("cuda", 0)
model = bench.make_model()
x_host = bench.make_frames(model, bench.BATCH, bench.SEQ, 2001).pin_memory()
x = x_host.to(dev)
layer = hm.MixtureGaussianHMMLayer(bench.K_STATES, bench.FEAT, num_components=bench.N_MIX).to(dev).eval()
layer.load_state_dict({k: v.to(dev) for k, v in model.items()})
names = ("posterior", "forward", "backward", "log_delta", "states", "score", "loglik")
ref = None
for fused, pdl, shard, streams in ((False, False, 256, 1), (True, False, 256, 1), (True, True, 256, 1), (True, False, 32, 4), (True, True, 32, 4),
                                   (False, False, 32, 4), (True, True, 32, 1), (True, True, 64, 2)):
    eng = HMMInferenceEngine(layer, bench.BATCH, bench.SEQ, shard=shard, n_streams=streams, device=dev, host_io=True, fused=fused, pdl=pdl)
    outs = {k: torch.empty(eng.out[k].shape, dtype=eng.out[k].dtype).pin_memory() for k in names}
    bad_runs = {}
    for rep in range(6):
        eng.run_host(x_host, outs, join=True)
        torch.cuda.synchronize()
        cur = {k: v.clone() for k, v in outs.items()}
        if ref is None:
            ref = cur
        for k in names:
            if not torch.equal(cur[k], ref[k]):
                nbad = int((cur[k] != ref[k]).sum())
                bad_runs.setdefault(k, []).append((rep, nbad))
    print(f"fused={fused} pdl={pdl} shard={shard} streams={streams}: mismatches vs first config {bad_runs if bad_runs else 'none'}")
    del eng
