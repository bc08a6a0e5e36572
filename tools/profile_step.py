"""Runs a few headline steps (emission -> forward-backward -> Viterbi at K=12, C=4, D=80, B=256, T=2000) and exits.
Small driver for ncu / compute-sanitizer captures; prints per-kernel CUDA-event times when run plainly."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=bench.BATCH)
    ap.add_argument("--seq", type=int, default=bench.SEQ)
    args = ap.parse_args()
    bench.BATCH, bench.SEQ = args.batch, args.seq
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    model = bench.make_model()
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(args.batch, args.seq, bench.FEAT, generator=g)).to(dev)
    torch.set_grad_enabled(False)
    h = bench.Headline(model, dev)
    for _ in range(args.steps):
        h.step(x)                                  # emission -> fused forward + backward + Viterbi -> posteriors
    torch.cuda.synchronize()
    if os.environ.get("PROFILE_STANDALONE"):
        for name, fn in (("emission", lambda: h.emission(x)), ("fused", h.fused), ("fb", h.fb), ("fb_sweeps_only", lambda: h.fb(want=())),
                         ("viterbi", h.vit)):
            print(name, round(bench.event_ms(fn, 5), 4), "ms")


if __name__ == "__main__":
    main()
