"""Timing experiments on the fused forward + backward + Viterbi kernel (debug build: warp layout and pipeline skip masks)."""
import os, sys
os.environ["HMMB200_DEBUG_BUILD"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda", 0)
model = bench.make_model()
x = bench.make_frames(model, bench.BATCH, bench.SEQ, 2001).to(dev)
h = bench.Headline(model, dev)
h.emission(x)
torch.cuda.synchronize()


def t(fn, it=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    return bench.event_ms(fn, it)


for name, val in (("layout0 all", 0), ("layout1 all", 1), ("layout0 F only", 0x60), ("layout0 B only", 0x50), ("layout0 V only", 0x30),
                  ("layout0 F+B", 0x40), ("layout0 F+V", 0x20), ("layout0 B+V", 0x10), ("layout1 F+B", 0x41), ("layout1 V only", 0x31),
                  ("layout1 F only", 0x61)):
    os.environ["HMMB200_FUSED_DBG"] = str(val)
    print(f"{name:18s} fused kernel {t(lambda: h.fused(want=())):.4f} ms")
os.environ["HMMB200_FUSED_DBG"] = "0"
print("stand-alone fb_sweep", t(lambda: h.fb(want=())), "viterbi", t(lambda: h.vit()))
