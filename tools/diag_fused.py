"""Timing experiments on the fused forward + backward + Viterbi kernel (debug build: run-time warp maps, pipeline skip masks)."""
import os, sys
os.environ["HMMB200_DEBUG_BUILD"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.set_grad_enabled(False)

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


F, B, V = 0, 1, 2
# role: 0 consumer; F/B: 1-2 loaders, 3-4 drainers; V: 1-2 loaders, 3-7 drainers.  warp w issues from SM sub-partition w % 4.
LAYOUTS = {
    "L0 consumers 15/16/17, 4 V drainers on SMSP2": [(F,1),(B,1),(V,3),(B,2),(F,3),(B,3),(V,4),(F,4),(V,1),(V,2),(V,5),(B,4),(V,7),(F,2),(V,6),(V,0),(B,0),(F,0)],
    "L2 spread drainers, consumers 14/15/16": [(V,5),(V,7),(V,3),(V,4),(V,6),(B,1),(F,1),(F,3),(F,2),(F,4),(B,2),(B,3),(V,1),(B,4),(V,0),(F,0),(B,0),(V,2)],
    "L3 contiguous blocks, consumers first (F 0-4, B 5-9, V 10-17)": [(F,0),(F,1),(F,2),(F,3),(F,4),(B,0),(B,1),(B,2),(B,3),(B,4),(V,0),(V,1),(V,2),(V,3),(V,4),(V,5),(V,6),(V,7)],
    "L4 contiguous blocks, V first (V 0-7, F 8-12, B 13-17)": [(V,0),(V,1),(V,2),(V,3),(V,4),(V,5),(V,6),(V,7),(F,0),(F,1),(F,2),(F,3),(F,4),(B,0),(B,1),(B,2),(B,3),(B,4)],
    "L5 consumers 0/1/2, V drainers 3,7,11,15 + 4": [(F,0),(B,0),(V,0),(V,3),(V,7),(F,1),(B,1),(V,4),(V,1),(F,2),(B,2),(V,5),(V,2),(F,3),(B,3),(V,6),(F,4),(B,4)],
    "L6 consumers 1/2/3 (SMSP 1,2,3), helpers round-robin": [(V,3),(F,0),(B,0),(V,0),(V,4),(F,1),(B,1),(V,1),(V,5),(F,2),(B,2),(V,2),(V,6),(F,3),(B,3),(F,4),(V,7),(B,4)],
}


def pack(layout):
    assert len(layout) == 18 and sorted(layout) == sorted([(F, r) for r in range(5)] + [(B, r) for r in range(5)] + [(V, r) for r in range(8)])
    pm = sum(p << (2 * i) for i, (p, r) in enumerate(layout))
    rm = sum(r << (3 * i) for i, (p, r) in enumerate(layout))
    return f"{pm},{rm}"


for name, lay in LAYOUTS.items():
    os.environ["HMMB200_FUSED_MAP"] = pack(lay)
    res = []
    for skip, tag in ((0, "all"), (0x60, "F"), (0x50, "B"), (0x30, "V"), (0x40, "F+B")):
        os.environ["HMMB200_FUSED_DBG"] = str(skip)
        res.append(f"{tag} {t(lambda: h.fused(want=())):.4f}")
    print(f"{name:62s} " + "  ".join(res), flush=True)
os.environ.pop("HMMB200_FUSED_MAP")
os.environ["HMMB200_FUSED_DBG"] = "0"
print("stand-alone (bulk feed): fb_sweep", round(t(lambda: h.fb(want=())), 4), "viterbi", round(t(lambda: h.vit()), 4))
os.environ["HMMB200_NO_BULK"] = "1"
print("stand-alone (per-lane loads): fb_sweep", round(t(lambda: h.fb(want=())), 4), "viterbi", round(t(lambda: h.vit()), 4))
os.environ.pop("HMMB200_NO_BULK")

# determinism of the stand-alone Viterbi kernel, alone and beside the sweeps on a second stream
e, o = h.eng, h.eng.out
ref = None
aux = torch.cuda.Stream(dev)
for mode in ("alone", "beside fb_sweep", "alone, per-lane loads", "beside fb_sweep, per-lane loads"):
    if "per-lane" in mode:
        os.environ["HMMB200_NO_BULK"] = "1"
    bad = 0
    for rep in range(150):
        if "beside" in mode:
            aux.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(aux):
                h.fb()
        h.vit()
        torch.cuda.current_stream(dev).wait_stream(aux)
        torch.cuda.synchronize()
        cur = (o["log_delta"].clone(), o["states"].clone())
        if ref is None:
            ref = cur
        elif not (torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1])):
            bad += 1
    print(f"stand-alone viterbi {mode}: {bad} of 150 runs differ from the first")
os.environ.pop("HMMB200_NO_BULK", None)
