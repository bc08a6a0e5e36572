"""Warp-placement search for the fused forward + backward + Viterbi kernel (debug build: run-time warp maps).
Warp w issues from SM sub-partition w % 4; SMSP0 = w0,4,8,12,16  SMSP1 = w1,5,9,13,17  SMSP2 = w2,6,10,14  SMSP3 = w3,7,11,15."""
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
F, B, V = "F", "B", "V"
PIPE = {"F": 0, "B": 1, "V": 2}


def by_smsp(s0, s1, s2, s3):
    """four lists of 'F0'..'V7' role names -> 18-entry warp map"""
    assert len(s0) == 5 and len(s1) == 5 and len(s2) == 4 and len(s3) == 4
    lay = [None] * 18
    for sm, names in enumerate((s0, s1, s2, s3)):
        for i, n in enumerate(names):
            lay[sm + 4 * i] = (PIPE[n[0]], int(n[1]))
    return lay


LAYOUTS = {
    "L6  own-pipeline helpers beside each consumer, V drainers on SMSP0": by_smsp(["V3", "V4", "V5", "V6", "V7"], ["F0", "F1", "F2", "F3", "B4"], ["B0", "B1", "B2", "B3"], ["V0", "V1", "V2", "F4"]),
    "L8  one pipeline per SMSP (F | B | V0-3 | V4-7)": by_smsp(["F0", "F1", "F2", "F3", "F4"], ["B0", "B1", "B2", "B3", "B4"], ["V0", "V1", "V2", "V3"], ["V4", "V5", "V6", "V7"]),
    "L10 consumers beside the light V loaders": by_smsp(["F1", "F2", "F3", "F4", "V3"], ["B3", "B4", "V4", "V5", "V6"], ["F0", "V1", "V2", "V7"], ["B0", "V0", "B1", "B2"]),
    "L11 three consumers on one SMSP": by_smsp(["F1", "F2", "F3", "F4", "V1"], ["B1", "B2", "B3", "B4", "V2"], ["F0", "B0", "V0", "V7"], ["V3", "V4", "V5", "V6"]),
    "L12 F+B consumers together, V apart": by_smsp(["B1", "B2", "F3", "F4", "B3"], ["B4", "V3", "V4", "V5", "V6"], ["F0", "B0", "V1", "V2"], ["V0", "V7", "F1", "F2"]),
    "L13 L6 with the idle V drainer beside the F consumer": by_smsp(["V3", "V4", "V5", "V6", "B4"], ["F0", "F1", "F2", "F3", "V7"], ["B0", "B1", "B2", "B3"], ["V0", "V1", "V2", "F4"]),
    "L14 L6 with drainers swapped (F3/F4 away from the F consumer)": by_smsp(["V3", "V4", "V5", "V6", "F3"], ["F0", "F1", "F2", "V7", "B4"], ["B0", "B1", "B2", "B3"], ["V0", "V1", "V2", "F4"]),
}


def pack(layout):
    assert sorted(layout) == sorted([(0, r) for r in range(5)] + [(1, r) for r in range(5)] + [(2, r) for r in range(8)])
    pm = sum(p << (2 * i) for i, (p, r) in enumerate(layout))
    rm = sum(r << (3 * i) for i, (p, r) in enumerate(layout))
    return pm, rm


def t(fn, it=40):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    return bench.event_ms(fn, it)


for name, lay in LAYOUTS.items():
    pm, rm = pack(lay)
    os.environ["HMMB200_FUSED_MAP"] = f"{pm},{rm}"
    os.environ["HMMB200_FUSED_DBG"] = "0"
    a = t(lambda: h.fused(want=()))
    b = t(lambda: h.fused(want=()))
    print(f"{name:70s} all {a:.4f} {b:.4f}   pm={pm} rm={rm}", flush=True)
