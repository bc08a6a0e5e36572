"""Run-to-run differences of the small-K kernels when they run next to each other (debug build: per-kernel feed switches).
Round 2 found a write-after-read race on the raw emission stage with this script (a reader warp released a buffer before its
shared-memory loads had been performed); it is kept as the regression check: every line must report 0 differences."""
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
e, o = h.eng, h.eng.out
aux = torch.cuda.Stream(dev)
ENV = ("HMMB200_NO_BULK", "HMMB200_NO_BULK_FB", "HMMB200_NO_BULK_VIT")


def setenv(env):
    for k in ENV:
        os.environ.pop(k, None)
    os.environ.update(env)


# references: each kernel alone (per-lane feed; the two feeds compute the same numbers)
setenv({"HMMB200_NO_BULK": "1"})
h.vit(); h.fb(); torch.cuda.synchronize()
ref_v = (o["log_delta"].clone(), o["states"].clone())
ref_f = (o["posterior"].clone(), o["forward"].clone(), o["backward"].clone())


def first_bad(cur, ref):
    d = (cur != ref).any(-1)
    seqs = d.any(-1).nonzero().flatten().tolist()
    return [(s, int(d[s].nonzero()[0])) for s in seqs[:4]]


def run(tag, beside, env, reps=60):
    setenv(env)
    bad_v = bad_f = 0
    first = None
    for rep in range(reps):
        o["log_delta"].zero_(); o["posterior"].zero_()
        if beside:
            aux.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(aux):
                h.fb()
        h.vit()
        torch.cuda.current_stream(dev).wait_stream(aux)
        torch.cuda.synchronize()
        if not (torch.equal(o["log_delta"], ref_v[0]) and torch.equal(o["states"], ref_v[1])):
            bad_v += 1
            first = first or ("V", first_bad(o["log_delta"], ref_v[0]))
        if beside and not all(torch.equal(a, b) for a, b in zip((o["posterior"], o["forward"], o["backward"]), ref_f)):
            bad_f += 1
            first = first or ("F", first_bad(o["posterior"], ref_f[0]))
    print(f"{tag}: viterbi {bad_v} / fb {bad_f} of {reps} differ; first bad (sequence, first frame): {first}", flush=True)


def run_fused(tag, env, reps=60):
    setenv(env)
    bad_v = bad_f = 0
    first = None
    for rep in range(reps):
        o["log_delta"].zero_(); o["posterior"].zero_()
        h.fused()
        torch.cuda.synchronize()
        if not (torch.equal(o["log_delta"], ref_v[0]) and torch.equal(o["states"], ref_v[1])):
            bad_v += 1
            first = first or ("V", first_bad(o["log_delta"], ref_v[0]))
        if not all(torch.equal(a, b) for a, b in zip((o["posterior"], o["forward"], o["backward"]), ref_f)):
            bad_f += 1
            first = first or ("F", first_bad(o["posterior"], ref_f[0]))
    print(f"{tag}: viterbi {bad_v} / fb {bad_f} of {reps} differ; first bad: {first}", flush=True)


run("viterbi per-lane beside fb per-lane", True, {"HMMB200_NO_BULK": "1"})
run("viterbi bulk, alone", False, {})
run("viterbi bulk beside fb bulk", True, {})
run("viterbi bulk beside fb per-lane", True, {"HMMB200_NO_BULK_FB": "1"})
run("viterbi per-lane beside fb bulk", True, {"HMMB200_NO_BULK_VIT": "1"})
run_fused("fused kernel", {})
