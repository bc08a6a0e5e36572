"""Isolate run-to-run differences of the small-K kernels when they run next to each other (debug build: per-kernel feed switches,
HMMB200_RAW_REL = how a loader warp releases a raw-stage buffer: 0 syncwarp + arrive, 1 data-dependent vote + arrive, 2 syncwarp +
fence.proxy.async + arrive)."""
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
ENV = ("HMMB200_NO_BULK", "HMMB200_NO_BULK_FB", "HMMB200_NO_BULK_VIT", "HMMB200_RAW_REL")


def setenv(env):
    for k in ENV:
        os.environ.pop(k, None)
    os.environ.update(env)


# references from the per-lane feed, each kernel alone
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
for rel in ("0", "1", "2"):
    run(f"rel={rel} viterbi bulk, alone", False, {"HMMB200_RAW_REL": rel})
    run(f"rel={rel} viterbi bulk beside fb bulk", True, {"HMMB200_RAW_REL": rel})
    run(f"rel={rel} viterbi bulk beside fb per-lane", True, {"HMMB200_RAW_REL": rel, "HMMB200_NO_BULK_FB": "1"})
    run(f"rel={rel} viterbi per-lane beside fb bulk", True, {"HMMB200_RAW_REL": rel, "HMMB200_NO_BULK_VIT": "1"})
    run_fused(f"rel={rel} fused kernel", {"HMMB200_RAW_REL": rel})

# cost of the release variants (fused kernel, no posterior pass)
for rel in ("1", "2", "1", "2"):
    setenv({"HMMB200_RAW_REL": rel})
    print(f"rel={rel} fused kernel {bench.event_ms(lambda: h.fused(want=()), 40):.4f} ms", flush=True)
