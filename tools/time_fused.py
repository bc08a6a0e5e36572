"""Times the fused recursion kernel and the two stand-alone recursion kernels on the headline shape (resident log b)."""
import os, sys, torch
sys.path.insert(0, "/root/repo")
import bench
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
model = bench.make_model()
x = bench.make_frames(model, bench.BATCH, bench.SEQ, 2001).to(dev)
h = bench.Headline(model, dev)
h.emission(x); torch.cuda.synchronize()
for _ in range(3):
    print("fused", round(bench.event_ms(lambda: h.fused(want=()), 40), 4), "fb", round(bench.event_ms(lambda: h.fb(want=()), 40), 4), "vit", round(bench.event_ms(h.vit, 40), 4))
for _ in range(3):
    print("fused+posteriors", round(bench.event_ms(lambda: h.fused(), 40), 4), "step", round(bench.event_ms(lambda: h.step(x), 40), 4))
