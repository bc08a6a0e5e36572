"""Dumps the role hand-off timeline of the tcgen05 emission kernel (CTA 0, first tiles).  Debug aid."""
import ctypes, os, sys
os.environ["HMMB200_DEBUG_BUILD"] = "1"      # the trace hooks exist in debug builds only (libhmm_b200_dbg.so)
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["HMMB200_TC_DBG"] = str(8 | int(os.environ.get("TC_EXTRA", "0")))
import bench
from pytorch_hmm_b200 import _lib
dev = torch.device("cuda", 0)
model = bench.make_model()
x = torch.randn(bench.BATCH, bench.SEQ, bench.FEAT, device=dev)
h = bench.Headline(model, dev)
for _ in range(3):
    h.emission(x)
torch.cuda.synchronize()
buf = np.zeros((6, 48, 2), np.int64)
lib = ctypes.CDLL(_lib.lib_path())
lib.hmmb200_debug_tc_trace(buf.ctypes.data_as(ctypes.c_void_p))
t0 = buf[0, 0, 0]
names = ["producer(wait_empty,issued)", "mma(start,committed)", "xfA(x_full,done)", "xfB(x_full,done)", "epi(d_full,released)", "epi2(ld0,drained)"]
for it in range(0, 28):
    print(it, " | ".join(f"{names[r].split('(')[0]} {buf[r, it, 0] - t0:7d} {buf[r, it, 1] - t0:7d}" for r in range(6)))
