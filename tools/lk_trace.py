"""Phase timeline (clock64) of the large-K forward sweep, CTA 0, steps 64..71.  Debug aid."""
import ctypes, os, sys
os.environ["HMMB200_DEBUG_BUILD"] = "1"      # the trace hooks exist in debug builds only (libhmm_b200_dbg.so)
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["HMMB200_LK_TRACE"] = "1"
import pytorch_hmm_b200 as hm
from pytorch_hmm_b200 import _lib
K, B, T = int(os.environ.get("K", 512)), int(os.environ.get("B", 64)), 200
dev = torch.device("cuda", 0)
hmm = hm.HMMPyTorch(hm.create_transition_matrix(K, "ergodic"), None, device="cuda")
obs = torch.softmax(torch.randn(B, T, K, device=dev), -1)
trans, init = hmm._effective_probs(dev)
MODE = os.environ.get("MODE", "fwd")            # fwd | fb | vit | all
for _ in range(2):
    if MODE == "all":     # the three sweeps in one launch (cluster 0 = the forward sweep)
        hm.ops.forward_backward_viterbi(torch.log(obs), hm.ops.EMIS_LOG, hm.ops.EMIS_LOG, trans, init, torch.log(trans), torch.log(init))
    elif MODE == "fb":
        hm.ops.forward_backward(obs, hm.ops.EMIS_PROB_FLOOR, trans, init)
    elif MODE == "vit":
        hm.ops.viterbi(torch.log(obs), hm.ops.EMIS_LOG, torch.log(trans), torch.log(init))
    else:
        hm.ops.forward_backward(obs, hm.ops.EMIS_PROB_FLOOR, trans, init, want=())
torch.cuda.synchronize()
buf = np.zeros(64, np.int64)
ctypes.CDLL(_lib.lib_path()).hmmb200_debug_lk_trace(buf.ctypes.data_as(ctypes.c_void_p))
buf = buf.reshape(8, 8)
names = ["top", "waited", "computed", "-", "synced", "stored", "arrived", "pushed"]
buf[:, 3] = buf[:, 2]      # slot 3 unused
for i in range(8):
    base = buf[i, 0]
    nxt = buf[i + 1, 0] - base if i < 7 else -1
    print(f"step {64 + i}: " + " ".join(f"{n}={buf[i, j] - base}" for j, n in enumerate(names)) + f" | next_top={nxt}")
lib = ctypes.CDLL(_lib.lib_path())
print("max co-resident clusters by cluster size:", {cs: lib.hmmb200_debug_lk_max_clusters(cs) for cs in (1, 2, 4, 8)})
