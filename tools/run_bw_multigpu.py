"""Baum-Welch EM over sharded synthetic utterances on N GPUs (BASELINE.json configs[2], scaled by --utts).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_bw_multigpu.py --utts 4096 --iters 3

Utterance u is generated from seed 3_000_000 + u (a function of the utterance id only), so every sharding sees the same
data; rank 0 prints one JSON line with frames/s per EM iteration and the per-iteration log-likelihood (identical for
every N up to fp64 summation order)."""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pytorch_hmm_b200 import baum_welch as bw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--seq", type=int, default=bench.SEQ)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = bench.make_model(3001)
    lo, hi = bw.shard_range(a.utts, rank, world)
    batches = []
    P = torch.softmax(model["transition_logits"], -1).to(dev)
    cdf = torch.cumsum(P, -1)
    for s in range(lo, hi, a.batch):
        e = min(hi, s + a.batch)
        # every random number of utterance u comes from torch.Generator().manual_seed(3_000_000 + u): the data are a function
        # of the utterance id only, whatever the sharding; the Markov chain itself is stepped for the whole
        # batch at once by inverse-CDF sampling
        us, cs, zs = [], [], []
        for u in range(s, e):
            g = torch.Generator().manual_seed(3_000_000 + u)
            us.append(torch.rand(a.seq, generator=g)); cs.append(torch.randint(0, bench.N_MIX, (a.seq,), generator=g))
            zs.append(torch.randn(a.seq, bench.FEAT, generator=g))
        ur, comp, z = torch.stack(us).to(dev), torch.stack(cs).to(dev), torch.stack(zs).to(dev)
        st = torch.empty(e - s, a.seq, dtype=torch.long, device=dev)
        st[:, 0] = (ur[:, 0] * bench.K_STATES).long().clamp_(max=bench.K_STATES - 1)
        for t in range(1, a.seq):
            st[:, t] = (ur[:, t, None] > cdf[st[:, t - 1]]).sum(-1).clamp_(max=bench.K_STATES - 1)
        means, lv = model["means"].to(dev), model["log_vars"].to(dev)
        batches.append((means[st, comp] + torch.exp(0.5 * lv[st, comp]) * z).contiguous())
    K, C, D = bench.K_STATES, bench.N_MIX, bench.FEAT
    g = torch.Generator().manual_seed(3001)
    start = bw.GMMHMMParams(torch.softmax(model["transition_logits"], -1), torch.full((K,), 1.0 / K),
                            torch.softmax(model["mixture_weights_logits"], -1),
                            model["means"] + 0.05 * torch.randn(K, C, D, generator=g), torch.ones(K, C, D))
    tr = bw.BaumWelch(start, device=dev)
    hist, times = [], []
    for it in range(a.iters):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for xb in batches:
            tr.e_step(xb)
        hist.append(tr.m_step())
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    if rank == 0:
        print(json.dumps({"what": "Baum-Welch EM, utterance-sharded, one stats all-reduce per iteration", "n_gpus": world,
                          "utterances": a.utts, "T": a.seq, "iters": a.iters, "loglik_per_frame": hist,
                          "frames_per_s_per_iter": [a.utts * a.seq / t for t in times]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
