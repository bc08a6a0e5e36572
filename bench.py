#!/usr/bin/env python
"""bench.py -- frames/sec of forward-backward + Viterbi (+ the GMM emissions that feed them) on B200.

Contract (task brief): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line from rank 0.
  step      one pass of the hot path over one batch: GMM emission -> forward-backward (posterior, forward, backward)
            + Viterbi (states, log_delta) on BASELINE.json configs[1]: K=12 states, 4 mixtures, D=80, B=256, T=2000.
  value     frames/s with inputs resident in HBM (whole job: all ranks' frames / max-over-ranks device time).
  e2e       same metric through the Python drop-in surface with HOST buffers: pinned host x -> device, step,
            every API-visible output -> pinned host, inside the timed region.
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM copy bandwidth.
  cpu_baseline   oracle/ref_port.py (op-for-op torch port of the reference, kind="port") on the host cores, on a
            bounded sample of the same workload.
`--impl reference` times that CPU port instead (rank 0 only) and prints the same line with "impl": "reference".
Multi-GPU: the batch of utterances is sharded across ranks (independent sequences, no data-path collective):
weak scaling, B=256 per GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_STATES, N_MIX, FEAT, BATCH, SEQ = 12, 4, 80, 256, 2000
BYTES_PER_FRAME = 4 * FEAT + 4 * K_STATES * 4 + 8          # x in; posterior, forward, backward, log_delta + int64 state out
METRIC = "frames/sec forward-backward+Viterbi (K=12,T=2000,B=256)"
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this workload
NCU_TRAFFIC_SRC = "profiles/r01_ncu_full_summary.md"
NCU_TRAFFIC = {"gmm_emission_tc_kernel": 175.8e6, "fb_sweep_kernel": 30.0e6, "fb_combine_kernel": 73.4e6, "viterbi_kernel": 24.7e6}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
# synthetic workload (seeded; SURVEY.md 8(d) C2 "soft" regime: the layer's own initialisation, x sampled from the model)
# ------------------------------------------------------------------------------------------------------------
def make_model(seed=2001):
    g = torch.Generator().manual_seed(seed)
    import math
    return {
        "transition_logits": torch.randn(K_STATES, K_STATES, generator=g) * 0.1,
        "mixture_weights_logits": torch.randn(K_STATES, N_MIX, generator=g) * 0.1,
        "means": torch.randn(K_STATES, N_MIX, FEAT, generator=g) * math.sqrt(2.0 / FEAT),
        "log_vars": torch.zeros(K_STATES, N_MIX, FEAT),
    }


def make_frames(model, batch, seq, seed):
    g = torch.Generator().manual_seed(seed)
    P = torch.softmax(model["transition_logits"], -1)
    s = torch.empty(batch, seq, dtype=torch.long)
    s[:, 0] = torch.randint(0, K_STATES, (batch,), generator=g)
    for t in range(1, seq):
        s[:, t] = torch.multinomial(P[s[:, t - 1]], 1, generator=g).squeeze(1)
    c = torch.randint(0, N_MIX, (batch, seq), generator=g)
    x = model["means"][s, c] + torch.exp(0.5 * model["log_vars"][s, c]) * torch.randn(batch, seq, FEAT, generator=g)
    return x.contiguous()


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,utilization.gpu,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                clk, cmax, util = float(f[0]), float(f[1]), float(f[2])
            except ValueError:
                continue
            mx.append(cmax)
            if util > 0:
                sm.append(clk)
                for n, v in zip(names, f[4:8]):
                    if v == "Active":
                        reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples_under_load": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the op-for-op torch port of the reference
# ------------------------------------------------------------------------------------------------------------
def cpu_port_step(model, x):
    from oracle import ref_port
    P = torch.softmax(model["transition_logits"], -1)
    log_P, log_p0 = ref_port.prepare_hmm(P, None)
    log_trans = ref_port.safe_log(P)
    with torch.no_grad():
        return ref_port.headline_step(x, model["means"], model["log_vars"], model["mixture_weights_logits"],
                                      log_P, log_p0, log_trans)


def time_cpu_port(model, x_sample, steps, warmup):
    for _ in range(warmup):
        cpu_port_step(model, x_sample)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_port_step(model, x_sample)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return x_sample.shape[0] * x_sample.shape[1] / dt, dt


def run_reference_arm(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    model = make_model()
    # bounded sample: the recursion is dispatch-bound, so keep the per-step cost ~1-2 s and the whole run in minutes
    budget_s = 150.0
    bs = 32
    x = make_frames(model, bs, SEQ, 2001)
    t0 = time.perf_counter(); cpu_port_step(model, x[:8]); probe = time.perf_counter() - t0
    per_step = probe * 2.5                                   # B=32 costs roughly 2-3x the B=8 probe
    while bs > 4 and per_step * (args.steps + args.warmup) > budget_s:
        bs //= 2; per_step *= 0.6
    x = x[:bs].contiguous()
    fps, dt = time_cpu_port(model, x, args.steps, args.warmup)
    cores = torch.get_num_threads()
    sample = f"{bs} of {BATCH} sequences x T={SEQ} per step (op-for-op torch port of the reference, all host threads)"
    emit_result({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: mixture_gaussian K=12, 4 mixtures, D=80, B=256, T=2000 (CPU arm runs a bounded sample)",
                   "B": BATCH, "T": SEQ, "K": K_STATES, "C": N_MIX, "D": FEAT},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
class Headline:
    """Device-side state of the headline step, through the public engine / ops of pytorch_hmm_b200."""

    def __init__(self, model, dev, shard=None, n_streams=1, host_io=False):
        import pytorch_hmm_b200 as hm
        from pytorch_hmm_b200.engine import HMMInferenceEngine
        self.hm, self.dev = hm, dev
        self.layer = hm.MixtureGaussianHMMLayer(K_STATES, FEAT, num_components=N_MIX).to(dev)
        self.layer.load_state_dict({k: v.to(dev) for k, v in model.items()})
        self.layer.eval()
        self.eng = HMMInferenceEngine(self.layer, BATCH, SEQ, shard=shard or BATCH, n_streams=n_streams, device=dev,
                                      host_io=host_io)
        e = self.eng
        self.logb = e.slots[0].logb
        self.launches_per_step = e.kernels_per_shard * e.n_shards

    # single kernels on the current stream (per-kernel timing only)
    def emission(self, x):
        e = self.eng
        self.hm.ops.gmm_emission(x, e.packed, K_STATES, N_MIX, FEAT, out=self.logb, tc_known=e.tc_known)

    def fb(self, want=("gamma", "fwd", "bwd")):
        e, o = self.eng, self.eng.out
        out = {"loglik": o["loglik"]}
        if want:
            out.update({"gamma": o["posterior"], "fwd": o["forward"], "bwd": o["backward"]})
        self.hm.ops.forward_backward(self.logb, self.hm.ops.EMIS_LOG_NORM_FLOOR, e.trans, e.init, want=want, out=out,
                                     workspace=e.slots[0].fb_ws)

    def vit(self):
        e, o = self.eng, self.eng.out
        self.hm.ops.viterbi(self.logb, self.hm.ops.EMIS_LOG, e.log_trans, e.prior,
                            out={"states": o["states"], "delta": o["log_delta"], "score": o["score"]},
                            workspace=e.slots[0].vit_ws)

    def step(self, x):
        self.eng.run_device(x)

    def outputs(self):
        o = self.eng.out
        return [o[k] for k in ("posterior", "forward", "backward", "log_delta", "states")]


def event_ms(fn, iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / iters


_ORIG_AFFINITY = None


def pin_to_gpu_numa_node(local_rank):
    """Host side of the end-to-end path: run this rank (and allocate its pinned buffers) on the NUMA node its GPU hangs
    off, so that PCIe copies do not cross the socket interconnect.  Best effort; returns a note for the JSON line."""
    global _ORIG_AFFINITY
    try:
        _ORIG_AFFINITY = os.sched_getaffinity(0)
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"numa: rank pinned to node {node} ({len(cpus)} cpus)"
        return f"numa: node {node} has no allowed cpus"
    except Exception as exc:                                      # noqa: BLE001
        return f"numa: not pinned ({type(exc).__name__})"


def run_gpu_arm(args, rank, world, local_rank):
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; pytorch_hmm_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa_note = pin_to_gpu_numa_node(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = make_model()
    x_host = make_frames(model, BATCH, SEQ, 2001 + 7919 * rank).pin_memory()     # this rank's shard of utterances
    x = x_host.to(dev, non_blocking=True)
    h = Headline(model, dev)
    torch.cuda.synchronize()
    # one pass = one CUDA-graph launch (emission, then forward-backward || Viterbi on two streams); falls back to
    # eager stream launches if capture is refused
    graph = None
    if not args.no_graph:
        try:
            graph = h.eng.capture_device(x)
        except Exception as exc:                                  # noqa: BLE001
            print(f"bench.py: CUDA-graph capture failed ({exc}); timing eager launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    step = (lambda: graph.replay()) if graph is not None else (lambda: h.step(x))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    # ---- timed region: K steps, inputs resident in HBM; x (164 MB) + logb + outputs exceed the 126 MB L2 ----
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        step()
    e.record()
    barrier()
    ms_total = s.elapsed_time(e)
    if world > 1:
        t = torch.tensor([ms_total], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * BATCH * SEQ / (ms_step * 1e-3)

    # keep the same loop running ~1.5 s so the 50 ms clock sampler sees the GPU under this load
    t_end = time.perf_counter() + 1.5
    while time.perf_counter() < t_end:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel durations (CUDA events on the launch stream, same loop) -> dominant kernel + roofline ----
    it = max(10, min(args.steps, 50))
    k_ms = {
        "gmm_emission_tc_kernel": event_ms(lambda: h.emission(x), it),
        "fb_sweep_kernel": event_ms(lambda: h.fb(want=()), it),
        "fb_sweep_kernel+fb_combine_kernel": event_ms(lambda: h.fb(), it),
        "viterbi_kernel": event_ms(lambda: h.vit(), it),
    }
    k_ms["fb_combine_kernel"] = max(k_ms["fb_sweep_kernel+fb_combine_kernel"] - k_ms["fb_sweep_kernel"], 0.0)
    frames = BATCH * SEQ
    alg_bytes = {   # algorithmic bytes per launch (per-frame figure x frames per launch), DESIGN.md "Kernels"
        "gmm_emission_tc_kernel": (4 * FEAT + 4 * K_STATES) * frames,
        "fb_sweep_kernel": (4 * K_STATES + 2 * (4 * K_STATES + 4)) * frames,
        "fb_combine_kernel": (2 * (4 * K_STATES + 4) + 3 * 4 * K_STATES) * frames,
        "viterbi_kernel": (4 * K_STATES + 4 * K_STATES + 8) * frames,
    }
    dom = max(alg_bytes, key=lambda k: k_ms[k])
    peak, peak_src = measured_peaks()
    achieved = alg_bytes[dom] / (k_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC.get(dom), "traffic_source": NCU_TRAFFIC_SRC, "peak_source": peak_src,
                "kernel_ms": {k: round(v, 4) for k, v in k_ms.items()},
                "note": "the recursion kernels are bound by the latency of T dependent steps, not by HBM (DESIGN.md 4.2); "
                        "per-kernel fractions: " + ", ".join(
                            f"{k} {alg_bytes[k] / (k_ms[k] * 1e-3) / 1e9 / peak:.3f}" for k in alg_bytes)
                        + " (fb_combine = the posterior kernel fb_combine_warp_kernel, timed as fb minus the sweeps; its inputs are"
                          " still in L2 from the sweeps and its outputs are written back after it ends, so its figure can exceed the HBM peak)",
                "path": {"bytes_per_frame": BYTES_PER_FRAME, "achieved": value / world * BYTES_PER_FRAME / 1e9,
                         "frac": value / world * BYTES_PER_FRAME / 1e9 / peak}}

    # ---- end to end: pinned host buffers in, pinned host buffers out, copies inside the timed region.  The public
    #      engine shards the batch over 4 streams so that H2D, kernels and D2H of different shards overlap. ----
    he = Headline(model, dev, shard=args.e2e_shard, n_streams=args.e2e_streams, host_io=True)
    names = ("posterior", "forward", "backward", "log_delta", "states")
    outs_host = [{k: torch.empty(he.eng.out[k].shape, dtype=he.eng.out[k].dtype).pin_memory() for k in names} for _ in range(2)]
    h2d = x_host.numel() * 4
    d2h = sum(o.numel() * o.element_size() for o in outs_host[0].values())

    def e2e_step(i):
        he.eng.run_host(x_host, outs_host[i & 1], join=False)     # consecutive passes pipeline; results alternate host sets

    for i in range(3):
        e2e_step(i)
    he.eng.join(); torch.cuda.synchronize()
    n_e2e = max(3, min(args.steps, 20))
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        e2e_step(i)
    he.eng.join()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_ms = float(t.item())
    e2e = {"value": world * BATCH * SEQ / (e2e_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e,
           "api": f"HMMInferenceEngine.run_host (shard={he.eng.shard}, streams={he.eng.n_streams})", "host": numa_note}
    # spot check: the pipelined host path returns what the single-pass device path computed
    torch.cuda.synchronize()
    same = all(torch.equal(outs_host[(n_e2e - 1) & 1][k], h.eng.out[k].cpu()) for k in names)
    e2e["matches_device_pass"] = bool(same)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "configs[1]: create_speech_hmm mixture_gaussian K=12, 4 mixtures, D=80, B=256, T=2000 "
                                   "(GMM emission -> forward_backward + viterbi_decode), B=256 per GPU",
                       "B": BATCH, "T": SEQ, "K": K_STATES, "C": N_MIX, "D": FEAT, "parallelism": f"utterance-sharded x{world}",
                       "l2": "no flush: per-step inputs+outputs (164 MB x, 25 MB log b, 102 MB outputs) exceed the 126 MB L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": h.launches_per_step * args.steps, "roofline": roofline,
        }
        line["config"]["launch"] = "cuda-graph replay per step" if not args.no_graph else "eager stream launches"
        if world == 1:
            if _ORIG_AFFINITY:
                os.sched_setaffinity(0, _ORIG_AFFINITY)          # the CPU baseline gets every host core again
            torch.set_num_threads(os.cpu_count() or 1)
            bs = 32
            xs = x_host[:bs].clone()
            fps, dt = time_cpu_port(model, xs, 3, 1)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{bs} of {BATCH} sequences x T={SEQ}, 1 warm-up + 3 timed passes of oracle/ref_port.py "
                                              f"({dt:.2f} s per pass)"}
        emit_result(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT_FD = None


def emit_result(line: dict):
    """Print the result line on the process's real stdout (see main)."""
    sys.stdout.flush()
    if _REAL_STDOUT_FD is not None:
        os.dup2(_REAL_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)
    if _REAL_STDOUT_FD is not None:
        os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    ap.add_argument("--e2e-shard", type=int, default=32, help="utterances per in-flight shard on the host path")
    ap.add_argument("--e2e-streams", type=int, default=4)
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: until it is printed, file descriptor 1 points at stderr so that
    # native libraries (NCCL prints its version banner on stdout at communicator creation) cannot add lines to it
    sys.stdout.flush()
    global _REAL_STDOUT_FD
    _REAL_STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
