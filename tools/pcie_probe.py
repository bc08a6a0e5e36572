"""Platform bound for the end-to-end number: every rank copies the bench's per-step bytes (164 MB pinned host -> device,
102 MB device -> pinned host) concurrently on two streams, nothing else.  Prints per-rank and aggregate GB/s."""
import os, time, torch, torch.distributed as dist
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
h_in = torch.empty(163_840_000 // 4).pin_memory(); d_in = torch.empty_like(h_in, device=dev)
d_out = torch.empty(102_400_000 // 4, device=dev); h_out = torch.empty(102_400_000 // 4).pin_memory()
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def step():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(20):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = (time.perf_counter() - t0) / 20
t = torch.tensor([dt], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print({"n_gpus": world, "ms_per_step_copies_only": t.item() * 1e3, "h2d_GBps_per_gpu": 0.16384 / t.item(), "d2h_GBps_per_gpu": 0.1024 / t.item(),
           "aggregate_GBps": world * 0.26624 / t.item()})
if world > 1:
    dist.destroy_process_group()
