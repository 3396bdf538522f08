"""GPU box helper (torchrun, N ranks): what the host can move per second when every rank copies the end-to-end path's buffers
(26.7 MB pinned host -> device, 67.4 MB device -> pinned host per step, on two streams, nothing else running) -- the ceiling of
`e2e` at N GPUs.  Rank 0 prints one JSON line."""
import json, os, time
import torch
import torch.distributed as dist

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
h2d_bytes, d2h_bytes = 26738688, 67371008
hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
hout = [torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
din = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
dout = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def loop(steps):
    for k in range(steps):
        with torch.cuda.stream(s_in):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s_out):
            hout[k & 1].copy_(dout, non_blocking=True)
    torch.cuda.synchronize()


loop(8)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
steps = 200
t0 = time.perf_counter()
loop(steps)
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    per_rank = steps * (h2d_bytes + d2h_bytes) / float(dt[0]) / 1e9
    print(json.dumps({"n_gpus": world, "steps_per_s_per_rank": steps / float(dt[0]), "GBps_per_rank": per_rank, "GBps_aggregate": per_rank * world,
                      "h2d_GBps_per_rank": steps * h2d_bytes / float(dt[0]) / 1e9, "d2h_GBps_per_rank": steps * d2h_bytes / float(dt[0]) / 1e9}), flush=True)
if world > 1:
    dist.destroy_process_group()
