"""Aggregate host<->device copy bandwidth of the box with ALL ranks copying at once (the floor of the N-GPU `e2e` numbers): pinned H2D,
D2H and both directions, 256 MB chunks, one process per GPU.  Launch with torchrun (launcher only; torch is used for the copies
and the barrier):  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 tools/pcie_multi.py"""
import json, os, time
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
total, chunk = 8 << 30, 256 << 20
h_in = torch.empty(total, dtype=torch.uint8).pin_memory()
h_out = torch.empty(total, dtype=torch.uint8).pin_memory()
d_in = torch.empty(chunk * 2, dtype=torch.uint8, device="cuda")
d_out = torch.empty(chunk * 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
n = total // chunk


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(do_in, do_out):
    barrier()
    t0 = time.perf_counter()
    for i in range(n):
        if do_in:
            with torch.cuda.stream(s1):
                d_in[(i & 1) * chunk:(i & 1) * chunk + chunk].copy_(h_in[i * chunk:(i + 1) * chunk], non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2):
                h_out[i * chunk:(i + 1) * chunk].copy_(d_out[(i & 1) * chunk:(i & 1) * chunk + chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return total / float(t.item()) / 1e9     # per rank, at the pace of the slowest rank


res = {}
for _ in range(2):
    res["h2d_GBps_per_gpu"] = run(True, False)
    res["d2h_GBps_per_gpu"] = run(False, True)
    res["both_each_GBps_per_gpu"] = run(True, True)
if rank == 0:
    res = {k: round(v, 2) for k, v in res.items()}
    res.update(n_gpus=world, aggregate_h2d_GBps=round(res["h2d_GBps_per_gpu"] * world, 1), aggregate_d2h_GBps=round(res["d2h_GBps_per_gpu"] * world, 1),
               aggregate_both_each_GBps=round(res["both_each_GBps_per_gpu"] * world, 1))
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
