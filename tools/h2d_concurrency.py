"""Aggregate page-locked host <-> device bandwidth of one box with 1, 2, 4, 8 GPUs copying AT ONCE (one process per GPU,
launched under torchrun) -- the ceiling of bench.py's end-to-end leg, whose inputs live in pinned host memory.

  python -m torch.distributed.run --nproc-per-node N tools/h2d_concurrency.py [MiB per copy]

Per direction (H2D, D2H) and for both at once (H2D on one stream, D2H on another, the mix of the host-buffer call: 3 bytes
in for 1 byte out): every rank times `reps` back-to-back cudaMemcpyAsync of one buffer between two barriers; aggregate =
total bytes / max over ranks.  Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = mib << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n // 3, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n // 3, dtype=torch.uint8, device=dev)
h_in.fill_(1)
s2 = torch.cuda.Stream()
reps = 8


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed(fn):
    fn()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sync_all()
    return float(t.item()), ms


def both():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    d_in.copy_(h_in, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


res = {"n_gpus": world, "mib_per_copy": mib}
for name, fn, byts in (("h2d", lambda: d_in.copy_(h_in, non_blocking=True), n),
                       ("d2h", lambda: h_out.copy_(d_out, non_blocking=True), n // 3),
                       ("h2d_and_d2h", both, n + n // 3)):
    mx, mine = timed(fn)
    per = [None] * world
    if world > 1:
        dist.all_gather_object(per, byts / mine / 1e6)
    else:
        per = [byts / mine / 1e6]
    res[name] = {"aggregate_GBps": world * byts / mx / 1e6, "per_gpu_GBps": [round(x, 2) for x in per]}
if rank == 0:
    try:
        node = open(f"/sys/bus/pci/devices/{torch.cuda.get_device_properties(0).pci_bus_id:02x}/numa_node").read().strip()
    except Exception:  # noqa: BLE001
        node = None
    res["host_cpus"] = os.cpu_count()
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
