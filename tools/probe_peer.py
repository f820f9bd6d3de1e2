"""2-GPU probe (torchrun): CUDA IPC peer mapping through librsa_b200 -- each rank writes into the other's buffer."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "rectified-spaattn_b200")]
from rsa_b200 import native as N  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
L = N.lib()
ptr = C.c_void_p()
N.check(L.rsa_peer_alloc(1 << 20, C.byref(ptr)), "alloc")
h = (C.c_char * 64)()
N.check(L.rsa_peer_export(ptr, h), "export")
handles = [None] * world
dist.all_gather_object(handles, bytes(h))
peers = []
for r in range(world):
    if r == rank:
        peers.append(ptr.value)
        continue
    pp = C.c_void_p()
    N.check(L.rsa_peer_open(C.create_string_buffer(handles[r], 64), C.byref(pp)), "open")
    peers.append(pp.value)


class Raw:
    def __init__(self, p, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (p, False), "version": 3}


views = [torch.as_tensor(Raw(p, 1024), device="cuda") for p in peers]
views[(rank + 1) % world].fill_(float(rank + 1))     # write into the neighbour's buffer
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
print(rank, "my buffer now holds", views[rank][:3].tolist(), "expected", float((rank - 1) % world + 1), flush=True)
dist.destroy_process_group()
