"""Multi-GPU check + measurement of the fused Ulysses exchange (run under torchrun on 2, 4 or 8 GPUs of one box):
  * rank r holds tokens [r*S/P, (r+1)*S/P) of the projection outputs; FusedUlysses.run() must return exactly the rows a
    single-GPU call over all heads produces (heads are independent, same kernels => bit-identical),
  * the NCCL form (two all-to-alls around kernel 0 + the pooled call, rsa_b200.parallel.ulysses_attention-style) gives
    the same bits and is timed beside it.
Prints one JSON line from rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "rectified-spaattn_b200")]
sys.argv, argv = [sys.argv[0]], sys.argv[1:]
import bench  # noqa: E402
from rsa_b200 import ops, parallel  # noqa: E402

name = argv[0] if argv else "c3a"
check = "--no-check" not in argv
rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
if name == "odd":
    # tokens AND heads that do not divide by the ranks (Wan form, 5 x 9 x 23 = 1035 tokens, 3 heads): the last rank owns
    # fewer rows, the first rank(s) compute one head more
    bench.WORKLOADS["odd"] = dict(desc="uneven token and head shards", fam="wan", grid=(5, 9, 23), text=0, text_valid=0,
                                  heads=3, drop=0.6, ffb=True)
wp = bench.workload_params(name)
heads, s, nv = wp["heads"], wp["s"], wp["nv"]
geo = bench.product_geometry(wp)
t, h, w = wp["grid"]
nbr = ops.gilbert_block_neighbors(t, h, w)
rows, hl = -(-s // world), heads // world          # rows per rank (the last rank may own fewer)
even = rows * world == s and hl * world == heads      # the NCCL form below is written for even shards only
g = torch.Generator(device=dev).manual_seed(1234)                       # same stream on every rank: the full tensors
full = [torch.randn(1, s, heads * 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(3)]
if wp["fam"] != "wan":                                                   # give the pooled scores some structure
    mu = torch.cumsum(torch.randn(1, (s + 127) // 128, heads * 128, generator=g, device=dev) * 0.35, dim=1)
    for x in full[:2]:
        x.add_(mu.repeat_interleave(128, dim=1)[:, :s].to(torch.bfloat16))
wan = wp["fam"] == "wan"     # Wan: RMSNorm over all heads*128 channels of a token, rotary embedding on every token
nw = heads * 128 if wan else 128
wq = (1 + 0.1 * torch.randn(nw, generator=g, device=dev)).to(torch.bfloat16)
wk = (1 + 0.1 * torch.randn(nw, generator=g, device=dev)).to(torch.bfloat16)
ang = torch.outer(torch.arange(nv, dtype=torch.float32, device=dev), 1.0 / (256.0 ** (torch.arange(0, 128, 2, device=dev) / 128)))
rope = (ang.cos().repeat_interleave(2, 1).contiguous(), ang.sin().repeat_interleave(2, 1).contiguous())
mine = slice(rank * rows, min((rank + 1) * rows, s))

fu = parallel.FusedUlysses(1, heads, geo, wp["top_k"], bench.P_REMAIN, nbr)
for dst, src in zip((fu.q_src, fu.k_src, fu.v_src), full):
    dst.copy_(src[:, mine])
out = fu.run(wq, wk, 1e-6, rope, nv).clone()
torch.cuda.synchronize()

# NCCL form: all-to-all the projection rows to head shards, kernel 0 + pooled call locally, all-to-all the result back
q, k, v = (torch.empty(1, max(hl, 1), s, 128, dtype=torch.bfloat16, device=dev) for _ in range(3))
plan = ops.Plan(q, k, v, geo, wp["top_k"], bench.P_REMAIN, nbr)


def wan_norm(x, w):
    """diffusers RMSNorm over the whole row (what the Wan processor does before the head split), in PyTorch."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    return (x * torch.rsqrt(var + 1e-6)).to(torch.bfloat16) * w


def nccl_form():
    srcs = []
    for i, x in enumerate(full):
        loc = x[:, mine]
        if wan and i < 2:      # the norm needs whole rows: before the exchange, on the tokens this rank owns
            loc = wan_norm(loc, wq if i == 0 else wk)
        loc = loc.reshape(1, rows, world, hl * 128).permute(2, 0, 1, 3).contiguous()   # [P(dst), 1, rows, hl*128]
        rcv = torch.empty_like(loc)
        dist.all_to_all_single(rcv, loc)
        srcs.append(rcv.permute(1, 0, 2, 3).reshape(1, s, hl * 128))                         # all tokens, my heads
    if wan:
        plan.qkv_prep(*srcs, dst_row=0, rope=rope, rope_rows=nv)
    elif geo.gap:
        plan.qkv_prep(*(x[:, :nv] for x in srcs), dst_row=0, q_weight=wq, k_weight=wk, eps=1e-6, rope=rope)
        plan.qkv_prep(*(x[:, nv:] for x in srcs), dst_row=nv, q_weight=wq, k_weight=wk, eps=1e-6)
    else:
        plan.qkv_prep(*srcs, dst_row=0, q_weight=wq, k_weight=wk, eps=1e-6, rope=rope, rope_rows=nv)
    o = plan.run_pooled()                                                                     # [1, S, hl, 128]
    return parallel.head_to_seq_shard(o).reshape(1, rows, heads * 128)


# (the NCCL form below is written for even shards; with uneven ones the single-GPU call is the only reference)
ref = nccl_form().clone() if even else out
torch.cuda.synchronize()
# Wan: the NCCL form's norm is PyTorch's (other reduction order: a bf16 rounding may flip, and with it a block selection),
# so it is a timing reference there, not a bit reference -- the bit reference is the single-GPU call below
same_as_nccl = bool(torch.equal(out.view(torch.int16), ref.view(torch.int16)))
frac_equal_nccl = float((out.view(torch.int16) == ref.view(torch.int16)).float().mean())
same_as_single = None
if check and (rank == 0 or not even):       # uneven shards: every rank checks its own rows (the shapes are small)
    # one GPU, all heads, no exchange at all
    qa, ka, va = (torch.empty(1, heads, s, 128, dtype=torch.bfloat16, device=dev) for _ in range(3))
    pa = ops.Plan(qa, ka, va, geo, wp["top_k"], bench.P_REMAIN, nbr)
    if geo.gap:
        pa.qkv_prep(*(x[:, :nv] for x in full), dst_row=0, q_weight=wq, k_weight=wk, eps=1e-6, rope=rope)
        pa.qkv_prep(*(x[:, nv:] for x in full), dst_row=nv, q_weight=wq, k_weight=wk, eps=1e-6)
    else:
        pa.qkv_prep(*full, dst_row=0, q_weight=wq, k_weight=wk, eps=1e-6, rope=rope, rope_rows=nv)
    single = pa.run_pooled().reshape(1, s, heads * 128)[:, mine]
    torch.cuda.synchronize()
    same_as_single = bool(torch.equal(out.view(torch.int16), single.view(torch.int16)))
    del qa, ka, va, pa


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item())


ms_fused = timed(lambda: fu.run(wq, wk, 1e-6, rope, nv))
ms_nccl = timed(nccl_form) if even else None
oks = [None] * world
dist.all_gather_object(oks, same_as_nccl)
singles = [None] * world
dist.all_gather_object(singles, same_as_single)
if rank == 0:
    print(json.dumps({"workload": name, "n_gpus": world, "tokens": s, "heads": heads,
                      "fused_ms_per_layer_attention": ms_fused, "nccl_all_to_all_form_ms": ms_nccl,
                      "norm": "across heads (Wan)" if wan else "per head",
                      "fused_equals_nccl_form_bitwise_all_ranks": all(oks),
                      "fraction_of_elements_equal_to_nccl_form_rank0": frac_equal_nccl,
                      "fused_equals_single_gpu_bitwise_rank0": same_as_single,
                      "fused_equals_single_gpu_bitwise_checked_ranks": [x for x in singles if x is not None],
                      "rows_per_rank": rows, "rows_last_rank": s - (world - 1) * rows, "heads_rank0": fu.heads,
                      "note": "both forms run kernel 0 + kernels 3a-4; fused = gather inside kernel 0 and scatter inside "
                              "kernel 4's epilogue over peer memory, barriers only; nccl = all_to_all_single x4 + staging copies"}))
fu.close()
dist.destroy_process_group()
