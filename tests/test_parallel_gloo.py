"""CPU tests of the N > 1 host logic (rsa_b200/parallel.py) on the gloo backend, world_size 2: the head-shard
arithmetic and the Ulysses all-to-all pair around a stand-in attention function (dense SDPA on CPU; the product
kernels need a GPU and are covered by the -m gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rsa_b200 import parallel as P


def test_shard_heads_partitions_contiguously():
    for heads in (12, 24, 40, 7):
        for world in (1, 2, 4, 8):
            spans = [P.shard_heads(heads, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == heads
            for (f0, n0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + n0 == f1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    with pytest.raises(ValueError):
        P.shard_heads(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dense(q, k, v):
    # [B, h, S, D] -> [B, S, h, D]  (the shape_xfuse=True layout of the reference entry points)
    return torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).contiguous()


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        b, h, s, d = 1, 4, 64, 16
        q, k, v = (torch.randn(b, h, s, d, generator=g) for _ in range(3))
        s_loc = s // world
        sl = slice(rank * s_loc, (rank + 1) * s_loc)
        # 1. the first exchange: sequence shards -> head shards holding the whole sequence
        ql = P.seq_to_head_shard(q[:, :, sl].contiguous())
        h0, hn = P.shard_heads(h, world, rank)
        assert torch.equal(ql, q[:, h0:h0 + hn])
        # 2. the second exchange is its inverse on the output layout
        o_full = _dense(q, k, v)                                 # [B, S, H, D]
        back = P.head_to_seq_shard(o_full[:, :, h0:h0 + hn].contiguous())
        assert torch.equal(back, o_full[:, sl])
        # 3. the whole wrapper equals the single-process result on this rank's token slice
        out = P.ulysses_attention(q[:, :, sl].contiguous(), k[:, :, sl].contiguous(), v[:, :, sl].contiguous(), _dense)
        ref = o_full[:, sl].reshape(b, s_loc, h * d)
        assert out.shape == ref.shape and torch.allclose(out, ref, atol=1e-6)
        # 4. head-parallel needs no exchange at all: concatenating the ranks' head slices is the full result
        mine = _dense(q[:, h0:h0 + hn], k[:, h0:h0 + hn], v[:, h0:h0 + hn])
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        assert torch.allclose(torch.cat(parts, dim=2), o_full, atol=1e-6)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_ulysses_exchange_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_bench_cpulist_parser_and_unbound_fallback():
    """bench.py binds the ranks of a multi-rank run to their GPU's NUMA node; the sysfs cpulist parser, and the fallback:
    without a CUDA device (this container) the binding returns None and leaves the affinity alone."""
    import os
    import sys
    argv, sys.argv = sys.argv, [sys.argv[0]]
    try:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
    finally:
        sys.argv = argv
    assert bench.parse_cpulist("0-7\n") == set(range(8))
    assert bench.parse_cpulist("0-3,8-11,40") == {0, 1, 2, 3, 8, 9, 10, 11, 40}
    assert bench.parse_cpulist("5") == {5} and bench.parse_cpulist("") == set()
    import torch
    if not torch.cuda.is_available():
        before = os.sched_getaffinity(0)
        assert bench.bind_to_gpu_numa_node(0) is None and os.sched_getaffinity(0) == before
