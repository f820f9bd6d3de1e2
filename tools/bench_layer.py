"""Layer-level measurement (run on the GPU box): one HunyuanVideo dual-stream attention layer at the C3a shape
(115 200 video + 256 text tokens, dim 3072, 24 heads) through the diffusers processor protocol --
  ours, fused       Rectified...Processor2_0 with kernel 0 (default)
  ours, op by op    the same processor with fuse_prep = False (projections / norm / RoPE in PyTorch, kernels 2-4)
  reference         the UNMODIFIED reference processor from baseline/_ref (or /root/reference) on the same module and
                    inputs, if it can be loaded; diffusers is not installed, so `diffusers.models.embeddings.
                    apply_rotary_emb` is stubbed with its restated expression (rectified_spaattn/_processors.rope_real)
on a stand-in `Attention` module (tests/fake_attention.py; bf16 Linear projections + RMSNorm(128)).  One JSON line."""
import json
import os
import sys
import types

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "rectified-spaattn_b200"), os.path.join(REPO, "tests")]
sys.argv, argv = [sys.argv[0]], sys.argv[1:]
import bench  # noqa: E402
from fake_attention import FakeAttention  # noqa: E402
from rectified_spaattn import _processors as P  # noqa: E402
from rectified_spaattn import rectified_hunyuan_attn as hun  # noqa: E402
from rsa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
wp = bench.workload_params("c3a")
nv, n_txt, dim, heads = wp["nv"], wp["text"], 3072, wp["heads"]
torch.manual_seed(0)
attn = FakeAttention(dim, heads, added=True).to(dev).to(torch.bfloat16)
# hidden states with the block structure of SURVEY 8d's "walk" regime (neighbouring blocks similar, like video): the
# projections are linear, so Q and K inherit it; iid hidden states would make every pooled score a near-tie
mu = torch.cumsum(torch.randn(nv // 128, dim, device=dev) * 0.35, dim=0).repeat_interleave(128, dim=0)
x = (torch.randn(nv, dim, device=dev) + mu)[None].to(torch.bfloat16)
del mu
txt = torch.randn(1, n_txt, dim, device=dev).to(torch.bfloat16)
mask = (torch.arange(nv + n_txt, device=dev) < wp["num_true"]).view(1, 1, 1, -1)
ang = torch.outer(torch.arange(nv, dtype=torch.float32, device=dev), 1.0 / (256.0 ** (torch.arange(0, 128, 2, device=dev) / 128)))
rope = (ang.cos().repeat_interleave(2, 1).contiguous(), ang.sin().repeat_interleave(2, 1).contiguous())
t, h, w = wp["grid"]
nbr = ops.gilbert_block_neighbors(t, h, w)


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {"workload": "c3a HunyuanVideo dual-stream layer", "tokens": nv + n_txt, "dim": dim, "heads": heads}
outs = {}
with torch.no_grad():
    for key, fuse in (("ours_fused_ms", True), ("ours_op_by_op_ms", False)):
        pr = hun.RectifiedHunyuanVideoSpaAttnProcessor2_0("sparse", wp["top_k"], nbr, bench.P_REMAIN)
        pr.fuse_prep = fuse
        pr.num_true = wp["num_true"]                       # host int: no device read-back per call
        res[key] = timed(lambda: outs.__setitem__(key, pr(attn, x, txt, mask, rope)), 5)
    proj = timed(lambda: [f(x) for f in (attn.to_q, attn.to_k, attn.to_v)] + [attn.to_out[0](x)], 5)
    res["projections_alone_ms"] = proj
    try:
        from oracle import ref_loader
        if "diffusers.models.embeddings" not in sys.modules:
            ref_loader.load(["rectified_hunyuan_attn"])    # installs the diffusers stubs
            m = types.ModuleType("diffusers.models.embeddings")
            m.apply_rotary_emb = lambda t_, freqs, **kw: P.rope_real(t_, freqs)
            sys.modules["diffusers.models.embeddings"] = m
        ref = ref_loader.load(["rectified_hunyuan_attn"])["rectified_hunyuan_attn"]
        rp = ref.RectifiedHunyuanVideoSpaAttnProcessor2_0("sparse", wp["top_k"], nbr, bench.P_REMAIN, 0)
        res["reference_ms"] = timed(lambda: outs.__setitem__("ref", rp(attn, x, txt, mask, rope)), 3)
        a, b = outs["ours_fused_ms"][0].float(), outs["ref"][0].float()
        res["video_rows_cosine_vs_reference"] = float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0))
    except Exception as e:  # noqa: BLE001
        res["reference_ms"] = None
        res["reference_error"] = repr(e)[:300]
a, b = outs["ours_fused_ms"][0].float(), outs["ours_op_by_op_ms"][0].float()
res["fused_vs_op_by_op_cosine"] = float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0))
print(json.dumps(res))
