"""Prints, per case, how the CUDA path compares with tests/golden/golden_gpu_<case>.npz (the unmodified reference run in
bf16 on a B200): mask agreement, and output error on query blocks whose mask row agrees.  Used to set the bounds in
tests/test_gpu_parity.py::test_against_unmodified_reference_on_b200."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "rectified-spaattn_b200"), os.path.join(REPO, "tests")]
from helpers import cos_sim, load_case, product_geometry  # noqa: E402
from oracle import cases as C  # noqa: E402
from rsa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for name in C.CASES:
    if name in C.REFERENCE_NEEDS_PADDED_LAYOUT:
        continue
    case = load_case(name)
    g = np.load(os.path.join(REPO, "tests", "golden", f"golden_gpu_{name}.npz"))
    mask_ref = np.unpackbits(g["mask"])[: int(np.prod(g["mask_shape"]))].reshape(g["mask_shape"]).astype(bool)
    q, k, v = (torch.from_numpy(case[n]).to(dev).to(torch.bfloat16) for n in ("q", "k", "v"))
    geo = product_geometry(case["fam"], case["nv"], case["s"], case["text_len"], case["ntrue_d"], case["grid"][0])
    plan = ops.Plan(q, k, v, geo, case["top_k"], case["p"], torch.from_numpy(case["nbr"]))
    out = plan.run().float().cpu().numpy()[0]
    ref = g["out"].astype(np.float32).reshape(out.shape)
    og = case["ogeo"]
    nq = og.nq_blocks
    mask = plan.dense_mask().cpu().numpy()[:, :nq]
    agree = mask == mask_ref
    rows_ok = agree.all(axis=2)
    rows = min(nq * 128, og.seq)
    keep = np.repeat(rows_ok, 128, axis=1)[:, :rows].T
    d = np.abs(out[:rows] - ref[:rows])
    dk = d[keep]
    oracle = None
    from oracle import rsa_oracle as O
    orc = O.forward(case["q"], case["k"], case["v"], og, case["nbr"]).reshape(1, og.seq, case["heads"], 128)[0]
    do = np.abs(orc[:rows] - ref[:rows])[keep]
    print(f"{name:18s} mask agree {agree.mean():.4f} rows ok {rows_ok.mean():.3f} | ours-vs-ref on ok rows: max {dk.max():.4f} "
          f"p99.9 {np.quantile(dk, 0.999):.4f} mean {dk.mean():.5f} frac<=2e-2 {np.mean(dk <= 2e-2):.5f} "
          f"cos {cos_sim(out[:rows][keep], ref[:rows][keep]):.5f} | oracle-vs-ref max {do.max():.4f} cos "
          f"{cos_sim(orc[:rows][keep], ref[:rows][keep]):.5f} | all rows cos {cos_sim(out[:rows], ref[:rows]):.5f}", flush=True)
    t0, nt = nq * 128, og.text_q_valid
    if nt:
        dt = np.abs(out[t0:t0 + nt] - ref[t0:t0 + nt])
        print(f"{'':18s} text rows max {dt.max():.4f} cos {cos_sim(out[t0:t0 + nt], ref[t0:t0 + nt]):.5f}")
