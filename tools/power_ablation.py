"""Where does kernel 4's power go?  Runs the debug instantiation of the tcgen05 attention kernel back to back for a
few seconds per setting on a C3b-like load (8 heads x 65 536 tokens, random block mask of density 0.22) while sampling
SM clock and board power with nvidia-smi:
    flags 0  the whole kernel
    flags 1  softmax arithmetic removed (S read, nothing computed, P handed back as is): TMA + tensor pipe + barriers
    flags 3  additionally no K/V TMA traffic after the first ring fill: tensor pipe + barriers alone
Results are garbage with flags != 0; only time, clock and power are read.  Prints one line per setting.
Run on the GPU box: python tools/power_ablation.py"""
import ctypes as C
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "rectified-spaattn_b200"))
from rsa_b200 import native, ops  # noqa: E402


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], False

    def run(self):
        while not self.stop:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                  "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout
            try:
                c, p, r = [x.strip() for x in out.strip().split(",")]
                self.rows.append((float(c), float(p), r))
            except Exception:
                pass
            time.sleep(0.05)


def main():
    dev = torch.device("cuda:0")
    L = native.lib()
    h, s, dens = 8, 65536, 0.22
    g = torch.Generator(device=dev).manual_seed(1)
    q, k, v = (torch.randn(1, h, s, 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(3))
    nb = s // 128
    mask = (torch.rand(1, h, nb, nb, generator=g, device=dev) < dens) | torch.eye(nb, dtype=torch.bool, device=dev)
    pairs = int(mask.sum())
    dbg = torch.zeros(34048, dtype=torch.float32, device=dev)
    for flags, name in ((None, "product kernel"), (0, "debug kernel, whole"), (1, "no softmax arithmetic"),
                        (3, "no softmax, no K/V traffic")):
        if flags is not None:
            L.rsa_debug_set_attention_dump(C.c_void_p(dbg.data_ptr()))
            L.rsa_debug_set_attention_flags(flags)
        for _ in range(3):
            ops.masked_attention(q, k, v, mask, s)
        torch.cuda.synchronize()
        sm = Sampler()
        sm.start()
        n = 150
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            ops.masked_attention(q, k, v, mask, s)
        e1.record()
        torch.cuda.synchronize()
        sm.stop = True
        sm.join()
        L.rsa_debug_set_attention_flags(0)
        L.rsa_debug_set_attention_dump(None)
        ms = e0.elapsed_time(e1) / n
        rows = sm.rows[len(sm.rows) // 3:]          # steady state: drop the first third of the samples
        clk = sorted(r[0] for r in rows)[len(rows) // 2] if rows else float("nan")
        pw = sorted(r[1] for r in rows)[len(rows) // 2] if rows else float("nan")
        cap = sum(r[2].lower().startswith("active") for r in rows) / max(1, len(rows))
        print(f"{name:28s} {ms:7.3f} ms/launch (incl. mask->lists)  {pairs * 8388608 / ms / 1e9:6.0f} TFLOP/s-equivalent  "
              f"SM clock {clk:.0f} MHz  board power {pw:.0f} W  power-cap active in {100 * cap:.0f} % of samples", flush=True)


if __name__ == "__main__":
    main()
