"""stdin: one bench.py JSON line -> 'label ms_per_step kernel4_ms sm_mhz' (A/B runs inside one gpurun call)."""
import json
import sys

d = json.loads(sys.stdin.read())
print(sys.argv[1] if len(sys.argv) > 1 else "-", round(d["ms_per_step"], 3),
      round(d["stages_ms"]["sparse_attention"], 3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
