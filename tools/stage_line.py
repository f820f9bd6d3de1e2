"""stdin: one bench.py JSON line -> 'label ms_per_step {stage: ms}' (stage comparisons inside one gpurun call)."""
import json
import sys

d = json.loads(sys.stdin.read())
print(sys.argv[1] if len(sys.argv) > 1 else "-", round(d["ms_per_step"], 3),
      {k: round(v, 3) for k, v in d["stages_ms"].items()}, d["clocks"]["sm_mhz"])
