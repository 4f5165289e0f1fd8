"""Prints ms per step and the stage breakdown of a bench.py JSON line read from stdin (helper for sweeps on the GPU box)."""
import json
import sys

for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        print("%.4f ms  %s" % (d["ms_per_step"], {k: round(v, 3) for k, v in (d.get("stage_ms_per_step") or {}).items() if v}), flush=True)
