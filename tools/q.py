"""One-line summary of a bench.py JSON line: value, e2e, ms per step and the per-stage ms per step."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
st = {r["kernel"]: round(r["avg_launch_ms"] * r["launches"] / d["steps"], 3) for r in d["roofline_all"]}
print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3), d["gpu_launches"], st)
