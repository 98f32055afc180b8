"""Match GEMM sweep alone (bench.py's match_gemm leg) for quick iteration on the GPU box."""
import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import vo_b200
import bench

sizes = tuple(int(x) for x in sys.argv[1:]) or (8192, 32768, 65536)
ctx = vo_b200.Context(0)
r = bench.match_gemm_leg(ctx, torch, bench.peaks(), sizes=sizes)
for d in r["sweep"]:
    print(d["n1"], "match: %.3f ms %.0f TF (%.2f of bf16 burst), pairs %d | top2: %.3f ms %.0f TF" % (
        d["match"]["kernel_ms"], d["match"]["tflops"], d["match"]["frac_of_burst_peak"], d["match"]["pairs"],
        d["top2"]["kernel_ms"], d["top2"]["tflops"]), "| call ms", round(d["match"]["call_ms"], 3), round(d["top2"]["call_ms"], 3))
print(json.dumps(ctx.match_stats()))
