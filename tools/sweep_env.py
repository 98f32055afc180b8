"""A/B harness for kernel variants selected by an environment variable: runs a short bench.py pass per value and
prints the per-stage times of the serial profiled pass next to the headline value.  One gpurun call measures
several variants in about ten seconds each:

    gpurun -- 'python tools/sweep_env.py VO_DESC_VAR 0 1 2 3 --stage sift_descriptor'

(The variants themselves are temporary template instantiations picked with getenv() at the launch site; the
round-1 descriptor, blur and extrema experiments in DESIGN.md section 8 were measured this way.)"""
import argparse
import json
import os
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("var")
    ap.add_argument("values", nargs="+")
    ap.add_argument("--stage", action="append", default=[], help="stage name(s) to print; default: the five largest")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    for v in a.values:
        env = dict(os.environ, **{a.var: v})
        out = subprocess.run([sys.executable, os.path.join(R, "bench.py"), "--no-cpu", "--quick", "--steps", str(a.steps),
                              "--warmup", "3", "--batch", str(a.batch)], env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(f"{a.var}={v}: bench.py failed\n{out.stderr[-800:]}")
            continue
        d = json.loads(out.stdout.strip().splitlines()[-1])
        st = {k["kernel"]: k["avg_launch_ms"] * k["launches"] / d["steps"] for k in d["roofline_all"]}
        names = a.stage or [k for k, _ in sorted(st.items(), key=lambda kv: -kv[1])[:5]]
        print(f"{a.var}={v}: {d['value']:.0f} frames/s, {d['ms_per_step']:.2f} ms/step (serial {d['ms_per_step_profiled_serial']:.2f}) | "
              + ", ".join(f"{n} {st.get(n, float('nan')):.3f} ms" for n in names), flush=True)


if __name__ == "__main__":
    main()
