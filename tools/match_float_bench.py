"""General-float match (matchFeatures mode, vo_match_dev) on unit-norm float descriptors: GEMM + top-k kernel time and
rows sent to the exact scan, for the one-term form (default with a score bound) and VO_MATCH_FLOAT_TERMS=3.
usage: python tools/match_float_bench.py [n ...]   (numbers printed here are not bench values)"""
import os, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
    import ctypes as C
    import numpy as np, torch
    import vo_b200
    from vo_b200 import _lib, synth
    ctx = vo_b200.Context(0); L = _lib.lib()
    for n in [int(x) for x in sys.argv[2:]]:
        a, b = synth.descriptor_sets("float", n, n, seed=1234)
        f1 = torch.from_numpy(a).cuda(); f2 = torch.from_numpy(b).cuda()
        i1 = torch.empty(n, dtype=torch.int32, device="cuda"); i2 = torch.empty_like(i1); mt = torch.empty(n, device="cuda")
        npairs = torch.zeros(1, dtype=torch.int32, device="cuda")
        def call():
            _lib.check(L.vo_match_dev(ctx.handle, C.c_void_p(f1.data_ptr()), n, C.c_void_p(f2.data_ptr()), n, 128, None,
                                      C.c_void_p(i1.data_ptr()), C.c_void_p(i2.data_ptr()), C.c_void_p(mt.data_ptr()),
                                      C.c_void_p(npairs.data_ptr()), C.c_void_p(ctx.stream)))
        for _ in range(3): call()
        ctx.sync(); ctx.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.ExternalStream(ctx.stream)
        e0.record(st)
        for _ in range(10): call()
        e1.record(st); ctx.sync()
        p = ctx.profile(); ctx.profile_enable(False)
        g = p["match_gemm_topk"]; ms = g["ms"] / g["launches"]
        # a vo_match (host) call of a slice reports the rows that needed the exact scan
        pairs = vo_b200.matchFeatures(a[:4096], b, ctx=ctx); stt = ctx.match_stats()
        print(f"terms={os.environ.get('VO_MATCH_FLOAT_TERMS', '1')} n={n}: kernel {ms:.3f} ms = {2.0 * n * n * 128 / ms / 1e9:.0f} TFLOP/s of algorithmic work, "
              f"call {e0.elapsed_time(e1) / 10:.3f} ms, pairs {int(npairs.item())}, k_extent {stt['k_extent']}, rowscan rows {stt['rowscan_rows']} of 4096", flush=True)
else:
    sizes = sys.argv[1:] or ["8192", "32768"]
    for terms in ("1", "3"):
        env = dict(os.environ); env["VO_MATCH_FLOAT_TERMS"] = terms
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"] + sizes, env=env, check=False)
