"""Turn the raw captures in gpurun_out/ into the tracked evidence under profiles/ (run here, no GPU).

Inputs (written by the gpurun command lines quoted in profiles/README.md):
  gpurun_out/r1_launches_traffic.csv   ncu launch list: time + DRAM bytes of every launch of tools/prof_targets.py 8
  gpurun_out/r1_sift2.ncu-rep          ncu --set full, SIFT / geometry kernels of one frame-loop step
  gpurun_out/r1_match_u8.ncu-rep       ncu --set full, match_topk_u8_kernel at 32768 x 32768 (matchFeatures mode)
  gpurun_out/r1_match_frames.ncu-rep   ncu --set full, match_topk_u8_kernel inside the frame loop
"""
import csv, io, json, os, re, shutil, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
IMAGES = 36.0   # tools/prof_targets.py 8: 2 steps x 9 stereo frames
RP = os.environ.get("VO_ROUND", "r2")          # file prefix of the round being written
RN = RP[1:]

STAGE_OF = [("sift_descriptor_kernel", "sift_descriptor"), ("sift_trig_kernel", "sift_descriptor"),
            ("sift_blur_tma_kernel", "sift_blur_tma_kernel"), ("sift_refine_kernel", "sift_refine_orient"), ("sift_orient_kernel", "sift_refine_orient"),
            ("sift_extrema_kernel", "sift_extrema"), ("sift_base_stream_kernel", "sift_base_upsample_blur"),
            ("sift_small_octaves_kernel", "sift_blur_dog_small"), ("sift_downsample_kernel", "sift_downsample"),
            ("sift_rank_bucket_kernel", "sift_sort_dedupe"), ("sift_dedupe_kernel", "sift_sort_dedupe")]


def run(*a):
    return subprocess.run(list(a), capture_output=True, text=True).stdout


def launches():
    src = os.path.join(G, RP + "_launches_traffic.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(P, RP + "_launches_traffic.csv"))
    md = run(sys.executable, os.path.join(R, "tools", "launches_summary.py"), src)
    per_kernel = json.loads(run(sys.executable, os.path.join(R, "tools", "launches_summary.py"), src, str(IMAGES), "--json"))
    stage = {}
    for k, v in per_kernel.items():
        for pat, st in STAGE_OF:
            if k.startswith(pat):
                stage[st] = stage.get(st, 0.0) + v
    json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                         "python tools/prof_targets.py 8 (2 steps x 9 stereo frames = 36 images of 1241x376), B200, round " + RN + "; "
                         "dram__bytes_read.sum + dram__bytes_write.sum per image, launches grouped by bench.py stage",
               "dram_bytes_per_image": stage}, open(os.path.join(P, RP + "_traffic.json"), "w"), indent=1)
    open(os.path.join(P, RP + "_launches_summary.md"), "w").write(
        "# Round " + RN + " -- ncu launch list of `python tools/prof_targets.py 8`\n\n"
        "2 steps of `vo_frames_dev` on 9 stereo frames (18 images of 1241x376 per step) followed by two 32768 x 32768 x 128 "
        "exact top-2 matches.  `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`. "
        "Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event numbers.\n\n" + md)


def full(rep, out, title, cmd):
    src = os.path.join(G, rep)
    if not os.path.exists(src):
        return
    raw = run("ncu", "-i", src, "--page", "raw", "--csv")
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    cols = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
            ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
            ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"),
            ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"), ("smsp__inst_executed.sum", "warp inst")]
    stall = [h for h in H if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    lines = [f"# {title}\n", f"Command: `{cmd}` (B200, `--clock-control none`).\n",
             "| kernel | " + " | ".join(n for _, n in cols) + " | top stalls (warps per issue) |", "|---|" + "---|" * (len(cols) + 1)]
    ki = H.index("Kernel Name")
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("vo::", "")
        vals = []
        for k, _ in cols:
            if k in H:
                v, u = r[H.index(k)], U[H.index(k)]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.3g}" if "%" in u or f < 1e4 else f"{f:.4g}"
                except ValueError:
                    pass
                vals.append(f"{v} {u if u not in ('%', '') and 'register' not in u else ''}".strip())
            else:
                vals.append("-")
        st = []
        for h in stall:
            try:
                st.append((float(r[H.index(h)].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
        top = ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:4] if n not in ("selected",))
        lines.append(f"| `{name}` | " + " | ".join(vals) + f" | {top} |")
    open(os.path.join(P, out), "w").write("\n".join(lines) + "\n")


def opmix(rep, rx, out, which="0"):
    src = os.path.join(G, rep)
    if os.path.exists(src):
        open(os.path.join(P, out), "w").write(run(sys.executable, os.path.join(R, "tools", "ncu_opmix.py"), src, rx, "24", which))


def bench_launches():
    """ncu launch list of bench.py itself, compared with the live stage shares of the same command run plain."""
    src, plain = os.path.join(G, RP + "_launches_bench.csv"), os.path.join(G, "bench_ll_plain.json")
    if not (os.path.exists(src) and os.path.exists(plain)):
        return
    shutil.copy(src, os.path.join(P, RP + "_launches_bench.csv"))
    md = run(sys.executable, os.path.join(R, "tools", "launches_summary.py"), src)
    live = json.load(open(plain))
    shares = "\n".join(f"| `{k}` | {100 * v:.1f} % |" for k, v in live["stage_share"].items())
    open(os.path.join(P, RP + "_launches_bench.md"), "w").write(
        "# Round " + RN + " -- ncu launch list of `bench.py` itself\n\n"
        "`python bench.py --steps 3 --warmup 3 --no-cpu --quick` (exit 0 plain, then the same command under "
        "`ncu --metrics gpu__time_duration.sum --clock-control none --csv`): every kernel launch of the warm-up, the "
        "device-resident pass, the host-buffer (e2e) pass and the profiled serial pass -- 30 frame-loop steps of 33 stereo frames.  "
        "Times under ncu are cold-cache and serialised; the SHARES are what to compare with the live CUDA-event shares below.\n\n"
        + md + "\n## Live stage shares of the same command run without ncu (`stage_share` of its JSON line)\n\n"
        "| stage | share of the step |\n|---|---|\n" + shares + "\n\n"
        f"Live: {live['value']:.0f} frames/s ({live['ms_per_step']:.2f} ms per step with {live['run']['batches_in_flight']} batches in flight, "
        f"{live['ms_per_step_profiled_serial']:.2f} ms per step in the serial profiled pass).\n")


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    launches()
    bench_launches()
    full(RP + "_sift.ncu-rep", RP + "_ncu_full_sift.md", "Round " + RN + " -- ncu `--set full`: SIFT, sort, prep and geometry kernels of one frame-loop step (18 images of 1241x376)",
         'ncu --set full --clock-control none --import-source on -k regex:"sift_descriptor_kernel|sift_refine_kernel|sift_orient_kernel|sift_extrema|sift_blur_tma_kernel|sift_base_stream|sift_small_oct|sift_rank_bucket|landmark" -c 33 python tools/prof_targets.py 8')
    full(RP + "_match_u8.ncu-rep", RP + "_ncu_full_match.md", "Round " + RN + " -- ncu `--set full`: match_topk_u8_kernel, 32768 x 32768 x 128, matchFeatures mode",
         "ncu --set full --clock-control none --import-source on -k regex:match_topk_u8 -s 2 -c 1 python tools/prof_match.py 32768 match")
    full(RP + "_match_float.ncu-rep", RP + "_ncu_full_match_float.md", "Round " + RN + " -- ncu `--set full`: match_topk_kernel (general-float path, one bf16 term), 32768 x 32768 x 128 unit-norm float rows, matchFeatures mode",
         "ncu --set full --clock-control none --import-source on -k regex:match_topk_kernel -s 1 -c 1 python tools/prof_float.py 32768")
    opmix(RP + "_match_float.ncu-rep", "match_topk_kernel", RP + "_opmix_match_float.txt")
    opmix(RP + "_match_u8.ncu-rep", "match_topk_u8", RP + "_opmix_match_u8.txt")
    opmix(RP + "_sift.ncu-rep", "sift_descriptor", RP + "_opmix_descriptor.txt")
    opmix(RP + "_sift.ncu-rep", "sift_blur_tma", RP + "_opmix_blur_tma_r13.txt", "4")
    opmix(RP + "_sift.ncu-rep", "sift_extrema", RP + "_opmix_extrema.txt")
    print(sorted(os.listdir(P)))
