"""Map relocalisation (BASELINE.json configs[4]): N1 query x N2 landmark SIFT-128 descriptors, queries
row-sharded over the ranks, landmarks replicated, one NCCL all-gather of 16-byte best-2 records, then
P3P-MSAC with exactly `--hyps` hypotheses on the surviving matches (rank 0).

  python tools/reloc_bench.py --queries 131072 --landmarks 1048576            # one rank's share, 1 GPU
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/reloc_bench.py --queries 262144 --landmarks 1048576                # 2 ranks

Descriptors are synthetic (OpenCV-SIFT statistics, generated on the device); half of the queries are
noisy copies of landmark rows.  Matched landmarks carry synthetic 3-D positions seen by a camera with
KITTI intrinsics, so the recovered pose is checked against the known one.  Prints one JSON line."""
import argparse, json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import torch


def sift_like(n, gen, device):
    g = torch.randn((n, 128), generator=gen, device=device).abs_().pow_(1.5)
    g *= 512.0 / g.norm(dim=1, keepdim=True)
    g.clamp_(max=0.2 * 512.0)
    g *= 512.0 / g.norm(dim=1, keepdim=True)
    return g.round_().clamp_(0, 255)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=131072)
    ap.add_argument("--landmarks", type=int, default=1048576)
    ap.add_argument("--hyps", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import vo_b200
    from vo_b200 import shard, synth
    ctx = vo_b200.Context(local)
    n1, n2 = a.queries, a.landmarks
    # replicated landmarks (same seed on every rank), this rank's query slice
    gl = torch.Generator(device=dev); gl.manual_seed(5678)
    land = sift_like(n2, gl, dev)
    lo, hi = shard.row_chunks(n1, world)[rank]
    gq = torch.Generator(device=dev); gq.manual_seed(1234)
    src_all = torch.randint(0, n2, (n1,), generator=gq, device=dev)          # landmark each query row imitates
    copy_all = torch.rand((n1,), generator=gq, device=dev) < 0.5             # ... for half of the rows
    gn = torch.Generator(device=dev); gn.manual_seed(99 + rank)
    q = sift_like(hi - lo, gn, dev)
    sel = copy_all[lo:hi]
    noisy = (land[src_all[lo:hi]] + 6.0 * torch.randn((hi - lo, 128), generator=gn, device=dev)).round_().clamp_(0, 255)
    q[sel] = noisy[sel]
    del noisy
    torch.cuda.synchronize()

    def one():
        return shard.relocalise_row_sharded_dev(ctx, q, land, rank, world, dist)

    for _ in range(2):
        rec, counts = one()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        rec, counts = one()
    e1.record()
    torch.cuda.synchronize()
    prof = ctx.profile(); ctx.profile_enable(False)
    ms = torch.tensor([e0.elapsed_time(e1) / a.reps, prof["match_gemm_topk"]["ms"] / prof["match_gemm_topk"]["launches"]], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_call, ms_gemm = float(ms[0]), float(ms[1])
    out = None
    if rank == 0:
        r = rec.cpu().numpy()
        keep = r[:, 3] == 1
        j1 = r[:, 0].view(np.uint32)
        truth = src_all.cpu().numpy()
        cp = copy_all.cpu().numpy()
        correct = int((j1[keep] == truth[keep]).sum())
        # synthetic geometry: landmark j sits at a fixed 3-D point; the query camera sees it through KITTI intrinsics
        rng = np.random.default_rng(7)
        K = (718.856, 718.856, 607.1928, 185.2157)
        ang = 0.05
        Rwc = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        twc = np.array([0.4, -0.1, 1.2])
        idx = np.nonzero(keep)[0]
        if len(idx) > 20000:
            idx = idx[rng.permutation(len(idx))[:20000]]
        cam = np.column_stack([rng.uniform(-8, 8, len(idx)), rng.uniform(-2, 2, len(idx)), rng.uniform(6, 40, len(idx))])
        world_pts = cam @ Rwc.T + twc                                   # X_w = R_wc X_c + t_wc
        img = np.column_stack([K[0] * cam[:, 0] / cam[:, 2] + K[2], K[1] * cam[:, 1] / cam[:, 2] + K[3]]) + rng.normal(0, 0.3, (len(idx), 2))
        wrong = j1[idx] != truth[idx]                                   # false matches become outliers
        img[wrong] = rng.uniform(0, 1200, (int(wrong.sum()), 2))
        t0 = time.time()
        res = vo_b200.estworldpose(img, world_pts, K, MaxNumTrials=a.hyps, Adaptive=False, Seed=3, full=True, ctx=ctx)
        A, inl, status = np.asarray(res["A"]), res["inliers"], res["status"]
        t_p3p = time.time() - t0
        err_t = float(np.linalg.norm(A[:3, 3] - twc)); err_R = float(np.linalg.norm(A[:3, :3] - Rwc))
        ops = 2.0 * n1 * n2 * 128
        out = dict(workload="map relocalisation (BASELINE configs[4])", n_gpus=world, queries=n1, landmarks=n2,
                   queries_per_rank=counts, ms_per_call_max_over_ranks=ms_call, ms_match_gemm_max_over_ranks=ms_gemm,
                   aggregate_tops_call=ops / ms_call / 1e9, aggregate_tops_gemm=ops / ms_gemm / 1e9,
                   allgather_bytes_per_rank=16 * max(counts), kept_rows=int(keep.sum()), kept_correct=correct,
                   copied_rows=int(cp.sum()), p3p=dict(points=int(len(idx)), hypotheses=a.hyps, status=int(status), inliers=int(np.sum(inl)),
                                                        ms_host_call=1e3 * t_p3p, pose_err_t=err_t, pose_err_R=err_R))
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
