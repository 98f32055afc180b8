"""Map relocalisation (BASELINE.json configs[4]) on its own; bench.py runs the same leg (vo_b200/reloc.py).

  python tools/reloc_bench.py --queries-per-rank 131072 --landmarks 1048576            # one rank's share, 1 GPU
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/reloc_bench.py --queries-per-rank 131072 --landmarks 1048576               # 2 ranks

Prints one JSON line (rank 0)."""
import argparse, json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries-per-rank", type=int, default=131072)
    ap.add_argument("--landmarks", type=int, default=1048576)
    ap.add_argument("--hyps", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import vo_b200
    from vo_b200 import reloc
    ctx = vo_b200.Context(local)
    out = reloc.run(ctx, rank, world, dist, a.queries_per_rank, a.landmarks, a.hyps, a.reps)
    if out is not None:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
