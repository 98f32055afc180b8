"""Golden oracle trajectory on the rendered street sequence (run here once, committed):

    python tests/golden/make_street_golden.py        ->  tests/golden/street_oracle_225.npz

Renders frames 0..224 of synth.StreetWorld(seed 7) along the reference's ground truth kitti/poses/00.txt (held in
kitti00_reference_data.npz), runs the CPU oracle over them on every host core (bench.oracle_sequence: the VO.m
loop with every toolbox call answered by oracle/*.c, MSAC seed 1) and stores the relative poses, the status
codes, the tracked counts and a SHA-256 of the rendered frames.  tests/test_vo_gpu.py compares the CUDA path's
KITTI t_err / r_err on the same >= 200 frames with these (north star: within 2 %), without re-running the oracle
(about 12 core-minutes).  If the renderer ever produces different pixels on another machine the hash differs and
the test re-runs the oracle instead of trusting the file.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N = 225


def frames_hash(left, right):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(left).tobytes()); h.update(np.ascontiguousarray(right).tobytes())
    return h.hexdigest()


def main():
    import bench
    left, right, gt = bench.street_frames(N)
    rel, status, dt, tracked = bench.oracle_sequence(left, right, N)
    out = os.path.join(HERE, "street_oracle_225.npz")
    np.savez_compressed(out, rel=rel, status=status, tracked=np.asarray(tracked), sha256=frames_hash(left, right),
                        seed=bench.SEQ_SEED, msac_seed=1)
    print(f"oracle over {N} frames in {dt:.0f} s; status ok {(status[1:] == 0).sum()}/{N - 1}; "
          f"tracked median {np.median(tracked):.0f}; errors {bench.trajectory_errors(rel, gt, N)}")
    if bench._POOL is not None:
        bench._POOL.terminate()


if __name__ == "__main__":
    main()
