import sys, json, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import bench, numpy as np
l, r, gt = bench.street_frames(385)
print(json.dumps(bench.png_leg(l, r, 32), indent=1))
