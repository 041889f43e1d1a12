"""Debug aid (GPU box): lock-step batch throughput for a few batch sizes (the bench's lockstep leg alone)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from lidar_odometry_b200 import api
K, W = int(os.environ.get("K", "60")), 5
scans, _ = bench.make_scans(K + W + 1, 42, "cuda:0")
dev = [torch.from_numpy(s).cuda() for s in scans]
def dev_args(i):
    return dev[i].data_ptr(), scans[i].shape[0], scans[i].shape[1]
for S in [int(x) for x in sys.argv[1:]]:
    print(json.dumps(bench.lockstep_leg(api, 0, dev_args, S, K, W, 0, "")), flush=True)

for G, S in [(3, 96), (4, 96), (6, 64), (8, 48), (6, 48), (8, 32)]:
    print(json.dumps(bench.lockstep_groups_leg(api, 0, dev_args, G, S, K, W)), flush=True)
