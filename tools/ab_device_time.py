"""Debug aid (GPU box): CUDA-event time per scan of the device-resident look-ahead path (bench.py's `value` pass), for A/B runs
of two library builds in one call: B2LO_LIB=<other build> python tools/ab_device_time.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidar_odometry_b200 import api, synth
scans, _ = synth.kitti_sequence(n_scans=106, seed=42, device="cuda")
dev = [torch.from_numpy(np.ascontiguousarray(s)).cuda() for s in scans]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rep in range(3):
    odo = api.Odometry()
    ms = []
    for i in range(105):
        t, u = dev[i], dev[i + 1]
        flush.zero_(); torch.cuda.synchronize()
        r = odo.process_dev(t.data_ptr(), t.shape[0], t.stride(0), lookahead=(u.data_ptr(), u.shape[0], u.stride(0)))
        if i >= 5:
            ms.append(r["device_ms"])
    print("%s: %.2f us/scan (median %.2f)" % (os.environ.get("B2LO_LIB", "in-tree"), 1e3 * float(np.mean(ms)), 1e3 * float(np.median(ms))), flush=True)
