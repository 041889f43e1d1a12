"""Debug aid (GPU box): where the wall clock of the host-buffer odometry call goes."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_odometry_b200 import api, capi, synth
scans, poses = synth.kitti_sequence(n_scans=30, seed=42, device="cuda")
import torch
pin = [torch.from_numpy(np.ascontiguousarray(s)).pin_memory() for s in scans]
scans = [t.numpy() for t in pin] if len(sys.argv) > 1 else [np.ascontiguousarray(s) for s in scans]
odo = api.Odometry()
for s in scans[:5]: odo.process(s)
out = (C.c_double * 8)(); capi.lib().b2lo_ctx_host_us(odo.ctx.h, out, 1)
t0 = time.perf_counter(); dev = 0.0
for s in scans[5:]:
    r = odo.process(s); dev += r["device_ms"]
wall = time.perf_counter() - t0
capi.lib().b2lo_ctx_host_us(odo.ctx.h, out, 1)
n = len(scans) - 5
print(odo.graph_stats()); print("per scan: wall %.1f us, device events %.1f us; host split [gather+h2d, enqueue K1+ICP, wait pose, host algebra, K6 enqueue+wait]:" % (1e6 * wall / n, 1e3 * dev / n), [round(v / n, 1) for v in list(out)[:5]], "purge per keyframe: voxels %.1f parents %.1f" % (out[5] / max(out[7], 1), out[6] / max(out[7], 1)))
