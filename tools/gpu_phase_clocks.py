"""Debug aid (GPU box): SM-clock cost of the single-CTA phases of the ICP kernels (PKO fit, Gauss-Newton finish)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_odometry_b200 import api, capi, synth
scans, poses = synth.kitti_sequence(n_scans=6, seed=7, n_rings=64, n_az=1200)
odo = api.Odometry()
for k, s in enumerate(scans):
    r = odo.process(s)
    out = (C.c_longlong * 32)()
    capi.lib().b2lo_ctx_debug_clocks(odo.ctx.h, out)
    v = list(out)
    print(k, r["n_iters"], r["n_corr"], "pko1[scan,scale,sample,kmeans,em,em_it,km_it] a:", v[0:7], "b:", v[8:15], "gn[partial_sum,finish]:", v[16:18],   f"dev_ms {r['device_ms']:.3f}")
