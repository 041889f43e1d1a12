"""Debug aid (GPU box): what the host-buffer call costs beyond its CUDA-event time, measured the way bench.py's e2e pass does
(L2 flush + synchronise outside the timer, page-locked scans, look-ahead)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidar_odometry_b200 import api, capi, synth
scans, poses = synth.kitti_sequence(n_scans=70, seed=42, device="cuda")
pin = [torch.from_numpy(np.ascontiguousarray(s)).pin_memory() for s in scans]
scans = [t.numpy() for t in pin]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
L = capi.lib()


def run(tag, call):
    odo = api.Odometry()
    for i in range(6):
        odo.process(scans[i], lookahead=scans[i + 1])
    out = (C.c_double * 8)(); L.b2lo_ctx_host_us(odo.ctx.h, out, 1)
    wall = dev = 0.0
    n = 0
    for i in range(6, 66):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        d = call(odo, i)
        wall += time.perf_counter() - t0
        dev += d; n += 1
    L.b2lo_ctx_host_us(odo.ctx.h, out, 1)
    print("%-28s wall %.1f us  device events %.1f us  gap %.1f us  host split [stage, enqueue, wait, absorb] %s" %
          (tag, 1e6 * wall / n, 1e3 * dev / n, 1e6 * wall / n - 1e3 * dev / n, [round(v / n, 1) for v in list(out)[:4]]), flush=True)


def via_api(odo, i):
    return odo.process(scans[i], lookahead=scans[i + 1])["device_ms"]


res = api.OdomResult()
ptr = [s.ctypes.data for s in scans]
npt = [s.shape[0] for s in scans]
sf = scans[0].strides[0] // 4
fn = L.b2lo_odom_process_la


def raw(odo, i):
    fn(odo.h, ptr[i], npt[i], sf, ptr[i + 1], npt[i + 1], sf, C.byref(res))
    return res.device_ms


def no_la(odo, i):
    return odo.process(scans[i])["device_ms"]


dev_scans = [torch.from_numpy(s).cuda() for s in scans]


def dev_no_la(odo, i):
    t = dev_scans[i]
    return odo.process_dev(t.data_ptr(), t.shape[0], t.stride(0))["device_ms"]


def dev_la(odo, i):
    t, u = dev_scans[i], dev_scans[i + 1]
    return odo.process_dev(t.data_ptr(), t.shape[0], t.stride(0), lookahead=(u.data_ptr(), u.shape[0], u.stride(0)))["device_ms"]


for rep in range(2):
    run("api.Odometry.process", via_api)
    run("ctypes b2lo_odom_process_la", raw)
    run("host scans, no look-ahead", no_la)
    run("device scans, no look-ahead", dev_no_la)
    run("device scans, look-ahead", dev_la)
# an empty ctypes call, for scale
t0 = time.perf_counter()
a, b, c = C.c_longlong(), C.c_longlong(), C.c_longlong()
odo = api.Odometry()
for _ in range(10000):
    L.b2lo_odom_graph_stats(odo.h, C.byref(a), C.byref(b), C.byref(c))
print("trivial ctypes call: %.2f us" % (1e6 * (time.perf_counter() - t0) / 10000))
