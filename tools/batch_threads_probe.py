"""Debug aid (GPU box): throughput mode with T host threads each driving a batch of S/T sequences (b2lo_odom_process_batch_dev)."""
import os, sys, threading, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lidar_odometry_b200 import api, synth
K, W = 60, 5
scans, _ = synth.kitti_sequence(n_scans=K + W + 1, seed=42, device="cuda")
dev = [torch.from_numpy(np.ascontiguousarray(s)).cuda() for s in scans]
def args(i, S): return [dev[i].data_ptr()] * S, [scans[i].shape[0]] * S, 4
for T, S in ((1, 32), (2, 32), (4, 32), (2, 64), (4, 64), (4, 128), (8, 128)):
    per = S // T
    bats = [api.OdometryBatch(per, 0) for _ in range(T)]
    bar = threading.Barrier(T + 1)
    def work(b):
        for i in range(W):
            p, n, sf = args(i, per); pn, nn, _ = args(i + 1, per); b.process_dev(p, n, sf, pn, nn)
        bar.wait()
        for i in range(W, W + K):
            p, n, sf = args(i, per); pn, nn, _ = args(i + 1, per); b.process_dev(p, n, sf, pn, nn)
        bar.wait()
    th = [threading.Thread(target=work, args=(b,)) for b in bats]
    [t.start() for t in th]
    bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
    [t.join() for t in th]
    print(f"threads {T} x {per} sequences = {S}: {S * K / dt:8.0f} scans/s", flush=True)
    del bats
