#!/bin/bash
# GPU box: same-box A/B of two library builds (gpurun_variants/libb2lo_old.so vs libb2lo_new.so): lone-sequence device time and
# lock-step throughput (one batch of 128, three batches of 96).  Usage: bash tools/ab_lockstep.sh
for v in old new old new; do
  echo "== $v"
  B2LO_LIB=gpurun_variants/libb2lo_$v.so timeout 300 python tools/ab_device_time.py 2>&1 | tail -1
  B2LO_LIB=gpurun_variants/libb2lo_$v.so K=40 timeout 300 python - <<'PY' 2>&1 | tail -3
import json, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import numpy as np, torch, bench
from lidar_odometry_b200 import api
K, W = 40, 5
scans, _ = bench.make_scans(K + W + 1, 42, "cuda:0")
dev = [torch.from_numpy(s).cuda() for s in scans]
def dev_args(i):
    return dev[i].data_ptr(), scans[i].shape[0], scans[i].shape[1]
r = bench.lockstep_leg(api, 0, dev_args, 128, K, W, 0, "")
print("lockstep 128:", round(r["scans_per_s"]), "scans/s")
for G, S in [(3, 96), (4, 96)]:
    r = bench.lockstep_groups_leg(api, 0, dev_args, G, S, K, W)
    print(f"groups {G}x{S}:", round(r["scans_per_s"]), "scans/s")
PY
done
