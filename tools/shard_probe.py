"""Debug aid (GPU box, one GPU): where a point-sharded Gauss-Newton iteration spends its time.  In-kernel timeline (instrumented build,
B2LO_LIB=lidar_odometry_b200/libb2lo_tl.so) of the fused single-GPU loop on the dense ~1.07 M-point scan, of the sharded loop (world 1:
no exchange, same kernels as every rank of a larger world runs) on the whole scan, and on half of it (what a rank of world 2 holds).

    B2LO_LIB=lidar_odometry_b200/libb2lo_tl.so python tools/shard_probe.py
"""
import collections, ctypes as C, os, re, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_odometry_b200 import api, capi
import bench

FILES = {1: "b2lo_filter.cu", 2: "b2lo_icp.cu", 3: "b2lo_odom.cu", 4: "b2lo_map.cu"}
src = {i: open(os.path.join(ROOT, "lidar_odometry_b200", "csrc", f)).read().split("\n") for i, f in FILES.items()}


def label(fid, line):
    text = src[fid][line - 1]
    if "TL_HERE" in text:
        return "  . " + text.split("//")[-1].strip()
    for k in range(line - 1, max(line - 12, -1), -1):
        m = re.search(r"\bstruct\s+(k_[a-z0-9_]+)", src[fid][k]) or re.search(r"\b(k_[a-z0-9_]+)\s*\(", src[fid][k])
        if m:
            return m.group(1)
    return f"{FILES[fid]}:{line}"


L = capi.lib()
L.b2lo_debug_timeline.restype = C.c_int
WORLD = int(os.environ.get("WORLD_SIZE", "1")); RANK = int(os.environ.get("RANK", "0"))
if WORLD > 1:      # under torchrun: the peer-memory exchange between the ranks, rank 0 prints
    import torch.distributed as dist
    torch.cuda.set_device(RANK)
    dist.init_process_group("nccl", device_id=torch.device("cuda", RANK))
scans, poses = bench.make_scans(11, 42, f"cuda:{RANK}")
odo = api.Odometry(api.Context(RANK))
for s in scans[:10]:
    r = odo.process(s)
guess = r["pose"]
rng = np.random.default_rng(99)
base = scans[10][:, :3]
dense = np.concatenate([base + rng.normal(0, 0.01, base.shape).astype(np.float32) for _ in range(9)]).astype(np.float32)
cfg = api.ICPConfig(max_iterations=4, translation_tolerance=0.0, rotation_tolerance=0.0)
ame = api.AdaptiveMEstimator()
vmap = odo.map()
buf = (C.c_ulonglong * (2 * 8192))()


def run(name, icp, cloud):
    for _ in range(3):
        icp.optimize(vmap, cloud, guess)
    torch.cuda.synchronize()
    L.b2lo_debug_timeline(buf, 8192)
    icp.optimize(vmap, cloud, guess)
    n = L.b2lo_debug_timeline(buf, 8192)
    marks = sorted(((buf[2 * k + 1], buf[2 * k] >> 32, buf[2 * k] & 0xffffffff) for k in range(n)))
    t0 = marks[0][0]
    rows = [(tm - t0, label(fid, ln)) for tm, fid, ln in marks]
    agg = collections.defaultdict(list)
    for (a, nm), (b, _) in zip(rows, rows[1:] + [(rows[-1][0], "")]):
        agg[nm].append((b - a) * 1e-3)
    st = icp.get_last_stats()
    if RANK:
        return
    print(f"=== {name}: {len(cloud)} queries, {st.num_iterations} iterations, event time {st.optimization_time_ms * 1e3:.1f} us, first-to-last mark {rows[-1][0] * 1e-3:.1f} us")
    for nm, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"   {nm:40s} n {len(v):3d}  mean {np.mean(v):7.2f} us  total {sum(v):8.2f} us")


if WORLD > 1:
    from lidar_odometry_b200 import sharding
    big = np.concatenate([base + rng.normal(0, 0.01, base.shape).astype(np.float32) for _ in range(36)]).astype(np.float32)
    for nm, cloud in (("1.07 M", dense), ("4.3 M", big)):
        lo, hi = sharding.shard_bounds(len(cloud), WORLD, RANK)
        run(f"fused loop, this rank's share of {nm}", api.IterativeClosestPointOptimizer(cfg, ame), cloud[lo:hi])
        dist.barrier()
        run(f"peer-memory sharded loop, world {WORLD}, {nm}", api.PointShardedICP(cfg, ame, exchange="peer"), np.ascontiguousarray(cloud[lo:hi]))
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
run("fused loop, whole scan", api.IterativeClosestPointOptimizer(cfg, ame), dense)
run("fused loop, half scan", api.IterativeClosestPointOptimizer(cfg, ame), dense[: len(dense) // 2])
run("sharded loop (world 1), whole scan", api.PointShardedICP(cfg, ame), dense)
run("sharded loop (world 1), half scan", api.PointShardedICP(cfg, ame), dense[: len(dense) // 2])
