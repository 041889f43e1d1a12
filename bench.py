#!/usr/bin/env python
"""bench.py — scan-to-map registration throughput on synthetic KITTI-shaped scans (BASELINE.json configs[1]).

One "step" = one scan through the whole hot path: K1 voxel downsample -> K2..K5 surfel ICP with the Gauss-Newton
loop on the device -> K6 incremental surfel-map update on keyframes (nearly every scan at 1.2 m/scan).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI of include/b2lo.h)
  python bench.py --impl reference ...                     the reference arm: the CPU restatement of the reference's own
                                                           path (oracle/), timed on this box's host cores

value    : scans/s over all ranks, raw scans already resident in HBM (b2lo_odom_process_dev), CUDA-event time per scan
e2e      : same metric through the host-buffer call (b2lo_odom_process): strided gather to pinned memory + H2D + pose/counter D2H
           inside the timed region, wall clock
roofline : K2 surfel-correspondence kernel, 48 algorithmic bytes per query, CUDA-event timed per launch
Multi-GPU: independent sequences, one per rank, no data-path collective (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# before CUDA is initialised (torch does that first here): one hardware work queue per stream for the batched-sequence legs
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC, UNIT = "scan_to_map_scans_per_s", "scans/s"
WORKLOAD = "surfel scan-to-map ICP, KITTI shape (64x1900 HDL-64 model, ~120k pts/scan, stride 8, voxel 0.5 m, <=4 GN iters, PKO), 1.2 m/scan"
ALGO_BYTES_PER_QUERY = 48  # K2: 16 B query point + 32 B surfel record


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t = index, [], False, None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(self.rows)}


def k2_traffic():
    """DRAM traffic of K2 per launch as ncu measured it for the committed kernel (profiles/k2_traffic.json, written from the ncu summaries
    of the round); bench.py cannot count DRAM bytes itself, so a missing entry reads as null instead of a stale constant."""
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "k2_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def make_scans(n, seed, device):
    from lidar_odometry_b200 import synth
    scans, poses = synth.kitti_sequence(n_scans=n, seed=seed, device=device)
    return [np.ascontiguousarray(s) for s in scans], poses


def run_reference(args, rank, world):
    """The reference arm: the CPU restatement of the reference's own scan-to-map path on this box's host cores."""
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    scans, _ = make_scans(K + W, 42, "cuda" if _cuda() else None)
    # the workload at N GPUs is N independent sequences (one per GPU): the CPU arm runs them on N host threads (one single-threaded
    # pipeline each, as the reference's own hot path is; the libraries release the GIL), capped at the host's core count
    S = max(1, args.gpus)
    T = max(1, min(S, os.cpu_count() or 1))
    arms = cpu_arms(scans, K, W, S, T)
    kind, v = max(arms.items(), key=lambda kv: kv[1])
    dt = S * K / v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": v / 400.0 if S == 1 else None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "scans": K, "seed": 42, "sequences": S, "host_threads": T},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": T, "kind": kind, "host_cores_available": os.cpu_count(), "arms": arms, "arms_note": CPU_ARMS_NOTE,
                             "sample": f"{K} consecutive scans after {W} warm-up scans of the same synthetic sequence, {S} independent sequence(s) "
                                       f"(one per GPU of the b2lo arm) on {T} host thread(s); one sequence is single-threaded, as the reference hot path is"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


CPU_ARMS_NOTE = ("'reference' = the reference's own translation units (VoxelMap.cpp, IterativeClosestPointOptimizer.cpp, AdaptiveMEstimator.cpp, MathUtils.cpp, "
                 "PointCloudUtils.cpp, LidarFrame.cpp) compiled with the reference's own flags (-O3 -DNDEBUG, CMakeLists.txt:12,19-21) into oracle/_ref/libref_core.so against oracle/eigen_compat (Eigen is absent from the image; the stand-in "
                 "evaluates eagerly, so this build is somewhat slower than one against real Eigen), driven by the reference's own processing::Estimator::process_frame (the unmodified Estimator.cpp, loop detection and pose graph switched off); 'port' = the oracle restatement "
                 "(bit-identical poses, plain loops).  `value` is the FASTER of the two, so the GPU / CPU ratio is never flattered by the stand-in.")


def cpu_arms(scans, K, W, S=1, T=1, stage=None):
    """scans/s of the CPU path on this box: S independent single-threaded pipelines on T host threads, for the reference's own code
    (oracle/_ref, when it was built) and for the oracle port.  `stage`: dict that receives the per-stage ms/scan of each arm."""
    import threading
    from oracle import orc, ref
    orc.build()
    makers = {"port": orc.Pipeline}
    if ref.estimator_available():
        makers["reference"] = ref.Estimator      # processing::Estimator::process_frame itself (the unmodified Estimator.cpp)
    elif ref.available():
        makers["reference"] = ref.Pipeline       # the same control flow over the reference's classes
    out = {}
    for kind, make in makers.items():
        pipes = [make() for _ in range(S)]
        bar = threading.Barrier(T + 1)
        st = [np.zeros(4) for _ in range(T)]

        def work(t):
            mine = pipes[t::T]
            for s in scans[:W]:
                for pp in mine:
                    pp.process(s)
            bar.wait()
            for s in scans[W:W + K]:
                for pp in mine:
                    st[t] += pp.process(s).get("times_ms", 0.0)
            bar.wait()

        th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
        for x in th:
            x.start()
        bar.wait()
        t0 = time.perf_counter()
        bar.wait()
        dt = time.perf_counter() - t0
        for x in th:
            x.join()
        out[kind] = S * K / dt
        if stage is not None:
            tot = sum(st) / (S * K)
            stage[kind] = {"preprocess": tot[0], "icp": tot[1], "map_update": tot[2]}
        del pipes
    return out


def point_sharded_leg(args, rank, world, local, api, torch, dist, standalone=False):
    """BASELINE.json configs[4], second half: one dense multi-LiDAR scan (~1.08 M points = 9 merged HDL-64 returns), queries split
    contiguously across the ranks, map replicated; three tiny exchanges per Gauss-Newton iteration (3 x world + 128 + 28 doubles),
    NCCL calls enqueued from C on the context stream between the kernels (b2lo_icp_shard_optimize: no host round trip inside the loop).
    Time = CUDA events around the whole optimize on the context stream, max over ranks; the exchange share = CUDA-event time of one
    iteration's exchanges (the collectives plus the plan / sample kernels between them) x iterations.  The sharded pose is checked
    against the unsharded (fused, single-GPU) loop on the same dense cloud (< 1e-5 m).  The PKO fit is replicated, it does not shard."""
    from lidar_odometry_b200 import sharding, capi as capi_mod
    scans, poses = make_scans(11, 42, f"cuda:{local}")       # same seed on every rank: identical map replicas
    ctx = api.Context(local)
    odo = api.Odometry(ctx)
    for s in scans[:10]:
        r = odo.process(s)
    guess = r["pose"]
    rng = np.random.default_rng(99)
    base = scans[10][:, :3]
    dense = np.concatenate([base + rng.normal(0, 0.01, base.shape).astype(np.float32) for _ in range(9)]).astype(np.float32)
    lo, hi = sharding.shard_bounds(len(dense), world, rank)
    mine = np.ascontiguousarray(dense[lo:hi])
    cfg = api.ICPConfig(max_iterations=4, translation_tolerance=0.0, rotation_tolerance=0.0)   # four full iterations
    ame = api.AdaptiveMEstimator()
    vmap = odo.map()
    # the unsharded loop on the whole cloud (every rank; also the N = 1 reference time)
    fused = api.IterativeClosestPointOptimizer(cfg, ame)
    for _ in range(3):
        ok_f, T_f = fused.optimize(vmap, dense, guess)
    fused_ms = 0.0
    R = max(args.steps // 10, 5)
    for _ in range(R):
        ok_f, T_f = fused.optimize(vmap, dense, guess)
        fused_ms += fused.get_last_stats().optimization_time_ms
    out = {}
    legs = [("device_ordered", True, "nccl"), ("host_driven", False, "nccl")]
    if world > 1:
        legs.insert(0, ("peer_memory", True, "peer"))      # exchanges as stores into the peers' mailboxes over NVLink, no NCCL
    for name, dev_ordered, exch in legs:
        icp = api.PointShardedICP(cfg, ame, device_ordered=dev_ordered, exchange=exch)
        try:
            ok, T = icp.optimize(vmap, mine, guess)
        except capi_mod.B2loError as e:       # raised on every rank alike (api._communicator agrees on it first)
            if exch != "peer":
                raise
            out[name] = {"unavailable": str(e)}
            continue
        for _ in range(max(args.warmup, 3)):
            ok, T = icp.optimize(vmap, mine, guess)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter(); dev_ms = 0.0; coll_ms = 0.0; iters = 0
        for _ in range(R):
            ok, T = icp.optimize(vmap, mine, guess)
            it = icp.get_last_stats().num_iterations
            iters += it
            if dev_ordered:
                dev_ms += icp.device_ms; coll_ms += icp.collective_ms_last_iteration * it
            else:
                coll_ms += 1e3 * icp.collective_seconds
        torch.cuda.synchronize()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        if not dev_ordered:
            dev_ms = wall_ms          # host-driven phases: only the wall clock sees the whole optimize
        tt = torch.tensor([dev_ms, coll_ms, wall_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        err = float(np.linalg.norm(T[:3, 3].astype(np.float64) - T_f[:3, 3]))
        assert ok and ok_f and err < 1e-5, f"sharded pose differs from the fused loop by {err} m"
        out[name] = {"ms_per_optimize": float(tt[0]) / R, "ms_per_iteration": float(tt[0]) / max(iters, 1), "exchange_ms_per_iteration": float(tt[1]) / max(iters, 1),
                     "exchange_share": float(tt[1]) / max(float(tt[0]), 1e-12), "wall_ms_per_optimize": float(tt[2]) / R, "gn_iterations": iters // R,
                     "correspondences": icp.get_last_stats().num_correspondences, "pose_vs_fused_m": err,
                     "timing": "CUDA events on the context stream, max over ranks" if dev_ordered else "wall clock incl. host-driven phase boundaries (stream syncs around each collective), max over ranks"}
        del icp
    have_peer = "ms_per_iteration" in out.get("peer_memory", {})
    res = {"mode": "point_sharded", "n_gpus": world, "comm_nranks": world, "queries": int(len(dense)), "queries_per_rank": int(hi - lo), "optimizes": R,
           "unsharded_single_gpu_ms_per_iteration": fused_ms / (R * 4), "device_ordered": out["device_ordered"], "host_driven": out["host_driven"],
           "peer_memory": out.get("peer_memory"),
           "collectives_per_iteration": 3, "payload_doubles_per_iteration": [3 * world, 128, 28],
           "speedup_vs_unsharded": (fused_ms / (R * 4)) / max((out["peer_memory"] if have_peer else out["device_ordered"])["ms_per_iteration"], 1e-12),
           "speedup_vs_unsharded_nccl": (fused_ms / (R * 4)) / max(out["device_ordered"]["ms_per_iteration"], 1e-12),
           "note": "map replicated; PKO fit replicated on every rank (does not shard); pose asserted equal to the fused loop within 1e-5 m"}
    if world > 1 and have_peer:
        # where sharding starts to pay: the same exchange on a 4x denser scan (36 merged returns, ~4.3 M points); the replicated
        # PKO fit and the exchanges stay constant while K2 / K5 grow with the queries
        big = np.concatenate([base + rng.normal(0, 0.01, base.shape).astype(np.float32) for _ in range(36)]).astype(np.float32)
        blo, bhi = sharding.shard_bounds(len(big), world, rank)
        bmine = np.ascontiguousarray(big[blo:bhi])
        for _ in range(3):
            ok_f, T_f = fused.optimize(vmap, big, guess)
        big_fused = 0.0
        for _ in range(R):
            ok_f, T_f = fused.optimize(vmap, big, guess)
            big_fused += fused.get_last_stats().optimization_time_ms
        icp = api.PointShardedICP(cfg, ame, device_ordered=True, exchange="peer")
        for _ in range(3):
            ok, T = icp.optimize(vmap, bmine, guess)
        torch.cuda.synchronize(); dist.barrier()
        big_ms = 0.0; big_x = 0.0
        for _ in range(R):
            ok, T = icp.optimize(vmap, bmine, guess)
            big_ms += icp.device_ms; big_x += icp.collective_ms_last_iteration * 4
        tt = torch.tensor([big_ms, big_x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        err = float(np.linalg.norm(T[:3, 3].astype(np.float64) - T_f[:3, 3]))
        assert ok and ok_f and err < 1e-5, f"sharded pose (4.3 M points) differs from the fused loop by {err} m"
        res["dense_4x"] = {"queries": int(len(big)), "unsharded_single_gpu_ms_per_iteration": big_fused / (R * 4), "peer_memory_ms_per_iteration": float(tt[0]) / (R * 4),
                           "exchange_ms_per_iteration": float(tt[1]) / (R * 4), "speedup_vs_unsharded": big_fused / max(float(tt[0]), 1e-12), "pose_vs_fused_m": err}
        del icp
    if standalone:
        if rank == 0:
            print(json.dumps(res))
        if world > 1:
            dist.destroy_process_group()
    return res


def run_point_sharded(args, rank, world, local, api, torch, dist):
    return point_sharded_leg(args, rank, world, local, api, torch, dist, standalone=True)


def dropin_estimator_leg(scans, K, W):
    """The reference's OWN driver on the CUDA engine: the unmodified src/processing/Estimator.cpp, compiled with database/VoxelMap.h and
    optimization/IterativeClosestPointOptimizer.h replaced by the drop-in shim and linked with libb2lo.so (oracle/_ref/libref_estimator_gpu.so,
    built where the reference tree exists).  Wall clock per process_frame call in a child process (tools/estimator_gpu_bench.py); includes
    what the reference does around the hot path (cloud copies, keyframe bookkeeping, sliding-window cleanup).
    (oracle/_ref is only where compiled reference code lives: here the REFERENCE is the caller and libb2lo.so does every bit of the
    hot-path arithmetic on the GPU - the product does not route through the oracle; the number is informational, next to dropin_e2e.)"""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.abspath(__file__))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "libref_estimator_gpu.so")):
        return {"unavailable": "oracle/_ref/libref_estimator_gpu.so was not built (no reference tree where the repository was built)"}
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        for s in scans[:W + K]:
            a = np.ascontiguousarray(s[:, :4], np.float32)
            f.write(np.uint32(a.shape[0]).tobytes()); f.write(a.tobytes())
        path = f.name
    try:
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "estimator_gpu_bench.py"), path, str(W)], capture_output=True, text=True, timeout=300)
    except Exception as e:      # timeout etc.
        return {"unavailable": f"child failed: {e}"}
    finally:
        os.unlink(path)
    if r.returncode != 0:
        return {"unavailable": "child exited with " + str(r.returncode) + ": " + (r.stderr or r.stdout)[-300:]}
    try:
        out = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return {"unavailable": "child printed no JSON: " + r.stdout[-200:]}
    out.update({"unit": UNIT, "value": out["scans_per_s"], "driver": "processing::Estimator::process_frame of the unmodified Estimator.cpp over b2lo_dropin.h + libb2lo.so",
                "timing": "host wall clock per scan in a child process; pageable clouds; loop detection and pose graph off"})
    return out


def mid360_leg(ctx, api, capi, flush, torch):
    """BASELINE.json configs[2]: KDTree-mode correspondence (exact 5-NN over the L0 hash + per-query plane fit) on MID360-shaped
    non-repetitive scans (20 k points, stride 4, 0.4 m voxels), whole scan-to-map pipeline, next to the CPU oracle on the same scans."""
    import ctypes as C
    from lidar_odometry_b200 import synth
    from oracle import orc
    L = capi.lib()
    scans, _ = synth.mid360_sequence(n_scans=45, seed=42, device="cuda")
    scans = [torch.from_numpy(np.ascontiguousarray(s)).pin_memory().numpy() for s in scans]
    odo = api.Odometry(ctx, mid360=True)
    for s in scans[:5]:
        odo.process(s)
    dev_ms = 0.0; wall = 0.0; nq = 0; kf = 0
    for s in scans[5:]:
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = odo.process(s)
        wall += time.perf_counter() - t0
        dev_ms += r["device_ms"]; nq += r["n_features"] * r["n_iters"]; kf += int(r["keyframe"])
    n = len(scans) - 5
    # K3 alone, CUDA events (plain launches)
    odo2 = api.Odometry(ctx, mid360=True)
    for s in scans[:5]:
        odo2.process(s)
    L.b2lo_ctx_profile(ctx.h, 1)
    q2 = 0
    for s in scans[5:]:
        r = odo2.process(s); q2 += r["n_features"] * r["n_iters"]
    ms, cnt = C.c_double(), C.c_longlong()
    L.b2lo_ctx_profile_read(ctx.h, 7, C.byref(ms), C.byref(cnt))
    L.b2lo_ctx_profile(ctx.h, 0)
    pipe = orc.Pipeline(orc.default_pipe_cfg(True))
    for s in scans[:5]:
        pipe.process(s)
    t0 = time.perf_counter()
    for s in scans[5:]:
        pipe.process(s)
    cpu = time.perf_counter() - t0
    return {"workload": "MID360-shaped rosette, 20 k pts/scan, stride 4, voxel 0.4 m, KDTree correspondence (5-NN + plane fit), 0.1 m/scan",
            "scans": n, "keyframes": kf, "ms_per_scan_device": dev_ms / n, "e2e_scans_per_s": n / wall, "queries_per_scan_iteration": nq / max(n, 1),
            "k3_knn_plane_fit": {"avg_us_per_iteration": 1e3 * ms.value / max(cnt.value, 1), "queries_per_s": q2 / max(ms.value * 1e-3, 1e-12),
                                 "algorithmic_bytes_per_query": 96},
            "cpu_oracle_scans_per_s": n / cpu}


def stress_leg(ctx, api, capi, target_voxels, peak, peak_kind):
    """BASELINE.json configs[3]: ~10^7-voxel hierarchical hash (a stack of planar slabs filling the 120 m cull sphere's
    bounding square), keyframe updates on top of it, and the K2 surfel probe on a table far larger than the 126 MB L2.
    Here the kernels ARE bandwidth-bound, so this is where the roofline fractions mean something."""
    import ctypes as C
    import torch
    L = capi.lib()
    rng = np.random.default_rng(1234)
    side = 440                                    # 220 m x 220 m of 0.5 m voxels per slab
    per_layer = side * side
    layers = max(1, int(round(target_voxels / per_layer)))
    vmap = api.VoxelMap(0.5, ctx, capacity_hint=int(per_layer * layers * 1.15))
    gx, gy = np.meshgrid(np.arange(side, dtype=np.float32), np.arange(side, dtype=np.float32), indexing="ij")
    base = np.stack([gx.ravel(), gy.ravel()], axis=1) * np.float32(0.5) - np.float32(110.0)
    t_build = time.perf_counter()
    for l in range(layers):
        z = np.float32(-39.0 + 1.5 * l + 0.7)   # one slab per 1.5 m L1 layer, so every L1 cell stays planar
        pts = np.empty((per_layer, 3), np.float32)
        pts[:, :2] = base + rng.uniform(0.05, 0.45, (per_layer, 2)).astype(np.float32)
        pts[:, 2] = z + rng.normal(0.0, 0.01, per_layer).astype(np.float32)
        vmap.UpdateVoxelMap(pts, [0.0, 0.0, 0.0], 400.0)
    ctx.sync()
    t_build = time.perf_counter() - t_build
    v0, v1, nsurf = vmap.GetVoxelCount(), vmap.GetL1VoxelCount(), vmap.GetSurfelCount()
    # keyframe updates of ~10^4 points on the big map: cull scan (16 B / voxel) + insert + surfel refit
    L.b2lo_ctx_profile(ctx.h, 1)
    n_upd = 20
    for i in range(n_upd):
        ang = rng.uniform(0, 2 * np.pi, 10000); rad = rng.uniform(2, 80, 10000)
        pts = np.stack([rad * np.cos(ang), rad * np.sin(ang), -39.0 + 1.5 * rng.integers(0, layers, 10000) + 0.7 + rng.normal(0, 0.01, 10000)], axis=1).astype(np.float32)
        vmap.UpdateVoxelMap(pts, [0.1 * i, 0.0, 0.0], 400.0)
    ms_cull, n_cull, ms_upd, n_updl = C.c_double(), C.c_longlong(), C.c_double(), C.c_longlong()
    L.b2lo_ctx_profile_read(ctx.h, 8, C.byref(ms_cull), C.byref(n_cull))
    L.b2lo_ctx_profile_read(ctx.h, 5, C.byref(ms_upd), C.byref(n_updl))
    cull_gbs = 16.0 * v0 * n_cull.value / max(ms_cull.value * 1e-3, 1e-12) / 1e9
    # K2 probe: 2^20 queries, one correspondence pass per optimize call.  "random": uniformly spread over the map (every query
    # its own 32 B table sector -> the worst case for DRAM); "coherent": a raster sweep over the slabs like a real scan, where
    # the ~9 queries of one 1.5 m L1 cell share a sector through L2.
    nq = 1 << 20
    probes = {}
    icp = api.IterativeClosestPointOptimizer(api.ICPConfig(max_iterations=1), api.AdaptiveMEstimator())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms_gn, n_gn = C.c_double(), C.c_longlong()
    for kind in ("random", "coherent"):
        if kind == "random":
            q = np.stack([rng.uniform(-109, 109, nq), rng.uniform(-109, 109, nq), -39.0 + 1.5 * rng.integers(0, layers, nq) + 0.7 + rng.normal(0, 0.02, nq)], axis=1).astype(np.float32)
        else:
            ax = np.arange(-109.0, 109.0, 0.5)          # one query per 0.5 m L0 cell, raster order: 9 queries per 1.5 m L1 cell and layer
            xx, yy = np.meshgrid(ax, ax, indexing="ij")
            xy = np.stack([xx.ravel(), yy.ravel()], axis=1)
            need = (nq + len(xy) - 1) // len(xy)
            q = np.concatenate([np.concatenate([xy, np.full((len(xy), 1), -39.0 + 1.5 * l + 0.7)], axis=1) for l in range(need)])[:nq]
            q = (q + rng.normal(0, 0.01, q.shape)).astype(np.float32)
        icp.optimize(vmap, q, np.eye(4, dtype=np.float32))   # warm-up (allocations, PKO tables)
        L.b2lo_ctx_profile(ctx.h, 1)
        reps = 10
        for _ in range(reps):
            flush.zero_(); torch.cuda.synchronize()
            ok, _T = icp.optimize(vmap, q, np.eye(4, dtype=np.float32))
        ncorr = icp.get_last_stats().num_correspondences
        ms_k2, n_k2 = C.c_double(), C.c_longlong()
        L.b2lo_ctx_profile_read(ctx.h, 1, C.byref(ms_k2), C.byref(n_k2))
        L.b2lo_ctx_profile_read(ctx.h, 4, C.byref(ms_gn), C.byref(n_gn))
        L.b2lo_ctx_profile(ctx.h, 0)
        k2_gbs = ALGO_BYTES_PER_QUERY * nq * n_k2.value / max(ms_k2.value * 1e-3, 1e-12) / 1e9
        probes[kind] = {"queries": nq, "accepted": ncorr, "avg_launch_us": 1e3 * ms_k2.value / max(n_k2.value, 1), "achieved_gbs": k2_gbs,
                        "frac_of_hbm_peak": k2_gbs / peak, "peak_kind": f"of {peak_kind}", "queries_per_s": nq * n_k2.value / max(ms_k2.value * 1e-3, 1e-12),
                        "algorithmic_bytes_per_query": ALGO_BYTES_PER_QUERY, "l2": "flushed between launches; L1 hash table > L2"}
    # bare ceiling of the access pattern on this GPU (tools/gather_probe.cu, profiles/r01_gather_probe.txt): 2^20 independent random 32 B-sector
    # gathers from a 128 MiB table, nothing else read or written, take 26.6 us (39.4 G sectors/s)
    probes["random"].update({"random_sector_gather_ceiling_us": 26.6, "frac_of_random_gather_ceiling": 26.6 / max(probes["random"]["avg_launch_us"], 1e-9),
                             "ceiling_source": "tools/gather_probe.cu on the same GPU model, profiles/r01_gather_probe.txt"})
    probes["coherent"].update({"coherent_gather_ceiling_us": 10.2, "ceiling_source": "tools/gather_probe.cu, profiles/r01_gather_probe.txt (bare gathers only; K2 also streams 28 B/query)"})
    tr = k2_traffic()
    for kind, key in (("random", "k2_stress_random"), ("coherent", "k2_stress_coherent")):
        t = tr.get(key)
        probes[kind].update({"traffic_bytes_per_launch": t["bytes_per_launch"] if t else None, "traffic_source": (t["source"] + " (ncu, cold caches)") if t else None,
                             "traffic_over_algorithmic": (t["bytes_per_launch"] / (ALGO_BYTES_PER_QUERY * nq)) if t else None})
    # bulk rebuild (ApplyTransformAndRehash + RecomputeAllSurfels, VoxelMap.cpp:264-366; runs after pose-graph corrections): transform every
    # L0 centroid, re-key, merge collisions, rebuild L1 and refit every surfel.  Algorithmic bytes per L0 voxel: 16 B read + 16 B written
    # + one 32 B slot written + 16 B gathered by its parent's refit = 80 B.
    import ctypes as C2
    from lidar_odometry_b200 import synth
    Tr = synth.pose_matrix(0.31, -0.22, 0.05, 0.02, 0.003, -0.004).astype(np.float32)
    vmap.ApplyTransformAndRehash(Tr)          # warm-up: sizes the scratch tables
    ctx.sync()
    v0b = vmap.GetVoxelCount()
    t_re = time.perf_counter()
    vmap.ApplyTransformAndRehash(np.linalg.inv(Tr.astype(np.float64)).astype(np.float32))
    ctx.sync()
    t_re = time.perf_counter() - t_re
    rehash = {"voxels_in": v0b, "voxels_out": vmap.GetVoxelCount(), "ms": 1e3 * t_re, "algorithmic_bytes_per_voxel": 80,
              "achieved_gbs": 80.0 * v0b / t_re / 1e9, "frac_of_hbm_peak": 80.0 * v0b / t_re / 1e9 / peak, "timing": "host wall clock around the call, synchronised"}
    del vmap
    return {"rehash_rebuild": rehash, "workload": f"{layers} planar slabs of {side}x{side} voxels (0.5 m), one point per voxel", "l0_voxels": v0, "l1_voxels": v1, "surfels": nsurf,
            "build_s": t_build,
            "k2_probe": probes["random"], "k2_probe_coherent": probes["coherent"],
            "k5_normal_eq": {"avg_launch_us": 1e3 * ms_gn.value / max(n_gn.value, 1)},
            "k6_cull_scan": {"avg_launch_us": 1e3 * ms_cull.value / max(n_cull.value, 1), "achieved_gbs": cull_gbs, "frac_of_hbm_peak": cull_gbs / peak,
                             "algorithmic_bytes_per_voxel": 16},
            "k6_update_rest": {"avg_ms_per_keyframe": ms_upd.value / max(n_upd, 1), "points_per_keyframe": 10000}}


def concurrent_leg(api, local, dev_args, S, K, W):
    """Throughput mode for batches of recorded sequences on ONE GPU: S independent sequences (own context, stream, map and CUDA
    graphs each) driven by S host threads.  One sequence keeps a single SM busy most of the time (the PKO fit and the Gauss-Newton
    finish are one-CTA latency chains), so independent sequences overlap almost freely.  Wall clock between two thread barriers."""
    import torch
    ctxs = [api.Context(local) for _ in range(S)]
    odos = [api.Odometry(c) for c in ctxs]
    bar = threading.Barrier(S + 1)
    errs = []

    def work(j):
        try:
            o = odos[j]
            for i in range(W):
                o.process_dev(*dev_args(i), lookahead=dev_args(i + 1))
            bar.wait()
            for i in range(W, W + K):
                o.process_dev(*dev_args(i), lookahead=dev_args(i + 1))
            bar.wait()
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
            bar.abort()

    th = [threading.Thread(target=work, args=(j,)) for j in range(S)]
    for t in th:
        t.start()
    try:
        bar.wait()
        t0 = time.perf_counter()
        bar.wait()
        dt = time.perf_counter() - t0
    except threading.BrokenBarrierError:
        dt = None
    for t in th:
        t.join()
    torch.cuda.synchronize()
    if errs or dt is None:
        return {"sequences": S, "error": errs[:1]}
    return {"sequences": S, "scans_per_sequence": K, "scans_per_s": S * K / dt, "ms_per_scan_per_sequence": 1e3 * dt / K,
            "timing": "host wall clock between thread barriers, all sequences resident in HBM, no L2 flush (S sequences x 1.9 MB scans stream through)",
            "note": "aggregate over independent sequences sharing one B200; not the per-sequence latency of `value`"}


def batch_leg(api, local, dev_args, S, K, W):
    """The same throughput mode without a host thread per sequence: b2lo_odom_process_batch_dev enqueues one scan of every sequence
    (one CUDA graph replay each, own stream) before it waits for the first result."""
    import torch
    bat = api.OdometryBatch(S, local)
    def args(i):
        p, n, sf = dev_args(i)
        return [p] * S, [n] * S, sf
    for i in range(W):
        p, n, sf = args(i); pn, nn, _ = args(i + 1)
        bat.process_dev(p, n, sf, pn, nn)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(W, W + K):
        p, n, sf = args(i); pn, nn, _ = args(i + 1)
        bat.process_dev(p, n, sf, pn, nn)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    r = bat.results()
    return {"sequences": S, "scans_per_sequence": K, "scans_per_s": S * K / dt, "ms_per_round": 1e3 * dt / K, "driver": "one host thread, b2lo_odom_process_batch_dev",
            "last_pose_t": [float(v) for v in r[0]["pose"][:3, 3]], "keyframes_last_round": int(sum(x["keyframe"] for x in r)),
            "timing": "host wall clock around K batch calls, all sequences resident in HBM, no L2 flush (S sequences x 1.9 MB scans stream through)",
            "note": "aggregate over independent sequences sharing one B200; not the per-sequence latency of `value`"}


def lockstep_leg(api, local, dev_args, S, K, W, peak, peak_kind):
    """Lock-step batch (b2lo_lockstep_*): S independent sequences, ONE graph replay per step, every kernel started once per step with
    blockIdx.y = sequence.  Sequence j runs the scans rotated by j, so the S maps and poses differ inside a step.  Time = CUDA events around
    each step's graph on the batch stream (device-resident scans, no L2 flush between steps: S maps + S scans are far larger than L2)."""
    import torch
    n_have = K + W + 1
    odos = [api.Odometry(api.Context(local)) for _ in range(S)]
    ls = api.LockstepBatch(odos)

    plans = [[dev_args((k + j) % n_have)[:2] for j in range(S)] for k in range(n_have)]   # argument lists built outside the timed loop

    def step(k):
        return ls.process_dev(plans[k % n_have], dev_args(0)[2], raw=True)

    for k in range(W):
        step(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); dev_ms = 0.0; nq = 0; iters = 0
    for k in range(W, W + K):
        res, ms = step(k)
        dev_ms += ms
        for r in res:
            nq += r.n_features * r.n_iters; iters += r.n_iters
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    st = ls.stats()
    out = {"sequences": S, "scans": S * K, "scans_per_s": S * K / (dev_ms * 1e-3), "ms_per_step": dev_ms / K, "scans_per_s_wall": S * K / wall,
           "wall_ms_per_step": 1e3 * wall / K, "gn_iterations_per_scan": iters / (S * K), "queries_per_step_iteration": nq / max(iters, 1) * S,
           "kernels_per_step": st["kernels_per_step"], "graph_builds": st["builds"], "fallback_steps": st["fallbacks"],
           "driver": "b2lo_lockstep_process_dev: one CUDA graph replay per step for all sequences, grid.y = sequence",
           "timing": "CUDA events around every step on the batch stream; wall = host clock around the K calls"}
    del ls, odos
    return out


def lockstep_groups_leg(api, local, dev_args, G, S, K, W):
    """G lock-step batches of S sequences each, one host thread and one stream per batch, running side by side on one GPU: while one
    batch sits in a latency-bound step (the S PKO fits: one CTA per sequence), the kernels of the other batches fill the remaining SMs.
    Whole-leg wall clock between barriers (the batches overlap, so per-step event times would double count)."""
    import threading
    import torch
    n_have = K + W + 1
    groups = []
    for g in range(G):
        odos = [api.Odometry(api.Context(local)) for _ in range(S)]
        groups.append((odos, api.LockstepBatch(odos)))

    plans = [[[dev_args((k + j + 7 * g) % n_have)[:2] for j in range(S)] for k in range(n_have)] for g in range(G)]

    def run(g, k0, k1):
        odos, ls = groups[g]
        for k in range(k0, k1):
            ls.process_dev(plans[g][k % n_have], dev_args(0)[2], raw=True)

    for g in range(G):
        run(g, 0, W)
    torch.cuda.synchronize()
    th = [threading.Thread(target=run, args=(g, W, W + K)) for g in range(G)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    out = {"batches": G, "sequences_per_batch": S, "sequences": G * S, "scans": G * S * K, "scans_per_s": G * S * K / wall, "wall_ms_per_step": 1e3 * wall / K,
           "driver": "G x b2lo_lockstep_process_dev from G host threads (one stream each)", "timing": "host wall clock around the whole leg, device synchronised on both sides"}
    del groups
    return out


def dropin_leg(scans, K, W, show_stderr=False):
    """The class-by-class drop-in through the C++ shim (lidar_odometry_b200/shim/b2lo_dropin.h) in Estimator::process_frame order
    (Estimator.cpp:116-233, 449-470): FastVoxelFilter::filter -> optimize -> host transform -> UpdateVoxelMap -> GetPointCloud, pageable
    std::vector clouds, every hand-over through the host.  A C++ program (shim/test/dropin_bench.cpp) compiled here with g++; wall clock
    per scan, stage split like the reference's TimingStats.  This is what an UNMODIFIED Estimator.cpp pays; `e2e` is the widened call."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.abspath(__file__))
    shim = os.path.join(root, "lidar_odometry_b200", "shim")
    exe = os.path.join(shim, "test", "dropin_bench")
    cmd = ["g++", "-std=c++17", "-O2", "-I" + os.path.join(root, "include"), "-I" + shim, os.path.join(shim, "test", "dropin_bench.cpp"), "-o", exe,
           "-L" + os.path.join(root, "lidar_odometry_b200"), "-lb2lo", "-Wl,-rpath," + os.path.join(root, "lidar_odometry_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return {"unavailable": "g++ failed: " + r.stderr[-300:]}
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        for s in scans[:W + K]:
            a = np.ascontiguousarray(s[:, :4], np.float32)
            f.write(np.uint32(a.shape[0]).tobytes()); f.write(a.tobytes())
        path = f.name
    try:
        r = subprocess.run([exe, path, str(W)], stdout=subprocess.PIPE, stderr=None if show_stderr else subprocess.PIPE, text=True, timeout=600)
    finally:
        os.unlink(path)
    if r.returncode != 0:
        return {"unavailable": "dropin_bench failed: " + (r.stderr or r.stdout)[-300:]}
    out = json.loads(r.stdout.strip().splitlines()[-1])
    out.update({"unit": UNIT, "value": out["scans_per_s"], "driver": "shim/test/dropin_bench.cpp over b2lo_dropin.h (the reference's three classes), pageable host clouds",
                "timing": "host wall clock per scan inside the C++ program; no L2 flush between scans (a separate process)"})
    return out


def export_leg(ctx, api, scans, poses, K):
    """the final-map downsample of Estimator::save_map_to_ply (util::VoxelGrid, leaf 0.4 m) over the accumulated world
    cloud of the benchmark sequence, through the host-buffer call (H2D + D2H inside), next to the CPU oracle on the same cloud."""
    from oracle import orc
    world = []
    for s, T in zip(scans[:K], poses[:K]):
        T = np.asarray(T, np.float32)
        world.append((s[::8, :3] @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
    acc = np.ascontiguousarray(np.concatenate(world))
    g = api.VoxelGrid(ctx)
    g.setLeafSize(0.4)
    g.setInputCloud(acc)
    out = g.filter()                 # warm-up: sizes the K1 scratch for this cloud
    ctx.sync()
    t0 = time.perf_counter()
    out = g.filter()
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = orc.voxel_grid_filter(acc, 0.4)
    t_cpu = time.perf_counter() - t0
    return {"workload": f"util::VoxelGrid leaf 0.4 m over {len(acc)} accumulated world points ({K} scans, every 8th point)", "points_in": int(len(acc)),
            "voxels_out": int(len(out)), "ms": 1e3 * t_gpu, "points_per_s": len(acc) / t_gpu, "cpu_oracle_ms": 1e3 * t_cpu,
            "bit_identical_to_oracle": bool(out.shape == ref.shape and np.array_equal(out.view(np.uint32), ref.view(np.uint32))),
            "algorithmic_bytes": 12 * int(len(acc)) + 12 * int(len(out)), "timing": "host wall clock around the host-buffer call (pageable input, H2D + D2H inside)"}


def _cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b2lo", choices=["b2lo", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="sequences", choices=["sequences", "sharded"],
                    help="sequences: one independent sequence per GPU (the contract line); sharded: one dense ~1M-point scan, queries split across ranks")
    ap.add_argument("--no-stress", action="store_true", help="skip the 10^7-voxel map leg (BASELINE.json configs[3])")
    ap.add_argument("--stress-voxels", type=float, default=1.0e7)
    ap.add_argument("--stress-only", action="store_true")
    ap.add_argument("--concurrent", type=int, nargs="*", default=[16, 32, 96], help="sequences sharing one GPU in the throughput-mode leg (empty: skip); the first size is also run with one host thread per sequence")
    ap.add_argument("--lockstep", type=int, nargs="*", default=[32, 128, 256, 384], help="batch sizes of the lock-step throughput leg (empty: skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lidar_odometry_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from lidar_odometry_b200 import api, capi
    import ctypes as C
    if args.mode == "sharded":
        return run_point_sharded(args, rank, world, local, api, torch, dist)
    K, W = args.steps, args.warmup
    if args.stress_only:
        peak, peak_kind = peaks()
        st = stress_leg(api.Context(local), api, capi, int(args.stress_voxels), peak, peak_kind)
        print(json.dumps({"k2_probe": st["k2_probe"], "k2_probe_coherent": st["k2_probe_coherent"], "rehash_rebuild": st["rehash_rebuild"]}))
        return
    scans, true_poses = make_scans(K + W + 1, 42, f"cuda:{local}")   # weak scaling: every rank processes its own copy of the same sequence; +1: the look-ahead of the last timed scan
    ctx = api.Context(local)
    dev_scans = [torch.from_numpy(s).cuda() for s in scans]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: scans resident in HBM, CUDA-event time per scan, L2 flushed between scans ------------------
    # Recorded-sequence mode: scan i+1 is announced (b2lo_odom_lookahead) before scan i is processed, so its K1 runs beside the
    # registration of scan i.  Every timed step still holds exactly one K1, one ICP and one map update.
    def dev_args(i):
        return dev_scans[i].data_ptr(), scans[i].shape[0], scans[i].shape[1]

    odo = api.Odometry(ctx)
    for i in range(W):
        odo.process_dev(*dev_args(i), lookahead=dev_args(i + 1))
    barrier()
    l0 = ctx.launch_count
    dev_ms = 0.0; ncorr = nq = 0; iters = 0; kf = 0
    clk = ClockSampler(local)
    if rank == 0:
        clk.__enter__()      # keeps sampling through the value and the e2e pass (each is only tens of ms long); rank 0 only: eight
                             # processes polling nvidia-smi at once stall each other's driver calls inside the timed region
    if True:
        t_wall0 = time.perf_counter()
        for i in range(W, W + K):
            flush.zero_()
            torch.cuda.synchronize()
            r = odo.process_dev(*dev_args(i), lookahead=dev_args(i + 1))
            dev_ms += r["device_ms"]; ncorr += r["n_corr"]; iters += r["n_iters"]; nq += r["n_features"] * r["n_iters"]; kf += int(r["keyframe"])
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count - l0
    if os.environ.get("B2LO_BENCH_VERBOSE"):
        print(f"[rank {rank}] value pass: {dev_ms / K:.4f} ms/scan device, {1e3 * t_wall / K:.4f} ms/scan wall incl. flush, "
              f"{iters / K:.2f} GN iterations/scan, {nq / max(iters, 1):.0f} queries/iteration", file=sys.stderr, flush=True)
    tt = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    max_ms = float(tt.item())
    value = world * K / (max_ms * 1e-3)
    final_l0, final_l1 = r["l0"], r["l1"]

    # ---- e2e: host scans (page-locked, as the contract asks) through the public host-buffer call, wall clock, copies inside ------
    pinned = [torch.from_numpy(s).pin_memory() for s in scans]
    scans = [t.numpy() for t in pinned]
    odo2 = api.Odometry(ctx)
    for i in range(W):
        odo2.process(scans[i], lookahead=scans[i + 1])
    barrier()
    h0, d0 = ctx.io_bytes()
    e2e_s = 0.0
    for i in range(W, W + K):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        odo2.process(scans[i], lookahead=scans[i + 1])
        e2e_s += time.perf_counter() - t0
    h1, d1 = ctx.io_bytes()
    if rank == 0:
        clk.__exit__(None, None, None)
    # live-sensor call sequence (what the reference's own player does: load, then process, app/player/kitti_player.cpp:109-123): no
    # look-ahead, K1 in line.  Device-resident (CUDA events) and through the host-buffer call (wall clock, copies inside).
    odo_s = api.Odometry(ctx)
    for i in range(W):
        odo_s.process_dev(*dev_args(i))
    stream_ms = 0.0
    for i in range(W, W + K):
        flush.zero_()
        torch.cuda.synchronize()
        stream_ms += odo_s.process_dev(*dev_args(i))["device_ms"]
    del odo_s
    odo_s = api.Odometry(ctx)
    for i in range(W):
        odo_s.process(scans[i])
    stream_e2e_s = 0.0
    for i in range(W, W + K):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        odo_s.process(scans[i])
        stream_e2e_s += time.perf_counter() - t0
    del odo_s
    ts = torch.tensor([stream_ms, stream_e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    stream_ms, stream_e2e_s = float(ts[0].item()), float(ts[1].item())
    graph = odo2.graph_stats()
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * K / float(te.item())

    # ---- throughput mode on every rank (N > 1): three lock-step batches of 128 independent sequences per GPU (b2lo_lockstep_process_dev from three host
    # threads: by the wall clock the per-sequence host work of ONE 384-sequence call bounds the rate at ~117 k scans/s per GPU; three threads overlap
    # it: 136 k), no collective;
    # aggregate = all sequences of all ranks / the slowest rank's wall time between barriers (informational, `value` stays one sequence per GPU)
    batched_all = None
    if world > 1 and args.concurrent:
        G_b, S_b = 3, 128
        barrier()
        bl = lockstep_groups_leg(api, local, dev_args, G_b, S_b, K, W)
        tb = torch.tensor([G_b * S_b * K / bl["scans_per_s"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        batched_all = {"sequences_per_gpu": G_b * S_b, "lockstep_batches_per_gpu": G_b, "gpus": world, "scans_per_s": world * G_b * S_b * K / float(tb.item()),
                       "driver": bl["driver"], "timing": "slowest rank's wall clock around its leg"}

    # ---- point-sharded dense scan (configs[4], second half) at N > 1: every rank takes part, rank 0 reports -----------------------------
    sharded = None
    if world > 1:
        barrier()
        sharded = point_sharded_leg(args, rank, world, local, api, torch, dist)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel CUDA-event pass (rank 0): roofline of the K2 correspondence kernel + stage split ---------------
    L = capi.lib()
    odo3 = api.Odometry(ctx)
    for s in scans[:W]:
        odo3.process(s)
    L.b2lo_ctx_profile(ctx.h, 1)
    q3 = 0
    for i in range(W, W + K):
        flush.zero_()
        torch.cuda.synchronize()
        r = odo3.process_dev(*dev_args(i))
        q3 += r["n_features"] * r["n_iters"]
    names = ["K1_downsample", "K2_surfel_corr", "K4_pko_fit", "K4_pko_argmin", "K5_normal_eq_solve", "K6_map_update", "transform", "K3_knn"]
    stage = {}
    for i, nme in enumerate(names):
        ms, n = C.c_double(), C.c_longlong()
        L.b2lo_ctx_profile_read(ctx.h, i, C.byref(ms), C.byref(n))
        stage[nme] = {"ms_total": ms.value, "launches": n.value}
    L.b2lo_ctx_profile(ctx.h, 0)
    k2 = stage["K2_surfel_corr"]
    peak, peak_kind = peaks()
    k2_bytes_per_launch = ALGO_BYTES_PER_QUERY * q3 / max(k2["launches"], 1)
    k2_avg_s = 1e-3 * k2["ms_total"] / max(k2["launches"], 1)
    achieved = k2_bytes_per_launch / max(k2_avg_s, 1e-12) / 1e9
    roof = {"bound": "hbm", "kernel": "k_icp_corr (K2 surfel correspondence)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_kind": f"of {peak_kind}", "traffic": (k2_traffic().get("k2_kitti_scan") or {}).get("bytes_per_launch"),
            "traffic_source": (k2_traffic().get("k2_kitti_scan") or {}).get("source"),
            "algorithmic_bytes_per_launch": k2_bytes_per_launch,
            "avg_launch_us": 1e6 * k2_avg_s, "launches": k2["launches"],
            "note": "KITTI-shaped scans give ~4k queries per launch (~190 KB): the launch is latency-bound, not bandwidth-bound; see DESIGN.md"}
    dominant = max(stage.items(), key=lambda kv: kv[1]["ms_total"])[0]

    # ---- CPU baseline beside it: the oracle port on this box's host cores, same scans ------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        stage_cpu = {}
        arms = cpu_arms(scans, K, W, 1, 1, stage_cpu)
        kind, v = max(arms.items(), key=lambda kv: kv[1])
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "host_cores_available": os.cpu_count(), "arms": arms, "arms_note": CPU_ARMS_NOTE,
               "sample": f"the same {K} scans after {W} warm-up scans, single thread (the reference hot path is single-threaded)",
               "ms_per_scan": 1e3 / v, "stage_ms_per_scan": stage_cpu[kind], "stage_ms_per_scan_all": stage_cpu}

    dropin = dropin_leg(scans, K, W) if world == 1 else None
    if dropin is not None:
        dropin["reference_estimator"] = dropin_estimator_leg(scans, K, W)

    conc = None
    if world == 1 and args.concurrent:
        conc = [concurrent_leg(api, local, dev_args, S, K, W) for S in args.concurrent[:1]]
        conc += [batch_leg(api, local, dev_args, S, K, W) for S in args.concurrent]
    lockstep = None
    if world == 1 and args.lockstep:
        lockstep = [lockstep_leg(api, local, dev_args, S, K, W, peak, peak_kind) for S in args.lockstep]
        lockstep += [lockstep_groups_leg(api, local, dev_args, G, S, K, W) for G, S in ((3, 128),)]   # wall clock, host work included: three batches from three host threads

    stress = mid360 = export = None
    if world == 1 and not args.no_stress:
        export = export_leg(ctx, api, scans, true_poses, K)
        mid360 = mid360_leg(ctx, api, capi, flush, torch)
        stress = stress_leg(ctx, api, capi, int(args.stress_voxels), peak, peak_kind)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": max_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": value / 400.0, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "scans_per_rank": K, "seed": 42, "l2": "flushed between scans (256 MiB memset outside the timed region)",
                       "sequences": world, "parallelism": "one independent sequence per GPU (the same synthetic sequence replicated per rank), no collective",
                       "pipelining": "value and e2e are PIPELINED throughput of a recorded sequence: K1 of scan i+1 (announced with b2lo_odom_lookahead, an extension: the reference's player loads and processes serially) overlaps the registration of scan i; one K1 + one ICP + one map update per timed step, nothing skipped; `streaming` is the same sequence through the reference's own call pattern",
                       "launch": f"steady-state scans replay one CUDA graph of {graph['kernels_per_replay']} kernels ({graph['replays']} replays, {graph['builds']} build(s) in the e2e pass); one host sync per scan"},
            "ms_per_scan": max_ms / K, "correspondences_per_s": ncorr / (dev_ms * 1e-3), "queries_per_s": nq / (dev_ms * 1e-3),
            "gn_iterations_per_scan": iters / K, "keyframes": kf, "features_per_scan": nq / max(iters, 1), "map_l0": final_l0, "map_l1": final_l1,
            "wall_ms_per_scan_incl_flush": 1e3 * t_wall / K, "streaming_ms_per_scan": stream_ms / K,
            "streaming": {"what": "the reference's own call sequence (load a scan, then process it; no look-ahead): K1 in line with the registration",
                          "value": world * K / (stream_ms * 1e-3), "unit": UNIT, "ms_per_scan": stream_ms / K,
                          "e2e": {"value": world * K / stream_e2e_s, "unit": UNIT, "ms_per_scan": 1e3 * stream_e2e_s / K,
                                  "timing": "wall clock around b2lo_odom_process on page-locked host scans, copies inside"}},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_scan": 1e3 * float(te.item()) / K, "h2d_bytes_per_step": (h1 - h0) / K,
                    "d2h_bytes_per_step": (d1 - d0) / K},
            "dropin_e2e": dropin, "gpu_launches": launches, "clocks": clk.summary(), "roofline": roof, "stage_ms_per_scan": {k: v["ms_total"] / K for k, v in stage.items()},
            "dominant_kernel_group": dominant, "cpu_baseline": cpu,
            "concurrent_sequences_one_gpu": conc, "lockstep_sequences_one_gpu": lockstep, "batched_sequences_all_gpus": batched_all, "point_sharded": sharded, "final_map_export": export, "large_map_stress": stress, "kdtree_mid360": mid360}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
