"""The point-sharded mode: one dense multi-LiDAR scan registered by several GPUs of one node, one process per GPU.

Every rank holds the same map replica and the contiguous query slice ``shard_bounds(m, world, rank)``; per Gauss-Newton iteration the
ranks exchange three small payloads - (count, sum r, sum r^2) per rank, the <= 128-entry PKO sample, the 28 normal-equation sums - and
every rank solves the identical 6x6 system.  ``PointShardedICP`` is optimize() for that mode with its three exchange implementations
(peer-memory mailboxes over NVLink, NCCL enqueued from C, host-driven torch.distributed); the small pure functions below are the host
arithmetic of the host-driven path and of the callers that split a cloud.  (Independent sequences, the other multi-GPU mode, need no
code here: rank r runs sequence r, nothing is exchanged, bench.py max-reduces the times.)
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi
from .capi import B2LO_OK, IcpStats, check


def shard_bounds(m: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of m queries for `rank`; the first m % world ranks get one extra query, so the
    concatenation over ranks is the original order (the PKO sample is drawn by position in that order)."""
    base, extra = divmod(m, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_plan(counts, rank: int):
    """Global offset of this rank's accepted correspondences and the global count, from the all-gathered counts."""
    counts = [int(c) for c in counts]
    return sum(counts[:rank]), sum(counts)


def scale_from_moments(total: int, sum_r: float, sum_r2: float) -> float:
    """Residual normalisation of iteration 0 (ICP.cpp:304-316): population sigma / 6 from the global raw moments."""
    mean = sum_r / total
    return math.sqrt(max(sum_r2 / total - mean * mean, 0.0)) / 6.0


def _api():
    from . import api      # api.py re-exports PointShardedICP at its end: import it lazily to keep the cycle harmless
    return api


def _cloud(*a, **k): return _api()._cloud(*a, **k)
def _f32(*a, **k): return _api()._f32(*a, **k)
def _icp_cfg(*a, **k): return _api()._icp_cfg(*a, **k)
def _trace(*a, **k): return _api()._trace(*a, **k)
def _p(*a, **k): return _api()._p(*a, **k)
def ICPConfig(*a, **k): return _api().ICPConfig(*a, **k)
def OptimizationStats(*a, **k): return _api().OptimizationStats(*a, **k)


class PointShardedICP:
    """optimize() for a dense scan whose queries are split across ranks; the map is replicated.
    Collectives (3 tiny ones per Gauss-Newton iteration) go through torch.distributed on the caller's process group;
    with ``group=None`` and an uninitialised torch.distributed the exchange is the identity (single rank)."""

    def __init__(self, config=None, adaptive_estimator=None, group=None, device_ordered=True, exchange="nccl"):
        """exchange="peer" (with device_ordered=True): the three exchanges per iteration are one small kernel each that stores straight into
        the peers' mailboxes over NVLink (cudaIpc peer memory, b2lo_shard_comm_open_peers) - no NCCL at all; "nccl": NCCL calls on the stream.
        device_ordered=True: the whole loop is ONE C call (b2lo_icp_shard_optimize) whose three exchanges per iteration are NCCL calls
        enqueued on the context stream between the kernels - no host round trip inside the loop; torch.distributed only carries the 128-byte
        NCCL unique id once.  device_ordered=False: the host-driven phase API (b2lo_icp_shard_corr / _sample / _accumulate / _finish with
        torch.distributed collectives between them), kept as the readable specification of the exchange and for A/B runs."""
        self.m_config = config or ICPConfig()
        self.m_adaptive_estimator = adaptive_estimator
        self.group = group
        self.device_ordered = device_ordered
        if exchange not in ("nccl", "peer"):
            raise ValueError("exchange must be 'nccl' or 'peer'")
        self.exchange = exchange
        self.m_last_stats = OptimizationStats()
        self.collective_seconds = 0.0
        self._comm = None
        self._comm_ctx = None

    def __del__(self):
        if getattr(self, "_comm", None):
            try:
                capi.lib().b2lo_shard_comm_destroy(self._comm)
            except Exception:
                pass
            self._comm = None

    def _communicator(self, ctx):
        """b2lo_shard_comm of this rank on ``ctx`` (created once): rank 0's NCCL unique id travels through torch.distributed."""
        if self._comm is not None and self._comm_ctx is ctx:
            return self._comm
        import torch
        import torch.distributed as dist
        L = capi.lib()
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        rank = dist.get_rank(self.group) if multi else 0
        world = dist.get_world_size(self.group) if multi else 1
        uid = np.zeros(128, np.uint8)
        peer = multi and self.exchange == "peer"
        on_gpu = multi and dist.get_backend(self.group) == "nccl"
        if multi and not peer:
            if rank == 0:
                check(L.b2lo_shard_unique_id(_p(uid), 128))
            t = torch.from_numpy(uid).cuda(ctx.device) if on_gpu else torch.from_numpy(uid)
            dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            uid = t.cpu().numpy().copy()
        h = C.c_void_p()
        check(L.b2lo_shard_comm_create(ctx.h, world, rank, None if peer else _p(uid), 128, C.byref(h)))
        if peer:    # every rank's mailbox handle (64 bytes) to every rank, in rank order
            mine = np.zeros(64, np.uint8)
            check(L.b2lo_shard_comm_ipc_handle(h, _p(mine), 64))
            t = torch.from_numpy(mine).cuda(ctx.device) if on_gpu else torch.from_numpy(mine)
            got = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(got, t, group=self.group)
            handles = np.ascontiguousarray(np.stack([g.cpu().numpy() for g in got]))
            rc = L.b2lo_shard_comm_open_peers(h, _p(handles), 64)
            why = capi.last_error() if rc < 0 else ""
            # every rank learns whether ALL ranks mapped all mailboxes (this is also the barrier before the first store into one):
            # a rank that could not must not leave the others spinning on its payload
            flag = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32)
            flag = flag.cuda(ctx.device) if on_gpu else flag
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 0:
                L.b2lo_shard_comm_destroy(h)
                raise capi.B2loError(rc if rc < 0 else -1, "peer-memory exchange unavailable: " + (why or "another rank could not map the mailboxes (no P2P access between the GPUs?)"))
        self._comm, self._comm_ctx = h, ctx
        return h

    def _optimize_device_ordered(self, voxel_map, cloud_shard, initial_transform):
        a, n, sf = _cloud(cloud_shard)
        T0 = _f32(initial_transform).reshape(16)
        ame = self.m_adaptive_estimator.get_config() if self.m_adaptive_estimator else None
        cfg = _icp_cfg(self.m_config, ame)
        comm = self._communicator(voxel_map.ctx)
        Tout = np.zeros(16, np.float32)
        st = IcpStats()
        cms = C.c_float(0.0)
        rc = check(capi.lib().b2lo_icp_shard_optimize(voxel_map.h, comm, _p(a), n, sf, _p(T0), C.byref(cfg), _p(Tout), C.byref(st), C.byref(cms)))
        self.collective_ms_last_iteration = float(cms.value)
        self.device_ms = float(st.device_ms)
        self.collective_seconds = 0.0
        if rc != B2LO_OK:
            self.m_last_stats = OptimizationStats()
            return False, T0.reshape(4, 4).copy()
        self.m_last_stats = OptimizationStats(num_iterations=st.num_iterations, num_correspondences=st.num_correspondences,
                                              initial_cost=st.initial_cost, final_cost=st.final_cost, converged=True,
                                              optimization_time_ms=st.device_ms, iterations=_trace(st))
        return True, Tout.reshape(4, 4).copy()

    def get_last_stats(self):
        return self.m_last_stats

    def optimize(self, voxel_map, cloud_shard, initial_transform):
        if self.device_ordered:
            return self._optimize_device_ordered(voxel_map, cloud_shard, initial_transform)
        import time
        import torch
        import torch.distributed as dist
        L = capi.lib()
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        rank = dist.get_rank(self.group) if multi else 0
        world = dist.get_world_size(self.group) if multi else 1
        a, n, sf = _cloud(cloud_shard)
        T0 = _f32(initial_transform).reshape(16)
        ame = self.m_adaptive_estimator.get_config() if self.m_adaptive_estimator else None
        cfg = _icp_cfg(self.m_config, ame)
        dev = torch.device("cuda", voxel_map.ctx.device)
        stats3 = torch.zeros(3, dtype=torch.float64, device=dev)
        sample = torch.zeros(128, dtype=torch.float64, device=dev)
        acc28 = torch.zeros(28, dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        check(L.b2lo_icp_shard_begin(voxel_map.h, _p(a), n, sf, _p(T0), C.byref(cfg)))
        Tout = np.zeros(16, np.float32)
        st = IcpStats()
        scale = 1.0
        self.collective_seconds = 0.0
        ok = True
        for it in range(self.m_config.max_iterations):
            check(L.b2lo_icp_shard_corr(voxel_map.h, C.byref(cfg), C.c_void_p(stats3.data_ptr())))
            voxel_map.ctx.sync()
            t0 = time.perf_counter()
            if multi:
                gathered = [torch.zeros_like(stats3) for _ in range(world)]
                dist.all_gather(gathered, stats3, group=self.group)
                g = torch.stack(gathered).cpu().numpy()
            else:
                g = stats3.cpu().numpy()[None, :]
            self.collective_seconds += time.perf_counter() - t0
            offset, total = shard_plan(g[:, 0], rank)
            if total < self.m_config.min_correspondence_points:   # ICP.cpp:298-302: false, output = initial
                ok = False
                break
            if it == 0:
                scale = scale_from_moments(total, float(g[:, 1].sum()), float(g[:, 2].sum()))
            check(L.b2lo_icp_shard_sample(voxel_map.h, C.byref(cfg), offset, total, C.c_double(scale), C.c_void_p(sample.data_ptr())))
            voxel_map.ctx.sync()
            t0 = time.perf_counter()
            if multi:
                dist.all_reduce(sample, group=self.group)
                torch.cuda.synchronize(dev)
            self.collective_seconds += time.perf_counter() - t0
            check(L.b2lo_icp_shard_accumulate(voxel_map.h, C.byref(cfg), total, C.c_double(scale), C.c_void_p(sample.data_ptr()),
                                              C.c_void_p(acc28.data_ptr())))
            voxel_map.ctx.sync()
            t0 = time.perf_counter()
            if multi:
                dist.all_reduce(acc28, group=self.group)
                torch.cuda.synchronize(dev)
            self.collective_seconds += time.perf_counter() - t0
            done = C.c_int(0)
            check(L.b2lo_icp_shard_finish(voxel_map.h, C.byref(cfg), C.c_void_p(acc28.data_ptr()), _p(Tout), C.byref(done), C.byref(st)))
            if done.value:
                break
        if not ok:
            self.m_last_stats = OptimizationStats()
            return False, T0.reshape(4, 4).copy()
        self.m_last_stats = OptimizationStats(num_iterations=st.num_iterations, num_correspondences=st.num_correspondences,
                                              initial_cost=st.initial_cost, final_cost=st.final_cost, converged=True, iterations=_trace(st))
        return True, Tout.reshape(4, 4).copy()
