"""Landmark-sharded multi-GPU global bundle adjustment (BASELINE config 5) -- one process per GPU.

Everything numeric AND the collective live under the C-ABI (csrc/ba_dist.cu: NCCL bound with dlopen, all-reduce of the compact
block-CSR reduced camera system once per LM iteration, replicated bit-identical PCG).  What this module does is rendezvous only:
rank 0 asks the C-ABI for an NCCL unique id and torch.distributed (any backend; gloo works) carries the 128 bytes to the other
ranks.  The single-process front end (one process driving N GPUs, what the optimizer plugin uses) is api.ba_solve_multi.
"""
from __future__ import annotations

import numpy as np

from .synth import BAProblem


def shard_bounds(pb: BAProblem, world: int):
    """The landmark ranges the C-ABI uses (contiguous, balanced by observation count) -- restated for tests."""
    counts = np.bincount(pb.obs_point, minlength=pb.n_points).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = int(csum[-1])
    b = [0]
    for r in range(1, world):
        b.append(max(int(np.searchsorted(csum, total * r / world, side="left")), b[-1]))
    b.append(pb.n_points)
    return b


def shard_landmarks(pb: BAProblem, rank: int, world: int):
    """The rank's local problem as the C-ABI builds it internally: (local problem, global ids of the local points)."""
    b = shard_bounds(pb, world)
    lo, hi = b[rank], b[rank + 1]
    ids = np.arange(lo, hi, dtype=np.int64)
    sel = (pb.obs_point >= lo) & (pb.obs_point < hi)
    local = BAProblem(cam_pose_wc=pb.cam_pose_wc.copy(), cam_dof=pb.cam_dof.copy(),
                      points=np.ascontiguousarray(pb.points[lo:hi]), point_free=np.ascontiguousarray(pb.point_free[lo:hi]),
                      obs_cam=np.ascontiguousarray(pb.obs_cam[sel]), obs_point=np.ascontiguousarray(pb.obs_point[sel] - lo).astype(np.int32),
                      obs_xyz=np.ascontiguousarray(pb.obs_xyz[sel]),
                      obs_info=None if pb.obs_info is None else np.ascontiguousarray(pb.obs_info[sel]))
    return local, ids


def broadcast_unique_id(rank: int, world: int, group=None) -> bytes | None:
    """Rank 0 creates the NCCL unique id under the C-ABI; torch.distributed broadcasts it (CPU tensor on gloo, CUDA on nccl)."""
    if world == 1:
        return None
    import torch
    import torch.distributed as dist
    from .api import Comm
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.frombuffer(bytearray(Comm.unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().numpy().tobytes())


class DistributedBA:
    """One rank of the landmark-sharded solve: Comm + ShardedBAGraph."""

    def __init__(self, ctx, pb: BAProblem, rank: int, world: int, group=None, unique_id: bytes | None = None):
        from .api import Comm, ShardedBAGraph
        self.ctx, self.rank, self.world = ctx, rank, world
        if world > 1 and unique_id is None:
            unique_id = broadcast_unique_id(rank, world, group)
        self.comm = Comm(ctx, world, rank, unique_id)
        self.graph = ShardedBAGraph(self.comm, pb)
        self.ids = np.arange(self.graph.lo, self.graph.hi, dtype=np.int64)
        self.reduce_bytes = self.graph.reduce_bytes

    def solve(self, cfg):
        return self.graph.solve(cfg)

    def download(self):
        """(poses, local points, their global ids)"""
        pose, pts = self.graph.download()
        return pose, pts, self.ids

    def close(self):
        self.graph.close()
        self.comm.close()
