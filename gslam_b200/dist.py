"""Landmark-sharded multi-GPU global bundle adjustment (BASELINE config 5).

Every rank holds ALL cameras and a shard of the landmarks with all their edges.  One LM iteration has exactly one real
exchange: the sum over ranks of the shard's Schur contribution [S | g~ | diag U | cost] (one NCCL all-reduce, f64), plus a
1-double all-reduce of the candidate cost.  PCG on the reduced camera system is replicated and deterministic, so every rank
takes bit-identical accept/reject decisions without any broadcast.  torch.distributed is plumbing only: the buffers are
plain device memory handed to the C-ABI stepwise entry points (gb_ba_graph_reduce_local / _step / _commit).
"""
from __future__ import annotations

import numpy as np

from .synth import BAProblem


def shard_landmarks(pb: BAProblem, rank: int, world: int):
    """Contiguous landmark ranges balanced by observation count.  Returns (local problem, global ids of the local points)."""
    counts = np.bincount(pb.obs_point, minlength=pb.n_points).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = csum[-1]
    # boundary b_r = first point whose prefix count reaches r/world of the edges
    bounds = [int(np.searchsorted(csum, total * r / world, side="left")) for r in range(world + 1)]
    bounds[0], bounds[-1] = 0, pb.n_points
    for r in range(1, world + 1):
        bounds[r] = max(bounds[r], bounds[r - 1])
    lo, hi = bounds[rank], bounds[rank + 1]
    ids = np.arange(lo, hi, dtype=np.int64)
    sel = (pb.obs_point >= lo) & (pb.obs_point < hi)
    local = BAProblem(cam_pose_wc=pb.cam_pose_wc.copy(), cam_dof=pb.cam_dof.copy(),
                      points=np.ascontiguousarray(pb.points[lo:hi]), point_free=np.ascontiguousarray(pb.point_free[lo:hi]),
                      obs_cam=np.ascontiguousarray(pb.obs_cam[sel]), obs_point=np.ascontiguousarray(pb.obs_point[sel] - lo).astype(np.int32),
                      obs_xyz=np.ascontiguousarray(pb.obs_xyz[sel]),
                      obs_info=None if pb.obs_info is None else np.ascontiguousarray(pb.obs_info[sel]))
    return local, ids


class DistributedBA:
    """One rank of the landmark-sharded solve.  `group` is a torch.distributed process group (NCCL on GPUs)."""

    def __init__(self, ctx, pb: BAProblem, rank: int, world: int, group=None):
        import torch
        import torch.distributed as dist
        from .api import BAGraph
        self.torch, self.dist = torch, dist
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.n_points_global = pb.n_points
        self.local, self.ids = shard_landmarks(pb, rank, world)
        self.graph = BAGraph(ctx, self.local)
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", ctx.device))
        n = self.graph.reduce_size()
        self.buf = torch.zeros(n, dtype=torch.float64, device=torch.device("cuda", ctx.device))
        self.cost = torch.zeros(8, dtype=torch.float64, device=torch.device("cuda", ctx.device))
        self.reduce_bytes = n * 8

    def solve(self, cfg):
        torch, dist = self.torch, self.dist
        g = self.graph
        with torch.cuda.stream(self.stream):  # collectives are ordered after / before our kernels on the ctx stream
            g.begin(cfg)
            for _ in range(cfg.maxIterations):
                g.reduce_local(self.buf.data_ptr())
                if self.world > 1:
                    dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
                g.step(self.buf.data_ptr(), self.cost.data_ptr())
                if self.world > 1:
                    dist.all_reduce(self.cost, op=dist.ReduceOp.SUM, group=self.group)
                g.commit(self.buf.data_ptr(), self.cost.data_ptr())
            return g.finish()

    def download(self):
        """(poses, local points, their global ids)"""
        pose, pts = self.graph.download()
        return pose, pts, self.ids

    def close(self):
        self.graph.close()
