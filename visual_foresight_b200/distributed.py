"""Sample-parallel CEM over one process per GPU (torch.distributed).

The reference's multi-GPU scheme is in-graph towers: a contiguous slice of the action batch per
GPU, the context tiled, outputs concatenated in rank order, ``batch_size % ngpu == 0`` required
(``video_prediction/setup_predictor.py:34-44,70,117-123,155-162``).  Here the same contiguous split
runs as one rank per GPU.  Per CEM iteration the ONLY exchange is an all-gather of the M per-sample
scalar costs (float64); every rank then runs the identical stable top-K and refit, regenerating the
elites' actions from their global sample indices (counter-based noise), so no action tensor and no
broadcast crosses NVLink.

``ShardedCEMPlanner`` is backend-agnostic host logic: the product backend is ``EngineShard`` (device
scores, NCCL in-place all-gather on the engine's stream); CPU tests drive the same planner with an
oracle shard over ``gloo``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(global_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split, rank order (reference Tower: startidx = gpu_id * nsmp_per_gpu)."""
    if global_samples % world != 0:
        raise ValueError("num_samples (%d) must be divisible by the number of GPUs (%d)" % (global_samples, world))
    per = global_samples // world
    return rank * per, per


class ShardedCEMPlanner:
    """Drives begin / (rollout -> exchange -> select) x iterations / finish on one rank."""

    def __init__(self, shard, rank: int = 0, world: int = 1, group=None):
        self.shard, self.rank, self.world, self.group = shard, rank, world, group

    def plan(self, global_samples: int, iterations: int, **kw):
        offset, local = shard_range(global_samples, self.rank, self.world)
        self.shard.begin(global_samples=global_samples, offset=offset, local=local, iterations=iterations, **kw)
        for it in range(iterations):
            self.shard.rollout(it)
            if self.world > 1:
                self.shard.exchange(it, offset, local, self.group)
            self.shard.select(it)
        return self.shard.finish()


class EngineShard:
    """One rank's engine.  ``backend`` is a predictor.EngineBackend whose context is already set."""

    def __init__(self, backend, device_collective: bool = True, stream=None):
        self.backend = backend
        self.engine = backend.engine
        self.device_collective = device_collective
        self._scores_t = None
        self.stream = stream
        if device_collective:
            # The collective is ordered against a torch stream, so the engine must run on that same stream.  torch's
            # default stream has handle 0 (which vf_set_stream reads as "use your own stream"), hence a dedicated one.
            import torch
            if self.stream is None:
                self.stream = torch.cuda.Stream()
            self.engine.set_stream(self.stream.cuda_stream)

    def begin(self, *, global_samples, offset, local, iterations, goal, noise=None, **params):
        from .predictor import cem_params
        p = cem_params(self.backend.spec, num_samples=local, global_samples=global_samples, sample_offset=offset,
                       iterations=iterations, n_ctx_actions=self.backend._n_ctx_actions, **params)
        self._p = p
        if global_samples > local and self.device_collective:
            # the (iterations, global) float64 score matrix lives in a torch tensor the collective library owns;
            # the engine writes its shard straight into it (vf_cem_bind_scores)
            import torch
            if self._scores_t is None or tuple(self._scores_t.shape) != (iterations, global_samples):
                self._scores_t = torch.zeros((iterations, global_samples), dtype=torch.float64, device="cuda")
            self.engine.cem_bind_scores(self._scores_t.data_ptr())
        else:
            self.engine.cem_bind_scores(0)
        self.engine.cem_begin(p, goal, noise)

    def rollout(self, it):
        self.engine.cem_iter_rollout(it)

    def exchange(self, it, offset, local, group=None):
        import torch.distributed as dist
        if self.device_collective and dist.get_backend(group) == "nccl":
            import torch
            row = self._scores_t[it]
            # NCCL in-place all-gather, ordered on the engine's stream: each rank's segment already sits at
            # row[offset : offset+local]
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(row, row[offset:offset + local], group=group)
        else:           # host-staged exchange (gloo)
            import torch
            mine = torch.from_numpy(self.engine.cem_scores_read(it, offset, local))
            parts = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
            dist.all_gather(parts, mine, group=group)
            for r, t in enumerate(parts):
                self.engine.cem_scores_write(it, r * local, t.numpy())

    def select(self, it):
        self.engine.cem_iter_select(it)

    def finish(self):
        best, eidx, scores = self.engine.cem_finish()
        return {"best_actions": best, "elite_idx": eidx, "scores": scores}


def init_from_env(backend: Optional[str] = None):
    """torchrun environment -> (rank, world, local_rank); initialises the default process group."""
    import os

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        be = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if be == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=be, rank=rank, world_size=world)
    return rank, world, local_rank
