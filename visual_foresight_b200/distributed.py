"""Sample-parallel CEM over one process per GPU (torch.distributed).

The reference's multi-GPU scheme is in-graph towers: a contiguous slice of the action batch per
GPU, the context tiled, outputs concatenated in rank order, ``batch_size % ngpu == 0`` required
(``video_prediction/setup_predictor.py:34-44,70,117-123,155-162``).  Here the same contiguous split
runs as one rank per GPU.  Per CEM iteration the ONLY exchange is an all-gather of the M per-sample
scalar costs (float64); every rank then runs the identical stable top-K and refit, regenerating the
elites' actions from their global sample indices (counter-based noise), so no action tensor and no
broadcast crosses NVLink.

``ShardedCEMPlanner`` is backend-agnostic host logic: the product backend is ``EngineShard`` (device
scores; the exchange is the engine's own peer-memory kernel ``vf_cem_exchange``, with the NCCL in-place
all-gather kept as a comparison arm); CPU tests drive the same planner with an oracle shard over ``gloo``.
``ShardedBackend`` puts the planner behind the policy surface: ``PixelCostController(..., ngpu=N)`` under
torchrun shards its plans through it.
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


def connect_peer_exchange(engine, rank: int, world: int, max_iterations: int, max_global_samples: int, group=None):
    """Engine-owned exchange (include/vfengine.h, vf_comm_*): every rank exports its window descriptor, the 128-byte
    descriptors travel over the process group ONCE (setup time, any backend), every rank maps its peers' windows."""
    import torch.distributed as dist
    desc = engine.comm_export(max_iterations, max_global_samples)
    descs = [None] * world
    dist.all_gather_object(descs, desc, group=group)
    engine.comm_connect(rank, world, descs)


class EngineShard:
    """One rank's engine.  ``backend`` is a predictor.EngineBackend whose context is already set.

    collective: how the per-iteration scores cross ranks
      "peer" : the engine's own kernel over peer memory (vf_cem_exchange: P2P stores + arrival counters over NVLink,
               no collective library, no host round trip) — the product path; needs ``connect_peer_exchange`` (done
               lazily here on the first sharded plan);
      "nccl" : torch.distributed in-place all-gather on the engine's stream (comparison arm);
      "host" : host-staged all-gather (gloo; CPU-side tests of the host logic).
    """

    def __init__(self, backend, device_collective: bool = True, stream=None, collective=None, rank=0, world=1, group=None):
        self.backend = backend
        self.engine = backend.engine
        self.collective = collective or ("peer" if device_collective else "host")
        self.device_collective = self.collective != "host"
        self.rank, self.world, self.group = rank, world, group
        self._scores_t = None
        self._peer_cap = None
        self.stream = stream
        if self.collective == "nccl" or stream is not None:
            # The NCCL collective is ordered against a torch stream, so the engine must run on that same stream.  torch's
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
        sharded = global_samples > local
        if sharded and self.collective == "peer":
            cap = self._peer_cap
            if cap is None or iterations > cap[0] or iterations * global_samples > cap[0] * cap[1]:
                cap = (max(iterations, 8), max(global_samples, 1024))
                connect_peer_exchange(self.engine, self.rank, self.world, cap[0], cap[1], self.group)
                self._peer_cap = cap
            self.engine.cem_bind_scores(0)
        elif sharded and self.collective == "nccl":
            # the (iterations, global) float64 score matrix lives in a torch tensor the collective library owns;
            # the engine writes its shard straight into it (vf_cem_bind_scores)
            import torch
            if self._scores_t is None or tuple(self._scores_t.shape) != (iterations, global_samples):
                self._scores_t = torch.zeros((iterations, global_samples), dtype=torch.float64, device="cuda")
            self.engine.cem_bind_scores(self._scores_t.data_ptr())
        elif self._scores_t is not None and tuple(self._scores_t.shape) == (iterations, global_samples):
            self.engine.cem_bind_scores(self._scores_t.data_ptr())          # caller-provided shared matrix (same-device shards)
        else:
            self.engine.cem_bind_scores(0)
        self.engine.cem_begin(p, goal, noise)

    def rollout(self, it):
        self.engine.cem_iter_rollout(it)

    def exchange(self, it, offset, local, group=None):
        if self.collective == "peer":
            self.engine.cem_exchange(it)
            return
        import torch.distributed as dist
        if self.collective == "nccl" and dist.get_backend(group) == "nccl":
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


class ShardedBackend:
    """Policy-facing rollout evaluator of ONE rank of an ``ngpu``-way sample-parallel policy: the same surface as
    ``predictor.EngineBackend`` (plan / evaluate / fetch_*), every argument and result GLOBAL (all M samples), the work
    split in contiguous rank-order slices like the reference's towers (setup_predictor.py:34-44, M % ngpu == 0 :70)."""

    def __init__(self, backend, rank: int, world: int, group=None, collective: str = "peer"):
        self.backend, self.rank, self.world, self.group = backend, rank, world, group
        self.engine, self.spec = backend.engine, backend.spec
        self.n_context, self.sequence_length, self.n_cam = backend.n_context, backend.sequence_length, backend.n_cam
        self.shard = EngineShard(backend, collective=collective, rank=rank, world=world, group=group)
        self.planner = ShardedCEMPlanner(self.shard, rank, world, group)
        self._rollouts_per_rank = None

    @property
    def _n_ctx_actions(self):
        return self.backend._n_ctx_actions

    def set_context(self, context, legacy_actions=False):
        self.backend.set_context(context, legacy_actions)

    def plan(self, context, *, num_samples, iterations, goal_pix, noise=None, k_futures=1, **kw):
        shard_range(num_samples, self.rank, self.world)          # raises unless M % ngpu == 0 (setup_predictor.py:70)
        self.backend.set_context(context)
        self._rollouts_per_rank = (num_samples // self.world) * max(int(k_futures), 1)
        return self.planner.plan(num_samples, iterations, goal=goal_pix, noise=noise, k_futures=k_futures, **kw)

    # -- host plugin path: rank 0's sampled actions are the plan's actions (ranks need not share an RNG state) --------------
    def _bcast(self, obj, src=0):
        import torch.distributed as dist
        box = [obj if self.rank == src else None]
        dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def _gather_scores(self, local_scores):
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, local_scores, group=self.group)
        import numpy as np
        return np.concatenate(parts, axis=0)

    def _slice(self, actions):
        import numpy as np
        actions = np.asarray(self._bcast(np.asarray(actions)))
        off, local = shard_range(actions.shape[0], self.rank, self.world)
        self._rollouts_per_rank = local
        return actions[off:off + local]

    def evaluate(self, context, actions, goal_pix, finalweight, task_weights):
        return self._gather_scores(self.backend.evaluate(context, self._slice(actions), goal_pix, finalweight, task_weights))

    def evaluate_goal_image(self, context, actions, goal_image):
        return self._gather_scores(self.backend.evaluate_goal_image(context, self._slice(actions), goal_image))

    def score_external(self, distrib, goal_pix, finalweight, task_weights):
        return self.backend.score_external(distrib, goal_pix, finalweight, task_weights)

    # -- fetches by GLOBAL rollout index: the owning rank reads its device, everybody gets the arrays -----------------------
    def _owner(self, index):
        per = self._rollouts_per_rank
        if not per:
            raise RuntimeError("fetch before a rollout")
        return int(index) // per, int(index) % per

    def fetch_distrib(self, index: int):
        owner, local = self._owner(index)
        arr = self.backend.fetch_distrib(local) if owner == self.rank else None
        return self._bcast(arr, src=owner)

    def fetch_top(self, indices):
        import numpy as np
        mine = [(i, self._owner(g)[1]) for i, g in enumerate(indices) if self._owner(g)[0] == self.rank]
        got = self.backend.fetch_top(np.asarray([l for _, l in mine], np.int32)) if mine else (None, None)
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, ([i for i, _ in mine], got), group=self.group)
        frames = distrib = None
        for pos, (f, d) in parts:
            if not pos:
                continue
            if frames is None:
                frames = np.empty((len(indices),) + f.shape[1:], f.dtype)
                distrib = np.empty((len(indices),) + d.shape[1:], d.dtype)
            frames[pos], distrib[pos] = f, d
        return frames, distrib


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
