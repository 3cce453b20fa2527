"""N>1 host logic on CPU: world_size-2 gloo run of ShardedCEMPlanner with an oracle shard must give
bit-identical scores / elite sets / best actions to the single-rank plan (sample-index-keyed noise,
contiguous split, one score all-gather per iteration, redundant deterministic selection)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cem as OC
from visual_foresight_b200.distributed import ShardedCEMPlanner, shard_range

M, K, ITERS, NACT, REP, ADIM = 16, 4, 3, 5, 3, 4
STD = np.array([0.05, 0.05, 0.15, np.pi / 18])
SEED = 77


def _score(actions):
    """cheap deterministic stand-in for rollout+cost: per-sample function of its own actions only"""
    pos = np.cumsum(actions[:, :13, :2], axis=1) * 40.0
    return np.linalg.norm(pos[:, -1] - np.array([5.0, -3.0]), axis=1) + 0.1 * np.abs(actions[:, :, 2]).sum(1)


class OracleShard:
    """Oracle restatement of one rank: Philox noise keyed by GLOBAL sample index (oracle/cem.py)."""

    def begin(self, *, global_samples, offset, local, iterations, **kw):
        self.Mg, self.off, self.local, self.iters = global_samples, offset, local, iterations
        self.scores = np.zeros((iterations, global_samples))
        self.mean, self.factor = np.zeros(NACT * ADIM), None

    def _actions(self, it, gidx):
        D = NACT * ADIM
        n = D if it == 0 else K
        z = np.array([[OC.philox_normal(SEED, 0, it, int(g), j) for j in range(n)] for g in gidx])
        x = OC.sample_diag(self.mean, np.tile(STD, NACT), z) if it == 0 else OC.sample_from_factor(self.mean, self.factor, z)
        return OC.finish_actions(x, NACT, ADIM, REP, (np.array([-.1, -.1, -np.inf, -np.pi / 4]), np.array([.1, .1, np.inf, np.pi / 4])))

    def rollout(self, it):
        acts, _ = self._actions(it, np.arange(self.off, self.off + self.local))
        self.scores[it, self.off:self.off + self.local] = _score(acts)

    def exchange(self, it, offset, local, group=None):
        mine = torch.from_numpy(self.scores[it, offset:offset + local].copy())
        parts = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, mine, group=group)
        for r, t in enumerate(parts):
            self.scores[it, r * local:(r + 1) * local] = t.numpy()

    def select(self, it):
        self.idx = OC.elite_select(self.scores[it], K)
        self.best, x_nr = self._actions(it, self.idx)        # elites regenerated from their global indices
        if it < self.iters - 1:
            self.mean, self.factor = x_nr.mean(0), OC.elite_factor(x_nr)

    def finish(self):
        return {"best_actions": self.best, "elite_idx": self.idx, "scores": self.scores}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = ShardedCEMPlanner(OracleShard(), rank, world).plan(M, ITERS)
    q.put((rank, res["scores"], res["elite_idx"], res["best_actions"]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range():
    assert shard_range(200, 3, 8) == (75, 25)
    with pytest.raises(ValueError):
        shard_range(200, 0, 3)        # reference: assert batch_size % ngpu == 0 (setup_predictor.py:70)


@pytest.mark.timeout(300)
def test_two_rank_gloo_plan_equals_single_rank():
    single = ShardedCEMPlanner(OracleShard(), 0, 1).plan(M, ITERS)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, scores, idx, best in got:
        np.testing.assert_array_equal(scores, single["scores"])
        np.testing.assert_array_equal(idx, single["elite_idx"])
        np.testing.assert_array_equal(best, single["best_actions"])
