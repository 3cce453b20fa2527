"""Run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py
Every rank plans its contiguous shard (NCCL in-place all-gather of the float64 scores per CEM iteration); rank 0 also
plans the whole sample set alone.  Scores, elite indices and best actions must be BIT-identical on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from visual_foresight_b200 import spec as S  # noqa: E402
from visual_foresight_b200.distributed import EngineShard, ShardedCEMPlanner, init_from_env  # noqa: E402
from visual_foresight_b200.predictor import EngineBackend  # noqa: E402
from visual_foresight_b200.synthetic import synth_inputs  # noqa: E402
from visual_foresight_b200.hparams import HParams  # noqa: E402
from visual_foresight_b200.samplers import GaussianCEMSampler, action_bounds, per_dim_variance  # noqa: E402


def main():
    rank, world, local_rank = init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    sp = S.spec_64(height=32, width=32, seq_len=6)
    w = [S.init_weights(sp, 3, 0)]
    inp = synth_inputs(sp, 3)
    Mg, K, iters = 8 * world, 4, 3
    hp = HParams(**GaussianCEMSampler.get_default_hparams())
    lo, hi = action_bounds(hp, sp.adim)
    kw = dict(num_elites=K, nactions=5, repeat=3, std=np.sqrt(per_dim_variance(hp, sp.adim)), clip=(lo, hi), mean0=None,
              reduce_std_scale=1.0, finalweight=10.0, task_weights=None, seed=99, plan_index=2)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"]}
    goal = inp["goal"].astype(np.float32)

    def make(M):
        be = EngineBackend(sp, w, M, device=local_rank)
        be.set_context(ctx)
        be.engine.set_desig(inp["desig"].astype(np.float32))
        return be

    be = make(Mg // world)
    res = ShardedCEMPlanner(EngineShard(be), rank, world).plan(Mg, iters, goal=goal, **kw)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        full = make(Mg)
        ref = ShardedCEMPlanner(EngineShard(full), 0, 1).plan(Mg, iters, goal=goal, **kw)
        for k in ("scores", "elite_idx", "best_actions"):
            same = np.array_equal(res[k], ref[k])
            print("rank0 sharded vs single-GPU %-12s identical: %s (max abs diff %.3g)" % (k, same, float(np.abs(np.asarray(res[k], np.float64) - np.asarray(ref[k], np.float64)).max())))
            ok &= same
        ref_t = torch.from_numpy(ref["scores"]).cuda()
    else:
        ref_t = torch.empty((iters, Mg), dtype=torch.float64, device="cuda")
    dist.broadcast(ref_t, 0)
    same = np.array_equal(ref_t.cpu().numpy(), res["scores"])
    print("rank %d scores identical to the single-GPU plan: %s" % (rank, same))
    ok &= same
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if not int(flag.item()):
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU_CHECK_OK world=%d" % world)


if __name__ == "__main__":
    main()
