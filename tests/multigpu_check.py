"""Run under torchrun, one process per rank:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py
With N GPUs every rank drives its own GPU (NCCL process group).  With VF_CHECK_SAME_DEVICE=1 every rank drives cuda:0 (gloo
process group; the windows of the engine's peer exchange are then mapped ACROSS PROCESSES on one device with cudaIpc) —
this is how the 1-GPU round-end box exercises the multi-rank path.

Checked, bit for bit, on every rank:
  1. ShardedCEMPlanner over the engine's own peer-memory exchange (vf_cem_exchange), and over the comparison arm
     (NCCL in-place all-gather, or the host-staged exchange when the ranks share a device): scores, elite indices and best
     actions equal the single-rank plan of all samples;
  2. the policy surface: PixelCostController(ag, pp, gpu_id, ngpu=world).act() on all ranks returns the actions and
     plan_stat of the 1-GPU policy (setup_predictor.py:34-44 towers -> one process per GPU)."""
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


def bcast(obj, rank):
    box = [obj if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def main():
    same = os.environ.get("VF_CHECK_SAME_DEVICE", "0") == "1"
    rank, world, local_rank = init_from_env("gloo" if same else "nccl")
    dev = 0 if same else local_rank
    torch.cuda.set_device(dev)
    sp = S.spec_64(height=32, width=32, seq_len=6)
    w = [S.init_weights(sp, 3, 0)]
    inp = synth_inputs(sp, 3)
    Mg, K, iters = 8 * world, 4, 3
    hp = HParams(**GaussianCEMSampler.get_default_hparams())
    lo, hi = action_bounds(hp, sp.adim)
    kw = dict(num_elites=K, nactions=5, repeat=3, std=np.sqrt(per_dim_variance(hp, sp.adim)), clip=(lo, hi), mean0=None,
              reduce_std_scale=1.0, finalweight=10.0, task_weights=None, seed=99)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"]}
    goal = inp["goal"].astype(np.float32)
    ok = True

    def make(M):
        be = EngineBackend(sp, w, M, device=dev)
        be.set_context(ctx)
        be.engine.set_desig(inp["desig"].astype(np.float32))
        return be

    # ---- 1. planner level ------------------------------------------------------------------------------------------
    ref = None
    if rank == 0:
        full = make(Mg)
        single = ShardedCEMPlanner(EngineShard(full, collective="host"), 0, 1)
        ref = [single.plan(Mg, iters, goal=goal, plan_index=pi, **kw) for pi in (2, 3, 4)]
        full.engine.close()
    ref = bcast(ref, rank)
    for coll in (("peer", "host") if same else ("peer", "nccl")):
        be = make(Mg // world)
        planner = ShardedCEMPlanner(EngineShard(be, collective=coll, rank=rank, world=world), rank, world)
        for j, pi in enumerate((2, 3, 4)):                       # consecutive plans reuse the exchange window
            res = planner.plan(Mg, iters, goal=goal, plan_index=pi, **kw)
            for k in ("scores", "elite_idx", "best_actions"):
                good = np.array_equal(res[k], ref[j][k])
                ok &= good
                if not good or (rank == 0 and j == 0):
                    print("rank %d [%s] plan %d %-12s identical to the single-rank plan: %s" % (rank, coll, pi, k, good))
        torch.cuda.synchronize()
        dist.barrier()
        be.engine.close()

    # ---- 2. policy surface -----------------------------------------------------------------------------------------
    from visual_foresight_b200.cem_controller import PixelCostController
    from visual_foresight_b200.policy import get_policy_args
    ag = {"adim": 4, "sdim": 4, "image_height": 32, "image_width": 32, "gpu_id": 0}
    pp = {"rejection_sampling": False, "verbose": False, "num_samples": Mg, "minimum_selection": 4,
          "model_spec": {"seq_len": 6}, "cem_seed": 5, "predictor_propagation": True}
    rng = np.random.default_rng(0)
    images = rng.integers(0, 256, (4, 1, 32, 32, 3), dtype=np.uint8)
    state = rng.uniform(-.5, .5, (4, 4))

    def drive(pol):
        pol.reset()
        outs = []
        for t in range(4):
            obs = {"images": images[:t + 1], "state": state[:t + 1]}
            o = pol.act(**get_policy_args(pol, obs, t, 0, {"desig_pix": np.array([[8, 8]]), "goal_pix": np.array([[24, 20]])}))
            outs.append((o["actions"].copy(), {k: v.copy() for k, v in o["plan_stat"].items()}))
        return outs

    ref_pol = None
    if rank == 0:
        p1 = PixelCostController(ag, dict(pp), dev, 1)
        ref_pol = drive(p1)
        p1.predictor.backend.engine.close()
    ref_pol = bcast(ref_pol, rank)
    ppn = dict(pp)
    if same:
        ppn["shard_devices"] = [0] * world
    pol = PixelCostController(ag, ppn, 0, world)
    got = drive(pol)
    for t, ((a, st), (ra, rst)) in enumerate(zip(got, ref_pol)):
        good = np.array_equal(a, ra) and sorted(st) == sorted(rst) and all(np.array_equal(st[k], rst[k]) for k in st)
        ok &= good
        if not good or rank == 0:
            print("rank %d policy.act(t=%d) identical to the 1-GPU policy: %s" % (rank, t, good))
    torch.cuda.synchronize()

    flag = torch.tensor([int(ok)], device="cuda" if not same else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if not int(flag.item()):
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU_CHECK_OK world=%d%s" % (world, " (ranks share cuda:0)" if same else ""))


if __name__ == "__main__":
    main()
