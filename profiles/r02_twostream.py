"""Experiment: one GPU, the M samples of a plan split over NS engine handles on NS CUDA streams (sample-parallel shards on
ONE device, scores written into one shared matrix).  Question: do the kernels of one shard fill the SMs the other shard leaves
idle (conv tail rounds, latency-shaped pointwise kernels)?  Prints ms per plan for NS = 1, 2 (and 4)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from visual_foresight_b200 import spec as S  # noqa: E402
from visual_foresight_b200.distributed import EngineShard  # noqa: E402
from visual_foresight_b200.predictor import EngineBackend  # noqa: E402


def run(ns, M=200, plans=5, warm=3):
    cfg = B.CFG
    spec = S.spec_64(height=cfg["H"], width=cfg["W"], seq_len=cfg["S"], context_frames=cfg["C"], adim=cfg["adim"], sdim=cfg["sdim"])
    inp, w = B.synth(spec)
    kw = B.plan_kwargs(spec, M)
    goal = inp["goal"].astype(np.float32)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"]}
    local = M // ns
    streams = [torch.cuda.Stream() for _ in range(ns)]
    shards = []
    scores = torch.zeros((cfg["iters"], M), dtype=torch.float64, device="cuda")
    for i in range(ns):
        be = EngineBackend(spec, w, local, device=0)
        be.set_context(ctx)
        be.engine.set_desig(inp["desig"].astype(np.float32))
        sh = EngineShard(be, stream=streams[i])
        sh._scores_t = scores
        shards.append(sh)
    evs = [torch.cuda.Event() for _ in range(ns)]

    def plan(pi):
        for i, sh in enumerate(shards):
            if ns > 1:
                sh.begin(global_samples=M, offset=i * local, local=local, iterations=cfg["iters"], goal=goal, plan_index=pi, **kw)
            else:
                sh.begin(global_samples=M, offset=0, local=M, iterations=cfg["iters"], goal=goal, plan_index=pi, **kw)
        for it in range(cfg["iters"]):
            for sh in shards:
                sh.rollout(it)
            if ns > 1:
                for i in range(ns):
                    evs[i].record(streams[i])
                for i in range(ns):
                    for j in range(ns):
                        if i != j:
                            streams[i].wait_event(evs[j])
            for sh in shards:
                sh.select(it)
        return [sh.finish() for sh in shards]

    for i in range(warm):
        res = plan(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(plans):
        res = plan(10 + i)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / plans * 1e3
    for sh in shards:
        sh.engine.close()
    return ms, res[0]


if __name__ == "__main__":
    out = {}
    ref = None
    for ns in (1, 2, 4):
        ms, res = run(ns)
        out["ns%d_ms_per_plan" % ns] = ms
        if ref is None:
            ref = res
        else:
            out["ns%d_identical" % ns] = bool(np.array_equal(ref["scores"], res["scores"]) and np.array_equal(ref["best_actions"], res["best_actions"]))
    print(json.dumps(out))
