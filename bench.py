#!/usr/bin/env python
"""bench.py — CEM plans/sec & predicted frames/sec (BASELINE.json metric) for the visual-MPC hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision ...]

One "step" = one CEM plan: `iterations` x [sample M action sequences -> roll the CDNA/conv-LSTM predictor
S-1 cell steps -> pixel-distance cost of P predicted frames -> top-K elites -> refit].
N=1 workload = BASELINE config c2 (M=200, S=15, 64x64x3, 3 iterations, K=10).  N>1: weak scaling, every
rank keeps M=200 samples (global M = 200*N); the per-iteration exchange of the M float64 scores is the engine's own
peer-memory kernel (vf_cem_exchange; --collective nccl runs the NCCL all-gather arm instead).

value      : predicted frames/s, whole job, context already resident in HBM, timed with CUDA events on the
             engine's stream, max over ranks (plans_per_sec is reported beside it).
e2e        : same metric through the reference-facing plugin call, PixelCostController.act(**get_policy_args(...)) with
             HOST buffers (context H2D, chosen action / plan_stat D2H inside the timed region).
strong     : strong scaling at fixed global work, same events: c4 (BASELINE configs[3]: M=4096, K=205, sharded M/N per
             rank) and c2 (M=200 split over the N ranks); the driver's own scaling efficiency uses `value` (weak).
roofline   : conv-LSTM gate convolutions (dominant kernel class), CUDA-event time per launch from a
             profiled plan, ALGORITHMIC flops, vs MEASURED_PEAKS.json bf16 (sustained: timed inside a long step).
cpu_baseline: the oracle port (NumPy CEM + PyTorch-CPU predictor) on a bounded sample, rank 0, N=1 only.
--impl reference: the same oracle port as its own arm (the TF1 reference cannot run: SURVEY.md 8c); every step is ONE
             complete CEM iteration at the full M=200 (a measured third of a plan, no batch-size extrapolation).
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

CFG = dict(M=200, S=15, C=2, H=64, W=64, iters=3, K=10, nactions=5, repeat=3, adim=4, sdim=4)
METRIC = "CEM predicted frames/sec (M=200,H=15,64x64; plans/sec alongside)"


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def synth(spec, seed=0):
    from visual_foresight_b200 import spec as S
    from visual_foresight_b200.synthetic import synth_inputs
    inp = synth_inputs(spec, seed)
    w = [S.init_weights(spec, seed, v) for v in range(spec.ncam)]
    return inp, w


def plan_kwargs(spec, M):
    from visual_foresight_b200.hparams import HParams
    from visual_foresight_b200.samplers import GaussianCEMSampler, action_bounds, per_dim_variance
    hp = HParams(**GaussianCEMSampler.get_default_hparams())
    lo, hi = action_bounds(hp, spec.adim)
    return dict(num_elites=CFG["K"], nactions=CFG["nactions"], repeat=CFG["repeat"], std=np.sqrt(per_dim_variance(hp, spec.adim)),
                clip=(lo, hi), mean0=None, reduce_std_scale=1.0, finalweight=10.0, task_weights=None, seed=0)


# ---------------------------------------------------------------------------------------------------------
def cpu_port_plan_time(spec, inp, weights, sample_M, iters=1, threads=None):
    """Oracle port (reference NumPy CEM restated + PyTorch-CPU predictor) on `sample_M` samples for `iters`
    CEM iterations; returns seconds."""
    import torch
    import helpers as Hh
    from oracle import cem as OC
    if threads:
        torch.set_num_threads(threads)
    kw = plan_kwargs(spec, sample_M)
    K = min(CFG["K"], sample_M)
    noise = np.random.default_rng(0).standard_normal((iters, sample_M, 20)).astype(np.float32)

    def evaluate(actions):
        _, od, _ = Hh.oracle_rollout(spec, weights, inp, actions.astype(np.float32))
        return OC.eval_pixel_cost(od, inp["goal"])
    t0 = time.perf_counter()
    OC.cem_plan(evaluate, num_samples=sample_M, iterations=iters, num_elites_k=K, nactions=CFG["nactions"], repeat=CFG["repeat"],
                adim=spec.adim, std=kw["std"], noise=noise, clip=kw["clip"])
    return time.perf_counter() - t0


def cpu_baseline(spec, inp, weights, sample_M=200):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cpu_port_plan_time(spec, inp, weights, 2, 1)                       # warm-up (thread pool, oneDNN primitives)
    t = cpu_port_plan_time(spec, inp, weights, sample_M, 1)
    plan_s = t * (CFG["M"] / sample_M) * CFG["iters"]
    frames = CFG["iters"] * CFG["M"] * spec.n_pred * spec.ncam
    return {"value": frames / plan_s, "unit": "frames/s", "plans_per_sec": 1.0 / plan_s, "cores": torch.get_num_threads(),
            "kind": "port", "measured_s": t,
            "sample": "1 of the plan's %d CEM iterations on %d of %d samples (S=%d, %dx%d): %.1f s measured, x%.1f for the whole plan" %
            (CFG["iters"], sample_M, CFG["M"], spec.seq_len, spec.height, spec.width, t, CFG["M"] / sample_M * CFG["iters"])}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    from visual_foresight_b200 import spec as S
    spec = S.spec_64(height=CFG["H"], width=CFG["W"], seq_len=CFG["S"], context_frames=CFG["C"], adim=CFG["adim"], sdim=CFG["sdim"])
    inp, w = synth(spec)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_M = args.ref_samples
    for _ in range(args.warmup):
        cpu_port_plan_time(spec, inp, w, 2, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_plan_time(spec, inp, w, sample_M, 1)
    dt = (time.perf_counter() - t0) / args.steps                  # measured seconds of one step = one CEM iteration on sample_M samples
    scale = (CFG["M"] / sample_M) * CFG["iters"]
    plan_s = dt * scale
    frames = CFG["iters"] * CFG["M"] * spec.n_pred
    v = frames / plan_s
    sample = "each step = 1 of the plan's %d CEM iterations on %d of %d samples (measured %.2f s); plan time = x%.1f" % (CFG["iters"], sample_M, CFG["M"], dt, scale)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "plans_per_sec": 1.0 / plan_s, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "plan_ms": plan_s * 1e3, "step_fraction_of_plan": 1.0 / scale,
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c2: M=200 S=15 64x64x3 3 CEM iters K=10 pixel-distance cost", "sample": sample},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "TF1 reference cannot execute (tensorflow/video_prediction absent, SURVEY.md 8c): oracle port = reference NumPy CEM restated + PyTorch-CPU spec-P predictor; ms_per_step is the MEASURED step (one CEM iteration), value is per whole plan"}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
def make_policy(M_global, world, precision, collective):
    """The reference-facing plugin object: PixelCostController(ag_params, policyparams, gpu_id, ngpu) (sim/simulator.py:21)."""
    from visual_foresight_b200.cem_controller import PixelCostController
    ag = {"adim": CFG["adim"], "sdim": CFG["sdim"], "image_height": CFG["H"], "image_width": CFG["W"], "gpu_id": 0, "T": CFG["S"]}
    pp = {"rejection_sampling": False, "verbose": False}
    if M_global != 200:
        pp["num_samples"] = M_global
    if precision != "f16x3":
        pp["precision"] = precision
    if collective != "peer":
        pp["collective"] = collective
    pol = PixelCostController(ag, pp, 0, world)
    pol.reset()
    return pol


def timed_plans(planner, stream, M_global, kw, goal, steps, warmup, barrier, iters):
    """`steps` device-resident plans bracketed by CUDA events on the engine's stream; returns ms per plan (this rank)."""
    import torch
    for i in range(warmup):
        planner.plan(M_global, iters, goal=goal, plan_index=i, **kw)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        planner.plan(M_global, iters, goal=goal, plan_index=100 + i, **kw)
    e1.record(stream)
    barrier()
    return e0.elapsed_time(e1) / steps


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from visual_foresight_b200 import spec as S
    from visual_foresight_b200.distributed import EngineShard, ShardedCEMPlanner, init_from_env
    from visual_foresight_b200.policy import get_policy_args
    from visual_foresight_b200.predictor import EngineBackend
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, "visual_foresight_b200", "libvfengine.so")):
        ge.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    spec = S.spec_64(height=CFG["H"], width=CFG["W"], seq_len=CFG["S"], context_frames=CFG["C"], adim=CFG["adim"], sdim=CFG["sdim"])
    inp, w = synth(spec)
    M_local = CFG["M"] if not args.samples else args.samples
    M_global = M_local * world
    # ONE engine per rank, built by the policy object exactly as a user of the reference would (policy ctor -> predictor_class
    # -> restore); the device-resident leg drives the same handle below the policy surface
    pol = make_policy(M_global, world, args.precision, args.collective)
    be = pol._backend.backend if world > 1 else pol._backend
    stream = torch.cuda.Stream()                      # engine kernels, the exchange and the timing events share it
    be.engine.set_stream(stream.cuda_stream)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"]}
    desig = inp["desig"].astype(np.float32)
    kw = plan_kwargs(spec, M_local)
    goal = inp["goal"].astype(np.float32)
    if world > 1:
        shard = pol._backend.shard
        shard.stream = stream
    else:
        shard = EngineShard(be, collective="host", stream=stream)
    planner = ShardedCEMPlanner(shard, rank, world)
    frames_per_plan = CFG["iters"] * M_global * spec.n_pred * spec.ncam

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # e2e: the plugin call.  images / state histories as the agent hands them over (general_agent.py:85-153), t = C-1
    obs = {"images": inp["frames"], "state": np.asarray(inp["states"], np.float64)}
    step_data = {"desig_pix": inp["desig"].reshape(spec.ncam, spec.ndesig, 2), "goal_pix": inp["goal"].reshape(spec.ncam, spec.ndesig, 2)}

    def e2e_step():
        return pol.act(**get_policy_args(pol, obs, spec.context_frames - 1, 0, step_data))

    be.set_context(ctx)
    be.engine.set_desig(desig)
    # ---- value: device-resident ---------------------------------------------------------------------------
    for i in range(args.warmup):
        planner.plan(M_global, CFG["iters"], goal=goal, plan_index=i, **kw)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = be.engine.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        planner.plan(M_global, CFG["iters"], goal=goal, plan_index=100 + i, **kw)
    e1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = be.engine.launch_count() - l0
    ms = e0.elapsed_time(e1)
    # ---- e2e: host buffers in/out through PixelCostController.act --------------------------------------------------
    for i in range(2):
        out = e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_e2e0 = time.perf_counter()
    f0.record(stream)
    for i in range(args.steps):
        out = e2e_step()
    f1.record(stream)
    barrier()
    t_e2e_wall = time.perf_counter() - t_e2e0
    clocks = sampler.stop()
    ms_e2e = max(f0.elapsed_time(f1), t_e2e_wall * 1e3)          # act() returns host data: the wall clock bounds the event time
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    # ---- roofline: one profiled plan ---------------------------------------------------------------------------
    be.set_context(ctx)
    be.engine.set_desig(desig)
    be.engine.profile_enable(True)
    planner.plan(M_global, CFG["iters"], goal=goal, plan_index=300, **kw)
    prof = be.engine.profile_read()
    be.engine.profile_enable(False)
    peak_tf, peak_hbm, peak_src = measured_peaks()
    traffic, traffic_src = None, None                # DRAM bytes per gate-conv launch from the committed ncu --set full capture
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            traffic_src = "profiles/%s (ncu dram__bytes_read+write, mean of the 5 gate convs)" % name
            break
    fl = S.flops_per_sample_step(spec)
    # cell steps that run on all M samples: the shared-prefix steps (context frame + context action, identical for every
    # sample) run once on one sample and are NOT counted as work of the full-batch launches timed here
    gate_per_step = sum(1 for _, rnn in tuple(spec.encoder) + tuple(spec.decoder) if rnn)
    full_steps = prof["lstm_conv"]["launches"] // max(gate_per_step * CFG["iters"] * spec.ncam, 1)
    alg_lstm = fl["conv_lstm"] * full_steps * M_local * CFG["iters"]            # algorithmic flops of the timed gate-conv launches
    lstm_ms = prof["lstm_conv"]["ms"]
    ach = alg_lstm / (lstm_ms * 1e-3) / 1e12 if lstm_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "conv-LSTM gate convolution (%s)" % args.precision,
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "peak_source": peak_src + " bf16 sustained",
                "traffic": traffic, "traffic_source": traffic_src,
                "effective_peak_x3": peak_tf / 3.0, "frac_of_x3_peak": ach / (peak_tf / 3.0),
                "launches": prof["lstm_conv"]["launches"], "ms_per_launch": lstm_ms / max(prof["lstm_conv"]["launches"], 1),
                "share_of_step": lstm_ms / (ms / args.steps), "other_conv_ms": prof["other_conv"]["ms"],
                "algorithmic_flops_per_launch": alg_lstm / max(prof["lstm_conv"]["launches"], 1),
                "full_batch_steps": full_steps, "shared_prefix_steps": spec.n_steps - full_steps,
                "shared_prefix_conv_ms": prof["shared_prefix_conv"]["ms"],
                "whole_plan_frac": (S.flops_per_plan(spec, M_local, CFG["iters"]) * full_steps / spec.n_steps
                                    / (ms / args.steps * 1e-3) / 1e12) / peak_tf}
    # ---- strong scaling: fixed GLOBAL work, sharded M/N per rank (c4 = BASELINE configs[3]; c2 split over the ranks) -----------
    strong = None
    if not args.no_strong:
        strong = {}
        legs = [("c4", 4096, 205, max(2, min(args.steps, 3)))]
        if world > 1:
            legs.append(("c2", CFG["M"], CFG["K"], args.steps))
        for name, Mg, K, nsteps in legs:
            if Mg % world:
                strong[name] = {"skipped": "M=%d not divisible by %d ranks" % (Mg, world)}
                continue
            if world == 1 and os.environ.get("VF_BENCH_STRONG_M"):          # debug: a smaller single-GPU strong leg
                Mg = int(os.environ["VF_BENCH_STRONG_M"]); K = max(Mg // 20, 4)
            try:
                sbe = EngineBackend(spec, w, Mg // world, device=local_rank, precision=args.precision)
            except Exception as ex:                              # e.g. out of memory: report, keep the headline line
                strong[name] = {"error": str(ex)[:200]}
                continue
            sbe.engine.set_stream(stream.cuda_stream)
            sbe.set_context(ctx)
            sbe.engine.set_desig(desig)
            ssh = EngineShard(sbe, collective=args.collective if world > 1 else "host", stream=stream, rank=rank, world=world)
            skw = dict(kw)
            skw["num_elites"] = K
            sms = timed_plans(ShardedCEMPlanner(ssh, rank, world), stream, Mg, skw, goal, nsteps, 2, barrier, CFG["iters"])
            if world > 1:
                t = torch.tensor([sms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sms = float(t[0])
            fpp = CFG["iters"] * Mg * spec.n_pred * spec.ncam
            strong[name] = {"global_samples": Mg, "samples_per_gpu": Mg // world, "num_elites": K, "steps": nsteps, "ms_per_plan": sms,
                            "plans_per_sec": 1e3 / sms, "value": fpp / (sms * 1e-3), "unit": "frames/s", "scaling": "strong",
                            "whole_plan_frac_of_peak": S.flops_per_plan(spec, Mg // world, CFG["iters"]) / (sms * 1e-3) / 1e12 / peak_tf}
            sbe.engine.close()
    if rank == 0:
        step_ms = ms / args.steps
        value = frames_per_plan / (step_ms * 1e-3)
        e2e_v = frames_per_plan / (ms_e2e / args.steps * 1e-3)
        onehot_bytes = spec.context_frames * spec.ncam * spec.height * spec.width * spec.ndesig * 4
        h2d = inp["frames"].nbytes + np.asarray(inp["states"], np.float32).nbytes + np.asarray(inp["ctx_actions"], np.float32).nbytes + onehot_bytes + goal.nbytes
        d2h = CFG["K"] * CFG["nactions"] * CFG["repeat"] * spec.adim * 8 + CFG["K"] * 4 + CFG["iters"] * M_global * 8
        assert out["actions"].shape == (spec.adim,) and len(out["plan_stat"]) == CFG["iters"]
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "plans_per_sec": 1e3 / step_ms, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32_simt": "f32", "f16x3": "f32 (fp16 hi/lo split x3 on tcgen05, fp32 accumulate)", "f16x1": "f16 (fp32 accumulate)"}[args.precision],
                "data": "synthetic",
                "config": {"workload": "c2: M=%d/GPU (global %d) S=15 C=2 64x64x3 3 CEM iters K=10 pixel-distance cost, random-init spec-P CDNA/conv-LSTM" % (M_local, M_global),
                           "parallelism": "sample-parallel dp%d, per CEM iteration one exchange of M f64 scores (%s)" % (world, "engine peer-memory kernel over NVLink" if args.collective == "peer" else args.collective),
                           "l2": "per-step working set (~%d MB activations) exceeds the 126 MB L2; no flush" % int(M_local * 9.5),
                           "precision": args.precision},
                "e2e": {"value": e2e_v, "unit": "frames/s", "plans_per_sec": 1e3 / (ms_e2e / args.steps), "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "call": "PixelCostController.act(**get_policy_args(policy, obs, t, i_tr, step_data))"},
                "gpu_launches": int(launches), "wall_s": t_wall, "clocks": clocks, "roofline": roofline}
        if strong is not None:
            line["strong"] = strong
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(spec, inp, w, args.cpu_samples)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# optional lines for the other BASELINE configs (python bench.py --config c3|c5): same timing rules, planner-level calls
EXTRA = {
    # Sawyer two-view: M=600 S=13 2x(48x64), registration-weighted cost (ndesig=2, 1/warp-error trade-off), K = 5 % = 30, 13x1 actions
    "c3": dict(family="64", spec=dict(height=48, width=64, ncam=2, ndesig=2, adim=4, sdim=5, seq_len=13), M=600, K=30, iters=3,
               nactions=13, repeat=1, futures=1, task_err=[0.8, 2.5, 1.3, 0.4],
               workload="c3: M=600 S=13 2 views x 48x64, ndesig=2, registration-weighted pixel cost, 3 CEM iters K=30"),
    # stochastic SAVP predictor: M=200 x K=10 futures, S=15 128x128, nz=8 (recurrent latent); on N GPUs the 200 sequences are sharded,
    # on ONE GPU this line runs one 1/8 shard (25 sequences x 10 futures = 250 rollouts) of the 8-GPU config
    "c5": dict(family="128", spec=dict(adim=4, sdim=4, seq_len=15, nz=8, rnn_z=True), M=200, K=10, iters=3, nactions=5, repeat=3,
               futures=10, task_err=None,
               workload="c5: M=200 x 10 futures S=15 128x128 nz=8 stochastic predictor, mean + 0.5 var over futures, 3 CEM iters K=10"),
}


def run_extra(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from visual_foresight_b200 import spec as S
    from visual_foresight_b200.distributed import EngineShard, ShardedCEMPlanner, init_from_env
    from visual_foresight_b200.hparams import HParams
    from visual_foresight_b200.predictor import EngineBackend
    from visual_foresight_b200.samplers import GaussianCEMSampler, action_bounds, per_dim_variance
    c = EXTRA[args.config]
    init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    spec = (S.spec_128 if c["family"] == "128" else S.spec_64)(**c["spec"])
    inp, w = synth(spec)
    Mg = c["M"]
    shard_note = ""
    if args.config == "c5" and world == 1:
        Mg = c["M"] // 8
        shard_note = " [ONE 1/8 shard of the 8-GPU config: %d sequences x %d futures]" % (Mg, c["futures"])
    if Mg % world:
        raise SystemExit("M=%d not divisible by %d ranks" % (Mg, world))
    local = Mg // world
    be = EngineBackend(spec, w, local * c["futures"], device=local_rank, precision=args.precision)
    stream = torch.cuda.Stream()
    be.engine.set_stream(stream.cuda_stream)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"]}
    desig, goal = inp["desig"].astype(np.float32), inp["goal"].astype(np.float32)
    hp = HParams(**GaussianCEMSampler.get_default_hparams())
    lo, hi = action_bounds(hp, spec.adim)
    tw = None
    if c["task_err"]:
        e = np.asarray(c["task_err"], np.float64)
        tw = (1.0 / e) / (1.0 / e).sum()
    kw = dict(num_elites=min(c["K"], Mg), nactions=c["nactions"], repeat=c["repeat"], std=np.sqrt(per_dim_variance(hp, spec.adim)), clip=(lo, hi),
              mean0=None, reduce_std_scale=1.0, finalweight=10.0, task_weights=tw, seed=0, k_futures=c["futures"], lambda_variance=0.5 if c["futures"] > 1 else 0.0)
    shard = EngineShard(be, collective=args.collective if world > 1 else "host", stream=stream, rank=rank, world=world)
    planner = ShardedCEMPlanner(shard, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    be.set_context(ctx)
    be.engine.set_desig(desig)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = be.engine.launch_count()
    ms = timed_plans(planner, stream, Mg, kw, goal, args.steps, max(args.warmup, 2), barrier, c["iters"])
    launches = be.engine.launch_count() - l0
    # e2e: host context in, plan out, per step
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record(stream)
    for i in range(args.steps):
        be.set_context(ctx)
        be.engine.set_desig(desig)
        res = planner.plan(Mg, c["iters"], goal=goal, plan_index=200 + i, **kw)
    f1.record(stream)
    barrier()
    ms_e2e = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3) / args.steps
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    be.engine.profile_enable(True)
    planner.plan(Mg, c["iters"], goal=goal, plan_index=300, **kw)
    prof = be.engine.profile_read()
    be.engine.profile_enable(False)
    peak_tf, _, peak_src = measured_peaks()
    fl = S.flops_per_sample_step(spec)
    gate_per_step = sum(1 for _, rnn in tuple(spec.encoder) + tuple(spec.decoder) if rnn)
    nroll = local * c["futures"]
    full_steps = prof["lstm_conv"]["launches"] // max(gate_per_step * c["iters"] * spec.ncam, 1)
    alg = fl["conv_lstm"] * full_steps * nroll * c["iters"] * spec.ncam
    lstm_ms = prof["lstm_conv"]["ms"]
    ach = alg / (lstm_ms * 1e-3) / 1e12 if lstm_ms > 0 else 0.0
    if rank == 0:
        frames = c["iters"] * Mg * c["futures"] * spec.n_pred * spec.ncam
        line = {"metric": "CEM predicted frames/sec (%s; plans/sec alongside)" % args.config, "value": frames / (ms * 1e-3), "unit": "frames/s",
                "plans_per_sec": 1e3 / ms, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 2), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (fp16 hi/lo split x3 on tcgen05, fp32 accumulate)" if args.precision == "f16x3" else args.precision, "data": "synthetic",
                "config": {"workload": c["workload"] + shard_note, "parallelism": "sample-parallel dp%d" % world, "precision": args.precision,
                           "l2": "per-step working set exceeds the 126 MB L2; no flush"},
                "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "frames/s", "plans_per_sec": 1e3 / ms_e2e,
                        "h2d_bytes_per_step": int(inp["frames"].nbytes + np.asarray(inp["states"], np.float32).nbytes + desig.nbytes + goal.nbytes),
                        "d2h_bytes_per_step": int(res["best_actions"].nbytes + res["elite_idx"].nbytes + res["scores"].nbytes),
                        "call": "EngineBackend.set_context + ShardedCEMPlanner.plan (host buffers)"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "tensor", "kernel": "conv-LSTM gate convolution (%s)" % args.precision, "achieved": ach, "peak": peak_tf,
                             "unit": "TFLOP/s", "frac": ach / peak_tf, "peak_source": peak_src + " bf16 sustained", "traffic": None,
                             "launches": prof["lstm_conv"]["launches"], "ms_per_launch": lstm_ms / max(prof["lstm_conv"]["launches"], 1),
                             "share_of_step": lstm_ms / ms, "other_conv_ms": prof["other_conv"]["ms"],
                             "whole_plan_frac": S.flops_per_plan(spec, nroll, c["iters"]) * full_steps / spec.n_steps / (ms * 1e-3) / 1e12 / peak_tf}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("VF_PRECISION", "f16x3"), choices=["fp32_simt", "f16x3", "f16x1"])
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"], help="N>1: the engine's peer-memory exchange (default) or the NCCL all-gather arm")
    ap.add_argument("--samples", type=int, default=0, help="override per-GPU M (debug)")
    ap.add_argument("--cpu-samples", type=int, default=200, help="samples of the cpu_baseline leg (one full CEM iteration by default)")
    ap.add_argument("--ref-samples", type=int, default=200, help="samples per step of the reference arm (one full CEM iteration by default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling legs (c4 M=4096, c2 split)")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c5"], help="c2 = the headline workload (default); c3 / c5 = optional lines for the other BASELINE configs")
    args = ap.parse_args()
    rank, world, local_rank = env_rank()
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config != "c2":
        run_extra(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
