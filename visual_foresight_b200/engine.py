"""ctypes binding of libvfengine.so (include/vfengine.h).  There is NO CPU path: importing works
anywhere (so host logic can be unit-tested), but constructing an ``Engine`` without the built
library or without a CUDA device raises ``EngineUnavailable``."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np

from .spec import MAX_LAYERS, PredictorSpec

VF_ABI_VERSION = 2
VF_MAX_TASKS = 16
VF_PEER_DESC_BYTES = 128
PREC_FP32_SIMT, PREC_F16X3, PREC_F16X1 = 0, 1, 2
COST_PIXEL_DISTANCE, COST_GOAL_IMAGE = 0, 1
SAMPLER_GAUSSIAN, SAMPLER_CORRELATED = 0, 1
PRECISIONS = {"fp32_simt": PREC_FP32_SIMT, "f16x3": PREC_F16X3, "f16x1": PREC_F16X1}

LIB_PATH = os.environ.get("VF_ENGINE_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvfengine.so")


class EngineUnavailable(RuntimeError):
    pass


class EngineError(RuntimeError):
    pass


class VfConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("ncam", C.c_int32), ("ndesig", C.c_int32),
        ("adim", C.c_int32), ("sdim", C.c_int32), ("nz", C.c_int32),
        ("seq_len", C.c_int32), ("context_frames", C.c_int32), ("ngf", C.c_int32),
        ("n_enc", C.c_int32), ("enc_channels", C.c_int32 * MAX_LAYERS), ("enc_rnn", C.c_int32 * MAX_LAYERS),
        ("n_dec", C.c_int32), ("dec_channels", C.c_int32 * MAX_LAYERS), ("dec_rnn", C.c_int32 * MAX_LAYERS),
        ("num_transformed", C.c_int32), ("cdna_ksize", C.c_int32), ("lstm_ksize", C.c_int32),
        ("norm_eps", C.c_float), ("forget_bias", C.c_float),
        ("max_samples", C.c_int32), ("device", C.c_int32), ("precision", C.c_int32),
        ("rnn_z", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class VfTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 6), ("data", C.c_void_p)]


class VfCemParams(C.Structure):
    _fields_ = [
        ("num_samples", C.c_int32), ("global_samples", C.c_int32), ("sample_offset", C.c_int32),
        ("iterations", C.c_int32), ("num_elites", C.c_int32), ("nactions", C.c_int32), ("repeat", C.c_int32),
        ("action_bound", C.c_int32), ("use_mean0", C.c_int32), ("cost_kind", C.c_int32),
        ("n_ctx_actions", C.c_int32), ("pad0", C.c_int32),
        ("initial_std", C.c_double * 8), ("clip_lo", C.c_double * 8), ("clip_hi", C.c_double * 8),
        ("mean0", C.c_double * 128), ("reduce_std_scale", C.c_double), ("finalweight", C.c_double),
        ("task_weights", C.c_double * VF_MAX_TASKS),
        ("seed", C.c_uint64), ("plan_index", C.c_uint32),
        ("k_futures", C.c_int32), ("lambda_variance", C.c_float), ("reserved", C.c_int32 * 6),
        ("sampler", C.c_int32), ("n_append", C.c_int32), ("discrete_mask", C.c_uint32), ("pad1", C.c_int32),
        ("append_action", C.c_double * 8), ("beta0", C.c_double), ("beta1", C.c_double), ("kappa", C.c_double),
        ("mean_bias", C.c_double * 8),
    ]


_lib = None

_SIGS = {
    "vf_abi_version": (C.c_int, []),
    "vf_create": (C.c_int, [C.POINTER(VfConfig), C.POINTER(C.c_void_p)]),
    "vf_destroy": (C.c_int, [C.c_void_p]),
    "vf_last_error": (C.c_char_p, [C.c_void_p]),
    "vf_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vf_synchronize": (C.c_int, [C.c_void_p]),
    "vf_load_weights": (C.c_int, [C.c_void_p, C.POINTER(VfTensor), C.c_int32]),
    "vf_set_context": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "vf_set_desig": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vf_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vf_score": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "vf_score_external": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "vf_fetch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "vf_cem_plan": (C.c_int, [C.c_void_p, C.POINTER(VfCemParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vf_cem_begin": (C.c_int, [C.c_void_p, C.POINTER(VfCemParams), C.c_void_p, C.c_void_p]),
    "vf_cem_iter_rollout": (C.c_int, [C.c_void_p, C.c_int32]),
    "vf_cem_iter_select": (C.c_int, [C.c_void_p, C.c_int32]),
    "vf_cem_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vf_cem_scores_dev": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "vf_cem_actions": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vf_cem_bind_scores": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vf_cem_scores_read": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "vf_cem_scores_write": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "vf_comm_export": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "vf_comm_connect": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "vf_cem_exchange": (C.c_int, [C.c_void_p, C.c_int32]),
    "vf_comm_close": (C.c_int, [C.c_void_p]),
    "vf_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "vf_refit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vf_debug_conv2d": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p]),
    "vf_debug_fetch": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_int64]),
    "vf_debug_conv_plan": (C.c_int, [C.c_int32] * 8 + [C.c_void_p]),
    "vf_debug_conv_time": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_void_p]),
    "vf_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "vf_profile_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "vf_launch_count": (C.c_int64, [C.c_void_p]),
}


def load_library(path: Optional[str] = None):
    """dlopen libvfengine.so and bind every symbol include/vfengine.h declares (no compute happens)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise EngineUnavailable(
            "libvfengine.so is not built (%s). Run `python -m visual_foresight_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback." % p)
    lib = C.CDLL(p)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.vf_abi_version() != VF_ABI_VERSION:
        raise EngineUnavailable("ABI mismatch: library %d, binding %d" % (lib.vf_abi_version(), VF_ABI_VERSION))
    if path is None:
        _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGS.keys())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def spec_to_config(spec: PredictorSpec, max_samples: int, device: int = 0, precision: int = PREC_FP32_SIMT) -> VfConfig:
    spec.validate()
    cfg = VfConfig()
    cfg.abi_version = VF_ABI_VERSION
    cfg.height, cfg.width, cfg.ncam, cfg.ndesig = spec.height, spec.width, spec.ncam, spec.ndesig
    cfg.adim, cfg.sdim, cfg.nz = spec.adim, spec.sdim, spec.nz
    cfg.seq_len, cfg.context_frames, cfg.ngf = spec.seq_len, spec.context_frames, spec.ngf
    cfg.n_enc = len(spec.encoder)
    cfg.n_dec = len(spec.decoder)
    for i, (oc, rnn) in enumerate(spec.encoder):
        cfg.enc_channels[i], cfg.enc_rnn[i] = oc, int(rnn)
    for i, (oc, rnn) in enumerate(spec.decoder):
        cfg.dec_channels[i], cfg.dec_rnn[i] = oc, int(rnn)
    cfg.num_transformed, cfg.cdna_ksize, cfg.lstm_ksize = spec.num_transformed, spec.cdna_ksize, spec.lstm_ksize
    cfg.norm_eps, cfg.forget_bias = spec.norm_eps, spec.forget_bias
    cfg.max_samples, cfg.device, cfg.precision = int(max_samples), int(device), int(precision)
    cfg.rnn_z = int(spec.rnn_z)
    return cfg


class Engine:
    """Owns one vf_engine handle.  All array arguments/results are NumPy (host) arrays."""

    def __init__(self, spec: PredictorSpec, max_samples: int, device: int = 0, precision="f16x3"):
        self.lib = load_library()
        self.spec = spec
        self.max_samples = int(max_samples)
        prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        cfg = spec_to_config(spec, max_samples, device, prec)
        h = C.c_void_p()
        rc = self.lib.vf_create(C.byref(cfg), C.byref(h))
        self._h = h
        if rc != 0:
            msg = self.lib.vf_last_error(h).decode() if h else "vf_create failed"
            if h:
                self.lib.vf_destroy(h)
            self._h = None
            if rc == -2:
                raise EngineUnavailable(msg)
            raise EngineError("vf_create: %s (code %d)" % (msg, rc))
        self._keep = []

    # -- plumbing ---------------------------------------------------------------------------------
    def _check(self, rc):
        if rc < 0:
            raise EngineError("%s (code %d)" % (self.lib.vf_last_error(self._h).decode(), rc))
        return rc

    def close(self):
        if getattr(self, "_h", None):
            self.lib.vf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.vf_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self.lib.vf_synchronize(self._h))

    def launch_count(self) -> int:
        return int(self.lib.vf_launch_count(self._h))

    # -- weights ------------------------------------------------------------------------------------
    def load_weights(self, weights_per_view: Sequence[Dict[str, np.ndarray]]):
        """weights_per_view[v]: name -> ndarray (spec.weight_shapes).  setup_predictor.py:130-145."""
        tensors, keep = [], []
        for v, wd in enumerate(weights_per_view):
            for name, arr in wd.items():
                a = _f32(arr)
                nm = ("view%d.%s" % (v, name)).encode()
                keep += [a, nm]
                t = VfTensor()
                t.name, t.dtype, t.ndim = nm, 0, a.ndim
                for i, s in enumerate(a.shape):
                    t.shape[i] = s
                t.data = a.ctypes.data
                tensors.append(t)
        arr_t = (VfTensor * len(tensors))(*tensors)
        self._check(self.lib.vf_load_weights(self._h, arr_t, len(tensors)))

    # -- predictor ----------------------------------------------------------------------------------
    def set_context(self, frames_u8, states=None, ctx_actions=None, pix_distrib=None):
        sp = self.spec
        f = np.ascontiguousarray(frames_u8, dtype=np.uint8)
        assert f.shape == (sp.context_frames, sp.ncam, sp.height, sp.width, 3), f.shape
        st = None if states is None or sp.sdim == 0 else _f32(states)
        if st is not None:
            assert st.shape == (sp.context_frames, sp.sdim), st.shape
        ca = None if ctx_actions is None or len(ctx_actions) == 0 else _f32(ctx_actions)
        nca = 0 if ca is None else ca.shape[0]
        pd = None if pix_distrib is None else _f32(pix_distrib)
        if pd is not None:
            assert pd.shape == (sp.context_frames, sp.ncam, sp.height, sp.width, sp.ndesig), pd.shape
        self._check(self.lib.vf_set_context(self._h, _ptr(f), _ptr(st), _ptr(ca), nca, _ptr(pd)))

    def set_desig(self, desig_pix):
        d = _f32(np.asarray(desig_pix, dtype=np.float64).reshape(self.spec.ncam, self.spec.ndesig, 2))
        self._check(self.lib.vf_set_desig(self._h, _ptr(d)))

    def predict(self, actions, zs=None, fetch=True):
        sp = self.spec
        a = _f32(actions)
        M, T = a.shape[0], a.shape[1]
        assert a.shape[2] == sp.adim
        z = None if zs is None else _f32(zs)
        of = od = os_ = None
        if fetch:
            of = np.empty((M, sp.n_pred, sp.ncam, sp.height, sp.width, 3), np.float32)
            od = np.empty((M, sp.n_pred, sp.ncam, sp.height, sp.width, sp.ndesig), np.float32)
            os_ = np.empty((M, sp.n_pred, sp.sdim), np.float32) if sp.sdim else None
        self._check(self.lib.vf_predict(self._h, _ptr(a), M, T, _ptr(z), _ptr(of), _ptr(od), _ptr(os_)))
        return of, od, os_

    def score(self, goal, cost_kind=COST_PIXEL_DISTANCE, task_weights=None, finalweight=10.0, M=None):
        g = _f32(goal)
        tw = None if task_weights is None else _f32(task_weights)
        out = np.empty((self.max_samples if M is None else M,), np.float64)
        self._check(self.lib.vf_score(self._h, cost_kind, _ptr(g), _ptr(tw), float(finalweight), _ptr(out)))
        return out

    def score_external(self, distrib, goal_pix, task_weights=None, finalweight=10.0):
        d = _f32(distrib)
        M, P = d.shape[0], d.shape[1]
        g = _f32(goal_pix)
        tw = None if task_weights is None else _f32(task_weights)
        out = np.empty((M,), np.float64)
        self._check(self.lib.vf_score_external(self._h, _ptr(d), M, P, _ptr(g), _ptr(tw), float(finalweight), _ptr(out)))
        return out

    def fetch(self, indices, frames=True, distrib=True):
        sp = self.spec
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        n = idx.shape[0]
        of = np.empty((n, sp.n_pred, sp.ncam, sp.height, sp.width, 3), np.float32) if frames else None
        od = np.empty((n, sp.n_pred, sp.ncam, sp.height, sp.width, sp.ndesig), np.float32) if distrib else None
        self._check(self.lib.vf_fetch(self._h, _ptr(idx), n, _ptr(of), _ptr(od)))
        return of, od

    # -- CEM ------------------------------------------------------------------------------------------
    def cem_plan(self, params: VfCemParams, goal, noise=None):
        g = _f32(goal)
        nz = None if noise is None else _f32(noise)
        K, T = params.num_elites, params.nactions * params.repeat
        best = np.empty((K, T, self.spec.adim), np.float64)
        eidx = np.empty((K,), np.int32)
        scores = np.empty((params.iterations, params.global_samples), np.float64)
        self._cem_params = params
        self._check(self.lib.vf_cem_plan(self._h, C.byref(params), _ptr(g), _ptr(nz), _ptr(best), _ptr(eidx), _ptr(scores)))
        return best, eidx, scores

    def cem_begin(self, params: VfCemParams, goal, noise=None):
        g = _f32(goal)
        nz = None if noise is None else _f32(noise)
        self._cem_params = params
        self._check(self.lib.vf_cem_begin(self._h, C.byref(params), _ptr(g), _ptr(nz)))

    def cem_iter_rollout(self, it: int):
        self._check(self.lib.vf_cem_iter_rollout(self._h, it))

    def cem_iter_select(self, it: int):
        self._check(self.lib.vf_cem_iter_select(self._h, it))

    def cem_finish(self):
        p = self._cem_params
        K, T = p.num_elites, p.nactions * p.repeat
        best = np.empty((K, T, self.spec.adim), np.float64)
        eidx = np.empty((K,), np.int32)
        scores = np.empty((p.iterations, p.global_samples), np.float64)
        self._check(self.lib.vf_cem_finish(self._h, _ptr(best), _ptr(eidx), _ptr(scores)))
        return best, eidx, scores

    def cem_scores_dev(self) -> int:
        p = C.c_void_p()
        self._check(self.lib.vf_cem_scores_dev(self._h, C.byref(p)))
        return int(p.value)

    def cem_bind_scores(self, dev_ptr: int):
        self._check(self.lib.vf_cem_bind_scores(self._h, C.c_void_p(dev_ptr) if dev_ptr else None))

    def cem_scores_read(self, it: int, offset: int, n: int):
        out = np.empty((n,), np.float64)
        self._check(self.lib.vf_cem_scores_read(self._h, it, offset, n, _ptr(out)))
        return out

    def cem_scores_write(self, it: int, offset: int, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._check(self.lib.vf_cem_scores_write(self._h, it, offset, v.shape[0], _ptr(v)))

    # -- multi-GPU score exchange over peer memory (vf_comm_*) -----------------------------------------
    def comm_export(self, max_iterations: int, max_global_samples: int) -> bytes:
        """allocate this handle's exchange window; returns the 128-byte descriptor the other ranks need"""
        buf = C.create_string_buffer(VF_PEER_DESC_BYTES)
        self._check(self.lib.vf_comm_export(self._h, int(max_iterations), int(max_global_samples), buf))
        return bytes(buf.raw)

    def comm_connect(self, rank: int, world: int, descs):
        """descs: the descriptors of all ranks, in rank order"""
        blob = b"".join(bytes(d) for d in descs)
        assert len(blob) == world * VF_PEER_DESC_BYTES, "need one %d-byte descriptor per rank" % VF_PEER_DESC_BYTES
        self._check(self.lib.vf_comm_connect(self._h, int(rank), int(world), C.c_char_p(blob)))

    def cem_exchange(self, it: int):
        self._check(self.lib.vf_cem_exchange(self._h, int(it)))

    def comm_close(self):
        self._check(self.lib.vf_comm_close(self._h))

    def cem_actions(self):
        p = self._cem_params
        kf = max(int(p.k_futures), 1)        # the device holds k_futures consecutive copies of every action row
        out = np.empty((p.num_samples * kf, p.nactions * p.repeat, self.spec.adim), np.float64)
        self._check(self.lib.vf_cem_actions(self._h, _ptr(out)))
        return out[::kf]

    def topk(self, scores, k: int):
        s = np.ascontiguousarray(scores, dtype=np.float64)
        out = np.empty((k,), np.int32)
        self._check(self.lib.vf_topk(self._h, _ptr(s), s.shape[0], k, _ptr(out)))
        return out

    def refit(self, elites, nactions: int, repeat: int):
        e = np.ascontiguousarray(elites, dtype=np.float64)
        K, T, adim = e.shape
        assert T == nactions * repeat
        D = nactions * adim
        mean, cov, fac = np.empty(D), np.empty((D, D)), np.empty((D, K))
        self._check(self.lib.vf_refit(self._h, _ptr(e), K, nactions, repeat, adim, _ptr(mean), _ptr(cov), _ptr(fac)))
        return mean, cov, fac

    def profile_enable(self, on: bool):
        self._check(self.lib.vf_profile_enable(self._h, int(on)))

    def profile_read(self):
        ms, fl, n = np.zeros(3), np.zeros(3), np.zeros(3, np.int64)
        self._check(self.lib.vf_profile_read(self._h, _ptr(ms), _ptr(fl), _ptr(n), 3))
        return {"lstm_conv": {"ms": ms[0], "flops": fl[0], "launches": int(n[0])},
                "other_conv": {"ms": ms[1], "flops": fl[1], "launches": int(n[1])},
                "shared_prefix_conv": {"ms": ms[2], "flops": fl[2], "launches": int(n[2])}}

    # -- debug ---------------------------------------------------------------------------------------
    def debug_conv2d(self, x, w, bias=None, impl=PREC_FP32_SIMT):
        x = _f32(x)
        w = _f32(w)
        B, H, W, Cin = x.shape
        k, _, _, Cout = w.shape
        b = None if bias is None else _f32(bias)
        y = np.empty((B, H, W, Cout), np.float32)
        self._check(self.lib.vf_debug_conv2d(self._h, impl, _ptr(x), _ptr(w), _ptr(b), B, H, W, Cin, Cout, k, _ptr(y)))
        return y

    def debug_conv_time(self, B, H, W, Cin, Cout, k, impl=PREC_F16X3, reps=20) -> float:
        """average ms per launch of one convolution of this shape (tuning aid)"""
        ms = C.c_double(0.0)
        self._check(self.lib.vf_debug_conv_time(self._h, impl, B, H, W, Cin, Cout, k, reps, C.byref(ms)))
        return ms.value

    def debug_fetch(self, name: str, view: int = 0):
        n = self._check(self.lib.vf_debug_fetch(self._h, name.encode(), view, None, 0))
        out = np.empty((n,), np.float32)
        self._check(self.lib.vf_debug_fetch(self._h, name.encode(), view, _ptr(out), n))
        return out
