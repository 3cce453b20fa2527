"""Generates tests/golden/ref_registration_golden.npz from the UNMODIFIED reference
``policy/cem_controllers/register_gtruth_controller.py`` (``get_warp_err`` :114-173 and the trade-off normalisation of
``register_gtruth`` :88-91).  The module imports three files that are not in the reference tree (visualizer.render_utils,
visualizer.make_cem_visuals, registration_network.setup_registration); they are stubbed here ONLY so that the module
imports — none of the stubbed names is executed.  ``get_warp_err`` is called on a bare instance whose attributes are set
by hand (the class constructor needs the missing registration network).

    python tests/golden/make_registration_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402


def main():
    ref_shim.install()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    import visual_mpc.policy.cem_controllers.visualizer as viz  # noqa: F401
    mod("visual_mpc.policy.cem_controllers.visualizer.render_utils", resize_image=None)
    mod("visual_mpc.policy.cem_controllers.visualizer.make_cem_visuals", CEM_Visual_Preparation_Registration=object)
    mod("visual_mpc.registration_network")
    mod("visual_mpc.registration_network.setup_registration", setup_gdn=None)
    from visual_mpc.policy.cem_controllers.register_gtruth_controller import Register_Gtruth_Controller as RC

    rng = np.random.RandomState(5)
    G = {}
    for tag, (H, W, region) in {"pt48": (48, 64, False), "reg48": (48, 64, True), "reg96": (96, 128, True)}.items():
        ncam, ntask = 2, 2
        start = rng.rand(ncam, H, W, 3)
        goal = rng.rand(ncam, H, W, 3)
        wstart = np.clip(start + 0.05 * rng.randn(ncam, H, W, 3), 0, 1)
        wgoal = np.clip(goal + 0.1 * rng.randn(ncam, H, W, 3), 0, 1)
        yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        base = np.stack([xx, yy], -1).astype(np.float64)
        spts = base[None] + rng.randn(ncam, H, W, 2) * 1.5
        gpts = base[None] + rng.randn(ncam, H, W, 2) * 2.5
        # stored as float32 (fixture size); the reference sees the same float32-representable values in float64
        start, goal, wstart, wgoal, spts, gpts = [a.astype(np.float32).astype(np.float64) for a in (start, goal, wstart, wgoal, spts, gpts)]
        d0 = np.stack([rng.randint(0, H, (ncam, ntask)), rng.randint(0, W, (ncam, ntask))], -1)
        d0[0, 0] = (1, 2)                                   # window clipped at the border
        gp = np.stack([rng.randint(0, H, (ncam, ntask)), rng.randint(0, W, (ncam, ntask))], -1)
        gp[1, 1] = (H - 1, W - 2)
        o = object.__new__(RC)
        o._hp = ref_shim._HParams(register_gtruth=["start", "goal"], register_region=region)
        o.ntask = ntask
        o.agentparams = {"image_height": H, "image_width": W}
        o._img_height = 48
        o.desig_pix_t0, o.goal_pix_sel = d0, gp
        o.desig_pix_t0_med, o.goal_pix_med = d0, gp
        errs, pix = [], []
        for cam in range(ncam):
            e, d = o.get_warp_err(cam, start, goal, spts, gpts, wstart, wgoal)
            errs.append(e)
            pix.append(d)
        errs = np.stack(errs, 0)
        for k, v in dict(start=start, goal=goal, wstart=wstart, wgoal=wgoal, spts=spts, gpts=gpts, d0=d0, gp=gp,
                         errs=errs, pix=np.stack(pix, 0)).items():
            G["%s_%s" % (tag, k)] = v.astype(np.float32) if k in ("start", "goal", "wstart", "wgoal", "spts", "gpts") else v
        if region:                                           # the reference normalisation (register_gtruth :88-91)
            tr = 1 / errs
            tr = tr / np.sum(np.sum(tr, 0, keepdims=True), 2, keepdims=True)
            G["%s_tradeoff" % tag] = tr.reshape(ncam, ntask * 2)
    out = os.path.join(HERE, "ref_registration_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, len(G), "arrays", os.path.getsize(out) // 1024, "KB")


if __name__ == "__main__":
    main()
