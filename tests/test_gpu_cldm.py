"""ControlNet branch on the GPU: ``anyedit_b200.cldm`` against golden outputs of the reference's own
``ControlNet`` / ``ControlledUnetModel`` (tests/golden/make_golden_cldm.py; cldm.py:21-304), and the sampler-level
equivalence of ``ddim_hacked``'s two-call guidance with the batched call."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
FWD_TOL = 4e-3


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def nets():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from anyedit_b200.cldm import ControlledUnetModel, ControlNet
    from oracle import weights
    g = np.load(os.path.join(G, "cldm_tiny.npz"))
    meta = json.load(open(os.path.join(G, "cldm_tiny_keys.json")))
    cn, un = ControlNet(**meta["control_config"]), ControlledUnetModel(**meta["unet_config"])
    assert {k: list(v.shape) for k, v in cn.state_dict().items()} == meta["control_keys"]          # same keys and shapes
    csd = weights.make_state_dict({k: tuple(v) for k, v in meta["control_keys"].items()}, int(g["cseed"]))
    usd = weights.make_state_dict({k: tuple(v) for k, v in meta["unet_keys"].items()}, int(g["useed"]))
    assert weights.checksum(csd) == pytest.approx(float(g["cwsum"]), rel=1e-12)
    cn.load_state_dict(csd, strict=True)
    un.load_state_dict(usd, strict=True)
    return cn.cuda(), un.cuda(), g


def test_controlnet_residuals_vs_reference(nets):
    cn, un, g = nets
    f = lambda k: torch.from_numpy(g[k]).cuda()
    outs = cn(x=f("x"), hint=f("hint"), timesteps=f("t"), context=f("ctx"))
    assert len(outs) == 7 and getattr(outs, "nhwc", False)
    errs = []
    for i, o in enumerate(outs):
        ref = torch.from_numpy(g[f"control_{i}"])
        assert tuple(o.shape) == (ref.shape[0], ref.shape[2], ref.shape[3], ref.shape[1])
        errs.append(rel(o.permute(0, 3, 1, 2), ref))
    print("[controlnet tiny] residual rel-L2 vs reference: " + " ".join(f"{e:.2e}" for e in errs))
    assert max(errs) < FWD_TOL, errs
    # the hint stem is kept while the hint tensor is unchanged, and refilled in place when its values change
    hint = f("hint")
    a = [o.clone() for o in cn(x=f("x"), hint=hint, timesteps=f("t"), context=f("ctx"))]
    n0 = __import__("anyedit_b200").ops.launch_count
    b = cn(x=f("x"), hint=hint, timesteps=f("t"), context=f("ctx"))
    n1 = __import__("anyedit_b200").ops.launch_count
    assert all(torch.equal(p, q) for p, q in zip(a, b))
    hint.mul_(0.5)
    c = cn(x=f("x"), hint=hint, timesteps=f("t"), context=f("ctx"))
    n2 = __import__("anyedit_b200").ops.launch_count
    assert n2 - n1 > n1 - n0 and not torch.equal(c[0], a[0])


def test_controlled_unet_and_denoiser_vs_reference(nets):
    """eps of the controlled UNet with scaled residuals (cldm.py:336-338) through ControlDenoiser.apply_model; then a CFG
    sampling run: DDIMSampler's batched [uncond ; cond] call == ddim_hacked's two separate calls (ddim_hacked.py:181-232)."""
    from anyedit_b200.cldm import ControlDenoiser
    from anyedit_b200.ddim import DDIMSampler
    cn, un, g = nets
    f = lambda k: torch.from_numpy(g[k]).cuda()
    scales = [float(s) for s in g["scales"]]
    for only_mid in (False, True):
        den = ControlDenoiser(un, cn, only_mid_control=only_mid, control_scales=scales).cuda()
        eps = den.apply_model(f("x"), f("t"), {"c_concat": [f("hint")], "c_crossattn": [f("ctx")]})
        e = rel(eps, torch.from_numpy(g[f"eps_only_mid{int(only_mid)}"]))
        print(f"[controlled unet only_mid={only_mid}] eps rel-L2 vs reference = {e:.3e}")
        assert e < FWD_TOL, e
    den = ControlDenoiser(un, cn, control_scales=scales).cuda()
    gen = torch.Generator().manual_seed(5)
    x_T, uctx = torch.randn(2, 4, 16, 16, generator=gen).cuda(), torch.randn(2, 7, 64, generator=gen).cuda()
    cond = {"c_concat": [f("hint")], "c_crossattn": [f("ctx")]}
    unc = {"c_concat": [f("hint")], "c_crossattn": [uctx]}
    outs = []
    for graph in (True, False):
        o, _ = DDIMSampler(den, use_cuda_graph=graph).sample(5, 2, (4, 16, 16), cond, verbose=False, x_T=x_T, eta=0.0,
                                                             unconditional_guidance_scale=4.0, unconditional_conditioning=unc)
        outs.append(o)
    assert torch.equal(outs[0], outs[1])
    # two-call guidance (ddim_hacked.py:181-232: the model is called once per branch), then the product's own fused update
    from anyedit_b200 import ops
    smp = DDIMSampler(den, use_cuda_graph=False)
    smp.make_schedule(5, verbose=False)
    x = x_T.clone()
    for i, step in enumerate(np.flip(smp.ddim_timesteps)):
        index = 5 - i - 1
        t = torch.full((2,), int(step), device="cuda", dtype=torch.long)
        eps = torch.cat([den.apply_model(x, t, unc), den.apply_model(x, t, cond)]).float().contiguous()
        x_prev = torch.empty_like(x)
        ops.cfg_ddim_step(x, eps, smp.ddim_coef[index], 4.0, True, x_prev)
        x = x_prev
    assert torch.equal(outs[0], x)
