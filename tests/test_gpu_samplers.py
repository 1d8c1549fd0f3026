"""Sibling samplers and sampler options on the DDIM stepper:
PLMS and DPM-Solver++(2M) against golden runs of the reference's own PLMSSampler / DPMSolverSampler
(tests/golden/make_golden_plms.py, make_golden_dpm.py), the fused update kernels bit-exact against the reference's fp32
tensor expressions, and the DDIM options ``use_original_steps`` / v-parameterisation / ``noise_dropout`` / ``mirror_rng``."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
LATENT_TOL_CFG = 3e-3          # final latent under guidance (see tests/test_gpu_parity_shapes.py)
LATENT_TOL = 1e-3


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _tiny_c(key="crossattn"):
    from anyedit_b200.diffusion import LatentDenoiser
    from anyedit_b200.unet import UNetModel
    from oracle import weights
    meta = json.load(open(os.path.join(G, "tiny_c_keys.json")))
    net = UNetModel(**meta["config"])
    sd = weights.make_state_dict({k: tuple(v) for k, v in meta["keys"].items()}, 15)
    net.load_state_dict(sd)
    return LatentDenoiser(net.cuda(), key).cuda(), sd, meta["config"]


def test_plms_kernel_bit_exact():
    """anysd_cfg_plms_step_f32 == the reference's fp32 tensor expressions (plms.py:205-238), incl. the history push."""
    from anyedit_b200 import ops
    gen = torch.Generator().manual_seed(7)
    B = 2
    x, eps = torch.randn(B, 4, 8, 8, generator=gen), torch.randn(2 * B, 4, 8, 8, generator=gen)
    o = [torch.randn(B, 4, 8, 8, generator=gen) for _ in range(3)]
    e_u, e_c = eps.chunk(2)
    e = e_u + 5.0 * (e_c - e_u)
    ddim = [0.6, 0.8, 0.9, 0.3]
    for comb, expect in (((3.0, 1.0, 0.0, 0.0, 2.0), (3 * e - o[0]) / 2),
                         ((23.0, 16.0, 5.0, 0.0, 12.0), (23 * e - 16 * o[0] + 5 * o[1]) / 12),
                         ((55.0, 59.0, 37.0, 9.0, 24.0), (55 * e - 59 * o[0] + 37 * o[1] - 9 * o[2]) / 24),
                         ((1.0, -1.0, 0.0, 0.0, 2.0), (e + o[0]) / 2)):
        for push in (1.0, 0.0):
            hist = torch.stack(o).cuda().contiguous()
            coef = torch.tensor(ddim + list(comb) + [push]).cuda()
            xp, p0 = torch.empty(B, 4, 8, 8, device="cuda"), torch.empty(B, 4, 8, 8, device="cuda")
            ops.cfg_plms_step(x.cuda(), eps.cuda(), coef, 5.0, True, hist, xp, p0)
            pred = (x - torch.tensor(0.6) * expect) / torch.tensor(0.8)
            assert torch.equal(p0.cpu(), pred), comb
            assert torch.equal(xp.cpu(), torch.tensor(0.9) * pred + torch.tensor(0.3) * expect), comb
            want = [e, o[0], o[1]] if push else o
            assert all(torch.equal(hist[k].cpu(), want[k]) for k in range(3)), (comb, push)


def test_plms_vs_reference_golden():
    """PLMSSampler.sample through the product API vs the reference PLMSSampler (golden), CUDA graph and eager."""
    from anyedit_b200.plms import PLMSSampler
    g = np.load(os.path.join(G, "plms_tiny.npz"))
    model, _, _ = _tiny_c()
    f = lambda k: torch.from_numpy(g[k]).cuda()
    outs = {}
    for graph in (True, False):
        smp = PLMSSampler(model, use_cuda_graph=graph)
        for S, scale in ((10, 5.0), (20, 1.0)):
            img, inter = smp.sample(S, 2, (4, 8, 12), conditioning=f("c"), verbose=False, x_T=f("x_T"), eta=0.0,
                                    unconditional_guidance_scale=scale, unconditional_conditioning=f("uc") if scale != 1.0 else None)
            e, e0 = rel(img, f(f"x0_S{S}")), rel(inter["pred_x0"][-1], f(f"pred_x0_last_S{S}"))
            print(f"[plms tiny S={S} scale={scale} graph={graph}] final rel-L2 {e:.3e}, last pred_x0 {e0:.3e}")
            assert e < (LATENT_TOL_CFG if scale != 1.0 else 2 * LATENT_TOL), e
            outs[(graph, S)] = img
    for S in (10, 20):
        assert torch.equal(outs[(True, S)], outs[(False, S)])
    with pytest.raises(ValueError):
        PLMSSampler(model).sample(10, 2, (4, 8, 12), conditioning=f("c"), verbose=False, x_T=f("x_T"), eta=0.5)


def test_dpmpp_kernel_matches_formula():
    from anyedit_b200 import ops
    gen = torch.Generator().manual_seed(9)
    B = 2
    x, eps, mp = torch.randn(B, 4, 8, 8, generator=gen), torch.randn(2 * B, 4, 8, 8, generator=gen), torch.randn(B, 4, 8, 8, generator=gen)
    coef = torch.tensor([0.9958, 0.0913, 0.9970, -0.0285, -0.0142, 0.9307])
    e_u, e_c = eps.chunk(2)
    e = e_u + 7.5 * (e_c - e_u)
    m = (x - coef[0] * e) / coef[1]
    want = (coef[2] * x - coef[3] * m) - coef[4] * (coef[5] * (m - mp))
    mpd, xn, x0o = mp.cuda().clone(), torch.empty(B, 4, 8, 8, device="cuda"), torch.empty(B, 4, 8, 8, device="cuda")
    ops.cfg_dpmpp_step(x.cuda(), eps.cuda(), coef.cuda(), 7.5, True, mpd, xn, x0o)
    assert torch.equal(xn.cpu(), want) and torch.equal(mpd.cpu(), m) and torch.equal(x0o.cpu(), m)


def test_dpm_solver_vs_reference_golden():
    """DPMSolverSampler.sample (DPM-Solver++(2M), float model timesteps) vs the reference DPMSolverSampler (golden)."""
    from anyedit_b200.dpm_solver import DPMSolverSampler
    g = np.load(os.path.join(G, "dpm_tiny.npz"))
    model, _, _ = _tiny_c()
    f = lambda k: torch.from_numpy(g[k]).cuda()
    outs = {}
    for graph in (True, False):
        smp = DPMSolverSampler(model, use_cuda_graph=graph)
        for S, scale in ((10, 5.0), (20, 1.0)):
            img, none = smp.sample(S, 2, (4, 8, 12), conditioning=f("c"), verbose=False, x_T=f("x_T"),
                                   unconditional_guidance_scale=scale, unconditional_conditioning=f("uc"))
            assert none is None
            e = rel(img, f(f"x0_S{S}"))
            print(f"[dpm-solver++ tiny S={S} scale={scale} graph={graph}] final rel-L2 {e:.3e}")
            assert e < (LATENT_TOL_CFG if scale != 1.0 else 2 * LATENT_TOL), e
            outs[(graph, S)] = img
    for S in (10, 20):
        assert torch.equal(outs[(True, S)], outs[(False, S)])


def test_ddim_v_param_kernel_and_original_steps():
    """v-parameterisation (ddim.py:214-218, 224-226) in the fused update, bit-exact vs the tensor expressions; and
    ``ddim_use_original_steps`` (the DDPM grid, :137-146, 221-225) for 3 steps vs the formula with the oracle UNet."""
    from anyedit_b200 import ops
    from anyedit_b200.ddim import DDIMSampler
    from oracle import unet_oracle
    gen = torch.Generator().manual_seed(17)
    B = 2
    x, out = torch.randn(B, 4, 8, 8, generator=gen), torch.randn(2 * B, 4, 8, 8, generator=gen)
    coef = torch.tensor([0.6, 0.8, 0.9, 0.3, 0.05, 0.71, 0.704])
    noise = torch.randn(B, 4, 8, 8, generator=gen)
    v_u, v_c = out.chunk(2)
    v = v_u + 3.0 * (v_c - v_u)
    e = coef[5] * v + coef[6] * x
    pred = coef[5] * x - coef[6] * v
    want = coef[2] * pred + coef[3] * e + coef[4] * noise
    xp, p0 = torch.empty(B, 4, 8, 8, device="cuda"), torch.empty(B, 4, 8, 8, device="cuda")
    ops.cfg_ddim_step(x.cuda(), out.cuda(), coef.cuda(), 3.0, True, xp, p0, noise.cuda(), v_param=True)
    assert torch.equal(p0.cpu(), pred) and torch.equal(xp.cpu(), want)

    model, sd, cfg = _tiny_c()
    x_T, c = torch.randn(2, 4, 8, 12, generator=gen), torch.randn(2, 5, 96, generator=gen)
    smp = DDIMSampler(model)
    smp.make_schedule(10, verbose=False)
    got, inter = smp.ddim_sampling(c.cuda(), (2, 4, 8, 12), x_T=x_T.cuda(), ddim_use_original_steps=True, timesteps=3, log_every_t=1)
    acp, acp_prev = model.alphas_cumprod.cpu(), model.alphas_cumprod_prev.cpu()
    xx = x_T
    for step in (2, 1, 0):
        t = torch.full((2,), step, dtype=torch.long)
        ee = unet_oracle.unet_forward(sd, xx, t, c, None, num_head_channels=cfg["num_head_channels"])
        a_t, a_prev = acp[step], acp_prev[step]
        p = (xx - (1 - a_t).sqrt() * ee) / a_t.sqrt()
        xx = a_prev.sqrt() * p + (1 - a_prev).sqrt() * ee
    err = rel(got, xx)
    print(f"[ddim use_original_steps, 3 steps] rel-L2 vs formula with the oracle UNet = {err:.3e}")
    assert err < 2e-3 and len(inter["x_inter"]) == 4


def test_ddim_noise_options():
    """eta > 0 with ``noise_dropout`` runs; ``mirror_rng`` makes an eta = 0 run consume torch's RNG like the reference
    (one randn per step, ddim.py:247) without changing the result."""
    from anyedit_b200.ddim import DDIMSampler
    model, _, _ = _tiny_c()
    gen = torch.Generator().manual_seed(23)
    x_T, c = torch.randn(2, 4, 8, 12, generator=gen).cuda(), torch.randn(2, 5, 96, generator=gen).cuda()
    base, _ = DDIMSampler(model).sample(5, 2, (4, 8, 12), c, verbose=False, x_T=x_T, eta=0.0)
    torch.manual_seed(1)
    s0 = torch.cuda.get_rng_state()
    mir, _ = DDIMSampler(model, mirror_rng=True).sample(5, 2, (4, 8, 12), c, verbose=False, x_T=x_T, eta=0.0)
    assert torch.equal(base, mir) and not torch.equal(s0, torch.cuda.get_rng_state())
    noisy, _ = DDIMSampler(model).sample(5, 2, (4, 8, 12), c, verbose=False, x_T=x_T, eta=0.7, noise_dropout=0.3)
    assert torch.isfinite(noisy).all() and not torch.equal(noisy, base)
