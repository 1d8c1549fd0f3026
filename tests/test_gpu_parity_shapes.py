"""Parity at the shapes bench.py times ("parity first"): the CUDA path at the full SD-1.5 / IP2P geometry
against golden vectors the REAL reference produced in the build container (tests/golden/make_golden_sd15.py:
``ldm`` UNetModel + DDIMSampler, CPU fp32), and against the rounding-matched restatement (oracle/unet_emul16.py).

  test_bench_shape_eps_vs_reference        one UNet evaluation at a 64x64 latent, B_eff = 16, run exactly as bench.py runs
                                           it: CUDA-graph replay + shared CFG halves + kept context K/V
  test_c1_trajectory_vs_reference          BASELINE configs[1]: 50 DDIM steps, CFG 7.5, batch 8 -- final latent
  test_c3_trajectory_vs_reference          BASELINE configs[3] geometry: 96x96 latent, 100 DDIM steps, CFG 7.5
  test_rounding_matched_oracle_config0     BASELINE configs[0]: CUDA vs an ensemble of fp32 oracles with fp16 rounding injected
                                           where the kernels round -- the CFG-amplified distance is operand rounding
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, G)
from sd15_inputs import SD15, WEIGHT_SCHEME, WEIGHT_SEED, checksum, inputs_requests  # noqa: E402

# Tolerances (relative L2).  One evaluation: fp16 operands / fp32 accumulate vs fp32 (same bound as test_gpu_unet).
FWD_TOL = 4e-3
# Final latent after a full sampling run.  north_star asks <= 1e-3: met at guidance scale 1.0.  Under CFG 7.5 the
# guidance combine amplifies the decorrelated fp16 operand rounding of the two halves; the rounding-matched oracle
# below shows the CUDA path is statistically indistinguishable from "fp32 math + fp16 storage", i.e. the distance to the
# fp32 reference is the cost of fp16 tensor-core operands, not an implementation defect.
LATENT_TOL = 1e-3
LATENT_TOL_CFG = 3e-3
EXCESS = 1.15          # the CUDA path may sit at most 15 % further out than ideal fp16-storage realisations do


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def sd15():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from anyedit_b200.diffusion import LatentDenoiser
    from anyedit_b200.unet import UNetModel
    from oracle import weights
    meta = json.load(open(os.path.join(G, "sd15_keys.json")))
    assert meta["config"] == SD15
    sd = weights.make_state_dict({k: tuple(v) for k, v in meta["keys"].items()}, WEIGHT_SEED, scheme=WEIGHT_SCHEME)
    with torch.device("cuda"):
        net = UNetModel(**SD15)
    net.load_state_dict(sd)
    return net, LatentDenoiser(net, "hybrid").cuda(), sd


def _golden(name, inputs):
    path = os.path.join(G, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated (tests/golden/make_golden_sd15.py)")
    g = np.load(path)
    assert np.allclose(g["in_sum"], checksum(*inputs), rtol=1e-12), "input generator drifted from the golden run"
    return g


def _cond(c_cat, c_txt, u_txt):
    n = c_cat.shape[0]
    cu = lambda t: t.cuda()
    return ({"c_concat": [cu(c_cat)], "c_crossattn": [cu(c_txt)]},
            {"c_concat": [cu(c_cat)], "c_crossattn": [cu(u_txt.expand(n, -1, -1).contiguous())]})


def test_bench_shape_eps_vs_reference(sd15):
    """64x64 latent, 8 requests (4 distinct, each twice), CFG => B_eff = 16: the eps of the captured-graph step (shared
    CFG halves, kept context K/V -- the path bench.py times) vs the reference UNet row by row; duplicated requests and the
    plain eager forward must agree bit for bit (batch independence)."""
    from anyedit_b200.ddim import DDIMSampler
    net, model, _ = sd15
    x_T, c_cat, c_txt, u_txt = inputs_requests(4, 64, 2024)
    g = _golden("sd15_fwd64.npz", (x_T, c_cat, c_txt, u_txt))
    dup = lambda t: torch.cat([t, t])
    x8, cc8, ct8 = dup(x_T), dup(c_cat), dup(c_txt)
    cond, uncond = _cond(cc8, ct8, u_txt)
    smp = DDIMSampler(model)
    smp.sample(2, 8, (4, 64, 64), cond, verbose=False, x_T=x8.cuda(), eta=0.0, unconditional_guidance_scale=7.5,
               unconditional_conditioning=uncond)                       # step 1 eager (fills K/V), step 2 captured
    st = next(iter(smp._graphs.values()))
    assert st.graph is not None and st.shared and st.kv is not None and not st.kv_dirty
    st.step(x8.cuda(), 1, int(g["t"]), 7.5, None)                        # graph replay at t = 981
    eps = st.last_eps.clone()
    assert tuple(eps.shape) == (16, 4, 64, 64)
    errs = [rel(eps[i], torch.from_numpy(g["eps_u"][i])) for i in range(4)] + \
           [rel(eps[8 + i], torch.from_numpy(g["eps_c"][i])) for i in range(4)]
    print("[bench shape 64x64 B_eff=16, graph + shared halves + kept K/V] eps rel-L2 vs reference per row: " +
          " ".join(f"{e:.2e}" for e in errs))
    assert max(errs) < FWD_TOL, errs
    assert torch.equal(eps[4:8], eps[0:4]) and torch.equal(eps[12:16], eps[8:12])       # duplicated requests
    # plain eager forward of the same CFG batch (no sharing, no kept K/V, no graph)
    t16 = torch.full((16,), int(g["t"]), dtype=torch.long, device="cuda")
    x16 = torch.cat([torch.cat([x8, cc8], 1)] * 2).cuda()
    plain = net(x16, t16, context=torch.cat([uncond["c_crossattn"][0], cond["c_crossattn"][0]]))
    assert torch.equal(plain, eps)


def _trajectory(sd15, name, h, B, seed, scale, tol):
    from anyedit_b200.ddim import DDIMSampler
    net, model, _ = sd15
    x_T, c_cat, c_txt, u_txt = inputs_requests(1, h, seed)
    g = _golden(name, (x_T, c_cat, c_txt, u_txt))
    S, every = int(g["S"]), int(g["log_every_t"])
    assert float(g["scale"]) == scale
    if B > 1:                                            # the fixture's request rides in slot 0 of a full batch
        ox, oc, ot, _ = inputs_requests(B - 1, h, seed + 7)
        x_T, c_cat, c_txt = torch.cat([x_T, ox]), torch.cat([c_cat, oc]), torch.cat([c_txt, ot])
    cond, uncond = _cond(c_cat, c_txt, u_txt)
    out, inter = DDIMSampler(model).sample(S, B, (4, h, h), cond, verbose=False, x_T=x_T.cuda(), eta=0.0, log_every_t=every,
                                           unconditional_guidance_scale=scale, unconditional_conditioning=uncond)
    curve = [rel(inter["x_inter"][k][0], torch.from_numpy(g["x_inter"][k][0])) for k in range(1, len(inter["x_inter"]))]
    e = rel(out[0], torch.from_numpy(g["final"][0]))
    moved = rel(torch.from_numpy(g["final"][0]), x_T[0])
    print(f"[{name}: {8 * h}^2, {S} steps, scale {scale}, batch {B}] final-latent rel-L2 vs reference = {e:.3e} "
          f"(latent moved {moved:.2f}); along the trajectory: " + " ".join(f"{c:.2e}" for c in curve))
    assert len(inter["x_inter"]) == g["x_inter"].shape[0]
    assert moved > 0.05
    assert e < tol, e
    return e


def test_c1_trajectory_vs_reference(sd15):
    """BASELINE configs[1]: 512x512 (64x64 latent), 50 DDIM steps, CFG 7.5, batch 8 -- vs the reference DDIMSampler."""
    _trajectory(sd15, "sd15_c1.npz", 64, 8, 2025, 7.5, LATENT_TOL_CFG)


def test_c1_trajectory_no_guidance_meets_1e3(sd15):
    _trajectory(sd15, "sd15_c1_s1.npz", 64, 8, 2025, 1.0, LATENT_TOL)


def test_c3_trajectory_vs_reference(sd15):
    """BASELINE configs[3] geometry: 768x768 (96x96 latent, 9216-token self-attention), 100 DDIM steps, CFG 7.5, batch 2."""
    _trajectory(sd15, "sd15_c3.npz", 96, 2, 2026, 7.5, LATENT_TOL_CFG)


def test_c3_trajectory_no_guidance_meets_1e3(sd15):
    _trajectory(sd15, "sd15_c3_s1.npz", 96, 2, 2026, 1.0, LATENT_TOL)


def test_rounding_matched_oracle_config0(sd15):
    """BASELINE configs[0] (256x256, 20 steps, CFG 7.5, batch 1): is the ~2e-3 distance to the fp32 reference under
    guidance rounding, or a defect?  ``oracle/unet_emul16.py`` evaluates the reference graph in fp32 but rounds to fp16
    exactly where the kernels store fp16 -- an ideal "fp32 math + fp16 storage" implementation.

    Two such implementations cannot agree element by element: a 1-ulp difference in an fp32 accumulation flips the fp16
    rounding of a fraction delta/ulp of the elements of the next tensor, which re-injects sqrt(delta/ulp) * ulp/sqrt(12) of
    noise; that map has its fixed point at ~0.1 ulp per layer, i.e. any two evaluation orders decorrelate to about the
    rounding noise itself within a few layers (measured here: emul16 vs emul16 with inputs perturbed by 4e-6).  What CAN
    be tested is whether the CUDA path is statistically distinguishable from the ideal implementation.  With K emul16
    realisations (inputs jittered by 2^-18, which moves the fp32 result by ~1e-5):
        e_c = |cuda - fp32|   vs  e_i = |emul_i - fp32|                 (distance to the reference)
        d_c = |cuda - mean_i emul_i|  vs  d_i = |emul_i - mean_{j != i} emul_j|   (distance to the ensemble mean, in which the
                                                                          common weight-rounding term cancels)
    Norms of 4096-dimensional noise concentrate to ~1 %, so a systematic 1e-3 defect on top of 2e-3 of rounding noise would
    lift e_c and d_c by > 10 %; both are required to stay within 15 % of the ensemble's own values."""
    from anyedit_b200.ddim import DDIMSampler
    from oracle import cpu, ddim_oracle, unet_emul16, unet_oracle
    net, model, sd = sd15
    S, scale, h, K = 20, 7.5, 32, 3
    x_T, c_cat, c_txt, u_txt = inputs_requests(1, h, 1234)
    cond, uncond = _cond(c_cat, c_txt, u_txt)
    out, _ = DDIMSampler(model).sample(S, 1, (4, h, h), cond, verbose=False, x_T=x_T.cuda(), eta=0.0,
                                       unconditional_guidance_scale=scale, unconditional_conditioning=uncond)
    out = out.cpu()
    sched = ddim_oracle.register_schedule("linear", 1000, 0.00085, 0.012)
    torch.set_num_threads(cpu.usable_cores())

    def run(mod, jitter_seed=None):
        xs = [x_T, c_cat, c_txt, u_txt]
        if jitter_seed is not None:
            g = torch.Generator().manual_seed(jitter_seed)
            xs = [t * (1.0 + 2.0 ** -18 * torch.randn(t.shape, generator=g)) for t in xs]
        unet = lambda x, t, context=None, y=None: mod.unet_forward(sd, x, t, context, y, num_heads=SD15["num_heads"])
        model_fn = lambda x, t, c: ddim_oracle.apply_model(unet, "hybrid", x, t, c)
        with torch.no_grad():
            r, _ = ddim_oracle.ddim_sample(model_fn, sched, S, xs[0], {"c_concat": [xs[1]], "c_crossattn": [xs[2]]},
                                           {"c_concat": [xs[1]], "c_crossattn": [xs[3]]}, scale, eta=0.0)
        return r

    fp32 = run(unet_oracle)
    emul = [run(unet_emul16, None if i == 0 else 900 + i) for i in range(K)]
    e_c, e_i = rel(out, fp32), [rel(e, fp32) for e in emul]
    mean_all = sum(emul) / K
    d_c = rel(out, mean_all)
    d_i = [rel(emul[i], (sum(emul) - emul[i]) / (K - 1)) for i in range(K)]
    # leave-one-out means average K-1 members, the full mean K: rescale the members' distances to the same footing
    # (|a|^2 (1 + 1/(K-1)) vs |a|^2 (1 + 1/K))
    d_i = [d * ((1 + 1 / K) / (1 + 1 / (K - 1))) ** 0.5 for d in d_i]
    pair = rel(emul[1], emul[0])
    print(f"[config0 256^2 20 steps CFG 7.5] distance to fp32: cuda {e_c:.3e} | emul16 realisations " + " ".join(f"{e:.3e}" for e in e_i) +
          f" || distance to the emul16 ensemble mean: cuda {d_c:.3e} | members " + " ".join(f"{d:.3e}" for d in d_i) +
          f" || emul16 vs emul16 (inputs jittered 4e-6): {pair:.3e}")
    assert e_c < LATENT_TOL_CFG, e_c
    assert e_c < EXCESS * sum(e_i) / K, (e_c, e_i)
    assert d_c < EXCESS * sum(d_i) / K, (d_c, d_i)
