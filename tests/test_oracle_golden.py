"""Pin the oracle restatement against the golden vectors generated from the real reference
(tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ddim_oracle, unet_oracle, weights

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def _keys(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def test_schedule_tables_bit_exact():
    g = _load("schedule.npz")
    betas = ddim_oracle.make_beta_schedule("linear", 1000, 0.00085, 0.012)
    assert np.array_equal(betas, g["betas"])
    sched = ddim_oracle.register_schedule("linear", 1000, 0.00085, 0.012)
    assert np.array_equal(sched["alphas_cumprod_f64"], g["alphas_cumprod"])
    # KATs recorded in SURVEY.md 8c
    assert sched["alphas_cumprod_f64"][0] == pytest.approx(0.99915, abs=1e-12)
    assert sched["alphas_cumprod_f64"][999] == pytest.approx(0.004660098513077238, rel=1e-12)
    for S in (20, 50, 100):
        ts = ddim_oracle.make_ddim_timesteps("uniform", S, 1000)
        assert np.array_equal(ts, g[f"ts_{S}"])
        for eta in (0.0, 0.5):
            tag = f"{S}_{int(eta * 10)}"
            sig, a, ap = ddim_oracle.make_ddim_sampling_parameters(sched["alphas_cumprod"], ts, eta)
            assert np.array_equal(np.asarray(a, dtype=np.float32), g[f"alphas_{tag}"])
            assert np.array_equal(np.asarray(ap, dtype=np.float64), g[f"alphas_prev_{tag}"])
            assert np.array_equal(np.asarray(sig, dtype=np.float64), g[f"sigmas_{tag}"])
    assert list(ddim_oracle.make_ddim_timesteps("uniform", 50, 1000)[:3]) == [1, 21, 41]
    assert np.array_equal(ddim_oracle.make_ddim_timesteps("quad", 20, 1000), g["ts_quad_20"])


def test_timestep_embedding_golden():
    g = _load("timestep_embedding.npz")
    t = torch.from_numpy(g["t"])
    assert np.array_equal(unet_oracle.timestep_embedding(t, 320).numpy(), g["e320"])
    assert np.array_equal(unet_oracle.timestep_embedding(t, 64).numpy(), g["e64"])


@pytest.mark.parametrize("name", ["tiny_a", "tiny_b"])
def test_unet_forward_golden(name):
    meta = _keys(f"{name}_keys.json")
    g = _load(f"unet_{name}.npz")
    sd = weights.make_state_dict({k: tuple(v) for k, v in meta["keys"].items()}, int(g["seed"]))
    assert weights.checksum(sd) == pytest.approx(float(g["wsum"]), rel=1e-12)
    cfg = meta["config"]
    y = torch.from_numpy(g["y"]) if "y" in g.files else None
    out = unet_oracle.unet_forward(sd, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]),
                                   torch.from_numpy(g["ctx"]), y,
                                   num_heads=cfg.get("num_heads", -1),
                                   num_head_channels=cfg.get("num_head_channels", -1))
    ref = torch.from_numpy(g["out"])
    err = (out - ref).norm() / ref.norm()
    assert ref.abs().max() > 1e-2            # not the zero-init trap
    assert err < 2e-6, float(err)


def test_ddim_loop_golden():
    meta = _keys("tiny_a_keys.json")
    cfg = meta["config"]
    g = _load("ddim_tiny.npz")
    sd = weights.make_state_dict({k: tuple(v) for k, v in meta["keys"].items()}, 11)
    sched = ddim_oracle.register_schedule("linear", 1000, 0.00085, 0.012)
    unet = lambda x, t, context=None, y=None: unet_oracle.unet_forward(
        sd, x, t, context, y, num_heads=cfg["num_heads"])
    model_fn = lambda x, t, c: ddim_oracle.apply_model(unet, "hybrid", x, t, c)
    f = lambda k: torch.from_numpy(g[k])
    cond = {"c_concat": [f("c_cat")], "c_crossattn": [f("c_txt")]}
    uncond = {"c_concat": [f("c_cat")], "c_crossattn": [f("u_txt")]}
    for S, scale in ((10, 7.5), (20, 1.0)):
        out, inter = ddim_oracle.ddim_sample(model_fn, sched, S, f("x_T"), cond, uncond, scale, eta=0.0,
                                             log_every_t=3)
        ref = f(f"final_S{S}")
        err = (out - ref).norm() / ref.norm()
        assert err < 1e-5, (S, float(err))
        assert len(inter["x_inter"]) == int(g[f"n_inter_S{S}"])
        e2 = (inter["pred_x0"][-1] - f(f"pred_x0_last_S{S}")).norm() / f(f"pred_x0_last_S{S}").norm()
        assert e2 < 1e-5
        e3 = (inter["x_inter"][1] - f(f"x_inter_1_S{S}")).norm() / f(f"x_inter_1_S{S}").norm()
        assert e3 < 1e-5


def test_training_step_gradients_golden():
    """a24: autograd through the oracle forward == autograd through the reference UNet (tests/golden/make_golden_train.py):
    q_sample, loss value, and d loss / d {class-embedding table, cross-attention K/V weights, context tokens}."""
    import torch.nn.functional as F
    from oracle import train_oracle
    g = _load("train_tiny_b.npz")
    meta = _keys("tiny_b_keys.json")
    shapes = {k: tuple(v) for k, v in meta["keys"].items()}
    sd = weights.make_state_dict(shapes, int(g["seed"]))
    assert weights.checksum(sd) == pytest.approx(float(g["wsum"]), rel=1e-12)
    sched = ddim_oracle.register_schedule("linear", 1000, 0.00085, 0.012)
    t = torch.tensor(g["t"])
    noisy = train_oracle.q_sample(torch.tensor(g["x0"]), torch.tensor(g["noise"]), t, sched["alphas_cumprod"])
    assert np.abs(noisy.numpy() - g["noisy"]).max() < 1e-6
    gkeys = [k[len("grad__"):] for k in g.files if k.startswith("grad__")]
    leaves = dict(sd)
    for k in gkeys:
        leaves[k] = sd[k].clone().requires_grad_(True)
    ctx = torch.tensor(g["ctx"]).requires_grad_(True)
    cfg = meta["config"]
    with torch.enable_grad():
        pred = unet_oracle.unet_forward(leaves, noisy, t, ctx, torch.tensor(g["y"]), num_heads=cfg.get("num_heads", -1),
                                        num_head_channels=cfg.get("num_head_channels", -1))
        loss = F.mse_loss(pred.float(), torch.tensor(g["noise"]), reduction="mean")
        loss.backward()
    assert np.abs(pred.detach().numpy() - g["pred"]).max() < 5e-6
    assert float(loss.detach()) == pytest.approx(float(g["loss"]), rel=1e-6)
    for k in gkeys:
        ref = g["grad__" + k]
        err = np.abs(leaves[k].grad.numpy() - ref).max() / np.abs(ref).max()
        assert err < 2e-5, (k, err)
    err = np.abs(ctx.grad.numpy() - g["grad_ctx"]).max() / np.abs(g["grad_ctx"]).max()
    assert err < 2e-5, err


def test_adamw_restatement_matches_torch():
    """oracle/train_oracle.adamw_step == torch.optim.AdamW (the optimizer of train.py:486-492), 3 steps."""
    from oracle import train_oracle
    gen = torch.Generator().manual_seed(5)
    p0 = torch.randn(37, 11, generator=gen)
    grads = [torch.randn(37, 11, generator=gen) * 0.1 for _ in range(3)]
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    q, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for i, gr in enumerate(grads, 1):
        p.grad = gr.clone()
        opt.step()
        q, m, v = train_oracle.adamw_step(q, gr, m, v, i, 1e-3, 0.9, 0.999, 1e-8, 1e-2)
        assert torch.allclose(q, p.detach(), rtol=1e-6, atol=1e-7), i


def test_vae_oracle_golden():
    """Groundwork for the next scope row: the autoencoder restatement against the reference's own
    Encoder / Decoder (tests/golden/make_golden_vae.py): encode -> moments and decode, non-square input."""
    from oracle import vae_oracle
    g = _load("vae_tiny.npz")
    meta = _keys("vae_tiny_keys.json")
    sd = weights.make_state_dict({k: tuple(v) for k, v in meta["keys"].items()}, int(g["seed"]))
    assert weights.checksum(sd) == pytest.approx(float(g["wsum"]), rel=1e-12)
    mom = vae_oracle.vae_encode_moments(sd, torch.from_numpy(g["x"]))
    img = vae_oracle.vae_decode(sd, torch.from_numpy(g["z"]))
    assert tuple(mom.shape) == g["moments"].shape and tuple(img.shape) == g["img"].shape
    for got, ref in ((mom, g["moments"]), (img, g["img"])):
        ref = torch.from_numpy(ref)
        assert ref.abs().max() > 1e-2
        assert float((got - ref).norm() / ref.norm()) < 2e-6


def test_plms_oracle_golden():
    """Groundwork for the next scope row: the PLMS restatement against the reference's own
    PLMSSampler run (tests/golden/make_golden_plms.py), with CFG 5.0 (10 steps) and without guidance (20 steps)."""
    from oracle import plms_oracle
    g = _load("plms_tiny.npz")
    meta = _keys("tiny_c_keys.json")
    sd = weights.make_state_dict({k: tuple(v) for k, v in meta["keys"].items()}, int(g["seed"]))
    assert weights.checksum(sd) == pytest.approx(float(g["wsum"]), rel=1e-12)
    cfg = meta["config"]
    sched = ddim_oracle.register_schedule("linear", 1000, 0.00085, 0.012)
    fn = lambda x, t, c: unet_oracle.unet_forward(sd, x, t, c, None, num_heads=cfg.get("num_heads", -1),
                                                  num_head_channels=cfg.get("num_head_channels", -1))
    x_T, c, uc = (torch.from_numpy(g[k]) for k in ("x_T", "c", "uc"))
    for S, scale in ((10, 5.0), (20, 1.0)):
        img, inter = plms_oracle.plms_sample(fn, sched, S, x_T, c, uc if scale != 1.0 else None, scale)
        for got, key in ((img, f"x0_S{S}"), (inter["pred_x0"][-1], f"pred_x0_last_S{S}")):
            ref = torch.from_numpy(g[key])
            assert float((got - ref).norm() / ref.norm()) < 1e-5, (S, key)


def test_dpm_solver_oracle_golden():
    """Groundwork for the next scope row: DPM-Solver++(2M) as the reference configures it
    (tests/golden/make_golden_dpm.py): 10 steps with CFG 5.0 (order-1 final step) and 20 steps without guidance."""
    from oracle import dpm_oracle
    g = _load("dpm_tiny.npz")
    meta = _keys("tiny_c_keys.json")
    sd = weights.make_state_dict({k: tuple(v) for k, v in meta["keys"].items()}, int(g["seed"]))
    assert weights.checksum(sd) == pytest.approx(float(g["wsum"]), rel=1e-12)
    cfg = meta["config"]
    sched = ddim_oracle.register_schedule("linear", 1000, 0.00085, 0.012)
    fn = lambda x, t, c: unet_oracle.unet_forward(sd, x, t, c, None, num_heads=cfg.get("num_heads", -1),
                                                  num_head_channels=cfg.get("num_head_channels", -1))
    x_T, c, uc = (torch.from_numpy(g[k]) for k in ("x_T", "c", "uc"))
    for S, scale in ((10, 5.0), (20, 1.0)):
        img = dpm_oracle.dpm_solver_pp_2m(fn, sched["alphas_cumprod"], S, x_T, c, uc, scale)
        ref = torch.from_numpy(g[f"x0_S{S}"])
        assert float((img - ref).norm() / ref.norm()) < 2e-5, S
