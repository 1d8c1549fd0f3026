#!/usr/bin/env python
"""Generate the committed golden vectors by running the REAL reference modules on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference pins no vectors of its own for this path, so these files,
produced by importing the reference's own ``UNetModel`` / ``DDIMSampler`` / ``util`` helpers,
are what pins the oracle restatement (tests/test_oracle_golden.py) and, through it, the CUDA
path.  Weights come from ``oracle.weights`` (name-keyed numpy stream), so fixtures only store
inputs, outputs and a weight checksum.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_import, weights  # noqa: E402

TINY_A = dict(image_size=16, in_channels=8, model_channels=64, out_channels=4, num_res_blocks=1,
              attention_resolutions=[1, 2, 4], channel_mult=[1, 2, 4], num_heads=4,
              use_spatial_transformer=True, transformer_depth=1, context_dim=64, legacy=False)
TINY_B = dict(image_size=16, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=2,
              attention_resolutions=[2, 1], channel_mult=[1, 2], num_head_channels=32,
              use_spatial_transformer=True, use_linear_in_transformer=True, transformer_depth=1,
              context_dim=96, legacy=False, num_classes=5, use_checkpoint=True)
SD15 = dict(image_size=32, in_channels=8, model_channels=320, out_channels=4, num_res_blocks=2,
            attention_resolutions=[4, 2, 1], channel_mult=[1, 2, 4, 4], num_heads=8,
            use_spatial_transformer=True, transformer_depth=1, context_dim=768, legacy=False)
ANYDOOR = dict(image_size=32, in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2,
               attention_resolutions=[4, 2, 1], channel_mult=[1, 2, 4, 4], num_head_channels=64,
               use_spatial_transformer=True, use_linear_in_transformer=True, transformer_depth=1,
               context_dim=1024, legacy=False, use_checkpoint=True)


def randn(seed, *shape):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def build(UNetModel, cfg, seed):
    torch.manual_seed(0)
    net = UNetModel(**cfg).eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = weights.make_state_dict(shapes, seed)
    net.load_state_dict(sd)
    return net, shapes, sd


def main():
    UNetModel, DDIMSamplerCPU, util = ref_import.load()
    sys.path.insert(0, ref_import.REF_ROOT)
    from ldm.modules import attention as ref_attn
    from ldm.modules.diffusionmodules import openaimodel as ref_oai
    torch.set_grad_enabled(False)

    # ---- 1. schedule tables (util.py:21-74, ddpm.py:145-147) ----------------------------
    betas = util.make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
    acp = np.cumprod(1.0 - betas, axis=0)
    out = {"betas": betas, "alphas_cumprod": acp}
    acp32 = torch.tensor(acp, dtype=torch.float32)
    for S in (20, 50, 100):
        ts = util.make_ddim_timesteps("uniform", S, 1000, verbose=False)
        out[f"ts_{S}"] = ts.astype(np.int64)
        for eta in (0.0, 0.5):
            sig, a, ap = util.make_ddim_sampling_parameters(acp32, ts, eta, verbose=False)
            tag = f"{S}_{int(eta * 10)}"
            out[f"sigmas_{tag}"] = np.asarray(sig, dtype=np.float64)
            out[f"alphas_{tag}"] = np.asarray(a, dtype=np.float32)
            out[f"alphas_prev_{tag}"] = np.asarray(ap, dtype=np.float64)
    out["ts_quad_20"] = util.make_ddim_timesteps("quad", 20, 1000, verbose=False).astype(np.int64)
    np.savez(os.path.join(HERE, "schedule.npz"), **out)

    # ---- 2. timestep embedding (util.py:154-174) ----------------------------------------
    t = torch.tensor([1, 21, 501, 981, 999], dtype=torch.long)
    np.savez(os.path.join(HERE, "timestep_embedding.npz"), t=t.numpy(),
             e320=util.timestep_embedding(t, 320).numpy(), e64=util.timestep_embedding(t, 64).numpy())

    # ---- 3. state-dict key/shape listings ------------------------------------------------
    for name, cfg in (("sd15", SD15), ("anydoor", ANYDOOR), ("tiny_a", TINY_A), ("tiny_b", TINY_B)):
        with torch.device("meta"):
            net = UNetModel(**cfg)
        keys = {k: list(v.shape) for k, v in net.state_dict().items()}
        with open(os.path.join(HERE, f"{name}_keys.json"), "w") as f:
            json.dump({"config": cfg, "keys": keys, "n_params": sum(int(np.prod(s)) for s in keys.values())}, f)

    # ---- 4. tiny UNet forward (openaimodel.py:754-786) -----------------------------------
    net_a, shapes_a, sd_a = build(UNetModel, TINY_A, seed=11)
    x = randn(1, 2, 8, 16, 16)
    ctx = randn(2, 2, 7, 64)
    tt = torch.tensor([981, 21], dtype=torch.long)
    y_a = net_a(x, tt, context=ctx)
    np.savez(os.path.join(HERE, "unet_tiny_a.npz"), x=x.numpy(), t=tt.numpy(), ctx=ctx.numpy(),
             out=y_a.numpy(), wsum=weights.checksum(sd_a), seed=11)

    net_b, shapes_b, sd_b = build(UNetModel, TINY_B, seed=12)
    xb = randn(3, 3, 4, 8, 12)                       # non-square, odd batch
    ctxb = randn(4, 3, 5, 96)
    tb = torch.tensor([1, 500, 999], dtype=torch.long)
    yb = torch.tensor([0, 4, 2], dtype=torch.long)
    y_b = net_b(xb, tb, context=ctxb, y=yb)
    np.savez(os.path.join(HERE, "unet_tiny_b.npz"), x=xb.numpy(), t=tb.numpy(), ctx=ctxb.numpy(),
             y=yb.numpy(), out=y_b.numpy(), wsum=weights.checksum(sd_b), seed=12)

    # ---- 5. per-op KATs from the reference's own module classes --------------------------
    kat = {}
    # GroupNorm32 + SiLU (util.py:202-219, openaimodel.py:200-203)
    gx = randn(5, 2, 64, 6, 10) * 2.0 + 0.5
    gn = util.normalization(64)
    gsd = weights.make_state_dict({"in_layers.0.weight": (64,), "in_layers.0.bias": (64,)}, 21)
    gn.load_state_dict({"weight": gsd["in_layers.0.weight"], "bias": gsd["in_layers.0.bias"]})
    kat["gn_x"] = gx.numpy()
    kat["gn_silu_out"] = torch.nn.SiLU()(gn(gx)).numpy()
    # CrossAttention (attention.py:145-194): self and cross
    ca = ref_attn.CrossAttention(query_dim=64, context_dim=48, heads=4, dim_head=16)
    ca_shapes = {k: tuple(v.shape) for k, v in ca.state_dict().items()}
    ca_sd = weights.make_state_dict({"attn2." + k: s for k, s in ca_shapes.items()}, 22)
    ca.load_state_dict({k[len("attn2."):]: v for k, v in ca_sd.items()})
    ax = randn(6, 2, 40, 64)
    actx = randn(7, 2, 9, 48)
    kat["ca_x"], kat["ca_ctx"], kat["ca_out"] = ax.numpy(), actx.numpy(), ca(ax, context=actx).numpy()
    # FeedForward GEGLU (attention.py:49-76)
    ff = ref_attn.FeedForward(64, glu=True)
    ff_shapes = {k: tuple(v.shape) for k, v in ff.state_dict().items()}
    ff_sd = weights.make_state_dict({"ff." + k: s for k, s in ff_shapes.items()}, 23)
    ff.load_state_dict({k[len("ff."):]: v for k, v in ff_sd.items()})
    kat["ff_out"] = ff(ax).numpy()
    # ResBlock with channel change (openaimodel.py:162-274)
    rb = ref_oai.ResBlock(64, 128, 0.0, out_channels=96)
    rb_shapes = {k: tuple(v.shape) for k, v in rb.state_dict().items()}
    rb_sd = weights.make_state_dict({"rb.0." + k: s for k, s in rb_shapes.items()}, 24)
    rb.load_state_dict({k[len("rb.0."):]: v for k, v in rb_sd.items()})
    remb = randn(8, 2, 128)
    kat["rb_x"], kat["rb_emb"], kat["rb_out"] = gx.numpy(), remb.numpy(), rb(gx, remb).numpy()
    # Downsample / Upsample (openaimodel.py:90-159)
    dn = ref_oai.Downsample(64, True, out_channels=64)
    dn_sd = weights.make_state_dict({"dn.0.op.weight": (64, 64, 3, 3), "dn.0.op.bias": (64,)}, 25)
    dn.load_state_dict({"op.weight": dn_sd["dn.0.op.weight"], "op.bias": dn_sd["dn.0.op.bias"]})
    up = ref_oai.Upsample(64, True, out_channels=64)
    up_sd = weights.make_state_dict({"up.0.conv.weight": (64, 64, 3, 3), "up.0.conv.bias": (64,)}, 26)
    up.load_state_dict({"conv.weight": up_sd["up.0.conv.weight"], "conv.bias": up_sd["up.0.conv.bias"]})
    kat["down_out"], kat["up_out"] = dn(gx).numpy(), up(gx).numpy()
    np.savez(os.path.join(HERE, "op_kats.npz"), **kat)

    # ---- 6. DDIM + CFG end to end on the tiny UNet (ddim.py:55-251) ----------------------
    from oracle import ddim_oracle
    sched = ddim_oracle.register_schedule("linear", 1000, 0.00085, 0.012)
    shim = ref_import.RefModelShim(net_a, sched, "hybrid")
    sampler = DDIMSamplerCPU(shim)
    B = 2
    x_T = randn(31, B, 4, 16, 16)
    c_cat = randn(32, B, 4, 16, 16)
    c_txt = randn(33, B, 7, 64)
    u_txt = randn(34, 1, 7, 64).repeat(B, 1, 1)
    cond = {"c_concat": [c_cat], "c_crossattn": [c_txt]}
    uncond = {"c_concat": [c_cat], "c_crossattn": [u_txt]}
    res = {"x_T": x_T.numpy(), "c_cat": c_cat.numpy(), "c_txt": c_txt.numpy(), "u_txt": u_txt.numpy()}
    for S, scale in ((10, 7.5), (20, 1.0)):
        samples, inter = sampler.sample(S, B, (4, 16, 16), cond, verbose=False, x_T=x_T, eta=0.0,
                                        unconditional_guidance_scale=scale,
                                        unconditional_conditioning=uncond, log_every_t=3)
        res[f"final_S{S}"] = samples.numpy()
        res[f"n_inter_S{S}"] = len(inter["x_inter"])
        res[f"pred_x0_last_S{S}"] = inter["pred_x0"][-1].numpy()
        res[f"x_inter_1_S{S}"] = inter["x_inter"][1].numpy()
    np.savez(os.path.join(HERE, "ddim_tiny.npz"), **res)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
