"""fp32 CPU restatement of the reference UNet forward (TEST INFRASTRUCTURE).

Functional: takes a reference-layout ``state_dict`` and walks it.  The block
structure is inferred from the *parameter names* (not from constructor kwargs)
so that it is derived independently of ``anyedit_b200.unet``.

Reference (relative to /root/reference):
  ldm/modules/diffusionmodules/openaimodel.py  UNetModel.forward :754-786,
      ResBlock._forward :254-274, Upsample :90-118, Downsample :133-159,
      TimestepEmbedSequential :79-87
  ldm/modules/attention.py  SpatialTransformer.forward :321-340,
      BasicTransformerBlock._forward :271-275, CrossAttention.forward :163-194,
      GEGLU :49-57, Normalize :88-89
  ldm/modules/diffusionmodules/util.py  timestep_embedding :154-174,
      GroupNorm32 :217-219
"""
import math
import re

import torch
import torch.nn.functional as F


def timestep_embedding(timesteps, dim, max_period=10000):
    # util.py:154-174 -- cos half first, then sin; fp32.
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def group_norm32(x, w, b, eps=1e-5):
    # util.py:217-219: GroupNorm(32, C) evaluated in fp32, cast back.
    return F.group_norm(x.float(), 32, w, b, eps).type(x.dtype)


def _attention(sd, p, x, context, heads, extra=None):
    # attention.py:163-194 (eager CrossAttention; ATTN_PRECISION=fp32 path).
    # ``extra(q_heads) -> [b*h, n, d]`` is added before the head merge / to_out, the place
    # IP-Adapter-style decoupled attention adds its stream
    # (AnyEdit_Collection/other_modules/ip_adapter/attention_processor.py:160-176).
    q = F.linear(x, sd[p + "to_q.weight"])
    ctx = x if context is None else context
    k = F.linear(ctx, sd[p + "to_k.weight"])
    v = F.linear(ctx, sd[p + "to_v.weight"])
    b, n, c = q.shape
    d = c // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q.float(), k.float()) * (d ** -0.5)
    sim = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", sim, v)
    if extra is not None:
        out = out + extra(q)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, c)
    return F.linear(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def _transformer_block(sd, p, x, context, heads, hooks=None, site=None):
    # attention.py:271-275 ; LayerNorm eps 1e-5 (nn.LayerNorm default).
    c = x.shape[-1]
    h = F.layer_norm(x, (c,), sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    x = _attention(sd, p + "attn1.", h, None, heads) + x
    h = F.layer_norm(x, (c,), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    extra = None
    if hooks is not None and "cross_extra" in hooks:
        # AnySD visual-expert stream (oracle/anysd_oracle.py)
        layer = hooks["_layer"][0]
        hooks["_layer"][0] += 1
        extra = lambda q: hooks["cross_extra"](layer, q, heads)
    x = _attention(sd, p + "attn2.", h, context, heads, extra) + x
    h = F.layer_norm(x, (c,), sd[p + "norm3.weight"], sd[p + "norm3.bias"])
    # GEGLU (attention.py:49-57): proj -> chunk(2) -> a * gelu(gate) (exact erf gelu)
    g = F.linear(h, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"])
    a, gate = g.chunk(2, dim=-1)
    h = a * F.gelu(gate)
    x = F.linear(h, sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"]) + x
    return x


def _spatial_transformer(sd, p, x, context, heads_of, hooks=None):
    # attention.py:321-340
    b, c, hh, ww = x.shape
    x_in = x
    x = F.group_norm(x, 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    use_linear = sd[p + "proj_in.weight"].dim() == 2
    if not use_linear:
        x = F.conv2d(x, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    x = x.permute(0, 2, 3, 1).reshape(b, hh * ww, -1)
    if use_linear:
        x = F.linear(x, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    inner = x.shape[-1]
    heads = heads_of(inner)
    depth = 0
    while (p + f"transformer_blocks.{depth}.norm1.weight") in sd:
        ctx = context[depth] if isinstance(context, (list, tuple)) else context
        x = _transformer_block(sd, p + f"transformer_blocks.{depth}.", x, ctx, heads, hooks,
                               site=p + f"transformer_blocks.{depth}.")
        depth += 1
    if use_linear:
        x = F.linear(x, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    x = x.reshape(b, hh, ww, -1).permute(0, 3, 1, 2)
    if not use_linear:
        x = F.conv2d(x, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return x + x_in


def _resblock(sd, p, x, emb):
    # openaimodel.py:254-274 (no up/down, no scale-shift norm: the SD configuration)
    h = group_norm32(x, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"])
    h = F.silu(h)
    h = F.conv2d(h, sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"]).type(h.dtype)
    h = h + e[:, :, None, None]
    h = group_norm32(h, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"])
    h = F.silu(h)
    h = F.conv2d(h, sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
    if (p + "skip_connection.weight") in sd:
        w = sd[p + "skip_connection.weight"]
        x = F.conv2d(x, w, sd[p + "skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def _run_block(sd, p, h, emb, context, heads_of, hooks=None):
    """One TimestepEmbedSequential (openaimodel.py:79-87): children p+'0.', p+'1.', ..."""
    j = 0
    while True:
        q = f"{p}{j}."
        if (q + "in_layers.0.weight") in sd:
            h = _resblock(sd, q, h, emb)
        elif (q + "transformer_blocks.0.norm1.weight") in sd:
            h = _spatial_transformer(sd, q, h, context, heads_of, hooks)
        elif (q + "op.weight") in sd:        # Downsample conv3x3 stride 2 pad 1 (:133-159)
            h = F.conv2d(h, sd[q + "op.weight"], sd[q + "op.bias"], stride=2, padding=1)
        elif (q + "conv.weight") in sd:      # Upsample nearest x2 + conv3x3 (:90-118)
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[q + "conv.weight"], sd[q + "conv.bias"], padding=1)
        else:
            break
        j += 1
    return h


def _count(sd, prefix):
    idx = set()
    pat = re.compile(re.escape(prefix) + r"\.(\d+)\.")
    for k in sd:
        m = pat.match(k)
        if m:
            idx.add(int(m.group(1)))
    return (max(idx) + 1) if idx else 0


def unet_forward(sd, x, timesteps, context=None, y=None, *, num_heads=-1, num_head_channels=-1,
                 hooks=None, control=None, only_mid_control=False):
    """openaimodel.py:754-786.  ``sd`` holds fp32 CPU tensors in the reference key layout.

    ``control`` (list of 13 residuals) follows ControlledUnetModel.forward,
    AnyEdit_Collection/other_modules/cldm/cldm.py:22-44.
    ``hooks`` lets anysd_oracle add the task-embedding / visual-expert stream.
    """
    def heads_of(inner):
        if num_head_channels == -1:
            return num_heads
        return inner // num_head_channels

    mc = sd["time_embed.0.weight"].shape[1]
    t_emb = timestep_embedding(timesteps, mc)
    emb = F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    has_label = "label_emb.weight" in sd
    assert (y is not None) == has_label, "must specify y if and only if the model is class-conditional"
    if has_label:
        emb = emb + F.embedding(y, sd["label_emb.weight"])         # :770-772
    if hooks is not None:
        hooks["_layer"] = [0]
        if "emb_extra" in hooks:
            emb = emb + hooks["emb_extra"](emb)

    n_in = _count(sd, "input_blocks")
    n_out = _count(sd, "output_blocks")
    hs = []
    h = x.float()
    for i in range(n_in):
        if i == 0:
            h = F.conv2d(h, sd["input_blocks.0.0.weight"], sd["input_blocks.0.0.bias"], padding=1)
        else:
            h = _run_block(sd, f"input_blocks.{i}.", h, emb, context, heads_of, hooks)
        hs.append(h)
    h = _run_block(sd, "middle_block.", h, emb, context, heads_of, hooks)
    if control is not None:
        control = list(control)
        h = h + control.pop()
    for i in range(n_out):
        skip = hs.pop()
        if control is not None and not only_mid_control:
            skip = skip + control.pop()
        h = torch.cat([h, skip], dim=1)                              # :780 current features first
        h = _run_block(sd, f"output_blocks.{i}.", h, emb, context, heads_of, hooks)
    h = h.type(x.dtype)
    h = group_norm32(h, sd["out.0.weight"], sd["out.0.bias"])
    h = F.silu(h)
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


# ---- per-op helpers used by kernel-level parity tests --------------------------------------

def groupnorm_silu_nchw(x, w, b, eps, silu=True):
    h = F.group_norm(x.float(), 32, w, b, eps)
    return F.silu(h) if silu else h


def attention_bhnd(q, k, v):
    """softmax(q k^T / sqrt(d)) v on [B*h, n, d] fp32 (attention.py:171-193)."""
    d = q.shape[-1]
    sim = torch.einsum("bid,bjd->bij", q.float(), k.float()) * (d ** -0.5)
    return torch.einsum("bij,bjd->bid", sim.softmax(dim=-1), v.float())
