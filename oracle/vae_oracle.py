"""CPU restatement of the first-stage autoencoder (SURVEY.md 8f rank 1 -- the step right after / before the denoising
loop).  TEST INFRASTRUCTURE (see oracle/__init__.py): groundwork for the next scope row, no product code uses it yet.

Reference (relative to /root/reference):
  ldm/modules/diffusionmodules/model.py   nonlinearity :40-42 (swish), Normalize :45-46 (GroupNorm 32, eps 1e-6),
      Upsample :49-65 (nearest x2 + conv3x3), Downsample :68-88 (zero pad right/bottom by 1, conv3x3 stride 2, pad 0),
      ResnetBlock :91-150, AttnBlock :152-203 (single head, d = C, scale C^-0.5), Encoder :368-545, Decoder :546-641
  ldm/models/autoencoder.py               AutoencoderKL.encode :82-86 (quant_conv -> moments), decode :88-91
      (post_quant_conv -> Decoder); DiagonalGaussianDistribution: mean | logvar = chunk(moments, 2, dim=1), mode = mean
Functional, structure inferred from the state-dict keys (like unet_oracle.py); fp32.
"""
import re

import torch
import torch.nn.functional as F


def _swish(x):
    return x * torch.sigmoid(x)


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _conv(sd, p, x, **kw):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], **kw)


def resnet_block(sd, p, x):
    """model.py:131-150 with temb = None (the autoencoder has no timestep embedding, temb_ch = 0)."""
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)), padding=1)
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)), padding=1)
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(sd, p + ".nin_shortcut", x)
    elif p + ".conv_shortcut.weight" in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=1)
    return x + h


def attn_block(sd, p, x):
    """model.py:176-203: softmax over keys of (q^T k) C^-0.5, one head of width C."""
    h = _gn(sd, p + ".norm", x)
    q, k, v = (_conv(sd, p + "." + n, h) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w = torch.bmm(q, k) * (int(c) ** -0.5)
    w = F.softmax(w, dim=2)
    v = v.reshape(b, c, hh * ww)
    o = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + ".proj_out", o)


def _indices(sd, pattern):
    rx = re.compile(pattern)
    return sorted({int(m.group(1)) for k in sd for m in [rx.match(k)] if m})


def decoder_forward(sd, z, prefix="", tanh_out=False):
    """Decoder.forward (model.py:617-641)."""
    p = prefix
    h = _conv(sd, p + "conv_in", z, padding=1)
    h = resnet_block(sd, p + "mid.block_1", h)
    h = attn_block(sd, p + "mid.attn_1", h)
    h = resnet_block(sd, p + "mid.block_2", h)
    levels = _indices(sd, re.escape(p) + r"up\.(\d+)\.")
    for lvl in reversed(levels):
        for ib in _indices(sd, re.escape(p) + rf"up\.{lvl}\.block\.(\d+)\."):
            h = resnet_block(sd, f"{p}up.{lvl}.block.{ib}", h)
            if f"{p}up.{lvl}.attn.{ib}.norm.weight" in sd:
                h = attn_block(sd, f"{p}up.{lvl}.attn.{ib}", h)
        if f"{p}up.{lvl}.upsample.conv.weight" in sd:
            h = _conv(sd, f"{p}up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"), padding=1)
    h = _conv(sd, p + "conv_out", _swish(_gn(sd, p + "norm_out", h)), padding=1)
    return torch.tanh(h) if tanh_out else h


def encoder_forward(sd, x, prefix=""):
    """Encoder.forward (model.py:517-545)."""
    p = prefix
    h = _conv(sd, p + "conv_in", x, padding=1)
    for lvl in _indices(sd, re.escape(p) + r"down\.(\d+)\."):
        for ib in _indices(sd, re.escape(p) + rf"down\.{lvl}\.block\.(\d+)\."):
            h = resnet_block(sd, f"{p}down.{lvl}.block.{ib}", h)
            if f"{p}down.{lvl}.attn.{ib}.norm.weight" in sd:
                h = attn_block(sd, f"{p}down.{lvl}.attn.{ib}", h)
        if f"{p}down.{lvl}.downsample.conv.weight" in sd:
            h = _conv(sd, f"{p}down.{lvl}.downsample.conv", F.pad(h, (0, 1, 0, 1)), stride=2)    # model.py:83-85
    h = resnet_block(sd, p + "mid.block_1", h)
    h = attn_block(sd, p + "mid.attn_1", h)
    h = resnet_block(sd, p + "mid.block_2", h)
    return _conv(sd, p + "conv_out", _swish(_gn(sd, p + "norm_out", h)), padding=1)


def vae_decode(sd, z):
    """AutoencoderKL.decode (autoencoder.py:88-91): post_quant_conv then the decoder."""
    return decoder_forward(sd, _conv(sd, "post_quant_conv", z), prefix="decoder.")


def vae_encode_moments(sd, x):
    """AutoencoderKL.encode up to the moments (autoencoder.py:82-85); mean = moments[:, :z] is the posterior mode."""
    return _conv(sd, "quant_conv", encoder_forward(sd, x, prefix="encoder."))
