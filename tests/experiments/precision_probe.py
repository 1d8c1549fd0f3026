#!/usr/bin/env python
"""Precision budget probe (CPU, not a test): which fp16 rounding points dominate the single-forward error
of the SD-1.5-geometry UNet?  Emulates rounding policies on top of the fp32 oracle by swapping the functional
namespace the oracle calls (F.linear / F.conv2d inputs = MMA operands, outputs = activation storage).

  P0  weights rounded to fp16 only
  P1  P0 + GEMM/conv input operands rounded (the floor for fp16 MMA with everything else fp32)
  P2  P1 + GEMM/conv outputs rounded (fp16 activation storage ~ the CUDA path / fp16 autocast)
"""
import json
import os
import sys
import types

import torch
import torch.nn.functional as RF

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from oracle import unet_oracle, weights  # noqa: E402

r16 = lambda t: t.half().float()


def make_F(round_w, round_in, round_out):
    ns = types.SimpleNamespace(**{k: getattr(RF, k) for k in dir(RF) if not k.startswith("_")})

    def linear(x, w, b=None):
        y = RF.linear(r16(x) if round_in else x, r16(w) if round_w else w, b)
        return r16(y) if round_out else y

    def conv2d(x, w, b=None, **kw):
        y = RF.conv2d(r16(x) if round_in else x, r16(w) if round_w else w, b, **kw)
        return r16(y) if round_out else y

    ns.linear, ns.conv2d = linear, conv2d
    return ns


def main():
    scheme = sys.argv[1] if len(sys.argv) > 1 else "unit"
    h = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    meta = json.load(open(os.path.join(os.path.dirname(__file__), "..", "golden", "sd15_keys.json")))
    sd = weights.make_state_dict({k: tuple(v) for k, v in meta["keys"].items()}, 3, scheme=scheme)
    gen = torch.Generator().manual_seed(99)
    x = torch.randn(1, 8, h, h, generator=gen).repeat(2, 1, 1, 1)          # CFG pair: same x, different text
    ctx = torch.randn(2, 77, 768, generator=gen)
    t = torch.tensor([981, 981])
    rel = lambda a, b: float((a - b).norm() / b.norm())
    with torch.no_grad():
        ref = unet_oracle.unet_forward(sd, x, t, ctx, None, num_heads=8)
        cfg = lambda e: e[0] + 7.5 * (e[1] - e[0])
        print(f"scheme={scheme} latent={h}: |eps| rms {float(ref.pow(2).mean().sqrt()):.3f}  |cfg eps| rms {float(cfg(ref).pow(2).mean().sqrt()):.3f}")
        for name, pol in (("P0 weights", (1, 0, 0)), ("P1 +operands", (1, 1, 0)), ("P2 +storage", (1, 1, 1)), ("acts only (in+out)", (0, 1, 1))):
            unet_oracle.F = make_F(*pol)
            out = unet_oracle.unet_forward(sd, x, t, ctx, None, num_heads=8)
            unet_oracle.F = RF
            print(f"  {name:22s} forward rel {rel(out, ref):.3e}   after CFG 7.5 rel {rel(cfg(out), cfg(ref)):.3e}")


if __name__ == "__main__":
    main()
