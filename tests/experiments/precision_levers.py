#!/usr/bin/env python
"""Precision levers (CPU, not a test): what would each change to the CUDA path's storage precision buy
against the <= 1e-3 final-latent bar under CFG 7.5?  Uses oracle/unet_emul16.py (fp32 math + fp16 rounding at the kernels'
storage points, statistically equivalent to the CUDA path: tests/test_gpu_parity_shapes.py) with rounding points switched off:

    as built            fp16 weights, fp16 activations everywhere
    fp32 residual       the residual stream (every residual-add epilogue output and what feeds it) kept in fp32
    exact weights       weights not rounded (= hi + lo split-weight contraction: every GEMM twice)
    exact weights + fp32 residual
    fp32 everything     sanity: must reproduce the fp32 oracle

SD-1.5 geometry, one CFG pair (same x, two contexts), single forward and after the guidance combine e_u + 7.5 (e_c - e_u).
    python tests/experiments/precision_levers.py [latent=16] [scheme=torch]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from oracle import unet_emul16, unet_oracle, weights  # noqa: E402


def main():
    h = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    scheme = sys.argv[2] if len(sys.argv) > 2 else "torch"
    meta = json.load(open(os.path.join(os.path.dirname(__file__), "..", "golden", "sd15_keys.json")))
    sd = weights.make_state_dict({k: tuple(v) for k, v in meta["keys"].items()}, 3, scheme=scheme)
    gen = torch.Generator().manual_seed(99)
    x = torch.randn(1, 8, h, h, generator=gen).repeat(2, 1, 1, 1)
    ctx = torch.randn(2, 77, 768, generator=gen)
    t = torch.tensor([981, 981])
    rel = lambda a, b: float((a - b).norm() / b.norm())
    cfg = lambda e: e[0] + 7.5 * (e[1] - e[0])
    with torch.no_grad():
        ref = unet_oracle.unet_forward(sd, x, t, ctx, None, num_heads=8)
        print(f"scheme={scheme} latent={h}")
        for name, pol in (("as built (fp16 weights + activations)", dict(weights=True, residual=True, acts=True)),
                          ("fp32 residual stream", dict(weights=True, residual=False, acts=True)),
                          ("exact weights (hi+lo split, 2x GEMM work)", dict(weights=False, residual=True, acts=True)),
                          ("exact weights + fp32 residual stream", dict(weights=False, residual=False, acts=True)),
                          ("fp16 weights only (activations fp32)", dict(weights=True, residual=False, acts=False)),
                          ("fp32 everything (sanity)", dict(weights=False, residual=False, acts=False))):
            unet_emul16.POLICY.update(pol)
            unet_emul16._W16.clear()
            out = unet_emul16.unet_forward(sd, x, t, ctx, None, num_heads=8)
            print(f"  {name:46s} forward rel {rel(out, ref):.3e}   after CFG 7.5 rel {rel(cfg(out), cfg(ref)):.3e}", flush=True)
    unet_emul16.POLICY.update(dict(weights=True, residual=True, acts=True))


if __name__ == "__main__":
    main()
