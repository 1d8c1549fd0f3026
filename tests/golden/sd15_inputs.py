"""Seeded inputs shared by tests/golden/make_golden_sd15.py (which runs the real reference on them) and the GPU
parity tests (which regenerate them on the GPU box): the fixtures then only hold the reference's OUTPUTS plus a
float64 checksum of every input, so drift of the generator is detected instead of silently compared."""
import numpy as np
import torch

SD15 = dict(image_size=32, in_channels=8, model_channels=320, out_channels=4, num_res_blocks=2,
            attention_resolutions=[4, 2, 1], channel_mult=[1, 2, 4, 4], num_heads=8,
            use_spatial_transformer=True, transformer_depth=1, context_dim=768, legacy=False)
WEIGHT_SEED, WEIGHT_SCHEME = 3, "torch"


def inputs_requests(n, h, seed):
    """n edit requests at an h x h latent: x_T, c_concat (source-image latent), text context; one shared null text."""
    g = torch.Generator().manual_seed(seed)
    x_T, c_cat = torch.randn(n, 4, h, h, generator=g), torch.randn(n, 4, h, h, generator=g)
    c_txt = torch.randn(n, 77, 768, generator=g)
    u_txt = torch.randn(1, 77, 768, generator=g)
    return x_T, c_cat, c_txt, u_txt


def checksum(*ts):
    return np.asarray([float(t.double().sum()) + 0.5 * float((t.double() ** 2).sum()) for t in ts])
