#!/usr/bin/env python
"""Golden vectors for the condition encoders.

The CLIP towers' arithmetic is not in the reference tree: ``FrozenCLIPEmbedder`` (ldm/modules/encoders/modules.py:107-150) and
train.py:688-691 call ``transformers``' ``CLIPTextModel`` / ``CLIPVisionModelWithProjection`` (unpinned in requirements.txt;
this image has the version printed below).  Their outputs on seeded small configurations are produced here with that library's
own modules; the Resampler / ImageProjModel outputs with the reference's own classes
(AnyEdit_Collection/other_modules/ip_adapter/resampler.py, ip_adapter.py).  Inputs are regenerated from seeds by the tests.
Usage: python tests/golden/make_golden_encoders.py"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)

from oracle import ref_import, weights  # noqa: E402
from make_golden import randn  # noqa: E402

TEXT = dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2,
            max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=2)
VISION = dict(hidden_size=256, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4, image_size=168, patch_size=14,
              num_channels=3, projection_dim=64, hidden_act="gelu")
RESAMPLER = dict(dim=256, depth=2, dim_head=64, heads=4, num_queries=16, embedding_dim=256, output_dim=64, ff_mult=4)
IMGPROJ = dict(cross_attention_dim=64, clip_embeddings_dim=64, clip_extra_context_tokens=4)


def token_ids():
    g = torch.Generator().manual_seed(71)
    ids = torch.randint(3, 999, (2, 77), generator=g)
    ids[0, 20:], ids[1, 50:] = 999, 999          # end-of-text token = the largest id (argmax pooling, legacy eos id 2) + padding
    return ids


def load(m, seed):
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
    sd = weights.make_state_dict(shapes, seed)
    m.load_state_dict(sd, strict=False)
    return shapes, sd


def main():
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPVisionConfig, CLIPVisionModelWithProjection
    import importlib.util        # the ip_adapter package __init__ pulls in diffusers: load resampler.py by path
    spec = importlib.util.spec_from_file_location(
        "ref_resampler", os.path.join(ref_import.SRC_ROOT, "AnyEdit_Collection", "other_modules", "ip_adapter", "resampler.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    Resampler = mod.Resampler
    torch.set_grad_enabled(False)
    out, keys = {"transformers": transformers.__version__}, {}
    tm = CLIPTextModel(CLIPTextConfig(**TEXT)).eval()
    keys["text"], sd = load(tm, 81)
    o = tm(input_ids=token_ids(), output_hidden_states=True)
    out.update(text_last=o.last_hidden_state.numpy(), text_pooled=o.pooler_output.numpy(), text_hidden_m2=o.hidden_states[-2].numpy(),
               text_wsum=weights.checksum(sd))
    vm = CLIPVisionModelWithProjection(CLIPVisionConfig(**VISION)).eval()
    keys["vision"], sd = load(vm, 82)
    ov = vm(pixel_values=randn(72, 2, 3, 168, 168), output_hidden_states=True)
    out.update(vision_hidden_m2=ov.hidden_states[-2].numpy(), vision_embeds=ov.image_embeds.numpy(), vision_wsum=weights.checksum(sd))
    rs = Resampler(**RESAMPLER).eval()
    keys["resampler"], sd = load(rs, 83)
    out.update(resampler_out=rs(ov.hidden_states[-2]).numpy(), resampler_wsum=weights.checksum(sd))
    # ImageProjModel (ip_adapter.py:28-46) written out with torch modules of the same names (ip_adapter.py imports diffusers)
    proj, norm = torch.nn.Linear(64, 4 * 64), torch.nn.LayerNorm(64)
    shapes = {"proj.weight": (256, 64), "proj.bias": (256,), "norm.weight": (64,), "norm.bias": (64,)}
    sd = weights.make_state_dict(shapes, 84)
    proj.load_state_dict({"weight": sd["proj.weight"], "bias": sd["proj.bias"]})
    norm.load_state_dict({"weight": sd["norm.weight"], "bias": sd["norm.bias"]})
    keys["imgproj"] = shapes
    out["imgproj_out"] = norm(proj(ov.image_embeds).reshape(-1, 4, 64)).numpy()
    with open(os.path.join(HERE, "encoders_keys.json"), "w") as f:
        json.dump({"text_config": TEXT, "vision_config": VISION, "resampler_config": RESAMPLER, "imgproj_config": IMGPROJ,
                   "keys": {n: {k: list(v) for k, v in d.items()} for n, d in keys.items()}}, f)
    np.savez(os.path.join(HERE, "encoders_tiny.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
