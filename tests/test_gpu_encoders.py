"""Condition encoders on the GPU: ``anyedit_b200.encoders`` against golden outputs of
``transformers``' own CLIP towers and the reference's own Resampler (tests/golden/make_golden_encoders.py)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, G)
FWD_TOL = 4e-3


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def gold():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return np.load(os.path.join(G, "encoders_tiny.npz")), json.load(open(os.path.join(G, "encoders_keys.json")))


def _load(mod, keys, seed, wsum):
    from oracle import weights
    sd = weights.make_state_dict({k: tuple(v) for k, v in keys.items()}, seed)
    assert weights.checksum(sd) == pytest.approx(float(wsum), rel=1e-12)
    assert {k: list(v.shape) for k, v in mod.state_dict().items()} == keys          # same keys and shapes as the source library
    mod.load_state_dict(sd, strict=True)
    return mod.cuda()


def test_attention_small_causal_and_embed_tokens():
    from anyedit_b200 import ops
    gen = torch.Generator().manual_seed(2)
    B, H, n, d = 2, 3, 77, 64
    qkv = torch.randn(B * n, 3 * H * d, generator=gen).half()
    out = torch.empty(B * n, H * d, dtype=torch.float16, device="cuda")
    for causal in (True, False):
        ops.attention_small(qkv.cuda(), qkv.cuda()[:, H * d:], qkv.cuda()[:, 2 * H * d:], out, B, H, n, n, d, 3 * H * d, 3 * H * d, 3 * H * d,
                            H * d, causal=causal)
        q, k, v = (qkv.float()[:, i * H * d:(i + 1) * H * d].reshape(B, n, H, d).permute(0, 2, 1, 3) for i in range(3))
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal).permute(0, 2, 1, 3).reshape(B * n, H * d)
        assert rel(out, ref) < 1e-3, causal
    ids = torch.randint(0, 50, (2, 9), generator=gen)
    tok, pos = torch.randn(50, 32, generator=gen).half(), torch.randn(12, 32, generator=gen).half()
    o = torch.empty(18, 32, dtype=torch.float16, device="cuda")
    ops.embed_tokens(ids.cuda(), tok.cuda(), pos.cuda(), o)
    assert torch.equal(o.cpu(), (tok[ids].float() + pos[:9].float()).half().reshape(18, 32))


def test_clip_text_tower_vs_transformers(gold):
    """FrozenCLIPEmbedder / CLIPTextModel: last hidden state, pooled (end-of-text) features, hidden_states[-2]."""
    from anyedit_b200.encoders import FrozenCLIPEmbedder
    from make_golden_encoders import token_ids
    g, meta = gold
    emb = FrozenCLIPEmbedder(version=meta["text_config"], layer="hidden", layer_idx=-2)
    _load(emb.transformer, meta["keys"]["text"], 81, g["text_wsum"])
    ids = token_ids().cuda()
    out = emb.transformer(input_ids=ids, output_hidden_states=True)
    errs = {"last": rel(out.last_hidden_state, torch.from_numpy(g["text_last"])), "pooled": rel(out.pooler_output, torch.from_numpy(g["text_pooled"])),
            "hidden[-2]": rel(out.hidden_states[-2], torch.from_numpy(g["text_hidden_m2"]))}
    print("[clip text tower] rel-L2 vs transformers:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < FWD_TOL, errs
    assert rel(emb(ids), torch.from_numpy(g["text_hidden_m2"])) < FWD_TOL
    emb.layer = "last"
    assert rel(emb.encode(ids), torch.from_numpy(g["text_last"])) < FWD_TOL


def test_clip_vision_resampler_imgproj(gold):
    """train.py:688-694's visual stream: CLIP vision hidden_states[-2] -> Resampler -> visual tokens; and the pooled image
    embedding -> ImageProjModel."""
    from anyedit_b200.encoders import CLIPVisionModelWithProjection, ImageProjModel, Resampler
    from make_golden import randn
    g, meta = gold
    vm = _load(CLIPVisionModelWithProjection(meta["vision_config"]), meta["keys"]["vision"], 82, g["vision_wsum"])
    ov = vm(pixel_values=randn(72, 2, 3, 168, 168).cuda(), output_hidden_states=True)
    e_h, e_e = rel(ov.hidden_states[-2], torch.from_numpy(g["vision_hidden_m2"])), rel(ov.image_embeds, torch.from_numpy(g["vision_embeds"]))
    print(f"[clip vision tower] hidden_states[-2] rel-L2 {e_h:.2e}, image_embeds {e_e:.2e}")
    assert e_h < FWD_TOL and e_e < FWD_TOL
    rs = _load(Resampler(**meta["resampler_config"]), meta["keys"]["resampler"], 83, g["resampler_wsum"])
    tokens = rs(torch.from_numpy(g["vision_hidden_m2"]).cuda())
    e_r = rel(tokens, torch.from_numpy(g["resampler_out"]))
    print(f"[resampler] visual tokens rel-L2 vs reference = {e_r:.2e}")
    assert tuple(tokens.shape) == (2, 16, 64) and e_r < FWD_TOL
    from oracle import weights
    ip = ImageProjModel(**meta["imgproj_config"])
    ip.load_state_dict(weights.make_state_dict({k: tuple(v) for k, v in meta["keys"]["imgproj"].items()}, 84), strict=True)
    e_p = rel(ip.cuda()(torch.from_numpy(g["vision_embeds"]).cuda()), torch.from_numpy(g["imgproj_out"]))
    print(f"[image proj model] rel-L2 vs reference = {e_p:.2e}")
    assert e_p < FWD_TOL
