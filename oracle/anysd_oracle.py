"""CPU restatement of AnySD's task-aware router and learnable task embedding.

TEST INFRASTRUCTURE (see oracle/__init__.py).  **PARITY UNPINNED**: the AnySD model
code (``AnySD.model.MoE``, ``AnySD.unet.UNet2DConditionAnySD``, ``task_embs_book``) is
an empty, un-pinned git submodule in the reference (.gitmodules:1-4), so
there is no source, test or golden vector to pin against.  This file is therefore a
builder-written *specification by restatement*, assembled from what is in the tree:

  * call signature  train.py:420-424, 483-485, 694-695
        ``MoE(unet, image_encoder, expert_num)``; trainables ``image_proj_model``,
        ``adapter_modules``, ``task_embs``; ``moe(noisy8ch, t, text_ehs, ref_embeds, edit_code)``
  * decoupled cross-attention that each expert follows
        AnyEdit_Collection/other_modules/ip_adapter/attention_processor.py:82-188
        ``hidden = Attn(q, K_txt, V_txt) + scale * Attn(q, K_ip, V_ip)``  (before ``to_out``)
  * the in-tree slot for an embedding added to the time embedding
        ldm/modules/diffusionmodules/openaimodel.py:533-540, 770-772 (``emb += label_emb(y)``)

Specification (E experts, T edit types, D = time_embed_dim, L cross-attention layers in
forward order, visual tokens ``v`` of shape [B, N_vis, ctx_dim], possibly N_vis = 0):

    te    = task_embs[edit_code]                                   [B, D]
    emb   = time_embed(t_emb) + te                                 (a23)
    g_l   = softmax(router_l.weight @ te + router_l.bias)          [B, E]   per layer l
    K_le  = v @ to_k_ip_l[e].T ;  V_le = v @ to_v_ip_l[e].T        [B, N_vis, C_l]
    attn2 = Attn(q, K_txt, V_txt) + sum_e g_l[:, e] * Attn(q, K_le, V_le)   (a22)
    out   = to_out(attn2)

Adapter state-dict keys:  ``task_embs.weight [T, D]``;
``adapter_modules.{l}.router.{weight [E, D], bias [E]}``;
``adapter_modules.{l}.to_k_ip.weight [E*C_l, ctx_dim]``, ``...to_v_ip.weight`` (experts stacked
along rows, expert e = rows [e*C_l, (e+1)*C_l)).

Pinned consistency property (tests): with ``task_embs == 0`` and ``N_vis == 0`` the model
is exactly the reference UNet.
"""
import torch
import torch.nn.functional as F

from . import unet_oracle


def make_hooks(adapter_sd, edit_code, visual_tokens):
    te = F.embedding(edit_code, adapter_sd["task_embs.weight"])          # [B, D]
    hooks = {"emb_extra": lambda emb: te.to(emb.dtype)}
    n_vis = 0 if visual_tokens is None else visual_tokens.shape[1]
    if n_vis > 0:
        def cross_extra(layer, q, heads):
            p = f"adapter_modules.{layer}."
            wr, br = adapter_sd[p + "router.weight"], adapter_sd[p + "router.bias"]
            gate = F.linear(te, wr, br).softmax(dim=-1)                  # [B, E]
            E = wr.shape[0]
            bh, n, d = q.shape
            b = bh // heads
            c = heads * d
            wk = adapter_sd[p + "to_k_ip.weight"].reshape(E, c, -1)
            wv = adapter_sd[p + "to_v_ip.weight"].reshape(E, c, -1)
            out = torch.zeros_like(q)
            for e in range(E):
                k = F.linear(visual_tokens, wk[e])                        # [B, N_vis, C]
                v = F.linear(visual_tokens, wv[e])
                k = k.reshape(b, n_vis, heads, d).permute(0, 2, 1, 3).reshape(bh, n_vis, d)
                v = v.reshape(b, n_vis, heads, d).permute(0, 2, 1, 3).reshape(bh, n_vis, d)
                o = unet_oracle.attention_bhnd(q, k, v)
                ge = gate[:, e].repeat_interleave(heads)[:, None, None]
                out = out + ge * o
            return out
        hooks["cross_extra"] = cross_extra
    return hooks


def anysd_forward(sd, adapter_sd, x, timesteps, context, edit_code, visual_tokens=None, **kw):
    """The MoE.__call__ boundary (train.py:694-695) restated on top of unet_oracle."""
    hooks = make_hooks(adapter_sd, edit_code, visual_tokens)
    return unet_oracle.unet_forward(sd, x, timesteps, context, None, hooks=hooks, **kw)


def adapter_shapes(unet_sd_shapes, num_tasks, num_experts, ctx_dim):
    """Shapes of the adapter state dict for a UNet whose state-dict shapes are given."""
    d_emb = unet_sd_shapes["time_embed.2.weight"][0]
    shapes = {"task_embs.weight": (num_tasks, d_emb)}

    def order(k):
        parts = k.split(".")
        grp = {"input_blocks": 0, "middle_block": 1, "output_blocks": 2}[parts[0]]
        idx = int(parts[1])
        # transformer depth index for ordering within a site
        depth = int(parts[parts.index("transformer_blocks") + 1])
        return (grp, idx if grp != 1 else 0, depth)

    sites = sorted([k for k in unet_sd_shapes if k.endswith("attn2.to_q.weight")], key=order)
    for l, k in enumerate(sites):
        c = unet_sd_shapes[k][0]
        p = f"adapter_modules.{l}."
        shapes[p + "router.weight"] = (num_experts, d_emb)
        shapes[p + "router.bias"] = (num_experts,)
        shapes[p + "to_k_ip.weight"] = (num_experts * c, ctx_dim)
        shapes[p + "to_v_ip.weight"] = (num_experts * c, ctx_dim)
    return shapes
