"""Training step of the AnySD adapters (reference train.py:629-710) on the B200 kernels.

The reference runs ``loss = mse(MoE(cat(q_sample(z, eps, t), z_orig), t, text, ref_embeds, edit_code), eps)``,
``loss.backward()`` through a FROZEN UNet (train.py:415) and AdamW over the adapter tensors (:486-492).  Here the
backward pass is an explicit tape over the same kernels as the forward: every dX of a linear / conv is
``anysd_gemm_f16`` on a transposed / 180-degree-rotated weight pack, the non-contraction ops have their own backward
kernels (``csrc/backward.cu``, ``csrc/attention_bwd.cu``), parameter gradients exist only for the trainables
(``to_k_ip`` / ``to_v_ip`` experts, router, task-embedding table) plus the visual tokens (for the projector upstream).
PyTorch holds memory and the stream; autograd is not used.
"""
import torch

from .unet import _pack_conv3


def pack_conv3_dx(w, dev, cin_pad=None):
    """OIHW conv weight -> the pack whose conv3x3 (pad 1, stride 1) maps dY [.., Cout] to dX [.., Cin]:
    Wb[ci, co, a, b] = w[co, ci, 2 - a, 2 - b]."""
    return _pack_conv3(w.detach().transpose(0, 1).flip(2, 3), dev, cin_pad)


def pack_linear_dx(w, dev):
    """[out, in] linear weight -> [in, out] fp16 so that gemm(dY, pack) = dY @ w."""
    return w.detach().to(dev).t().contiguous().to(torch.float16)
