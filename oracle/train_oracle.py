"""CPU restatement of the AnySD training step (TEST INFRASTRUCTURE, see oracle/__init__.py).

Reference (relative to /root/reference), train.py:
  :631-641  noise ~ N(0, I); t ~ U{0..T-1}; noisy = scheduler.add_noise(latents, noise, t)
            (diffusers DDPMScheduler.add_noise == ldm q_sample, ddpm.py:356-359:
             sqrt(acp[t]) * x0 + sqrt(1 - acp[t]) * noise, scaled-linear betas 0.00085 -> 0.012)
  :651-669  conditioning dropout: one uniform draw p per sample;
            text := null text          where p < 2*P
            image latent := 0          where P <= p < 3*P      (image_mask = 1 - [p >= P][p < 3P])
  :672      x8 = cat([noisy, original_image_latent], dim=1)
  :675-676  target = noise  (epsilon prediction)
  :694-696  pred = MoE(x8, t, text, ref_embeds, edit_code);  loss = mse_loss(pred.float(), target.float(), "mean")
  :703-709  backward; AdamW over image_proj_model + adapter_modules + task_embs (:486-492); UNet frozen (:415)
Gradients come from torch autograd over the functional oracle forward (oracle/unet_oracle.py, anysd_oracle.py);
tests/golden/make_golden_train.py pins autograd-through-the-oracle against autograd through the imported reference
UNet.  The router / expert / task-embedding part stays PARITY UNPINNED (source absent, see anysd_oracle.py).
"""
import torch
import torch.nn.functional as F

from . import anysd_oracle


def q_sample(x0, noise, t, alphas_cumprod):
    """ddpm.py:356-359 / DDPMScheduler.add_noise: fp32 tables indexed by t."""
    acp = alphas_cumprod.to(torch.float32)
    a = acp[t].sqrt().reshape(-1, 1, 1, 1)
    b = (1.0 - acp[t]).sqrt().reshape(-1, 1, 1, 1)
    return a * x0 + b * noise


def conditioning_dropout(text, null_text, image_latent, random_p, prob):
    """train.py:651-669.  random_p: [B] uniform draws; prob: conditioning_dropout_prob."""
    b = text.shape[0]
    prompt_mask = (random_p < 2 * prob).reshape(b, 1, 1)
    text = torch.where(prompt_mask, null_text.expand_as(text), text)
    image_mask = 1 - ((random_p >= prob).to(image_latent.dtype) * (random_p < 3 * prob).to(image_latent.dtype))
    return text, image_mask.reshape(b, 1, 1, 1) * image_latent


def train_forward_loss(sd, adapter_sd, latents, noise, t, image_latent, text, edit_code, visual_tokens, alphas_cumprod,
                       **unet_kw):
    noisy = q_sample(latents, noise, t, alphas_cumprod)
    x8 = torch.cat([noisy, image_latent], dim=1)
    pred = anysd_oracle.anysd_forward(sd, adapter_sd, x8, t, text, edit_code, visual_tokens, **unet_kw)
    return F.mse_loss(pred.float(), noise.float(), reduction="mean"), pred


def train_step_grads(sd, adapter_sd, latents, noise, t, image_latent, text, edit_code, visual_tokens, alphas_cumprod,
                     **unet_kw):
    """loss and d loss / d {every adapter tensor, visual tokens}; the UNet (sd) is frozen (train.py:415)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in adapter_sd.items()}
    vis = None
    if visual_tokens is not None and visual_tokens.shape[1] > 0:
        vis = visual_tokens.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss, pred = train_forward_loss(sd, leaves, latents, noise, t, image_latent, text, edit_code, vis,
                                        alphas_cumprod, **unet_kw)
        loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    if vis is not None:
        grads["visual_tokens"] = vis.grad
    return loss.detach(), pred.detach(), grads


def adamw_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2):
    """torch.optim.AdamW (decoupled weight decay), single tensor, step counted from 1; returns the new tensors."""
    param = param * (1.0 - lr * weight_decay)
    exp_avg = beta1 * exp_avg + (1.0 - beta1) * grad
    exp_avg_sq = beta2 * exp_avg_sq + (1.0 - beta2) * grad * grad
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = exp_avg_sq.sqrt() / (bc2 ** 0.5) + eps
    param = param - (lr / bc1) * exp_avg / denom
    return param, exp_avg, exp_avg_sq
