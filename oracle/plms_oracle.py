"""CPU restatement of the PLMS sampler (sibling sampler on the same apply_model boundary).
TEST INFRASTRUCTURE (see oracle/__init__.py): groundwork for the next scope row, no product code uses it yet.

Reference: ldm/models/diffusion/plms.py -- make_schedule :25-57 (the DDIM tables, eta must be 0), plms_sampling
:118-176 (old_eps keeps the last 3 raw eps), p_sample_plms :178-244:
    first step   pseudo improved Euler: e' = (e_t + eps(x_prev(e_t), t_next)) / 2           (two model calls)
    then         Adams-Bashforth 2 / 3 / 4:  (3e - e1)/2, (23e - 16e1 + 5e2)/12, (55e - 59e1 + 37e2 - 9e3)/24
    x_prev, pred_x0 from e' with the DDIM update (:205-225)
CFG in the reference concatenates TENSOR conditionings only (:189-191).
"""
import numpy as np
import torch

from .ddim_oracle import make_ddim_sampling_parameters, make_ddim_timesteps


def plms_sample(model_fn, sched, S, x_T, cond, uncond=None, scale=1.0, ddpm_steps=1000, log_every_t=100):
    ts = make_ddim_timesteps("uniform", S, ddpm_steps)
    sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(sched["alphas_cumprod"].cpu(), ts, 0.0)
    sqrt_one_minus_alphas = np.sqrt(1.0 - alphas)
    b = x_T.shape[0]

    def eps(x, t):
        if uncond is None or scale == 1.0:
            return model_fn(x, t, cond)
        e_u, e_c = model_fn(torch.cat([x] * 2), torch.cat([t] * 2), torch.cat([uncond, cond])).chunk(2)
        return e_u + scale * (e_c - e_u)

    def update(x, e, index):
        a_t = torch.full((b, 1, 1, 1), float(alphas[index]))
        a_prev = torch.full((b, 1, 1, 1), float(alphas_prev[index]))
        sigma_t = torch.full((b, 1, 1, 1), float(sigmas[index]))
        somat = torch.full((b, 1, 1, 1), float(sqrt_one_minus_alphas[index]))
        pred_x0 = (x - somat * e) / a_t.sqrt()
        x_prev = a_prev.sqrt() * pred_x0 + (1.0 - a_prev - sigma_t ** 2).sqrt() * e      # sigma = 0: no noise term
        return x_prev, pred_x0

    img = x_T
    inter = {"x_inter": [img], "pred_x0": [img]}
    time_range = np.flip(ts)
    total = ts.shape[0]
    old = []
    for i, step in enumerate(time_range):
        index = total - i - 1
        t = torch.full((b,), int(step), dtype=torch.long)
        t_next = torch.full((b,), int(time_range[min(i + 1, len(time_range) - 1)]), dtype=torch.long)
        e_t = eps(img, t)
        if len(old) == 0:
            x_prev, _ = update(img, e_t, index)
            e_prime = (e_t + eps(x_prev, t_next)) / 2
        elif len(old) == 1:
            e_prime = (3 * e_t - old[-1]) / 2
        elif len(old) == 2:
            e_prime = (23 * e_t - 16 * old[-1] + 5 * old[-2]) / 12
        else:
            e_prime = (55 * e_t - 59 * old[-1] + 37 * old[-2] - 9 * old[-3]) / 24
        img, pred_x0 = update(img, e_prime, index)
        old.append(e_t)
        if len(old) >= 4:
            old.pop(0)
        if index % log_every_t == 0 or index == total - 1:
            inter["x_inter"].append(img)
            inter["pred_x0"].append(pred_x0)
    return img, inter
