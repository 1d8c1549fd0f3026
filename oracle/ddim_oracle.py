"""CPU restatement of the DDIM schedule and sampler loop (TEST INFRASTRUCTURE).

Reference (relative to /root/reference):
  ldm/modules/diffusionmodules/util.py   make_beta_schedule :21-43,
      make_ddim_timesteps :46-60, make_ddim_sampling_parameters :63-74
  ldm/models/diffusion/ddpm.py           register_schedule :138-166, q_sample :356-359,
      DiffusionWrapper.forward :1332-1363
  ldm/models/diffusion/ddim.py           make_schedule :23-52, ddim_sampling :122-178,
      p_sample_ddim :181-251
"""
import numpy as np
import torch


def make_beta_schedule(schedule="linear", n_timestep=1000, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    # util.py:21-43 -- float64 throughout; 'linear' is linear in sqrt(beta).
    if schedule == "linear":
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "cosine":
        ts = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        alphas = torch.cos(ts / (1 + cosine_s) * np.pi / 2).pow(2)
        alphas = alphas / alphas[0]
        betas = 1 - alphas[1:] / alphas[:-1]
        betas = np.clip(betas, a_min=0, a_max=0.999)
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy()


def register_schedule(beta_schedule="linear", timesteps=1000, linear_start=1e-4, linear_end=2e-2):
    """ddpm.py:138-166: float64 numpy cumprod, buffers cast to fp32."""
    betas = make_beta_schedule(beta_schedule, timesteps, linear_start, linear_end)
    alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
    alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return {
        "betas": f32(betas),
        "alphas_cumprod": f32(alphas_cumprod),
        "alphas_cumprod_prev": f32(alphas_cumprod_prev),
        "sqrt_alphas_cumprod": f32(np.sqrt(alphas_cumprod)),
        "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - alphas_cumprod)),
        "alphas_cumprod_f64": alphas_cumprod,
    }


def make_ddim_timesteps(method, num_ddim, num_ddpm):
    # util.py:46-60 -- integer arithmetic, +1 shift.
    if method == "uniform":
        c = num_ddpm // num_ddim
        ts = np.asarray(list(range(0, num_ddpm, c)))
    elif method == "quad":
        ts = ((np.linspace(0, np.sqrt(num_ddpm * 0.8), num_ddim)) ** 2).astype(int)
    else:
        raise NotImplementedError(method)
    return ts + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta):
    # util.py:63-74.  alphacums is the fp32 torch buffer moved to cpu (ddim.py:42);
    # indexing it with a numpy int array yields an fp32 tensor for `alphas`, while
    # alphas_prev goes through .tolist() and becomes float64 numpy.
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def q_sample(sched, x_start, t, noise):
    # ddpm.py:356-359
    a = sched["sqrt_alphas_cumprod"][t].reshape(-1, 1, 1, 1)
    b = sched["sqrt_one_minus_alphas_cumprod"][t].reshape(-1, 1, 1, 1)
    return a * x_start + b * noise


def apply_model(unet_fn, conditioning_key, x, t, cond):
    """LatentDiffusion.apply_model + DiffusionWrapper.forward (ddpm.py:854-869, 1332-1363).
    ``unet_fn(x, t, context=None, y=None)``."""
    if not isinstance(cond, dict):
        if not isinstance(cond, list):
            cond = [cond]
        cond = {("c_concat" if conditioning_key == "concat" else "c_crossattn"): cond}
    cc_list, ca_list, adm = cond.get("c_concat"), cond.get("c_crossattn"), cond.get("c_adm")
    if conditioning_key is None:
        return unet_fn(x, t)
    if conditioning_key == "concat":
        return unet_fn(torch.cat([x] + cc_list, dim=1), t)
    if conditioning_key == "crossattn":
        return unet_fn(x, t, context=torch.cat(ca_list, 1))
    if conditioning_key == "hybrid":
        return unet_fn(torch.cat([x] + cc_list, dim=1), t, context=torch.cat(ca_list, 1))
    if conditioning_key == "hybrid-adm":
        assert adm is not None
        return unet_fn(torch.cat([x] + cc_list, dim=1), t, context=torch.cat(ca_list, 1), y=adm)
    if conditioning_key == "crossattn-adm":
        assert adm is not None
        return unet_fn(x, t, context=torch.cat(ca_list, 1), y=adm)
    if conditioning_key == "adm":
        return unet_fn(x, t, y=ca_list[0])
    raise NotImplementedError(conditioning_key)


def _cat_cond(uc, c):
    # ddim.py:194-210 -- [uncond ; cond]
    if isinstance(c, dict):
        out = {}
        for k in c:
            if isinstance(c[k], list):
                out[k] = [torch.cat([uc[k][i], c[k][i]]) for i in range(len(c[k]))]
            else:
                out[k] = torch.cat([uc[k], c[k]])
        return out
    if isinstance(c, list):
        return [torch.cat([uc[i], c[i]]) for i in range(len(c))]
    return torch.cat([uc, c])


def ddim_sample(model_fn, sched, S, x_T, cond, uncond=None, scale=1.0, eta=0.0, ddpm_steps=1000,
                mask=None, x0=None, log_every_t=100, generator=None, img_cond=None, img_scale=None):
    """ddim.py:23-52 + 122-178 + 181-251, eps-parameterisation.
    ``model_fn(x, t, cond) -> eps`` (the apply_model boundary).  Returns (x0_latent, intermediates)."""
    ts = make_ddim_timesteps("uniform", S, ddpm_steps)
    sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(sched["alphas_cumprod"].cpu(), ts, eta)
    sqrt_one_minus_alphas = np.sqrt(1.0 - alphas)            # ddim.py:48 (fp32 tensor)
    img = x_T
    b = img.shape[0]
    inter = {"x_inter": [img], "pred_x0": [img]}
    time_range = np.flip(ts)
    total = ts.shape[0]
    for i, step in enumerate(time_range):
        index = total - i - 1
        t = torch.full((b,), int(step), dtype=torch.long)
        if mask is not None:
            noise = torch.randn(x0.shape, generator=generator)
            img = q_sample(sched, x0, t, noise) * mask + (1.0 - mask) * img
        if img_cond is not None:
            # InstructPix2Pix three-way guidance, AnyEdit_Collection/adaptive_editing_pipelines/tools/global_tool.py:160-177:
            # batch [text ; image ; uncond], e = e_unc + s_txt (e_txt - e_img) + s_img (e_img - e_unc)
            out = model_fn(torch.cat([img] * 3), torch.cat([t] * 3), _cat_cond(_cat_cond(cond, img_cond), uncond))
            e_txt, e_img, e_unc = out.chunk(3)
            e_t = e_unc + scale * (e_txt - e_img) + img_scale * (e_img - e_unc)
        elif uncond is None or scale == 1.0:
            e_t = model_fn(img, t, cond)
        else:
            out = model_fn(torch.cat([img] * 2), torch.cat([t] * 2), _cat_cond(uncond, cond))
            e_u, e_c = out.chunk(2)
            e_t = e_u + scale * (e_c - e_u)
        # ddim.py:228-231: python scalars -> fp32 tensors through torch.full
        a_t = torch.full((b, 1, 1, 1), float(alphas[index]))
        a_prev = torch.full((b, 1, 1, 1), float(alphas_prev[index]))
        sigma_t = torch.full((b, 1, 1, 1), float(sigmas[index]))
        somat = torch.full((b, 1, 1, 1), float(sqrt_one_minus_alphas[index]))
        pred_x0 = (img - somat * e_t) / a_t.sqrt()
        dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
        noise = sigma_t * torch.randn(img.shape, generator=generator)
        img = a_prev.sqrt() * pred_x0 + dir_xt + noise
        if index % log_every_t == 0 or index == total - 1:
            inter["x_inter"].append(img)
            inter["pred_x0"].append(pred_x0)
    return img, inter
