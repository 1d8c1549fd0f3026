"""``PLMSSampler`` -- same API as ``ldm.models.diffusion.plms.PLMSSampler`` (plms.py:12-244), on the DDIM stepper.

Pseudo linear multistep sampling (Liu et al. 2022) as the reference implements it: the DDIM tables with eta = 0
(make_schedule, plms.py:25-57), a pseudo improved Euler first step (two model calls, :226-228) and Adams-Bashforth
2 / 3 / 4 combinations of the last raw eps afterwards (:229-238), the DDIM update applied to the combined eps (:205-225).
Execution: every model call is the same captured CUDA graph as in ``DDIMSampler`` (conditioning mux + UNet on the
CFG-doubled batch) followed by ONE fused kernel (``anysd_cfg_plms_step_f32``: guidance combine, multistep combination,
update, history push) whose ten coefficients are refreshed by a tiny async copy -- S + 1 UNet evaluations for S steps.
CFG batching accepts the dict / list conditionings of ``DDIMSampler`` as well as the tensor-only form the reference
PLMS supports (:189-191).
"""
import numpy as np
import torch

from .ddim import DDIMSampler, step_coefficients

# (c0, c1, c2, c3, den): e' = (((c0 e - c1 o1) + c2 o2) - c3 o3) / den
_EULER_A = (1.0, 0.0, 0.0, 0.0, 1.0)          # e' = e_t                       (x_prev for the second model call)
_EULER_B = (1.0, -1.0, 0.0, 0.0, 2.0)         # e' = (e_t + e_next) / 2        (o1 holds e_t)
_AB = {1: (3.0, 1.0, 0.0, 0.0, 2.0), 2: (23.0, 16.0, 5.0, 0.0, 12.0), 3: (55.0, 59.0, 37.0, 9.0, 24.0)}


def plms_plan(time_range):
    """The model calls of one PLMS run (plms.py:154-172, 226-238), in order: dicts with the DDIM table ``index`` of the
    update, the timestep ``t`` the model is evaluated at, the combination ``comb`` = (c0, c1, c2, c3, den), whether the
    raw eps is pushed into the history, whether the model sees the provisional latent of the previous call
    (``use_tmp``: second half of the pseudo improved Euler step) and whether the call completes a step (``final``)."""
    total = len(time_range)
    plan = []
    for i, step in enumerate(time_range):
        index = total - i - 1
        if i == 0:
            t_next = int(time_range[min(i + 1, total - 1)])
            plan.append(dict(i=i, index=index, t=int(step), comb=_EULER_A, push=True, use_tmp=False, final=False))
            plan.append(dict(i=i, index=index, t=t_next, comb=_EULER_B, push=False, use_tmp=True, final=True))
        else:
            plan.append(dict(i=i, index=index, t=int(step), comb=_AB[min(i, 3)], push=True, use_tmp=False, final=True))
    return plan


class PLMSSampler(DDIMSampler):
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")
        super().make_schedule(ddim_num_steps, ddim_discretize=ddim_discretize, ddim_eta=0., verbose=verbose)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, dynamic_threshold=None, **kwargs):
        if conditioning is not None:
            ctmp = conditioning
            if isinstance(ctmp, dict):
                ctmp = ctmp[list(ctmp.keys())[0]]
            while isinstance(ctmp, list):
                ctmp = ctmp[0]
            if ctmp.shape[0] != batch_size:
                print(f"Warning: Got {ctmp.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        if verbose:
            print(f"Data shape for PLMS sampling is {size}")
        return self.plms_sampling(conditioning, size, callback=callback, img_callback=img_callback, quantize_denoised=quantize_x0,
                                  mask=mask, x0=x0, ddim_use_original_steps=False, noise_dropout=noise_dropout,
                                  temperature=temperature, score_corrector=score_corrector, corrector_kwargs=corrector_kwargs,
                                  x_T=x_T, log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, dynamic_threshold=dynamic_threshold)

    @torch.no_grad()
    def plms_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1.,
                      noise_dropout=0., score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, dynamic_threshold=None):
        """plms.py:118-176."""
        if ddim_use_original_steps or quantize_denoised or score_corrector is not None or dynamic_threshold is not None:
            raise NotImplementedError("option not on the AnySD path")
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device=device, dtype=torch.float32)
        img = img.contiguous().clone()
        if timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        intermediates = {"x_inter": [img.clone()], "pred_x0": [img.clone()]}
        scale = unconditional_guidance_scale
        use_cfg = not (unconditional_conditioning is None or scale == 1.)
        stepper = self._get_stepper(cond, unconditional_conditioning, use_cfg, b, tuple(shape), device, graph=self.use_cuda_graph,
                                    update="plms")
        stepper.reset()
        # coefficient rows on the device: [4 DDIM scalars (sigma = 0), c0..c3, den, push]
        ddim = [step_coefficients(self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sigmas, self.ddim_sqrt_one_minus_alphas, i)[:4]
                for i in range(len(self.ddim_timesteps))]
        row = lambda index, comb, push: torch.tensor(ddim[index] + list(comb) + [float(push)], dtype=torch.float32)
        x_tmp = None
        for call in plms_plan(time_range):
            i, index = call["i"], call["index"]
            if mask is not None and not call["use_tmp"]:
                assert x0 is not None
                ts = torch.full((b,), call["t"], device=device, dtype=torch.long)
                img_orig = self.model.q_sample(x0, ts)
                img = img_orig * mask + (1. - mask) * img
            out, pred_x0 = stepper.step(img, index, call["t"], scale, None, coef=row(index, call["comb"], call["push"]).to(device),
                                        x_model=x_tmp if call["use_tmp"] else None)
            if not call["final"]:
                x_tmp = out                              # pseudo improved Euler: provisional x_prev for the second model call
                continue
            img, x_tmp = out, None
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img.clone())
                intermediates["pred_x0"].append(pred_x0.clone())
        return img, intermediates
