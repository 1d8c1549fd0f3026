"""``DPMSolverSampler`` -- same API as ``ldm.models.diffusion.dpm_solver.sampler.DPMSolverSampler`` (sampler.py:14-87):
DPM-Solver++(2M) exactly as the reference configures it,

    NoiseScheduleVP('discrete', alphas_cumprod);  model_wrapper(..., guidance_type="classifier-free")
    DPM_Solver(model_fn, ns, predict_x0=True, thresholding=False).sample(x_T, steps=S, skip_type="time_uniform",
                                                                     method="multistep", order=2, lower_order_final=True)

(dpm_solver.py: discrete schedule :79-88 + interpolate_fn :1104-1143, float model time :246-255, classifier-free guidance
:303-311, data prediction :352-365, first-order update :469-513, second-order multistep update :723-778, driver :1044-1073).
The host evaluates the noise schedule at the S + 1 grid points in fp32 (a few dozen scalars); every step is the captured
CUDA graph of the DDIM stepper (conditioning mux + UNet on the CFG-doubled batch, float timesteps) followed by ONE fused
kernel (``anysd_cfg_dpmpp_step_f32``: guidance combine, data prediction, multistep update, history).  S UNet evaluations
for S steps: 20 such steps stand in for 50 DDIM steps.
"""
import torch

from .ddim import DDIMSampler


class NoiseScheduleVP:
    """The 'discrete' schedule of dpm_solver.py:17-160: log alpha_n = 0.5 log(acp_n) at t_n = (n + 1) / N, piecewise
    linear in t (linear extrapolation with the outermost segment), fp32."""

    def __init__(self, schedule="discrete", betas=None, alphas_cumprod=None, **kwargs):
        if schedule != "discrete":
            raise ValueError("only the 'discrete' schedule (what DPMSolverSampler uses) is implemented")
        if betas is not None:
            log_alphas = 0.5 * torch.log(1 - betas.detach().float().cpu()).cumsum(dim=0)
        else:
            assert alphas_cumprod is not None
            log_alphas = 0.5 * torch.log(alphas_cumprod.detach().float().cpu())
        self.schedule = schedule
        self.total_N = log_alphas.numel()
        self.T = 1.0
        self.t_array = torch.linspace(0.0, 1.0, self.total_N + 1)[1:]
        self.log_alpha_array = log_alphas

    def marginal_log_mean_coeff(self, t):
        t = torch.as_tensor(t, dtype=torch.float32).reshape(-1)
        idx = torch.searchsorted(self.t_array, t.contiguous(), right=False).clamp(1, self.total_N - 1)
        x0, x1 = self.t_array[idx - 1], self.t_array[idx]
        y0, y1 = self.log_alpha_array[idx - 1], self.log_alpha_array[idx]
        return y0 + (t - x0) * (y1 - y0) / (x1 - x0)

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        lm = self.marginal_log_mean_coeff(t)
        return lm - 0.5 * torch.log(1.0 - torch.exp(2.0 * lm))


def dpmpp_2m_tables(ns, S):
    """Per model call k = 0 .. S-1 (time s = ts[k], target t = ts[k+1]): the float model timestep and the six
    coefficients of ``anysd_cfg_dpmpp_step_f32``.  time_uniform grid from T = 1 to 1/N (dpm_solver.py:1046-1049); the
    first step is first-order, the last one too when S < 15 (lower_order_final, :1064-1067)."""
    ts = torch.linspace(ns.T, 1.0 / ns.total_N, S + 1)
    f = lambda v: torch.full((1,), float(v), dtype=torch.float32)
    t_model, rows = [], []
    for k in range(S):
        s, t = f(ts[k]), f(ts[k + 1])
        t_model.append(float((s - 1.0 / ns.total_N) * 1000.0))                   # :246-255
        sig_s, alp_s = ns.marginal_std(s), ns.marginal_alpha(s)
        lam_s, lam_t = ns.marginal_lambda(s), ns.marginal_lambda(t)
        ratio = ns.marginal_std(t) / sig_s
        alp_t = ns.marginal_alpha(t)
        h = lam_t - lam_s
        first_order = k == 0 or (S < 15 and k == S - 1)
        if first_order:
            c, half_c, inv_r0 = alp_t * torch.expm1(-h), torch.zeros(1), torch.zeros(1)
        else:
            lam_p = ns.marginal_lambda(f(ts[k - 1]))
            r0 = (lam_s - lam_p) / h
            c = alp_t * (torch.exp(-h) - 1.0)
            half_c, inv_r0 = 0.5 * c, 1.0 / r0
        rows.append([float(v) for v in (sig_s, alp_s, ratio, c, half_c, inv_r0)])
    return t_model, torch.tensor(rows, dtype=torch.float32)


class DPMSolverSampler(DDIMSampler):
    def __init__(self, model, **kwargs):
        super().__init__(model, **kwargs)
        self.register_buffer("alphas_cumprod", model.alphas_cumprod.clone().detach().to(torch.float32))

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        """sampler.py:27-87.  Returns ``(x, None)`` like the reference."""
        if conditioning is not None:
            ctmp = conditioning
            if isinstance(ctmp, dict):
                ctmp = ctmp[list(ctmp.keys())[0]]
            while isinstance(ctmp, list):
                ctmp = ctmp[0]
            if ctmp.shape[0] != batch_size:
                print(f"Warning: Got {ctmp.shape[0]} conditionings but batch-size is {batch_size}")
        if getattr(self.model, "parameterization", "eps") != "eps":
            raise NotImplementedError("DPMSolverSampler: only the eps parameterisation (model_type 'noise') is implemented")
        C, H, W = shape
        size = (batch_size, C, H, W)
        if verbose:
            print(f"Data shape for DPM-Solver sampling is {size}, sampling steps {S}")
        device = self.model.betas.device
        img = torch.randn(size, device=device) if x_T is None else x_T.to(device=device, dtype=torch.float32)
        img = img.contiguous().clone()
        key = ("dpmpp", S)
        if getattr(self, "_tab_key", None) != key:
            ns = NoiseScheduleVP("discrete", alphas_cumprod=self.alphas_cumprod)
            self._t_model, coef = dpmpp_2m_tables(ns, S)
            self._coef_dev, self._tab_key = coef.to(device), key
        scale = unconditional_guidance_scale
        # model_wrapper (dpm_solver.py:303-311): guidance_scale == 1 or no unconditional condition -> conditional branch only
        use_cfg = not (unconditional_conditioning is None or scale == 1.)
        stepper = self._get_stepper(conditioning, unconditional_conditioning, use_cfg, batch_size, size, device,
                                    graph=self.use_cuda_graph, update="dpmpp")
        stepper.reset()
        for k in range(S):
            img, _ = stepper.step(img, k, self._t_model[k], scale, None, coef=self._coef_dev[k])
            if callback:
                callback(k)
        return img, None
