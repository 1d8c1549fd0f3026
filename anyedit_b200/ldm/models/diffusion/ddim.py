"""Alias of ``ldm.models.diffusion.ddim``."""
from anyedit_b200.ddim import DDIMSampler, make_ddim_sampling_parameters, make_ddim_timesteps  # noqa: F401
