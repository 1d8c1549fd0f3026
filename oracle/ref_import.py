"""Import the real reference modules from /root/reference (BUILD CONTAINER ONLY).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so nothing at run time
may depend on this module; it is used by ``tests/golden/make_golden.py`` to generate the
committed golden vectors and by ``tests/test_oracle_vs_reference.py`` (skipped when the
reference tree is absent) to pin the restatement.

Work-arounds:
  * ``omegaconf`` is not installed and UNetModel.__init__ imports
    ``omegaconf.listconfig.ListConfig`` (openaimodel.py:477-481) -> stub module.
  * ``DDIMSampler.register_buffer`` forces CUDA (ddim.py:17-21) -> CPU subclass.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("ANYEDIT_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ldm"))


def _stub_omegaconf():
    if "omegaconf" in sys.modules:
        return
    m = types.ModuleType("omegaconf")
    lc = types.ModuleType("omegaconf.listconfig")

    class ListConfig(list):
        pass

    lc.ListConfig = ListConfig
    m.listconfig = lc
    sys.modules["omegaconf"] = m
    sys.modules["omegaconf.listconfig"] = lc


def load():
    """Returns (UNetModel, DDIMSamplerCPU, util_module) from the reference tree."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _stub_omegaconf()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from ldm.models.diffusion.ddim import DDIMSampler
    from ldm.modules.diffusionmodules import util

    class DDIMSamplerCPU(DDIMSampler):
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    return UNetModel, DDIMSamplerCPU, util


class RefModelShim:
    """Duck-typed ``model`` for the reference DDIMSampler: the attributes ddim.py reads
    (ddim.py:13-14, 27-34, 188-211) on top of a reference UNetModel, with the ``hybrid``
    conditioning mux of DiffusionWrapper.forward (ddpm.py:1344-1347)."""

    def __init__(self, unet, sched, conditioning_key="hybrid"):
        import torch
        self.unet = unet
        self.num_timesteps = 1000
        self.betas = sched["betas"]
        self.alphas_cumprod = sched["alphas_cumprod"]
        self.alphas_cumprod_prev = sched["alphas_cumprod_prev"]
        self.sqrt_alphas_cumprod = sched["sqrt_alphas_cumprod"]
        self.sqrt_one_minus_alphas_cumprod = sched["sqrt_one_minus_alphas_cumprod"]
        self.device = torch.device("cpu")
        self.parameterization = "eps"
        self.conditioning_key = conditioning_key

    def apply_model(self, x, t, cond):
        import torch
        if self.conditioning_key == "hybrid":
            xc = torch.cat([x] + cond["c_concat"], dim=1)
            cc = torch.cat(cond["c_crossattn"], 1)
            return self.unet(xc, t, context=cc)
        if self.conditioning_key == "crossattn":
            cc = torch.cat(cond["c_crossattn"], 1) if isinstance(cond, dict) else cond
            return self.unet(x, t, context=cc)
        raise NotImplementedError(self.conditioning_key)
