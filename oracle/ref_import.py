"""Import the real reference modules: from /root/reference in the build container, from the staged copy
``oracle/_ref/`` on the GPU box.

TEST INFRASTRUCTURE.  Used by ``tests/golden/make_golden*.py`` to generate the committed golden vectors, by
``tests/test_oracle_golden.py`` to pin the restatement, and by ``bench.py --impl reference`` / the ``cpu_baseline``
leg to time the reference's OWN ``UNetModel`` + ``DDIMSampler`` on the box's host cores.

``stage()`` (called by ``__graft_entry__.build()`` when /root/reference is present) copies the five source files of
the path -- and nothing else -- into ``oracle/_ref/ldm/...``.  ``oracle/_ref/`` is git-ignored (reference sources
never enter the history) but not gpurun-ignored, so it travels to the GPU box like the built ``.so`` does; when it
is absent the CPU legs fall back to the restatement (``cpu_baseline.kind = "port"``).

Work-arounds:
  * ``omegaconf`` is not installed and UNetModel.__init__ imports
    ``omegaconf.listconfig.ListConfig`` (openaimodel.py:477-481) -> stub module.
  * ``DDIMSampler.register_buffer`` forces CUDA (ddim.py:17-21) -> CPU subclass.
"""
import os
import sys
import types

SRC_ROOT = os.environ.get("ANYEDIT_REFERENCE", "/root/reference")
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# the files of the hot path; ldm/ is a namespace package (no top-level __init__.py)
STAGED_FILES = ("ldm/util.py", "ldm/modules/attention.py", "ldm/modules/diffusionmodules/__init__.py",
                "ldm/modules/diffusionmodules/util.py", "ldm/modules/diffusionmodules/openaimodel.py",
                "ldm/models/diffusion/__init__.py", "ldm/models/diffusion/ddim.py")


def _has(root) -> bool:
    return all(os.path.isfile(os.path.join(root, f)) for f in STAGED_FILES)


REF_ROOT = SRC_ROOT if os.path.isdir(os.path.join(SRC_ROOT, "ldm")) else STAGED_ROOT


def available() -> bool:
    return _has(REF_ROOT)


def stage() -> str:
    """Copy the path's reference sources into oracle/_ref/ (build container only; no-op elsewhere)."""
    import shutil
    if not os.path.isdir(os.path.join(SRC_ROOT, "ldm")):
        return STAGED_ROOT if _has(STAGED_ROOT) else ""
    for f in STAGED_FILES:
        dst = os.path.join(STAGED_ROOT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC_ROOT, f), dst)
    return STAGED_ROOT


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
