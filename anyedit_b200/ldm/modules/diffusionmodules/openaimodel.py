"""Alias of ``ldm.modules.diffusionmodules.openaimodel`` for yaml ``target:`` strings."""
from anyedit_b200.unet import UNetModel  # noqa: F401
