"""Alias of ``ldm.util.instantiate_from_config`` (ldm/util.py:74-89)."""
from anyedit_b200.diffusion import instantiate_from_config  # noqa: F401
