#!/usr/bin/env python
"""Golden vectors for the ControlNet branch, produced by the REFERENCE's own ``ControlNet`` and
``ControlledUnetModel`` (AnyEdit_Collection/other_modules/cldm/cldm.py:21-304) in the build container.  cldm.py imports
``ldm.models.diffusion.ddpm`` (pytorch_lightning, omegaconf): both are stubbed just enough for the import -- the two classes
used here do not touch them.  Usage: python tests/golden/make_golden_cldm.py"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)

from oracle import ref_import, weights  # noqa: E402
from make_golden import randn  # noqa: E402

CN = dict(image_size=16, in_channels=4, model_channels=64, hint_channels=3, num_res_blocks=1, attention_resolutions=[1, 2, 4],
          channel_mult=[1, 2, 4], num_heads=4, use_spatial_transformer=True, transformer_depth=1, context_dim=64, legacy=False)
UN = {**{k: v for k, v in CN.items() if k != "hint_channels"}, "out_channels": 4}


def _stubs():
    ref_import._stub_omegaconf()
    sys.modules["omegaconf"].ListConfig = sys.modules["omegaconf.listconfig"].ListConfig
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = torch.nn.Module
    plu = types.ModuleType("pytorch_lightning.utilities")
    plud = types.ModuleType("pytorch_lightning.utilities.distributed")
    plud.rank_zero_only = lambda f: f
    plr = types.ModuleType("pytorch_lightning.utilities.rank_zero")
    plr.rank_zero_only = lambda f: f
    sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.utilities": plu, "pytorch_lightning.utilities.distributed": plud,
                        "pytorch_lightning.utilities.rank_zero": plr})


def main():
    _stubs()
    sys.path.insert(0, ref_import.SRC_ROOT)
    sys.path.insert(0, os.path.join(ref_import.SRC_ROOT, "AnyEdit_Collection", "other_modules"))
    from cldm.cldm import ControlledUnetModel, ControlNet
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    cn, un = ControlNet(**CN).eval(), ControlledUnetModel(**UN).eval()
    cshapes = {k: tuple(v.shape) for k, v in cn.state_dict().items()}
    ushapes = {k: tuple(v.shape) for k, v in un.state_dict().items()}
    csd, usd = weights.make_state_dict(cshapes, 41), weights.make_state_dict(ushapes, 42)
    cn.load_state_dict(csd)
    un.load_state_dict(usd)
    x, hint, ctx = randn(61, 2, 4, 16, 16), randn(62, 2, 3, 128, 128), randn(63, 2, 7, 64)
    t = torch.tensor([981, 21], dtype=torch.long)
    control = cn(x=x, hint=hint, timesteps=t, context=ctx)
    out = {"x": x.numpy(), "hint": hint.numpy(), "ctx": ctx.numpy(), "t": t.numpy(), "cseed": 41, "useed": 42,
           "cwsum": weights.checksum(csd), "uwsum": weights.checksum(usd)}
    for i, c in enumerate(control):
        out[f"control_{i}"] = c.numpy()
    scales = [0.5 + 0.1 * i for i in range(len(control))]
    for only_mid in (False, True):
        eps = un(x=x, timesteps=t, context=ctx, control=[c * s for c, s in zip(control, scales)], only_mid_control=only_mid)
        out[f"eps_only_mid{int(only_mid)}"] = eps.numpy()
    out["scales"] = np.asarray(scales)
    with open(os.path.join(HERE, "cldm_tiny_keys.json"), "w") as f:
        json.dump({"control_config": CN, "unet_config": UN, "control_keys": {k: list(v) for k, v in cshapes.items()},
                   "unet_keys": {k: list(v) for k, v in ushapes.items()}}, f)
    np.savez(os.path.join(HERE, "cldm_tiny.npz"), **out)
    print(len(control), [tuple(c.shape) for c in control], float(out["eps_only_mid0"].std()))


if __name__ == "__main__":
    main()
