#!/usr/bin/env python
"""Diagnostic (not a pytest file): the register-resident GroupNorm kernel on the two small bench maps, for an ncu capture
(``ncu --set full -k regex:gn_res python tests/diag_gn_small.py``)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from anyedit_b200 import ops  # noqa: E402

N = 16
for HW, C in ((256, 1280), (64, 1280), (256, 2560)):
    x = torch.randn(N, HW, C, device="cuda").half()
    y = torch.empty_like(x)
    g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    ws = ops.groupnorm_workspace(N, 32, C)
    for _ in range(3):
        ops.groupnorm(x, g, b, y, N, HW, 1e-5, True, ws)
    torch.cuda.synchronize()
print("done")
