#!/usr/bin/env python
"""Diagnostic: short-K dense GEMMs of the level-0 transformer (for ncu --set full)."""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from anyedit_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
M = 65536
a = rn(M, 320).half()
w1, b1 = (rn(320, 320) * 320 ** -0.5).half(), rn(320) * 0.1
res = rn(M, 320).half()
o1 = torch.empty(M, 320, dtype=torch.float16, device="cuda")
wq = (rn(1152, 320) * 320 ** -0.5).half()
o2 = torch.empty(M, 1152, dtype=torch.float16, device="cuda")
for _ in range(2):
    ops.gemm(a, w1, o1, bias=b1, residual=res)      # to_out + residual
for _ in range(2):
    ops.gemm(a, wq, o2)                              # fused qkv (padded heads)
torch.cuda.synchronize()
print("done")
