#!/usr/bin/env python
"""Diagnostic (not a pytest file): GroupNorm(+SiLU) / LayerNorm timing on the bench shapes (ANYSD_GN_FUSED=0|1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from anyedit_b200 import ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    N = 16
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # L2 flush between calls
    for HW, C in ((4096, 320), (1024, 640), (256, 1280), (64, 1280), (4096, 960), (4096, 640), (1024, 1280), (1024, 1920), (256, 2560), (64, 2560)):
        x = torch.randn(N, HW, C, device="cuda").half()
        y = torch.empty_like(x)
        g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
        ws = ops.groupnorm_workspace(N, 32, C)
        fn = lambda: ops.groupnorm(x, g, b, y, N, HW, 1e-5, True, ws)
        t = timeit(fn)
        fn2 = lambda: (big.zero_(), ops.groupnorm(x, g, b, y, N, HW, 1e-5, True, ws))
        t2 = timeit(fn2) - timeit(lambda: big.zero_())
        mb = x.numel() * 2 / 1e6
        res = ops._lib.load().anysd_groupnorm_resident(C, 0, HW, 32)
        print(f"groupnorm fused={os.environ.get('ANYSD_GN_FUSED', '1')} resident={res} N={N} HW={HW} C={C}: warm {t * 1e6:7.1f} us  cold {t2 * 1e6:7.1f} us "
              f"({2 * mb / t2 / 1e6:5.2f} TB/s algorithmic, {mb:.0f} MB tensor)", flush=True)
        # statistics from a producer's epilogue (here: an identity-sized 1x1 contraction), then finalize + streaming apply
        if C <= 1280:
            w = (torch.eye(C, device="cuda") * 1.0).half()
            st = ops.gemm(x.view(N * HW, C), w, y.view(N * HW, C), rows_per_batch=HW, stats_images=N)
            if st is not None:
                fa = lambda: ops.groupnorm(x, g, b, y, N, HW, 1e-5, True, ws, stats=st)
                ta = timeit(fa)
                ta2 = timeit(lambda: (big.zero_(), fa())) - timeit(lambda: big.zero_())
                print(f"   epilogue-stats apply: warm {ta * 1e6:7.1f} us  cold {ta2 * 1e6:7.1f} us ({2 * mb / ta2 / 1e6:5.2f} TB/s algorithmic)", flush=True)
    for M, C in ((65536, 320), (16384, 640), (4096, 1280)):
        x = torch.randn(M, C, device="cuda").half()
        y = torch.empty_like(x)
        g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
        t = timeit(lambda: ops.layernorm(x, g, b, y))
        print(f"layernorm M={M} C={C}: warm {t * 1e6:7.1f} us ({2 * x.numel() * 2 / t / 1e12:5.2f} TB/s)", flush=True)


if __name__ == "__main__":
    main()
