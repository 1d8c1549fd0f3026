#!/usr/bin/env python
"""Diagnostic (not a pytest file): the LayerNorm fold (row moments from the producer's epilogue, gamma folded into the consumer)
against the separate layernorm kernel, per launch, on the SD-1.5 transformer shapes at B_eff = 16.
usage: python tests/diag_lnfold.py"""
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
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    g = torch.Generator(device="cuda").manual_seed(1)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    tot_plain = tot_fold = 0.0
    for M, C, Cp, reps in ((65536, 320, 384, 5), (16384, 640, 640, 5), (4096, 1280, 1280, 5), (1024, 1280, 1280, 1)):
        a = rn(M, C).half()
        x = torch.empty(M, C, dtype=torch.float16, device="cuda")
        res = rn(M, C).half()
        wo = (rn(C, C) * C ** -0.5).half()
        bo = rn(C)
        rs = ops.row_stats_buffer(M, C, "cuda")
        gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        ln = torch.empty_like(x)
        t_o = timeit(lambda: ops.gemm(a, wo, x, bias=bo, residual=res))
        t_os = timeit(lambda: ops.gemm(a, wo, x, bias=bo, residual=res, row_stats=rs))
        t_ln = timeit(lambda: ops.layernorm(x, gamma, beta, ln))
        print(f"M={M:6d} C={C:5d}  out-proj+res {t_o:7.1f} us   + row moments {t_os:7.1f} us   layernorm kernel {t_ln:6.1f} us", flush=True)
        plain = 3 * t_ln + 3 * t_o
        fold = 3 * t_os
        for name, N, act in (("q|k|v", 3 * Cp, 0), ("q", Cp, 0), ("geglu", 8 * C, 2)):
            w = (rn(N, C) * C ** -0.5).half()
            b = rn(N)
            cs = w.float().sum(1).contiguous()
            out = torch.empty(M, N // 2 if act == 2 else N, dtype=torch.float16, device="cuda")
            t0 = timeit(lambda: ops.gemm(ln, w, out, bias=b, act=act))
            t1 = timeit(lambda: ops.gemm(x, w, out, bias=b, act=act, ln=(rs, cs, 1e-5)))
            print(f"      {name:6s} N={N:5d}  plain {t0:7.1f} us   folded {t1:7.1f} us", flush=True)
            plain += t0
            fold += t1
        print(f"      per block: separate {plain:7.1f} us, folded {fold:7.1f} us", flush=True)
        tot_plain += reps * plain
        tot_fold += reps * fold
    print(f"per forward (16 blocks): separate {tot_plain / 1e3:.3f} ms, folded {tot_fold / 1e3:.3f} ms")


if __name__ == "__main__":
    main()
