#!/usr/bin/env python
"""Diagnostic (not a pytest file): contraction kernels vs fp32 reference + timing on the UNet's real shapes.
usage: python tests/diag_gemm.py [--quick]     (ANYSD_FORCE_MMA=1 selects the mma.sync kernel)"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from anyedit_b200 import ops  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def timeit(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def gemm_case(M, N, K, check=True, act=0, res=False):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).half()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    R = torch.randn(M, N if act != 2 else N // 2, device="cuda", generator=g).half() if res else None
    out = torch.empty(M, N if act != 2 else N // 2, dtype=torch.float16, device="cuda")
    fn = lambda: ops.gemm(A, W, out, bias=bias, act=act, residual=R)
    t = timeit(fn)
    err = float("nan")
    if check:
        ref = A.float() @ W.float().t() + bias
        if act == 2:
            ref = ref[:, 0::2] * F.gelu(ref[:, 1::2])
        if act == 1:
            ref = F.silu(ref)
        if res:
            ref = ref + R.float()
        err = rel(out.float(), ref)
    print(f"gemm  M={M:6d} N={N:5d} K={K:5d} act={act} res={int(res)}  {t * 1e6:9.1f} us  {2.0 * M * N * K / t / 1e12:7.1f} TFLOP/s  rel={err:.2e}",
          flush=True)
    return err


def conv_case(Nimg, Cin, Cout, H, W, check=True):
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(Nimg, H, W, Cin, device="cuda", generator=g).half()
    w = (torch.randn(Cout, 3, 3, Cin, device="cuda", generator=g) * (9 * Cin) ** -0.5).half()
    bias = torch.randn(Cout, device="cuda", generator=g) * 0.1
    emb = torch.randn(Nimg, Cout, device="cuda", generator=g)
    out = torch.empty(Nimg * H * W, Cout, dtype=torch.float16, device="cuda")
    fn = lambda: ops.conv3x3(x, w.view(Cout, -1), out, bias=bias, rowadd=emb)
    t = timeit(fn)
    err = float("nan")
    if check:
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=1) + emb[:, :, None, None]
        err = rel(out.float().view(Nimg, H, W, Cout).permute(0, 3, 1, 2), ref)
    fl = 2.0 * Nimg * H * W * 9 * Cin * Cout
    print(f"conv  N={Nimg:3d} {Cin:4d}->{Cout:4d} @{H:3d}x{W:<3d}  {t * 1e6:9.1f} us  {fl / t / 1e12:7.1f} TFLOP/s  rel={err:.2e}", flush=True)
    return err


def main():
    quick = "--quick" in sys.argv
    print("force_mma =", os.environ.get("ANYSD_FORCE_MMA"), "device", torch.cuda.get_device_name(0), flush=True)
    print("cpu: nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
    try:
        print("cgroup cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip())
    except Exception as e:
        print("cgroup cpu.max: n/a", e)
    worst = 0.0
    # small correctness cases first (a hang or garbage shows up before the big shapes)
    for M, N, K in ((128, 128, 64), (128, 160, 128), (256, 256, 192), (200, 320, 320), (77, 640, 768), (16, 1280, 1280),
                    (513, 136, 72), (4096, 320, 320)):
        worst = max(worst, gemm_case(M, N, K))
    worst = max(worst, gemm_case(384, 512, 192, act=2, res=True))
    worst = max(worst, gemm_case(384, 256, 192, act=1))
    for c in ((2, 64, 64, 6, 10), (1, 128, 320, 16, 16), (2, 320, 320, 8, 8), (3, 64, 128, 4, 4), (2, 64, 96, 12, 24), (1, 64, 64, 96, 96)):
        worst = max(worst, conv_case(*c))
    print("worst rel err (small cases):", worst, flush=True)
    if quick:
        return
    B = 16
    for M, N, K in ((B * 4096, 320, 320), (B * 4096, 960, 320), (B * 4096, 2560, 320), (B * 4096, 320, 1280),
                    (B * 1024, 640, 640), (B * 1024, 5120, 640), (B * 256, 1280, 1280), (B * 256, 10240, 1280),
                    (B * 256, 1280, 5120), (B * 64, 1280, 1280), (B * 77, 640, 768), (B, 20160, 1280)):
        gemm_case(M, N, K, check=False)
    gemm_case(B * 4096, 2560, 320, check=False, act=2)
    gemm_case(B * 1024, 5120, 640, check=False, act=2)
    gemm_case(B * 256, 10240, 1280, check=False, act=2)
    gemm_case(B * 4096, 320, 1280, check=False, res=True)
    gemm_case(B * 4096, 320, 320, check=False, res=True)
    for c in ((B, 320, 320, 64, 64), (B, 640, 320, 64, 64), (B, 960, 320, 64, 64), (B, 640, 640, 32, 32), (B, 1280, 640, 32, 32),
              (B, 1920, 640, 32, 32), (B, 1280, 1280, 16, 16), (B, 2560, 1280, 16, 16), (B, 1280, 1280, 8, 8), (B, 2560, 1280, 8, 8)):
        conv_case(*c, check=False)


if __name__ == "__main__":
    main()
