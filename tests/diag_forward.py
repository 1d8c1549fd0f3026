#!/usr/bin/env python
"""Diagnostic (not a pytest file): ONE eager UNet forward of the bench workload (SD-1.5 geometry,
B_eff=16 @ 64x64 latent) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
Also prints the per-kind event-timed breakdown (ops.trace) when run without ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from anyedit_b200 import ops  # noqa: E402
from anyedit_b200.unet import UNetModel  # noqa: E402
from bench import SD15  # noqa: E402


def main():
    B = int(os.environ.get("DIAG_B", "16"))
    h = int(os.environ.get("DIAG_H", "64"))
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with torch.device(dev):
        net = UNetModel(**SD15)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() > 1 and float(p.abs().max()) == 0.0:
                p.uniform_(-0.02, 0.02)
    x = torch.randn(B, 8, h, h, device=dev)
    ctx = torch.randn(B, 77, 768, device=dev)
    t = torch.full((B,), 500, device=dev, dtype=torch.long)
    for _ in range(2):
        net(x, t, context=ctx)
    torch.cuda.synchronize()
    if "--trace" in sys.argv:
        ops.trace = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        net(x, t, context=ctx)
        e1.record()
        torch.cuda.synchronize()
        agg = {}
        shapes = {}
        for kind, fl, a, b, tag in ops.trace:
            d = agg.setdefault(kind, [0.0, 0.0, 0])
            ms = a.elapsed_time(b)
            d[0] += fl
            d[1] += ms
            d[2] += 1
            s = shapes.setdefault((kind, tag), [0.0, 0.0, 0])
            s[0] += fl
            s[1] += ms
            s[2] += 1
        ops.trace = None
        tot = e0.elapsed_time(e1)
        print(f"forward total {tot:.2f} ms (eager, with events)")
        for k, v in agg.items():
            print(f"  {k:10s} {v[2]:4d} launches {v[1]:8.2f} ms {v[0] / v[1] / 1e9:8.1f} TFLOP/s")
        print(f"  other      {tot - sum(v[1] for v in agg.values()):8.2f} ms")
        if "--shapes" in sys.argv:
            for (kind, tag), v in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
                print(f"    {v[1]:7.3f} ms {v[2]:3d}x {v[1] / v[2] * 1e3:8.1f} us {v[0] / v[1] / 1e9:8.1f} TF  {kind} {tag}")
        # plain timing without events
        e0.record()
        for _ in range(3):
            net(x, t, context=ctx)
        e1.record()
        torch.cuda.synchronize()
        print(f"forward (eager, no per-launch events): {e0.elapsed_time(e1) / 3:.2f} ms")
        return
    torch.cuda.cudart().cudaProfilerStart()
    net(x, t, context=ctx)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
