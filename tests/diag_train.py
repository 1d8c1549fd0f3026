#!/usr/bin/env python
"""Diagnostic (not a pytest file): one AnySD training step at BASELINE config 4's per-GPU shape (SD-1.5 geometry,
16 requests @ 64x64 latent, 11 experts, 16 visual tokens): ms per step (forward + backward + AdamW) and launches."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from anyedit_b200 import ops  # noqa: E402
from anyedit_b200.anysd import MoE  # noqa: E402
from anyedit_b200.training import AdapterTrainer  # noqa: E402
from anyedit_b200.unet import UNetModel  # noqa: E402
from bench import SD15  # noqa: E402


def main():
    B = int(os.environ.get("DIAG_B", "16"))
    hw = int(os.environ.get("DIAG_H", "64"))
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with torch.device(dev):
        net = UNetModel(**SD15)
        moe = MoE(net, None, expert_num=11, num_tasks=20)
    with torch.no_grad():
        for p in moe.parameters():
            if p.dim() > 1 and float(p.abs().max()) == 0.0:
                p.uniform_(-0.02, 0.02)
    g = torch.Generator(device=dev).manual_seed(1)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    lat, noise, img = rn(B, 4, hw, hw), rn(B, 4, hw, hw), rn(B, 4, hw, hw)
    text, vis = rn(B, 77, 768), rn(B, 16, 768)
    t = torch.randint(0, 1000, (B,), device=dev, generator=g)
    code = torch.arange(B, device=dev) % 20
    tr = AdapterTrainer(moe, lr=1e-5, loss_scale=1024.0)
    for i in range(2):
        loss, _ = tr.step(lat, noise, t, img, text, vis, code)
    torch.cuda.synchronize()
    n0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    iters = 3
    for i in range(iters):
        loss, _ = tr.step(lat, noise, t, img, text, vis, code)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 3 * 803.37e9 * B * (hw / 64) ** 2
    print(f"training step B={B} @{hw}x{hw}: {ms:.1f} ms/step (host wall {(time.time() - t0) / iters * 1e3:.1f} ms), "
          f"{(ops.launch_count - n0) // iters} launches/step, loss {float(loss):.4f}, ~{fl / ms / 1e9:.0f} TFLOP/s of 3*F(h) per sample, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)


    if "--trace" in sys.argv:
        ops.trace = []
        e0.record()
        tr.step(lat, noise, t, img, text, vis, code)
        e1.record()
        torch.cuda.synchronize()
        agg, shapes = {}, {}
        for kind, fl, a, b, tag in ops.trace:
            d = agg.setdefault(kind, [0.0, 0])
            d[0] += a.elapsed_time(b)
            d[1] += 1
            if kind in ("attention_bwd", "expert_attention_bwd", "attention", "groupnorm_bwd"):
                sd = shapes.setdefault((kind, tag), [0.0, 0])
                sd[0] += a.elapsed_time(b)
                sd[1] += 1
        ops.trace = None
        tot = e0.elapsed_time(e1)
        print(f"traced step {tot:.1f} ms")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"  {k:16s} {v[1]:5d} launches {v[0]:8.2f} ms")
        print(f"  untraced        {tot - sum(v[0] for v in agg.values()):8.2f} ms")
        for (k, tag), v in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:14]:
            print(f"    {v[0]:7.2f} ms {v[1]:3d}x  {k} {tag}")


if __name__ == "__main__":
    main()
