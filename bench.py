#!/usr/bin/env python
"""bench.py -- edited images/sec at 512x512 / 50-step DDIM / CFG (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one complete CFG DDIM sampling run (``DDIMSampler.sample``: 50 UNet evaluations on the
CFG-doubled batch + 50 fused updates) over one batch of synthetic edit requests.  Workload at every
N = BASELINE configs[1] per GPU (512x512 "replace" edit -> 64x64 latent, 50 DDIM steps, CFG 7.5,
batch 8, SD-1.5/IP2P UNet geometry with 8 input channels and 77x768 text context); N GPUs run N
such batches (weak scaling, = configs[2] at N=8), no collective inside the loop.

Prints ONE JSON line (rank 0).  ``value`` is device-resident throughput, ``e2e`` the same metric
through the public API with pinned-host inputs/outputs inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SD15 = dict(image_size=32, in_channels=8, model_channels=320, out_channels=4, num_res_blocks=2,
            attention_resolutions=[4, 2, 1], channel_mult=[1, 2, 4, 4], num_heads=8,
            use_spatial_transformer=True, transformer_depth=1, context_dim=768, legacy=False)
F64 = 803.37e9          # algorithmic FLOPs of one UNet sample-forward at a 64x64 latent
METRIC = "edited images/sec (512x512, 50-step DDIM, CFG)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="edit requests per GPU")
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--scale", type=float, default=7.5)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--skip-roofline", action="store_true")
    return ap.parse_args()


# ---- clocks sampler (nvidia-smi during the timed region) ------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # "under load": the upper half of the samples (the sampler also sees the idle edges)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU arm: the oracle port of the reference's own UNet + DDIM loop ------------------------------
_CPU_SD = None


def cpu_arm(latent, ddim_steps, scale, n_steps_sample=3):
    """Times the reference's CPU implementation of the path (oracle port: eager torch fp32, all host
    threads) on a bounded sample of the workload: ONE edit request, CFG (B_eff=2), ``n_steps_sample``
    DDIM steps after one warm-up step; images/sec extrapolated linearly to ``ddim_steps`` steps."""
    import torch
    from oracle import ddim_oracle, unet_oracle, weights
    from anyedit_b200.unet import UNetModel
    from oracle.cpu import usable_cores
    cores = usable_cores()              # min(affinity, cgroup CPU quota): what the box lets us use
    torch.set_num_threads(cores)
    global _CPU_SD
    if _CPU_SD is None:                 # 859.5 M seeded weights, generated once per process (untimed)
        with torch.device("meta"):
            shapes = {k: tuple(v.shape) for k, v in UNetModel(**SD15).state_dict().items()}
        _CPU_SD = weights.make_state_dict(shapes, 3)
    sd = _CPU_SD
    gen = torch.Generator().manual_seed(1234)
    x_T, c_cat = torch.randn(1, 4, latent, latent, generator=gen), torch.randn(1, 4, latent, latent, generator=gen)
    c_txt, u_txt = torch.randn(1, 77, 768, generator=gen), torch.randn(1, 77, 768, generator=gen)
    unet = lambda x, t, context=None, y=None: unet_oracle.unet_forward(sd, x, t, context, y, num_heads=8)
    times = []

    def model_fn(x, t, c):
        t0 = time.perf_counter()
        out = ddim_oracle.apply_model(unet, "hybrid", x, t, c)
        times.append(time.perf_counter() - t0)
        return out

    sched = ddim_oracle.register_schedule("linear", 1000, 0.00085, 0.012)
    with torch.no_grad():
        # a (1 + n)-step DDIM run: the first UNet call is the warm-up
        ddim_oracle.ddim_sample(model_fn, sched, 1 + n_steps_sample, x_T,
                                {"c_concat": [c_cat], "c_crossattn": [c_txt]},
                                {"c_concat": [c_cat], "c_crossattn": [u_txt]}, scale, eta=0.0)
    per_step = sum(times[1:]) / max(1, len(times) - 1)
    ips = 1.0 / (per_step * ddim_steps)
    return {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"oracle port of the reference ldm UNetModel+DDIMSampler (fp32 eager torch, {cores} threads): "
                      f"1 request, CFG B_eff=2, {latent}x{latent} latent, {n_steps_sample} DDIM steps after 1 warm-up "
                      f"({per_step:.2f} s/step), extrapolated linearly to {ddim_steps} steps",
            "sec_per_unet_step": per_step}


def run_reference(args, rank, world):
    if rank != 0:
        return
    t_steps = []
    last = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        last = cpu_arm(args.latent, args.ddim_steps, args.scale, n_steps_sample=1 if i < args.warmup else 3)
        if i >= args.warmup:
            t_steps.append(time.perf_counter() - t0)
    ips = last["value"]
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t_steps) / len(t_steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world), "cpu_baseline": {**last, "value": ips},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"BASELINE configs[1] per GPU: {8 * args.latent}x{8 * args.latent} replace edit, "
                        f"{args.ddim_steps} DDIM steps + CFG {args.scale}, batch {args.batch}/GPU "
                        f"(SD-1.5/IP2P UNet geometry: in=8, 320ch, mult 1-2-4-4, 8 heads, ctx 77x768)",
            "global_batch": args.batch * world, "latent": args.latent, "ddim_steps": args.ddim_steps,
            "guidance_scale": args.scale, "parallelism": f"dp{world} (requests sharded, no in-loop collective)",
            "l2_policy": "working set > L2 every step (1.72 GB fp16 weights + >1 GB activations vs 126 MB L2)"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from anyedit_b200 import distributed as D
    from anyedit_b200 import ops
    from anyedit_b200.ddim import DDIMSampler
    from anyedit_b200.diffusion import LatentDenoiser
    from anyedit_b200.unet import UNetModel

    rank, local, world = D.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: anyedit_b200 has no CPU fallback"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    # ---- model: random-init weights of the named architecture, built on rank 0, broadcast once ----
    torch.manual_seed(0)
    with torch.device(dev):
        net = UNetModel(**SD15)
    with torch.no_grad():
        for name, p in net.named_parameters():      # re-randomise the zero-init tensors
            if float(p.abs().max()) == 0.0 and p.dim() > 1:
                fan_in = p[0].numel()
                p.uniform_(-(3.0 / fan_in) ** 0.5, (3.0 / fan_in) ** 0.5)
    model = LatentDenoiser(net, "hybrid").to(dev)
    D.broadcast_module_(model, src=0)                # the one collective of the inference path
    net.prepare()

    B, h, S = args.batch, args.latent, args.ddim_steps
    gen = torch.Generator().manual_seed(1234 + rank)
    host = {"x_T": torch.randn(B, 4, h, h, generator=gen), "c_cat": torch.randn(B, 4, h, h, generator=gen),
            "c_txt": torch.randn(B, 77, 768, generator=gen),
            "u_txt": torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(4321)).repeat(B, 1, 1)}
    host = {k: v.pin_memory() for k, v in host.items()}
    out_host = torch.empty(B, 4, h, h).pin_memory()
    sampler = DDIMSampler(model, use_cuda_graph=not args.no_graph)

    def run(devt):
        cond = {"c_concat": [devt["c_cat"]], "c_crossattn": [devt["c_txt"]]}
        uncond = {"c_concat": [devt["c_cat"]], "c_crossattn": [devt["u_txt"]]}
        out, _ = sampler.sample(S, B, (4, h, h), cond, verbose=False, x_T=devt["x_T"], eta=0.0,
                                unconditional_guidance_scale=args.scale, unconditional_conditioning=uncond)
        return out

    def step_resident(devt):
        return run(devt)

    def step_e2e():
        devt = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        out = run(devt)
        out_host.copy_(out, non_blocking=True)
        return out

    def timed(fn, n):
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        D.barrier()
        return D.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)

    devt = {k: v.to(dev) for k, v in host.items()}
    for _ in range(args.warmup):
        step_resident(devt)
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    if os.environ.get("ANYSD_NCU"):      # ncu --profile-from-start off: launch list of the timed region only
        torch.cuda.cudart().cudaProfilerStart()
    n0 = ops.launch_count
    t_res = timed(lambda: step_resident(devt), args.steps)
    launches = ops.launch_count - n0
    t_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop() if rank == 0 else None

    imgs = B * world * args.steps
    value, e2e = imgs / t_res, imgs / t_e2e
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * out_host.element_size()

    # ---- roofline of the dominant kernel (tensor-core contraction: conv3x3 + linear GEMMs) --------
    roof = None
    if not args.skip_roofline and rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        ops.trace = []
        x_in = torch.cat([devt["x_T"]] * 2)
        cond2 = {"c_concat": [torch.cat([devt["c_cat"]] * 2)], "c_crossattn": [torch.cat([devt["u_txt"], devt["c_txt"]])]}
        tt = torch.full((2 * B,), 500, device=dev, dtype=torch.long)
        for _ in range(2):
            ops.trace.clear()
            model.apply_model(x_in, tt, cond2)       # eager (no graph): events around every contraction launch
        torch.cuda.synchronize()
        agg = {}
        for kind, fl, e0, e1, _tag in ops.trace:
            a = agg.setdefault(kind, [0.0, 0.0, 0])
            a[0] += fl
            a[1] += e0.elapsed_time(e1) * 1e-3
            a[2] += 1
        ops.trace = None
        gemm_f = agg.get("gemm", [0, 0, 0])[0] + agg.get("conv3x3", [0, 0, 0])[0]
        gemm_t = agg.get("gemm", [0, 1e-9, 0])[1] + agg.get("conv3x3", [0, 0, 0])[1]
        ach = gemm_f / gemm_t / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json"))).get("bytes_per_launch")
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "implicit-GEMM conv3x3 + linear contraction (anysd_gemm_f16)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback",
                "traffic": traffic,
                "per_kind": {k: {"tflops": v[0] / v[1] / 1e12, "ms": v[1] * 1e3, "launches": v[2], "gflop": v[0] / 1e9}
                             for k, v in agg.items()},
                "whole_step_tflops": 2 * S * F64 * B * world * args.steps / t_res / 1e12 if h == 64 else None}

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_arm(h, S, args.scale)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": workload_config(args, world), "clocks": clk,
                "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * t_e2e / args.steps},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
