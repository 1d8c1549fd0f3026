"""the AnySD training step (train.py:629-710) on the CUDA path vs the oracle restatement
(oracle/train_oracle.py: the same loss through the fp32 CPU oracle forward, gradients by torch autograd; its autograd
is pinned against the reference UNet's in tests/test_oracle_golden.py)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
# fp16 operands and fp16 activation gradients through ~40 layers of backward: relative L2 per gradient tensor
GRAD_TOL = 2e-2
# router gradients are differences of per-expert gate gradients (softmax: they sum to zero over the experts), so the
# fp16 noise of the gate gradients is amplified by the cancellation; measured 5e-2 with two experts
ROUTER_TOL = 1e-1
LOSS_TOL = 2e-3


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _setup(name, seed, E, T, B, hw, n_vis, gain=2.0):
    from anyedit_b200.anysd import MoE
    from anyedit_b200.unet import UNetModel
    from oracle import anysd_oracle, ddim_oracle, weights
    meta = json.load(open(os.path.join(G, f"{name}_keys.json")))
    cfg = meta["config"]
    net = UNetModel(**cfg)
    sd = weights.make_state_dict({k: tuple(v) for k, v in meta["keys"].items()}, seed)
    net.load_state_dict(sd, strict=True)
    moe = MoE(net.cuda(), None, expert_num=E, num_tasks=T).cuda()
    shapes = anysd_oracle.adapter_shapes({k: tuple(v.shape) for k, v in sd.items()}, T, E, cfg["context_dim"])
    asd = weights.make_state_dict(shapes, 77, gain=gain)
    moe.load_state_dict(asd, strict=False)
    gen = torch.Generator().manual_seed(100 + seed)
    batch = dict(latents=torch.randn(B, 4, hw, hw, generator=gen), noise=torch.randn(B, 4, hw, hw, generator=gen),
                 image_latent=torch.randn(B, 4, hw, hw, generator=gen), text=torch.randn(B, 7, cfg["context_dim"], generator=gen),
                 vis=torch.randn(B, n_vis, cfg["context_dim"], generator=gen) if n_vis else None,
                 t=torch.randint(0, 1000, (B,), generator=gen), code=torch.randint(0, T, (B,), generator=gen))
    acp = ddim_oracle.register_schedule("linear", 1000, 0.00085, 0.012)["alphas_cumprod"]
    return moe, sd, asd, cfg, batch, acp


def _cuda(b):
    return {k: (v.cuda() if v is not None else None) for k, v in b.items()}


def test_training_step_gradients_vs_oracle():
    """loss, prediction and every trainable's gradient (experts' to_k_ip / to_v_ip, router weight / bias, task table) plus
    d loss / d visual tokens, tiny_a geometry (8-channel input, 3 levels incl. Down/Upsample, skip concats)."""
    from anyedit_b200.training import AdapterTrainer
    from oracle import train_oracle
    moe, sd, asd, cfg, b, acp = _setup("tiny_a", 11, E=3, T=6, B=3, hw=16, n_vis=5)
    tr = AdapterTrainer(moe, loss_scale=256.0)
    c = _cuda(b)
    loss, pred, grads = tr.loss_and_grads(c["latents"], c["noise"], c["t"], c["image_latent"], c["text"], c["vis"], c["code"])
    torch.cuda.synchronize()
    loss_ref, pred_ref, g_ref = train_oracle.train_step_grads(sd, asd, b["latents"], b["noise"], b["t"], b["image_latent"], b["text"],
                                                              b["code"], b["vis"], acp, num_heads=cfg["num_heads"])
    assert rel(pred, pred_ref) < 4e-3
    assert abs(float(loss) - float(loss_ref)) < LOSS_TOL * float(loss_ref)
    worst = {}
    for k, ref in g_ref.items():
        assert k in grads, k
        got = grads[k].float() / tr.loss_scale
        assert torch.isfinite(got).all(), k
        if float(ref.abs().max()) == 0.0:
            assert float(got.abs().max()) < 1e-6, k
            continue
        kind = k.split(".")[-2] + "." + k.split(".")[-1] if "." in k else k
        worst[kind] = max(worst.get(kind, 0.0), rel(got.reshape(ref.shape), ref))
    print("[training step] worst relative L2 per gradient kind:", {k: f"{v:.2e}" for k, v in worst.items()})
    for k, v in worst.items():
        assert v < (ROUTER_TOL if k.startswith("router") else GRAD_TOL), (k, v)


def test_training_step_without_visual_tokens_and_optimizer():
    """No visual stream: only the task table trains (through every ResBlock's time-embedding row add); three AdamW steps
    on the CUDA path track the oracle's (autograd + restated AdamW) parameter trajectory; the loss goes down."""
    from anyedit_b200.training import AdapterTrainer
    from oracle import train_oracle
    moe, sd, asd, cfg, b, acp = _setup("tiny_a", 11, E=3, T=6, B=2, hw=16, n_vis=0)
    tr = AdapterTrainer(moe, lr=2e-2, weight_decay=1e-2, loss_scale=256.0)
    c = _cuda(b)
    table = asd["task_embs.weight"].clone()
    m, v = torch.zeros_like(table), torch.zeros_like(table)
    losses = []
    for step in (1, 2, 3):
        loss, _ = tr.step(c["latents"], c["noise"], c["t"], c["image_latent"], c["text"], None, c["code"])
        cur = dict(asd)
        cur["task_embs.weight"] = table
        loss_ref, _, g_ref = train_oracle.train_step_grads(sd, cur, b["latents"], b["noise"], b["t"], b["image_latent"], b["text"],
                                                           b["code"], None, acp, num_heads=cfg["num_heads"])
        table, m, v = train_oracle.adamw_step(table, g_ref["task_embs.weight"], m, v, step, 2e-2, 0.9, 0.999, 1e-8, 1e-2)
        losses.append(float(loss))
        assert abs(float(loss) - float(loss_ref)) < 5e-3 * float(loss_ref), (step, float(loss), float(loss_ref))
    got = moe.task_embs.weight.detach().float().cpu()
    used = torch.unique(b["code"])
    assert float((got[used] - table[used]).abs().max()) < 2e-2 * 3 + 1e-3        # 3 Adam steps of lr 2e-2 (sign-like updates)
    assert losses[-1] < losses[0]


def test_training_step_linear_projection_geometry():
    """tiny_b-like geometry (num_head_channels, linear proj_in/out) without the class embedding: gradients vs oracle."""
    from anyedit_b200.anysd import MoE
    from anyedit_b200.training import AdapterTrainer
    from anyedit_b200.unet import UNetModel
    from oracle import anysd_oracle, ddim_oracle, train_oracle, weights
    cfg = dict(image_size=16, in_channels=8, model_channels=64, out_channels=4, num_res_blocks=2, attention_resolutions=[2, 1],
               channel_mult=[1, 2], num_head_channels=32, use_spatial_transformer=True, use_linear_in_transformer=True,
               transformer_depth=1, context_dim=96, legacy=False)
    net = UNetModel(**cfg)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = weights.make_state_dict(shapes, 14)
    net.load_state_dict(sd)
    E, T, B = 2, 4, 2
    moe = MoE(net.cuda(), None, expert_num=E, num_tasks=T).cuda()
    ashapes = anysd_oracle.adapter_shapes(shapes, T, E, 96)
    asd = weights.make_state_dict(ashapes, 79, gain=2.0)
    moe.load_state_dict(asd, strict=False)
    gen = torch.Generator().manual_seed(9)
    lat, noise, img = (torch.randn(B, 4, 8, 8, generator=gen) for _ in range(3))
    text, vis = torch.randn(B, 5, 96, generator=gen), torch.randn(B, 4, 96, generator=gen)
    t, code = torch.tensor([30, 870]), torch.tensor([3, 1])
    acp = ddim_oracle.register_schedule("linear", 1000, 0.00085, 0.012)["alphas_cumprod"]
    tr = AdapterTrainer(moe, loss_scale=256.0)
    loss, pred, grads = tr.loss_and_grads(lat.cuda(), noise.cuda(), t.cuda(), img.cuda(), text.cuda(), vis.cuda(), code.cuda())
    loss_ref, pred_ref, g_ref = train_oracle.train_step_grads(sd, asd, lat, noise, t, img, text, code, vis, acp, num_head_channels=32)
    assert rel(pred, pred_ref) < 4e-3
    worst = {}
    for k, ref in g_ref.items():
        if float(ref.abs().max()) == 0.0:
            continue
        kind = k.split(".")[-2] + "." + k.split(".")[-1] if "." in k else k
        worst[kind] = max(worst.get(kind, 0.0), rel((grads[k].float() / tr.loss_scale).reshape(ref.shape), ref))
    print("[training step, linear-projection geometry] worst relative L2 per gradient kind:", {k: f"{v:.2e}" for k, v in worst.items()})
    for k, v in worst.items():
        assert v < (ROUTER_TOL if k.startswith("router") else GRAD_TOL), (k, v)


def test_c4_training_step_sd15_geometry():
    """BASELINE config 4 flavour at a size the CPU oracle's autograd finishes in about a minute: SD-1.5 geometry
    (8 heads, d = 40 at level 0 -> the padded-head / aux_cols packing and its ln 2 score factor in the backward,
    d = 160 two-pass dk/dv), 11 experts, 16 visual tokens, 2 requests at 16x16 latent."""
    from anyedit_b200.training import AdapterTrainer
    from oracle import cpu, train_oracle
    moe, sd, asd, cfg, b, acp = _setup("sd15", 3, E=11, T=20, B=2, hw=16, n_vis=16, gain=1.5)
    tr = AdapterTrainer(moe, loss_scale=1024.0)
    c = _cuda(b)
    loss, pred, grads = tr.loss_and_grads(c["latents"], c["noise"], c["t"], c["image_latent"], c["text"], c["vis"], c["code"])
    torch.cuda.synchronize()
    torch.set_num_threads(cpu.usable_cores())
    loss_ref, pred_ref, g_ref = train_oracle.train_step_grads(sd, asd, b["latents"], b["noise"], b["t"], b["image_latent"], b["text"],
                                                              b["code"], b["vis"], acp, num_heads=cfg["num_heads"])
    assert rel(pred, pred_ref) < 4e-3
    assert abs(float(loss) - float(loss_ref)) < LOSS_TOL * float(loss_ref)
    worst = {}
    for k, ref in g_ref.items():
        got = grads[k].float() / tr.loss_scale
        assert torch.isfinite(got).all(), k
        if float(ref.abs().max()) == 0.0:
            continue
        kind = k.split(".")[-2] + "." + k.split(".")[-1] if "." in k else k
        worst[kind] = max(worst.get(kind, 0.0), rel(got.reshape(ref.shape), ref))
    print("[C4 training step, sd15 geometry, 11 experts] worst relative L2 per gradient kind:", {k: f"{v:.2e}" for k, v in worst.items()})
    for k, v in worst.items():
        assert v < (ROUTER_TOL if k.startswith("router") else 3e-2), (k, v)


def test_overflow_step_is_skipped_and_scale_adapts():
    """fp16 activation gradients under a loss scale: a step whose gradients overflow must not touch the
    parameters or the AdamW moments and must halve the scale (GradScaler semantics, kept on the device -- no host sync in
    ``step``); clean steps advance the step count and grow the scale after ``growth_interval`` of them."""
    from anyedit_b200.training import AdapterTrainer
    moe, sd, asd, cfg, b, acp = _setup("tiny_a", 11, E=3, T=6, B=2, hw=16, n_vis=5)
    c = _cuda(b)
    args = (c["latents"], c["noise"], c["t"], c["image_latent"], c["text"], c["vis"], c["code"])
    tr = AdapterTrainer(moe, lr=1e-2, loss_scale=2.0 ** 40)                 # 2^40 * dL/dpred overflows fp16 at once
    before = {k: v.detach().clone() for k, v in tr.trainables().items()}
    loss, _ = tr.step(*args)
    assert torch.isfinite(loss).all()
    assert tr.step_count == 0 and tr.loss_scale == 2.0 ** 39
    assert all(torch.equal(before[k], v.detach()) for k, v in tr.trainables().items())
    assert float(tr._flat["M"].abs().max()) == 0.0 and float(tr._flat["V"].abs().max()) == 0.0
    # the packed adapter tensors a sampler would use are still the old ones, bit for bit
    tr2 = AdapterTrainer(moe, lr=1e-2, loss_scale=256.0, growth_interval=2)
    for _ in range(2):
        tr2.step(*args)
    assert tr2.step_count == 2 and tr2.loss_scale == 512.0
    moved = [k for k, v in tr2.trainables().items() if not torch.equal(before[k], v.detach())]
    assert len(moved) == len(before), sorted(set(before) - set(moved))
    assert all(torch.isfinite(v).all() for v in tr2.trainables().values())
    # a static scale never moves
    tr3 = AdapterTrainer(moe, lr=1e-2, loss_scale=128.0, dynamic_loss_scale=False)
    tr3.step(*args)
    assert tr3.loss_scale == 128.0 and tr3.step_count == 1
    # fp16 / strided tensors must be refused by the raw-pointer optimizer kernels, not silently corrupted
    from anyedit_b200 import ops
    p = torch.zeros(8, 8, device="cuda")
    with pytest.raises(ValueError):
        ops.adamw_(p.half(), p, p.clone(), p.clone(), 1, 1e-3)
    with pytest.raises(ValueError):
        ops.adamw_(p.t(), p, p.clone(), p.clone(), 1, 1e-3)


def test_training_two_rank_nccl(tmp_path):
    """config 4 is data parallel over 8 GPUs: 2 ranks (skipped on a 1-GPU box) each back-propagate half of a 4-request
    batch, all-reduce the trainables' gradients over NCCL and step; both ranks end with identical parameters, equal
    (up to fp16 batch-composition noise) to one rank stepping on the whole batch's mean gradient."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    script = tmp_path / "t.py"
    script.write_text(
        "import os, sys, json, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "from anyedit_b200 import distributed as D\n"
        "from anyedit_b200.training import AdapterTrainer\n"
        "from test_gpu_train_step import _setup\n"
        "rank, local, world = D.init_from_env('nccl')\n"
        "moe, sd, asd, cfg, b, acp = _setup('tiny_a', 11, E=3, T=6, B=4, hw=16, n_vis=5)\n"
        "D.broadcast_module_(moe, src=0)\n"
        "lo, hi = D.shard_range(4, rank, world)\n"
        "c = {k: (v[lo:hi].cuda() if v is not None else None) for k, v in b.items()}\n"
        "tr = AdapterTrainer(moe, lr=1e-2, loss_scale=256.0)\n"
        "tr.step(c['latents'], c['noise'], c['t'], c['image_latent'], c['text'], c['vis'], c['code'])\n"
        "w = moe.adapter_modules[3].to_k_ip.weight.detach().float().flatten()[:4096].contiguous()\n"
        "ws = [torch.empty_like(w) for _ in range(world)]\n"
        "dist.all_gather(ws, w)\n"
        "assert torch.equal(ws[0], ws[1])\n"
        "if rank == 0:\n"
        "    moe2, *_ = _setup('tiny_a', 11, E=3, T=6, B=4, hw=16, n_vis=5)\n"
        "    full = {k: (v.cuda() if v is not None else None) for k, v in b.items()}\n"
        "    tr2 = AdapterTrainer(moe2, lr=1e-2, loss_scale=256.0)\n"
        "    l0, _, g0 = tr2.loss_and_grads(*(full[k][:2] for k in ('latents', 'noise', 't', 'image_latent', 'text', 'vis', 'code')))\n"
        "    l1, _, g1 = tr2.loss_and_grads(*(full[k][2:] for k in ('latents', 'noise', 't', 'image_latent', 'text', 'vis', 'code')))\n"
        "    g0.pop('visual_tokens'); g1.pop('visual_tokens')\n"
        "    tr2.apply_gradients({k: g0[k] + g1[k] for k in g0}, world=2)\n"
        "    w2 = moe2.adapter_modules[3].to_k_ip.weight.detach().float().flatten()[:4096]\n"
        "    assert float((w2 - w).abs().max()) < 1e-6, float((w2 - w).abs().max())\n"
        "    print('OK')\n"
        "dist.destroy_process_group()\n")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
