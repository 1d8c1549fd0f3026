"""Multi-GPU plumbing for the denoising path: one process per GPU, requests sharded by rank.

The path shards naturally: each edit request's S-step trajectory depends only on its
own latent, conditioning and the frozen weights, so there is NO collective inside the denoising
loop.  The only communication is a one-time broadcast of the frozen weights from rank 0
(the reference's DDP construction does the same through accelerate, train.py:536-538).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


def shard_range(n_requests, rank, world):
    """Contiguous slice [lo, hi) of the request list owned by ``rank`` (sizes differ by at most 1;
    a request's CFG pair always stays on one GPU)."""
    base, rem = divmod(n_requests, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@torch.no_grad()
def broadcast_module_(module, src=0):
    """One-time broadcast of every parameter and buffer from ``src`` (frozen UNet / adapter / task table)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)
    invalidate_caches_(module)


def invalidate_caches_(module):
    """``.data`` writes do not bump parameter versions: every module in the tree that caches packed weights (UNetModel,
    MoE) or keys CUDA graphs on them (LatentDenoiser / AnySDDenoiser) is told explicitly, so nothing packed or captured
    before the write can survive it."""
    seen = set()
    for m in module.modules():
        inv = getattr(m, "invalidate", None)
        if callable(inv) and id(m) not in seen:
            seen.add(id(m))
            inv()


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


@torch.no_grad()
def allreduce_sum_(tensors, async_op=True):
    """Training path (train.py:536, 703: DDP over the trainable set only): sum every gradient tensor
    over the ranks in place -- the only data-path collective of the training step.  The mean's 1/world is folded
    into the optimizer's gradient scale by the caller (no extra elementwise pass).  Returns the world size."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 1
    works = [dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op) for t in tensors if t is not None]
    if async_op:
        for w in works:
            w.wait()
    return dist.get_world_size()


def world_size() -> int:
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


@torch.no_grad()
def allreduce_sum_async(t):
    """Launch ``all_reduce(sum)`` of one gradient bucket and return the work handle (None on a single rank).  With the
    NCCL backend the collective runs on NCCL's stream behind everything issued so far on the current stream and overlaps
    whatever is issued next; ``handle.wait()`` makes the current stream wait for it."""
    if world_size() == 1:
        return None
    return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)
