"""Deterministic weights for parity tests (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference zero-initialises the last conv of every ResBlock, every
transformer ``proj_out`` and the output conv (openaimodel.py:228-230, 729;
attention.py:312-318), so a freshly built UNet outputs exactly 0 and proves
nothing.  Parity tests therefore fill *every* tensor from a counter-based numpy
stream keyed by the parameter name, which is identical on every machine and
independent of torch's RNG and of parameter creation order.
"""
import zlib

import numpy as np
import torch


def _uniform(name: str, n: int, seed: int) -> np.ndarray:
    key = (zlib.crc32(name.encode()) << 16) ^ (seed & 0xFFFFFFFF)
    rng = np.random.Generator(np.random.PCG64(key))
    return rng.random(n, dtype=np.float32) * 2.0 - 1.0  # U(-1, 1)


def fill_tensor(name: str, shape, seed: int = 0, gain: float = 1.0) -> torch.Tensor:
    shape = tuple(int(s) for s in shape)
    n = int(np.prod(shape)) if len(shape) else 1
    u = _uniform(name, n, seed)
    leaf = name.rsplit(".", 1)[-1]
    if len(shape) == 1:
        is_norm = any(t in name for t in (".norm", "in_layers.0.", "out_layers.0.", "out.0."))
        if leaf == "weight" and is_norm:
            v = 1.0 + 0.2 * u            # norm gains around 1
        else:
            v = 0.1 * u                  # biases / norm shifts
    else:
        if "label_emb" in name or "task_embs" in name:
            v = 0.5 * u                  # embedding tables
        else:
            fan_in = int(np.prod(shape[1:]))
            # U(-a, a) with a = sqrt(3/fan_in) -> unit-gain (variance 1/fan_in) layer
            v = u * np.sqrt(3.0 / fan_in) * gain
    return torch.from_numpy(v.reshape(shape).astype(np.float32))


def make_state_dict(shapes: dict, seed: int = 0, gain: float = 1.0) -> dict:
    """shapes: name -> shape (e.g. ``{k: v.shape for k, v in module.state_dict().items()}``)."""
    return {k: fill_tensor(k, s, seed, gain) for k, s in shapes.items()}


def checksum(sd: dict) -> float:
    """Order-independent float64 checksum, stored in golden fixtures to detect drift."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(v.sum()) + 0.5 * float((v * v).sum())
    return tot
