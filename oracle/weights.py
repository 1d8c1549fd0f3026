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


_ZERO_INIT = ("out_layers.3.", "proj_out.", "out.2.")     # zero_module'd tensors of the reference


def fill_tensor_torch_default(name: str, shape, seed: int = 0) -> torch.Tensor:
    """SURVEY.md 8d weight scheme: PyTorch default init (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv /
    linear weights and biases, ones / zeros for norms, N(0,1) embeddings) with the reference's zero-init
    tensors re-randomised N(0, 0.02) -- reproduced with the name-keyed numpy stream."""
    shape = tuple(int(s) for s in shape)
    n = int(np.prod(shape)) if len(shape) else 1
    is_norm = any(t in name for t in (".norm", "in_layers.0.", "out_layers.0.", "out.0."))
    leaf = name.rsplit(".", 1)[-1]
    if is_norm and len(shape) == 1:
        v = np.ones(n, np.float32) if leaf == "weight" else np.zeros(n, np.float32)
        return torch.from_numpy(v.reshape(shape))
    key = (zlib.crc32(name.encode()) << 16) ^ (seed & 0xFFFFFFFF)
    rng = np.random.Generator(np.random.PCG64(key))
    if any(z in name for z in _ZERO_INIT):
        v = rng.standard_normal(n, dtype=np.float32) * 0.02
    elif "label_emb" in name or "task_embs" in name:
        v = rng.standard_normal(n, dtype=np.float32)
    else:
        # a bias uses its layer's weight fan_in (torch default), looked up by module name
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else _FAN_IN.get(name.rsplit(".", 1)[0])
        bound = 1.0 / np.sqrt(fan_in) if fan_in else 0.0
        v = (rng.random(n, dtype=np.float32) * 2.0 - 1.0) * bound
    return torch.from_numpy(v.reshape(shape).astype(np.float32))


_FAN_IN = {}


def make_state_dict(shapes: dict, seed: int = 0, gain: float = 1.0, scheme: str = "unit") -> dict:
    """shapes: name -> shape (e.g. ``{k: v.shape for k, v in module.state_dict().items()}``).

    scheme "unit"  : unit-gain layers everywhere, every tensor random (the stress scheme: O(1) eps output,
                     errors are not damped by shrinking activations);
    scheme "torch" : the weight scheme SURVEY.md 8d prescribes for measurements (see
                     ``fill_tensor_torch_default``)."""
    if scheme == "unit":
        return {k: fill_tensor(k, s, seed, gain) for k, s in shapes.items()}
    assert scheme == "torch"
    _FAN_IN.clear()
    for k, s in shapes.items():
        if len(s) > 1:
            _FAN_IN[k.rsplit(".", 1)[0]] = int(np.prod(tuple(s)[1:]))
    return {k: fill_tensor_torch_default(k, s, seed) for k, s in shapes.items()}


def checksum(sd: dict) -> float:
    """Order-independent float64 checksum, stored in golden fixtures to detect drift."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(v.sum()) + 0.5 * float((v * v).sum())
    return tot
