"""Shard the naturally data-parallel axis (shifts λ, z, energies) over ranks — no exchange step on the data path
(SURVEY §8e): every factorization is independent, so a rank only needs its index set; results are gathered at the end.

Used by bench.py (one rank per GPU under torchrun) and testable on CPU with the gloo backend.
"""
from __future__ import annotations

import numpy as np


def interleaved(total: int, rank: int, world: int) -> np.ndarray:
    """Indices rank, rank+world, ... : round-robin, balances sorted-by-length batches (Bessel z) and grid rows."""
    return np.arange(rank, total, world, dtype=np.int64)


def contiguous(total: int, rank: int, world: int) -> np.ndarray:
    """Contiguous slab of ceil(total/world) indices (equal-cost batches)."""
    per = (total + world - 1) // world
    return np.arange(min(rank * per, total), min((rank + 1) * per, total), dtype=np.int64)


def grid_rows(nx: int, ny_total: int, rank: int, world: int, x0=-10.0, x1=7.0, y0=-7.0, y1=8.0) -> np.ndarray:
    """λ values of the rows `interleaved(ny_total, rank, world)` of the nx x ny_total grid over [x0,x1] x [y0,y1]
    (BASELINE config C3 window, examples/toeplitz.jl:72)."""
    x = np.linspace(x0, x1, nx)
    y = np.linspace(y0, y1, ny_total)[interleaved(ny_total, rank, world)]
    return (x[None, :] + 1j * y[:, None]).reshape(-1)


def scatter_back(parts, index_sets, total: int, dtype=None):
    """Inverse of the sharding: place rank r's results `parts[r]` at `index_sets[r]` of a length-`total` array."""
    out = np.empty(total, dtype=dtype if dtype is not None else np.asarray(parts[0]).dtype)
    for part, idx in zip(parts, index_sets):
        out[idx] = part
    return out


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Device-timed durations are reported as the max over ranks (every multi-GPU number in bench.py)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
