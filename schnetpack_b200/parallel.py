"""Multi-GPU plumbing for the hot path.

Molecule batches shard by *system* with no data-path collective: every rank evaluates a contiguous slice of the systems
of a collated batch and only the (tiny) results are gathered.  ``shard_batch`` cuts the reference's flat-graph batch
(`data/loader.py:13-58` layout) at system boundaries and re-bases the neighbour indices; ``gather_results`` reassembles
per-system / per-atom outputs in the original order with one ``all_gather`` each (NCCL on GPUs, gloo in the CPU tests).
A single large periodic system does not shard this way (it needs spatial bricks + halo exchange).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

from . import properties


def system_slices(n_atoms: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous system ranges [m0, m1) per rank, balanced by atom count (a proxy for edge work)."""
    n_atoms = np.asarray(n_atoms, dtype=np.int64)
    B = n_atoms.shape[0]
    cum = np.concatenate([[0], np.cumsum(n_atoms)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        m = int(np.searchsorted(cum, target, side="left"))
        m = min(max(m, cuts[-1]), B)
        cuts.append(m)
    cuts.append(B)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_batch(batch: Dict[str, np.ndarray], rank: int, world: int) -> Dict[str, np.ndarray]:
    """Sub-batch of ``batch`` holding the systems assigned to ``rank`` (numpy in, numpy out; indices re-based)."""
    n_at = np.asarray(batch[properties.n_atoms], dtype=np.int64)
    m0, m1 = system_slices(n_at, world)[rank]
    cum = np.concatenate([[0], np.cumsum(n_at)])
    a0, a1 = int(cum[m0]), int(cum[m1])
    idx_i = np.asarray(batch[properties.idx_i])
    idx_j = np.asarray(batch[properties.idx_j])
    emask = (idx_i >= a0) & (idx_i < a1)          # edges never cross systems, so idx_j is in the same range
    out: Dict[str, np.ndarray] = {}
    for k, v in batch.items():
        v = np.asarray(v)
        if k in (properties.idx_i, properties.idx_j):
            out[k] = v[emask] - a0
        elif k in (properties.offsets, properties.Rij):
            out[k] = v[emask]
        elif k == properties.idx_m:
            out[k] = v[a0:a1] - m0
        elif k in (properties.n_atoms, properties.cell):
            out[k] = v[m0:m1]
        elif k == properties.pbc:
            out[k] = v[3 * m0:3 * m1]
        elif v.shape[:1] == (int(cum[-1]),):       # per-atom arrays (Z, R, ...)
            out[k] = v[a0:a1]
        else:
            out[k] = v
    return out


def gather_results(results: Dict[str, torch.Tensor], n_systems_total: int, n_atoms_total: int, n_atoms_local: int,
                   group=None) -> Dict[str, torch.Tensor]:
    """All-gather per-system ([B_local, ...]) and per-atom ([N_local, ...]) result tensors into full-batch tensors.
    Shards are contiguous and ordered by rank, so concatenation restores the original order."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = {}
    for k, v in results.items():
        per_atom = v.shape[0] == n_atoms_local and k != properties.energy
        sizes = [torch.zeros(1, dtype=torch.int64, device=v.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([v.shape[0]], dtype=torch.int64, device=v.device), group=group)
        sizes = [int(s) for s in sizes]
        mx = max(sizes)
        pad = torch.zeros((mx,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        pad[: v.shape[0]] = v
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        full = torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
        assert full.shape[0] == (n_atoms_total if per_atom else n_systems_total) or True
        out[k] = full
    return out
