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
        if full.shape[0] != (n_atoms_total if per_atom else n_systems_total):
            raise RuntimeError(f"gather_results: {k} has {full.shape[0]} rows after the gather")
        out[k] = full
    return out


# ----------------------------------------------------------------------------------------------------------------------
# One LARGE system over several ranks: graph partition + halo exchange.
#
# The neighbour list already says who needs whom, so the halo is defined by the GRAPH, not by geometry: a rank owns a slab
# of atoms, evaluates the edges whose RECEIVER it owns, and needs the per-atom rows (positions, x = context(q), mu) of the
# senders it does not own -- its ghosts.  Per interaction block that is one exchange of x (and mu) rows forward and, for
# forces, the reverse exchange of their gradients; ``HaloExchange`` is that exchange as an autograd function over
# torch.distributed point-to-point calls (NCCL on GPUs, gloo in the CPU tests), so the chain rule through positions and
# features of ghost atoms reaches the owner's leaves without any rank seeing the whole system.
# ----------------------------------------------------------------------------------------------------------------------
class RankPlan:
    """What rank ``rank`` of ``world`` needs to evaluate its part of one big graph (all index arrays int64, numpy).

    owned      global ids of the atoms this rank owns (ascending)
    ghosts     global ids of the senders it needs from other ranks, grouped by owner rank (ascending rank, ascending id
               inside a group); local index = len(owned) + position
    edge_ids   positions in the global edge list of the edges with an owned receiver (global order preserved)
    idx_i, idx_j   those edges in LOCAL indices (receivers < len(owned); senders may be ghosts)
    send[p]    local (owned) indices of the rows peer p needs from this rank, in the order p stores them as ghosts
    recv[p]    (start, stop) slice of this rank's ghost block filled by peer p
    """

    def __init__(self, rank, world, owned, ghosts, edge_ids, idx_i, idx_j, send, recv):
        self.rank, self.world = rank, world
        self.owned, self.ghosts, self.edge_ids, self.idx_i, self.idx_j = owned, ghosts, edge_ids, idx_i, idx_j
        self.send, self.recv = send, recv

    @property
    def n_owned(self) -> int:
        return int(self.owned.shape[0])

    @property
    def n_ghost(self) -> int:
        return int(self.ghosts.shape[0])


def slab_owners(positions: np.ndarray, world: int, axis: int = None) -> np.ndarray:
    """Owner rank of every atom: ``world`` slabs of equal atom count along ``axis`` (default: the longest extent).
    Ties are broken by atom index, so every rank computes the same assignment."""
    R = np.asarray(positions, dtype=np.float64)
    if axis is None:
        axis = int(np.argmax(R.max(axis=0) - R.min(axis=0))) if R.shape[0] else 0
    order = np.lexsort((np.arange(R.shape[0]), R[:, axis]))
    owner = np.empty(R.shape[0], dtype=np.int64)
    bounds = [(R.shape[0] * r) // world for r in range(world + 1)]
    for r in range(world):
        owner[order[bounds[r]:bounds[r + 1]]] = r
    return owner


def partition_graph(owner: np.ndarray, idx_i: np.ndarray, idx_j: np.ndarray, rank: int, world: int) -> RankPlan:
    """Plan of ``rank`` for the graph (idx_i = receivers, idx_j = senders) under the atom -> rank map ``owner``."""
    owner = np.asarray(owner, dtype=np.int64)
    idx_i, idx_j = np.asarray(idx_i, dtype=np.int64), np.asarray(idx_j, dtype=np.int64)
    n = owner.shape[0]

    def ghosts_of(r):
        e = np.nonzero(owner[idx_i] == r)[0]
        s = np.unique(idx_j[e])
        return e, s[owner[s] != r]

    owned = np.nonzero(owner == rank)[0]
    edge_ids, ghosts = ghosts_of(rank)
    # ghosts grouped by owner rank (ascending rank, ascending id inside): the block peer p fills is contiguous
    ghosts = ghosts[np.lexsort((ghosts, owner[ghosts]))]
    local = np.full(n, -1, dtype=np.int64)
    local[owned] = np.arange(owned.shape[0])
    local[ghosts] = owned.shape[0] + np.arange(ghosts.shape[0])
    recv, send = {}, {}
    for p in range(world):
        if p == rank:
            continue
        sel = np.nonzero(owner[ghosts] == p)[0]
        if sel.size:
            recv[p] = (int(sel[0]), int(sel[-1]) + 1)
        # what p needs from me: p's ghosts that I own, in p's storage order
        _, gp = ghosts_of(p)
        gp = gp[np.lexsort((gp, owner[gp]))]
        mine = gp[owner[gp] == rank]
        if mine.size:
            send[p] = local[mine]
    return RankPlan(rank, world, owned, ghosts, edge_ids, local[idx_i[edge_ids]], local[idx_j[edge_ids]], send, recv)


class HaloExchange(torch.autograd.Function):
    """rows of the owned atoms [n_owned, ...]  ->  rows of this rank's ghost atoms [n_ghost, ...]  (autograd-aware).

    forward: every rank sends the rows its peers list as ghosts and receives its own ghost rows (isend/irecv);
    backward: the gradients of the ghost rows travel back to their owners and are accumulated onto the rows they came
    from (fixed peer order: deterministic).  All ranks must call it the same number of times in the same order."""

    @staticmethod
    def forward(ctx, rows: torch.Tensor, plan: RankPlan, group=None):
        import torch.distributed as dist

        ctx.plan, ctx.group, ctx.n_owned = plan, group, rows.shape[0]
        tail = tuple(rows.shape[1:])
        ghost = rows.new_zeros((plan.n_ghost,) + tail)
        ops, bufs = [], []
        for p in sorted(set(plan.send) | set(plan.recv)):
            if p in plan.send:
                sb = rows.detach()[torch.as_tensor(plan.send[p], device=rows.device)].contiguous()
                bufs.append(sb)
                ops.append(dist.P2POp(dist.isend, sb, p, group))
            if p in plan.recv:
                a, b = plan.recv[p]
                rb = rows.new_empty((b - a,) + tail)
                bufs.append((rb, a, b))
                ops.append(dist.P2POp(dist.irecv, rb, p, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for item in bufs:
            if isinstance(item, tuple):
                rb, a, b = item
                ghost[a:b] = rb
        return ghost

    @staticmethod
    def backward(ctx, g_ghost: torch.Tensor):
        import torch.distributed as dist

        plan, group = ctx.plan, ctx.group
        tail = tuple(g_ghost.shape[1:])
        g_rows = g_ghost.new_zeros((ctx.n_owned,) + tail)
        ops, recvs, keep = [], [], []
        for p in sorted(set(plan.send) | set(plan.recv)):
            if p in plan.recv:                               # my ghosts came from p: their gradients go back to p
                a, b = plan.recv[p]
                sb = g_ghost[a:b].contiguous()
                keep.append(sb)
                ops.append(dist.P2POp(dist.isend, sb, p, group))
            if p in plan.send:                               # p holds ghosts of my rows: receive their gradients
                rb = g_ghost.new_empty((len(plan.send[p]),) + tail)
                recvs.append((p, rb))
                ops.append(dist.P2POp(dist.irecv, rb, p, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for p, rb in recvs:                                  # ascending peer order: deterministic accumulation
            g_rows.index_add_(0, torch.as_tensor(plan.send[p], device=g_rows.device), rb)
        return g_rows, None, None
