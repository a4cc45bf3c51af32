"""ORACLE -- test infrastructure, NOT product code: CPU restatement (numpy, float64) of the reference's neighbour-list
semantics for the "next" row f1 of SURVEY.md section 8.

The reference's own torch implementation is ``transform/neighborlist.py:428-553`` (``TorchNeighborList``); its ASE /
matscipy / vesin front ends (``:213-286``) return the same SET of pairs.  Contract restated here:

  * a pair (i, j, S) is listed iff  | R[j] - R[i] + S @ cell | < cutoff  (strict, ``:500-501``), for integer image
    vectors S with |S_a| <= ceil(cutoff * |row a of inverse(cell).T|) on periodic axes and S_a = 0 elsewhere
    (``_get_shifts`` ``:515-553``), excluding (i == j, S == 0); self-image pairs i == j with S != 0 are listed;
  * the list is symmetric ((j, i, -S) is listed with (i, j, S), ``:444-456``), ordered by idx_i (``:450``);
  * offsets = S @ cell (``:457``).

Parity pin: ``tests/golden/make_golden_nl.py`` runs the UNMODIFIED reference class and stores its output in canonical
(i, j, S) order under ``tests/golden/nl_*.npz``; ``tests/test_oracle_golden.py`` holds this oracle to those fixtures.
"""
from __future__ import annotations

import numpy as np


def canonical(idx_i, idx_j, shifts):
    """Sort a pair list by (i, j, Sx, Sy, Sz) -- the comparison order of the reference's tests
    (tests/data/test_transforms.py:53-104 sort both lists before comparing)."""
    idx_i, idx_j, shifts = np.asarray(idx_i), np.asarray(idx_j), np.asarray(shifts).reshape(-1, 3)
    order = np.lexsort((shifts[:, 2], shifts[:, 1], shifts[:, 0], idx_j, idx_i))
    return idx_i[order], idx_j[order], shifts[order]


def image_reach(cell, pbc, cutoff):
    """transform/neighborlist.py:529-535 -- images needed per axis."""
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    pbc = np.asarray(pbc).astype(bool).reshape(3)
    if not pbc.any():
        return np.zeros(3, dtype=np.int64)
    inv_len = np.linalg.norm(np.linalg.inv(cell).T, axis=1)
    return np.where(pbc, np.ceil(cutoff * inv_len).astype(np.int64), 0)


def neighbor_list(positions, cell, pbc, cutoff):
    """One structure.  Returns idx_i, idx_j (int64), shifts S [E,3] (int64) in canonical order and offsets = S @ cell."""
    R = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    n = R.shape[0]
    reach = image_reach(cell, pbc, cutoff)
    out_i, out_j, out_s = [], [], []
    for sx in range(-reach[0], reach[0] + 1):
        for sy in range(-reach[1], reach[1] + 1):
            for sz in range(-reach[2], reach[2] + 1):
                S = np.array([sx, sy, sz], dtype=np.int64)
                d = R[None, :, :] + (S.astype(np.float64) @ cell)[None, None, :] - R[:, None, :]      # [i, j]
                mask = np.linalg.norm(d, axis=-1) < cutoff
                if not S.any():
                    mask &= ~np.eye(n, dtype=bool)
                ii, jj = np.nonzero(mask)
                out_i.append(ii)
                out_j.append(jj)
                out_s.append(np.broadcast_to(S, (ii.shape[0], 3)))
    ii = np.concatenate(out_i) if out_i else np.zeros(0, dtype=np.int64)
    jj = np.concatenate(out_j) if out_j else np.zeros(0, dtype=np.int64)
    ss = np.concatenate(out_s) if out_s else np.zeros((0, 3), dtype=np.int64)
    ii, jj, ss = canonical(ii.astype(np.int64), jj.astype(np.int64), ss.astype(np.int64))
    return ii, jj, ss, ss.astype(np.float64) @ cell


def batch_neighbor_list(positions, cells, pbcs, n_atoms, cutoff):
    """Collated batch (data/loader.py:35-46: indices of system b are shifted by the atoms before it)."""
    out_i, out_j, out_s, out_o = [], [], [], []
    start = 0
    for b, na in enumerate(np.asarray(n_atoms).tolist()):
        ii, jj, ss, oo = neighbor_list(np.asarray(positions)[start:start + na], np.asarray(cells).reshape(-1, 3, 3)[b],
                                       np.asarray(pbcs).reshape(-1, 3)[b], cutoff)
        out_i.append(ii + start)
        out_j.append(jj + start)
        out_s.append(ss)
        out_o.append(oo)
        start += na
    return (np.concatenate(out_i), np.concatenate(out_j), np.concatenate(out_s).reshape(-1, 3),
            np.concatenate(out_o).reshape(-1, 3))
