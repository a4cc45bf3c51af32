"""ORACLE -- test infrastructure, NOT product code.

CPU restatement (plain torch tensor algebra, fp32 or fp64) of the reference's message-passing hot path, written
as pure functions over a flat weight dictionary whose keys are the reference ``state_dict`` keys.  It exists only
to *check* the CUDA path: nothing under ``schnetpack_b200/`` imports it; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may.

Parity pin: this oracle is pinned against the UNMODIFIED reference modules imported from
``/root/reference/src/schnetpack`` (see ``tests/golden/make_golden.py`` -> fixtures under ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py``), including the reference's own known-answer vectors for GaussianRBF
(``tests/nn/test_radial.py:6-74``), CosineCutoff (``tests/nn/test_cutoff.py:7-22``) and shifted_softplus
(``tests/nn/test_activations.py:7-25``), and the shipped trained PaiNN model ``tests/testdata/md_ethanol.model``.

Every function cites the reference lines it follows (paths relative to /root/reference/src/schnetpack).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F_

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------------- nn primitives
def gaussian_rbf(d: Tensor, offsets: Tensor, widths: Tensor) -> Tensor:
    """nn/radial.py:11-15 -- exp(-0.5/w^2 (d-mu_k)^2), output [..., n_rbf]."""
    coeff = -0.5 / widths**2
    diff = d[..., None] - offsets
    return torch.exp(coeff * diff**2)


def bessel_rbf(d: Tensor, freqs: Tensor) -> Tensor:
    """nn/radial.py:105-110 -- sin(k pi d / rc) / d with d==0 -> divide by 1."""
    ax = d[..., None] * freqs
    norm = torch.where(d == 0, torch.ones_like(d), d)
    return torch.sin(ax) / norm[..., None]


def cosine_cutoff(d: Tensor, cutoff: float) -> Tensor:
    """nn/cutoff.py:14-33 -- 0.5 (cos(pi d / rc) + 1) * [d < rc]."""
    c = 0.5 * (torch.cos(d * math.pi / cutoff) + 1.0)
    return c * (d < cutoff).to(d.dtype)


def shifted_softplus(x: Tensor) -> Tensor:
    """nn/activations.py:9-22 -- softplus(x) - ln 2 (torch softplus, threshold 20)."""
    return F_.softplus(x) - math.log(2.0)


def dense(x: Tensor, w: Tensor, b: Optional[Tensor] = None, act=None) -> Tensor:
    """nn/base.py:52-55 -- activation(x W^T + b)."""
    y = x @ w.t()
    if b is not None:
        y = y + b
    return act(y) if act is not None else y


def scatter_add(x: Tensor, idx: Tensor, dim_size: int) -> Tensor:
    """nn/scatter.py:26-34 -- zeros(dim_size, ...).index_add(0, idx, x)."""
    out = torch.zeros((dim_size,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    return out.index_add(0, idx, x)


def pairwise_distances(R: Tensor, idx_i: Tensor, idx_j: Tensor, offsets: Tensor) -> Tensor:
    """atomistic/distances.py:14-26 -- Rij = R[idx_j] - R[idx_i] + offsets."""
    return R[idx_j] - R[idx_i] + offsets


def radial_basis(spec: dict, p: Dict[str, Tensor], d: Tensor) -> Tensor:
    pre = "representation.radial_basis."
    if spec["rbf"] == "gaussian":
        return gaussian_rbf(d, p[pre + "offsets"], p[pre + "widths"])
    return bessel_rbf(d, p[pre + "freqs"])


# ----------------------------------------------------------------------------------------------------- representations
def schnet_representation(spec: dict, p: Dict[str, Tensor], Z: Tensor, r_ij: Tensor, idx_i: Tensor,
                          idx_j: Tensor) -> Tensor:
    """representation/schnet.py:147-173 (SchNet.forward) with SchNetInteraction.forward :41-70 inlined."""
    n = Z.shape[0]
    d = torch.linalg.norm(r_ij, dim=1)                                   # :156
    f_ij = radial_basis(spec, p, d)                                      # :157
    rcut = cosine_cutoff(d, spec["cutoff"])                              # :158
    x = p["representation.embedding.weight"][Z]                          # :161
    for t in range(spec["n_interactions"]):
        tt = 0 if spec["shared_interactions"] else t
        b = f"representation.interactions.{tt}."
        h = dense(x, p[b + "in2f.weight"])                               # :60  (no bias)
        w = dense(f_ij, p[b + "filter_network.0.weight"], p[b + "filter_network.0.bias"], shifted_softplus)
        w = dense(w, p[b + "filter_network.1.weight"], p[b + "filter_network.1.bias"])    # :61
        w = w * rcut[:, None]                                            # :62
        m = scatter_add(h[idx_j] * w, idx_i, n)                          # :65-67
        v = dense(m, p[b + "f2out.0.weight"], p[b + "f2out.0.bias"], shifted_softplus)
        v = dense(v, p[b + "f2out.1.weight"], p[b + "f2out.1.bias"])     # :69
        x = x + v                                                        # :168
    return x


def painn_representation(spec: dict, p: Dict[str, Tensor], Z: Tensor, r_ij: Tensor, idx_i: Tensor,
                         idx_j: Tensor) -> Tuple[Tensor, Tensor]:
    """representation/painn.py:207-256 (PaiNN.forward); PaiNNInteraction.forward :31-67 and
    PaiNNMixing.forward :92-117 inlined.  Returns (scalar_representation [N,F], vector_representation [N,3,F])."""
    n = Z.shape[0]
    Fd = spec["n_atom_basis"]
    d = torch.linalg.norm(r_ij, dim=1, keepdim=True)                     # :227  [E,1]
    u = r_ij / d                                                         # :228  [E,3]
    phi = radial_basis(spec, p, d)                                       # :229  [E,1,R]
    fcut = cosine_cutoff(d, spec["cutoff"])                              # :230  [E,1]
    filters = dense(phi, p["representation.filter_net.weight"], p["representation.filter_net.bias"])
    filters = filters * fcut[..., None]                                  # :232  [E,1,T*3F]
    q = p["representation.embedding.weight"][Z][:, None, :]             # :239,:242  [N,1,F]
    mu = torch.zeros((n, 3, Fd), dtype=q.dtype, device=q.device)         # :246
    for t in range(spec["n_interactions"]):
        tt = 0 if spec["shared_interactions"] else t
        w_t = filters if spec["shared_filters"] else filters[..., t * 3 * Fd:(t + 1) * 3 * Fd]   # :233-236
        b = f"representation.interactions.{tt}.interatomic_context_net."
        x = dense(q, p[b + "0.weight"], p[b + "0.bias"], F_.silu)
        x = dense(x, p[b + "1.weight"], p[b + "1.bias"])                 # :54   [N,1,3F]
        xj = x[idx_j]                                                    # :55
        muj = mu[idx_j]                                                  # :56
        y = w_t * xj                                                     # :57
        dq, dmuR, dmumu = torch.split(y, Fd, dim=-1)                     # :59
        dq = scatter_add(dq, idx_i, n)                                   # :60
        dmu = dmuR * u[..., None] + dmumu * muj                          # :61   [E,3,F]
        dmu = scatter_add(dmu, idx_i, n)                                 # :62
        q = q + dq                                                       # :64
        mu = mu + dmu                                                    # :65
        m = f"representation.mixing.{tt}."
        mix = dense(mu, p[m + "mu_channel_mix.weight"])                  # :103  [N,3,2F]
        V, W = torch.split(mix, Fd, dim=-1)                              # :104
        Vn = torch.sqrt(torch.sum(V**2, dim=-2, keepdim=True) + spec["epsilon"])   # :105
        ctx = torch.cat([q, Vn], dim=-1)                                 # :107
        s = dense(ctx, p[m + "intraatomic_context_net.0.weight"], p[m + "intraatomic_context_net.0.bias"], F_.silu)
        s = dense(s, p[m + "intraatomic_context_net.1.weight"], p[m + "intraatomic_context_net.1.bias"])   # :108
        dq_i, dmu_i, dqmu_i = torch.split(s, Fd, dim=-1)                 # :110
        dmu_i = dmu_i * W                                                # :111
        dqmu_i = dqmu_i * torch.sum(V * W, dim=1, keepdim=True)          # :113
        q = q + dq_i + dqmu_i                                            # :115
        mu = mu + dmu_i                                                  # :116
    return q.squeeze(1), mu                                              # :250-254


def atomwise(p: Dict[str, Tensor], q: Tensor, idx_m: Tensor, n_mol: int, prefix: str = "output_modules.0.") -> Tensor:
    """atomistic/atomwise.py:69-88 with build_mlp (nn/blocks.py:38-76): Dense(F->F/2, silu) -> Dense(F/2->1), sum
    over idx_m."""
    y = dense(q, p[prefix + "outnet.0.weight"], p[prefix + "outnet.0.bias"], F_.silu)
    y = dense(y, p[prefix + "outnet.1.weight"], p[prefix + "outnet.1.bias"])
    return scatter_add(y, idx_m, n_mol).squeeze(-1)


# ----------------------------------------------------------------------------------------------------- whole model
def to_torch(d: dict, dtype=torch.float32, device=None) -> Dict[str, Tensor]:
    out = {}
    for k, v in d.items():
        t = torch.as_tensor(v)
        if t.is_floating_point():
            t = t.to(dtype)
        if device is not None:
            t = t.to(device)
        out[k] = t
    return out


def apply_strain(x: Dict[str, Tensor]):
    """atomistic/response.py:441-464 (Strain.forward): a zero strain leaf [B,3,3]; cell, positions and offsets are
    multiplied by (1 + strain^T).  Returns (strain leaf, strained cell, positions, offsets)."""
    strain = torch.zeros_like(x["_cell"]).requires_grad_(True)
    st = strain.transpose(1, 2)
    cell = x["_cell"] + torch.matmul(x["_cell"], st)                           # :448-450
    s_i = st[x["_idx_m"]]
    R = x["_positions"] + torch.matmul(x["_positions"][:, None, :], s_i).squeeze(1)      # :455-457
    s_ij = s_i[x["_idx_i"]]
    off = x["_offsets"] + torch.matmul(x["_offsets"][:, None, :], s_ij).squeeze(1)       # :461-463
    return strain, cell, R, off


def energy_forces(spec: dict, params: dict, inputs: dict, dtype=torch.float32, need_repr: bool = False, device=None,
                  n_mol: Optional[int] = None, stress: bool = False):
    """model/base.py:174-190 (NeuralNetworkPotential.forward) for the modules
    [PairwiseDistances] -> {SchNet|PaiNN} -> [Atomwise, Forces]; forces = -dE/dR (atomistic/response.py:59-76).

    ``inputs`` uses the reference keys; if ``_Rij`` is present it is used directly (padded neighbour lists) and forces
    are taken w.r.t. it pushed back through R only when ``_positions``/``_offsets`` are given.
    Returns dict(energy [B], forces [N,3] or None, scalar_representation, vector_representation?).
    """
    p = params if device is not None and all(torch.is_tensor(v) for v in params.values()) else to_torch(params, dtype, device)
    x = inputs if device is not None and all(torch.is_tensor(v) for v in inputs.values()) else to_torch(inputs, dtype, device)
    Z, idx_i, idx_j, idx_m = x["_atomic_numbers"], x["_idx_i"], x["_idx_j"], x["_idx_m"]
    if n_mol is None:
        n_mol = int(idx_m[-1]) + 1                                       # atomwise.py:80 (host sync, as in the reference)
    want_f = bool(spec.get("forces", True))
    direct_rij = "_Rij" in x
    strain = cell_s = None
    if direct_rij:
        r_ij = x["_Rij"].clone().requires_grad_(want_f)
        R = None
    elif stress:
        x = dict(x)
        x["_positions"] = x["_positions"].clone().requires_grad_(True)   # base.py:105-111
        x["_cell"] = x["_cell"].reshape(-1, 3, 3)
        strain, cell_s, R, off_s = apply_strain(x)                       # input module Strain before PairwiseDistances
        r_ij = pairwise_distances(R, idx_i, idx_j, off_s)
    else:
        R = x["_positions"].clone().requires_grad_(want_f)               # base.py:105-111
        r_ij = pairwise_distances(R, idx_i, idx_j, x["_offsets"])
    out = {}
    if spec["kind"] == "painn":
        q, mu = painn_representation(spec, p, Z, r_ij, idx_i, idx_j)
        out["vector_representation"] = mu.detach()
    else:
        q = schnet_representation(spec, p, Z, r_ij, idx_i, idx_j)
    out["scalar_representation"] = q.detach()
    e = atomwise(p, q, idx_m, n_mol)
    out["energy"] = e.detach()
    if stress:
        g, gs = torch.autograd.grad([e], [R, strain], grad_outputs=[torch.ones_like(e)])      # response.py:62-68
        out["forces"] = -g.detach()                                                           # wrt the strained positions
        vol = torch.sum(cell_s[:, 0, :] * torch.cross(cell_s[:, 1, :], cell_s[:, 2, :], dim=1), dim=1, keepdim=True)[:, :, None]
        out["stress"] = (gs / vol).detach()                                                   # :78-90
    elif want_f:
        wrt = r_ij if direct_rij else R
        (g,) = torch.autograd.grad([e], [wrt], grad_outputs=[torch.ones_like(e)])   # response.py:62-68
        out["dEdRij" if direct_rij else "forces"] = g.detach() if direct_rij else -g.detach()
    return out
