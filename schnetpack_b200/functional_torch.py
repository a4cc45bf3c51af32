"""TRAINING-MODE path: the hot-path math written with differentiable ATen operations.

The CUDA kernels of this package provide the forward pass and first-order gradients with respect to positions / strain --
everything inference, forces and MD need.  Training needs more: gradients with respect to every weight and, for force
losses, a second differentiation through the force computation (``create_graph=self.training``,
/root/reference/src/schnetpack/atomistic/response.py:62-68).  Those are NOT implemented as kernels.  So that the modules
still drop in under ``spktrain`` (``task.py:166-185``), a module that is in ``train()`` mode with trainable parameters under
grad mode evaluates the SAME formulas below with plain torch operations (autograd-complete, any device, reference speed);
``eval()`` mode -- validation, inference, MD, deployment -- always runs the kernels.  The switch is
``schnetpack_b200.nn.use_training_path`` and is exercised by tests/test_training_path.py against the reference's own weight
gradients.  Formulas and line citations as in oracle/spk_oracle.py (which this package never imports).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F_

from . import properties

Tensor = torch.Tensor


def _act(code_fn, x):
    return x if code_fn is None else code_fn(x)


def ssp(x: Tensor) -> Tensor:
    """nn/activations.py:9-22."""
    return F_.softplus(x) - math.log(2.0)


def _activation(act):
    from . import nn as snn
    from . import ops

    code = snn.activation_code(act)
    return {ops.ACT_NONE: None, ops.ACT_SILU: F_.silu, ops.ACT_SSP: ssp}[code]


def dense(x: Tensor, lin, act=None) -> Tensor:
    """nn/base.py:52-55."""
    y = F_.linear(x, lin.weight, lin.bias)
    return y if act is None else act(y)


def radial(rb, d: Tensor) -> Tensor:
    if type(rb).__name__ == "GaussianRBF":                                   # nn/radial.py:11-15
        coeff = -0.5 / torch.pow(rb.widths, 2)
        return torch.exp(coeff * torch.pow(d[..., None] - rb.offsets, 2))
    ax = d[..., None] * rb.freqs                                             # nn/radial.py:105-110
    norm = torch.where(d == 0, torch.ones_like(d), d)
    return torch.sin(ax) / norm[..., None]


def cutoff(cf, d: Tensor) -> Tensor:
    rc = cf.cutoff
    return 0.5 * (torch.cos(d * math.pi / rc) + 1.0) * (d < rc).to(d.dtype)   # nn/cutoff.py:14-33


def scatter_add(x: Tensor, idx: Tensor, dim_size: int) -> Tensor:
    """nn/scatter.py:26-34."""
    return torch.zeros((dim_size,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device).index_add(0, idx, x)


def pairwise(inputs: Dict[str, Tensor]) -> Dict[str, Tensor]:
    R, off = inputs[properties.R], inputs[properties.offsets]                # atomistic/distances.py:14-26
    inputs[properties.Rij] = R[inputs[properties.idx_j]] - R[inputs[properties.idx_i]] + off
    return inputs


def _embed(mod, inputs):
    x = mod.embedding(inputs[properties.Z])
    for e in mod.electronic_embeddings:
        x = x + e(x, inputs)
    return x


def painn(mod, inputs: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """representation/painn.py:207-256."""
    r_ij, idx_i, idx_j = inputs[properties.Rij], inputs[properties.idx_i], inputs[properties.idx_j]
    n_atoms = inputs[properties.Z].shape[0]
    F = mod.n_atom_basis
    act = _activation(mod.activation)
    d = torch.norm(r_ij, dim=1, keepdim=True)                                # :227
    u = r_ij / d                                                             # :228
    phi = radial(mod.radial_basis, d)                                        # :229
    fc = cutoff(mod.cutoff_fn, d)                                            # :230
    filters = dense(phi, mod.filter_net) * fc[..., None]                     # :232
    flist = [filters] * mod.n_interactions if mod.share_filters else torch.split(filters, 3 * F, dim=-1)
    q = _embed(mod, inputs)[:, None]                                         # :239-242
    mu = torch.zeros((n_atoms, 3, F), dtype=q.dtype, device=q.device)        # :246
    for it, mx, W in zip(mod.interactions, mod.mixing, flist):
        c = it.interatomic_context_net
        x = dense(dense(q, c[0], act), c[1])                                 # :54
        y = W * x[idx_j]                                                     # :55,:57
        dq, dmuR, dmumu = torch.split(y, F, dim=-1)                          # :59
        dq = scatter_add(dq, idx_i, n_atoms)                                 # :60
        dmu = scatter_add(dmuR * u[..., None] + dmumu * mu[idx_j], idx_i, n_atoms)     # :61-62
        q = q + dq
        mu = mu + dmu
        V, Wm = torch.split(dense(mu, mx.mu_channel_mix), F, dim=-1)         # :103-104
        Vn = torch.sqrt(torch.sum(V ** 2, dim=-2, keepdim=True) + mx.epsilon)  # :105
        m = mx.intraatomic_context_net
        s = dense(dense(torch.cat([q, Vn], dim=-1), m[0], act), m[1])        # :107-108
        dq_i, dmu_i, dqmu_i = torch.split(s, F, dim=-1)
        q = q + dq_i + dqmu_i * torch.sum(V * Wm, dim=1, keepdim=True)       # :113-115
        mu = mu + dmu_i * Wm                                                 # :111,:116
    inputs["scalar_representation"] = q.squeeze(1)
    inputs["vector_representation"] = mu
    return inputs


def schnet(mod, inputs: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """representation/schnet.py:147-173."""
    r_ij, idx_i, idx_j = inputs[properties.Rij], inputs[properties.idx_i], inputs[properties.idx_j]
    act = _activation(mod.activation)
    d = torch.norm(r_ij, dim=1)                                              # :156
    f_ij = radial(mod.radial_basis, d)                                       # :157
    rcut = cutoff(mod.cutoff_fn, d)                                          # :158
    x = _embed(mod, inputs)                                                  # :161
    for it in mod.interactions:
        h = dense(x, it.in2f)                                                # :60
        Wij = dense(dense(f_ij, it.filter_network[0], act), it.filter_network[1]) * rcut[:, None]   # :61-62
        m = scatter_add(h[idx_j] * Wij, idx_i, x.shape[0])                   # :65-67
        x = x + dense(dense(m, it.f2out[0], act), it.f2out[1])               # :69, :168
    inputs["scalar_representation"] = x
    return inputs


def atomwise(mod, inputs: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """atomistic/atomwise.py:69-88."""
    act = _activation(mod.activation)
    y = inputs["scalar_representation"]
    n = len(mod.outnet)
    for k, lin in enumerate(mod.outnet):
        y = dense(y, lin, act if k < n - 1 else None)
    if mod.per_atom_output_key is not None:
        inputs[mod.per_atom_output_key] = y
    if mod.aggregation_mode is not None:
        idx_m = inputs[properties.idx_m]
        n_mol = int(inputs[properties.n_atoms].shape[0]) if properties.n_atoms in inputs else int(idx_m[-1]) + 1
        y = scatter_add(y, idx_m, n_mol)
        y = torch.squeeze(y, -1)
        if mod.aggregation_mode == "avg":
            y = y / inputs[properties.n_atoms]
    inputs[mod.output_key] = y
    return inputs
