"""ORACLE -- test infrastructure, NOT product code: PaiNN energy + forces of ONE large system evaluated by several ranks
with the graph partition and halo exchange of ``schnetpack_b200.parallel``, written with the
primitives of ``oracle/spk_oracle.py`` (same reference line citations).  It exists to check the N > 1 host logic on CPU
(gloo): the per-rank engine of the product is the CUDA path, which plugs into the same plan / exchange."""
from __future__ import annotations

import torch
import torch.nn.functional as F_

from oracle import spk_oracle as O
from schnetpack_b200.parallel import HaloExchange, RankPlan


def painn_energy_forces(spec: dict, params: dict, data: dict, plan: RankPlan, dtype=torch.float64, group=None):
    """Returns (local partial energies [n_systems], forces of the OWNED atoms [n_owned, 3]).  The caller all-reduces the
    energies and concatenates the force blocks by ``plan.owned``."""
    p = O.to_torch(params, dtype)
    x_in = O.to_torch(data, dtype)
    own = torch.as_tensor(plan.owned)
    gho = torch.as_tensor(plan.ghosts)
    eid = torch.as_tensor(plan.edge_ids)
    i_loc, j_loc = torch.as_tensor(plan.idx_i), torch.as_tensor(plan.idx_j)
    n_o = plan.n_owned
    Fd = spec["n_atom_basis"]
    Z_own = x_in["_atomic_numbers"][own]
    R_own = x_in["_positions"][own].clone().requires_grad_(True)                  # model/base.py:105-111
    R_loc = torch.cat([R_own, HaloExchange.apply(R_own, plan, group)], dim=0)     # ghost positions, autograd-aware
    r_ij = R_loc[j_loc] - R_loc[i_loc] + x_in["_offsets"][eid]                    # atomistic/distances.py:14-26
    d = torch.linalg.norm(r_ij, dim=1, keepdim=True)                              # painn.py:227
    u = r_ij / d                                                                  # :228
    phi = O.radial_basis(spec, p, d)                                              # :229
    fcut = O.cosine_cutoff(d, spec["cutoff"])                                     # :230
    filters = O.dense(phi, p["representation.filter_net.weight"], p["representation.filter_net.bias"]) * fcut[..., None]
    q = p["representation.embedding.weight"][Z_own][:, None, :]                   # :239,:242
    mu = torch.zeros((n_o, 3, Fd), dtype=q.dtype)                                 # :246
    for t in range(spec["n_interactions"]):
        tt = 0 if spec["shared_interactions"] else t
        w_t = filters if spec["shared_filters"] else filters[..., t * 3 * Fd:(t + 1) * 3 * Fd]
        b = f"representation.interactions.{tt}.interatomic_context_net."
        x = O.dense(O.dense(q, p[b + "0.weight"], p[b + "0.bias"], F_.silu), p[b + "1.weight"], p[b + "1.bias"])   # :54
        x_loc = torch.cat([x, HaloExchange.apply(x, plan, group)], dim=0)         # senders' rows incl. ghosts
        mu_loc = torch.cat([mu, HaloExchange.apply(mu, plan, group)], dim=0)
        y = w_t * x_loc[j_loc]                                                    # :55,:57
        dq, dmuR, dmumu = torch.split(y, Fd, dim=-1)                              # :59
        dq = O.scatter_add(dq, i_loc, n_o)                                        # :60 (receivers are owned)
        dmu = O.scatter_add(dmuR * u[..., None] + dmumu * mu_loc[j_loc], i_loc, n_o)     # :61-62
        q = q + dq                                                                # :64
        mu = mu + dmu                                                             # :65
        m = f"representation.mixing.{tt}."
        mix = O.dense(mu, p[m + "mu_channel_mix.weight"])                         # :103
        V, W = torch.split(mix, Fd, dim=-1)
        Vn = torch.sqrt(torch.sum(V**2, dim=-2, keepdim=True) + spec["epsilon"])  # :105
        s = O.dense(O.dense(torch.cat([q, Vn], dim=-1), p[m + "intraatomic_context_net.0.weight"],
                            p[m + "intraatomic_context_net.0.bias"], F_.silu),
                    p[m + "intraatomic_context_net.1.weight"], p[m + "intraatomic_context_net.1.bias"])       # :108
        dq_i, dmu_i, dqmu_i = torch.split(s, Fd, dim=-1)
        q = q + dq_i + dqmu_i * torch.sum(V * W, dim=1, keepdim=True)             # :113-115
        mu = mu + dmu_i * W                                                       # :111,:116
    n_sys = int(x_in["_n_atoms"].shape[0])
    e_part = O.atomwise(p, q.squeeze(1), x_in["_idx_m"][own], n_sys)              # atomwise.py:69-88 on the owned atoms
    (g,) = torch.autograd.grad([e_part.sum()], [R_own])                           # response.py:62-68; halo grads inside
    return e_part.detach(), -g.detach()
