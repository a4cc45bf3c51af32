"""Mirror of the hot-path output/input modules of ``schnetpack.atomistic`` on the B200 kernels:
``PairwiseDistances`` (atomistic/distances.py:9-26), ``Atomwise`` (atomistic/atomwise.py:14-88) and ``Forces``
(atomistic/response.py:18-92)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import grad

from .. import functional as K
from .. import nn as snn
from .. import ops
from .. import properties

__all__ = ["PairwiseDistances", "Atomwise", "Forces", "Strain"]


class PairwiseDistances(nn.Module):
    """Rij = R[idx_j] - R[idx_i] + offsets (distances.py:14-26).  The backward assembles dE/dR per atom from the
    receiver- and sender-grouped edge views (deterministic; the reference uses index_put atomics)."""

    @ops.on_tensor_device
    def forward(self, inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if self.training and torch.is_grad_enabled():     # training graph: differentiable twice (functional_torch)
            from .. import functional_torch as T

            return T.pairwise(inputs)
        R = inputs[properties.R]
        offsets = inputs[properties.offsets]
        idx_i = inputs[properties.idx_i].long()
        idx_j = inputs[properties.idx_j].long()
        graph = ops.get_graph(inputs[properties.idx_i] if inputs[properties.idx_i].dtype == torch.int64 else idx_i,
                              inputs[properties.idx_j] if inputs[properties.idx_j].dtype == torch.int64 else idx_j,
                              R.shape[0])
        inputs[properties.Rij] = K.PairwiseDistancesFunction.apply(
            R, offsets, dict(idx_i=idx_i.contiguous(), idx_j=idx_j.contiguous(), graph=graph))
        return inputs


class Atomwise(nn.Module):
    """Atom-wise MLP + per-system sum (atomwise.py:14-88).  The fused kernel path covers the reference default
    ``n_out=1, n_layers=2`` head with ``aggregation_mode in {"sum","avg",None}``."""

    def __init__(self, n_in: int, n_out: int = 1, n_hidden: Optional[Union[int, Sequence[int]]] = None,
                 n_layers: int = 2, activation: Callable = F.silu, aggregation_mode: str = "sum",
                 output_key: str = "y", per_atom_output_key: Optional[str] = None):
        super().__init__()
        self.output_key = output_key
        self.model_outputs = [output_key]
        self.per_atom_output_key = per_atom_output_key
        if self.per_atom_output_key is not None:
            self.model_outputs.append(self.per_atom_output_key)
        self.n_out = n_out
        if aggregation_mode is None and self.per_atom_output_key is None:
            raise ValueError("If `aggregation_mode` is None, `per_atom_output_key` needs to be set,"
                             + " since no accumulated output will be returned!")
        self.outnet = snn.build_mlp(n_in=n_in, n_out=n_out, n_hidden=n_hidden, n_layers=n_layers,
                                    activation=activation)
        self.aggregation_mode = aggregation_mode
        self.activation = activation
        self._pk, self._sig = None, None

    def _pack(self):
        params = list(self.outnet.parameters())
        sig = K.ParamPack.signature(params)
        if self._sig != sig:
            l0, l1 = self.outnet[0], self.outnet[1]
            self._pk = dict(l0=ops.Lin(l0.weight, l0.bias),
                            w1=l1.weight.detach().reshape(-1).contiguous(),
                            b1=l1.bias.detach().contiguous() if l1.bias is not None else None)
            self._sig = sig
        return self._pk

    @ops.on_tensor_device
    def forward(self, inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if snn.use_training_path(self):           # training: differentiable ATen path (functional_torch), SURVEY 8 f3
            from .. import functional_torch as T

            return T.atomwise(self, inputs)
        if self.n_out != 1 or len(self.outnet) != 2:
            raise NotImplementedError("schnetpack_b200.Atomwise: fused head covers n_out=1, n_layers=2")
        q = inputs["scalar_representation"]
        idx_m, n_mol = None, 0
        if self.aggregation_mode is not None:
            idx_m = inputs[properties.idx_m]
            if properties.n_atoms in inputs:      # number of systems without the reference's host sync
                n_mol = int(inputs[properties.n_atoms].shape[0])
            else:                                 # atomwise.py:80 (forces a device->host read)
                n_mol = int(idx_m[-1]) + 1
        y, e = K.AtomwiseFunction.apply(q, dict(pack=self._pack(), idx_m=idx_m, n_mol=n_mol,
                                                act=snn.activation_code(self.activation)))
        if self.per_atom_output_key is not None:
            inputs[self.per_atom_output_key] = y.unsqueeze(-1)
        if self.aggregation_mode is not None:
            if self.aggregation_mode == "avg":
                e = e / inputs[properties.n_atoms]
            inputs[self.output_key] = e
        else:
            inputs[self.output_key] = y.unsqueeze(-1)
        return inputs


class Strain(nn.Module):
    """Adds a zero strain leaf and applies (1 + strain^T) to cell, positions and offsets so that ``Forces(calc_stress=True)``
    can differentiate the energy with respect to it (response.py:434-464).  Plain tensor algebra on the device; the
    derivative reaches the strain through the positions / offsets gradients of ``PairwiseDistances``."""

    @ops.on_tensor_device
    def forward(self, inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        cell = inputs[properties.cell].reshape(-1, 3, 3)
        eps = torch.zeros_like(cell).requires_grad_()
        inputs[properties.strain] = eps
        eps_t = eps.transpose(1, 2)
        inputs[properties.cell] = cell + cell @ eps_t
        per_atom = eps_t[inputs[properties.idx_m]]
        R = inputs[properties.R]
        inputs[properties.R] = R + torch.einsum("na,nab->nb", R, per_atom)
        off = inputs[properties.offsets]
        inputs[properties.offsets] = off + torch.einsum("ea,eab->eb", off, per_atom[inputs[properties.idx_i]])
        return inputs


class Forces(nn.Module):
    """forces = -dE/dR, stress = dE/dstrain / V via autograd over the kernel pipeline (response.py:18-92)."""

    def __init__(self, calc_forces: bool = True, calc_stress: bool = False, energy_key: str = properties.energy,
                 force_key: str = properties.forces, stress_key: str = properties.stress):
        super().__init__()
        self.calc_forces = calc_forces
        self.calc_stress = calc_stress
        self.energy_key = energy_key
        self.force_key = force_key
        self.stress_key = stress_key
        self.model_outputs = []
        if calc_forces:
            self.model_outputs.append(force_key)
        if calc_stress:
            self.model_outputs.append(stress_key)
        self.required_derivatives = []
        if self.calc_forces:
            self.required_derivatives.append(properties.R)
        if self.calc_stress:
            self.required_derivatives.append(properties.strain)

    @ops.on_tensor_device
    def forward(self, inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        Epred = inputs[self.energy_key]
        go: List[Optional[torch.Tensor]] = [torch.ones_like(Epred)]
        grads = grad([Epred], [inputs[prop] for prop in self.required_derivatives], grad_outputs=go,
                     create_graph=self.training)
        if self.calc_forces:
            dEdR = grads[0]
            if dEdR is None:
                dEdR = torch.zeros_like(inputs[properties.R])
            inputs[self.force_key] = -dEdR
        if self.calc_stress:
            stress = grads[-1]
            if stress is None:
                stress = torch.zeros_like(inputs[properties.cell])
            cell = inputs[properties.cell].reshape(-1, 3, 3)
            volume = torch.sum(cell[:, 0, :] * torch.cross(cell[:, 1, :], cell[:, 2, :], dim=1), dim=1,
                               keepdim=True)[:, :, None]
            inputs[self.stress_key] = stress / volume
        return inputs
