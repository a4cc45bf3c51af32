"""PaiNN representation on the B200 kernels -- drop-in for ``schnetpack.representation.PaiNN``.

Constructor signature, attributes, tensor-dict protocol and ``state_dict`` keys follow
/root/reference/src/schnetpack/representation/painn.py:120-256; the arithmetic is the fused kernel pipeline of
``schnetpack_b200.functional.painn_forward`` (no ``[E,3F]`` temporaries, filter recomputed per edge in-kernel).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as K
from .. import nn as snn
from .. import ops
from .. import properties

__all__ = ["PaiNN", "PaiNNInteraction", "PaiNNMixing"]


class PaiNNInteraction(nn.Module):
    """Parameter container of one interaction block (painn.py:14-67): interatomic_context_net = Dense(F,F,act) ->
    Dense(F,3F).  The block is executed by the fused edge kernel inside ``PaiNN.forward``."""

    def __init__(self, n_atom_basis: int, activation: Callable):
        super().__init__()
        self.n_atom_basis = n_atom_basis
        self.interatomic_context_net = nn.Sequential(
            snn.Dense(n_atom_basis, n_atom_basis, activation=activation),
            snn.Dense(n_atom_basis, 3 * n_atom_basis, activation=None),
        )

    def forward(self, q, mu, Wij, dir_ij, idx_i, idx_j, n_atoms: int):
        raise NotImplementedError(
            "PaiNNInteraction is executed inside PaiNN.forward by the fused edge kernel (spk_painn_edge_fwd); "
            "the materialised-filter block-level call of the reference is not provided")


class PaiNNMixing(nn.Module):
    """Parameter container of one mixing block (painn.py:70-117)."""

    def __init__(self, n_atom_basis: int, activation: Callable, epsilon: float = 1e-8):
        super().__init__()
        self.n_atom_basis = n_atom_basis
        self.intraatomic_context_net = nn.Sequential(
            snn.Dense(2 * n_atom_basis, n_atom_basis, activation=activation),
            snn.Dense(n_atom_basis, 3 * n_atom_basis, activation=None),
        )
        self.mu_channel_mix = snn.Dense(n_atom_basis, 2 * n_atom_basis, activation=None, bias=False)
        self.epsilon = epsilon

    def forward(self, q, mu):
        raise NotImplementedError("PaiNNMixing is executed inside PaiNN.forward")


class PaiNN(nn.Module):
    """PaiNN - polarizable interaction neural network (painn.py:120-256), B200 kernel path."""

    def __init__(self, n_atom_basis: int, n_interactions: int, radial_basis: nn.Module,
                 cutoff_fn: Optional[Callable] = None, activation: Optional[Callable] = F.silu,
                 shared_interactions: bool = False, shared_filters: bool = False, epsilon: float = 1e-8,
                 nuclear_embedding: Optional[nn.Module] = None, electronic_embeddings: Optional[List] = None):
        super().__init__()
        self.n_atom_basis = n_atom_basis
        self.n_interactions = n_interactions
        self.cutoff_fn = cutoff_fn
        self.cutoff = cutoff_fn.cutoff
        self.radial_basis = radial_basis
        self.activation = activation

        if nuclear_embedding is None:
            nuclear_embedding = nn.Embedding(100, n_atom_basis)
        self.embedding = nuclear_embedding
        if electronic_embeddings is None:
            electronic_embeddings = []
        self.electronic_embeddings = nn.ModuleList(electronic_embeddings)

        self.share_filters = shared_filters
        if shared_filters:
            self.filter_net = snn.Dense(self.radial_basis.n_rbf, 3 * n_atom_basis, activation=None)
        else:
            self.filter_net = snn.Dense(self.radial_basis.n_rbf, self.n_interactions * n_atom_basis * 3,
                                        activation=None)
        self.interactions = snn.replicate_module(
            lambda: PaiNNInteraction(n_atom_basis=self.n_atom_basis, activation=activation),
            self.n_interactions, shared_interactions)
        self.mixing = snn.replicate_module(
            lambda: PaiNNMixing(n_atom_basis=self.n_atom_basis, activation=activation, epsilon=epsilon),
            self.n_interactions, shared_interactions)
        self._pk = K.PaiNNPack()

    # ---- kernel-side views of the configuration ------------------------------------------------------------------
    @property
    def _act(self):
        return snn.activation_code(self.activation)

    @property
    def _n_rbf(self):
        return self.radial_basis.n_rbf

    @property
    def _rbf_kind(self):
        return self.radial_basis.kind

    @property
    def _rbf_p0(self):
        return self.radial_basis.kernel_params()[0]

    @property
    def _rbf_p1(self):
        return self.radial_basis.kernel_params()[1]

    @property
    def _cutoff_value(self):
        return self.cutoff_fn.value()

    def _pack(self) -> K.PaiNNPack:
        params = [p for p in self.parameters()]
        if self._pk.stale(params):
            self._pk.build(self)
            self._pk.mark(params)
        return self._pk

    def _apply(self, fn, *a, **k):  # .to()/.cuda()/.float(): drop the cached kernel-side weights
        self._pk = K.PaiNNPack()
        return super()._apply(fn, *a, **k)

    def forward(self, inputs: Dict[str, torch.Tensor]):
        atomic_numbers = inputs[properties.Z]
        r_ij = inputs[properties.Rij]
        idx_i = inputs[properties.idx_i]
        idx_j = inputs[properties.idx_j]
        n_atoms = atomic_numbers.shape[0]
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "schnetpack_b200.PaiNN: weight gradients / double backward (training) are not implemented in the "
                "CUDA path (SURVEY.md §8 f3); call model.eval() for inference, forces and MD")
        if not isinstance(self.radial_basis, (snn.GaussianRBF, snn.BesselRBF)) or not isinstance(
                self.cutoff_fn, snn.CosineCutoff):
            raise NotImplementedError("fused PaiNN kernels support GaussianRBF/BesselRBF x CosineCutoff")
        graph = ops.get_graph(idx_i, idx_j, n_atoms)

        if isinstance(self.embedding, nn.Embedding) and len(self.electronic_embeddings) == 0:
            q0 = ops.embedding(self.embedding.weight.detach().contiguous(), atomic_numbers)   # painn.py:239
        else:  # user-supplied embedding modules run as given (painn.py:239-241)
            q0 = self.embedding(atomic_numbers)
            for embedding in self.electronic_embeddings:
                q0 = q0 + embedding(q0, inputs)
            q0 = q0.detach().contiguous()

        # system boundaries (if the batch carries them) let the edge kernels keep a small system's rows in shared memory
        holder = dict(module=self, graph=graph)
        if properties.idx_m in inputs and properties.n_atoms in inputs:
            n_mol = int(inputs[properties.n_atoms].shape[0])
            if n_mol > 0 and n_atoms / n_mol <= ops.SYS_MAX_AVG_ATOMS and ops.EDGE_IMPL == "sys":
                holder["mol_ptr"] = ops.segment_ptr(inputs[properties.idx_m], n_mol)
                holder["n_mol"] = n_mol
        q, mu = K.PaiNNFunction.apply(r_ij if r_ij.is_contiguous() else r_ij.contiguous(), q0, holder)
        inputs["scalar_representation"] = q
        inputs["vector_representation"] = mu
        return inputs
