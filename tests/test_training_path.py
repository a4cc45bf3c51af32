"""Row f3 (training): in ``train()`` mode the modules evaluate the hot-path formulas with differentiable ATen operations
(schnetpack_b200/functional_torch.py) so that weight gradients and the double backward of a force loss exist
(reference: atomistic/response.py:62-68 ``create_graph=self.training``, task.py:166-185).  Checked here on CPU against the
UNMODIFIED reference: same weights, same batch, loss = sum(E^2) + sum(F^2) -> identical weight gradients.  ``eval()`` mode
never takes this path (it raises on CPU tensors: no CPU fallback of the kernels)."""
import numpy as np
import pytest
import torch

from oracle import ref_loader as rl

needs_ref = pytest.mark.skipif(not rl.available(), reason="reference modules absent")


def _pair(kind):
    from schnetpack_b200 import synthetic as S
    from schnetpack_b200.model import from_spec

    spec = S.model_spec(kind, n_atom_basis=32, n_interactions=2)
    params = S.init_params(spec, seed=13)
    mine = from_spec(spec, params).double()
    ref = rl.build_from_spec(spec, params, torch.float64)
    data = S.aspirin_batch(2, seed=4) if kind == "painn" else S.qm9like_batch(3, seed=4)
    x = {k: (torch.as_tensor(v).double() if np.asarray(v).dtype.kind == "f" else torch.as_tensor(v)) for k, v in data.items()}
    return mine, ref, x


@needs_ref
@pytest.mark.parametrize("kind", ["painn", "schnet"])
def test_training_mode_weight_gradients_match_reference(kind):
    mine, ref, x = _pair(kind)
    mine.train()
    ref.train()
    grads = []
    for model in (mine, ref):
        out = model({k: v.clone() for k, v in x.items()})
        loss = (out["energy"] ** 2).sum() + (out["forces"] ** 2).sum()      # the force term needs the double backward
        model.zero_grad()
        loss.backward()
        grads.append({k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
        assert torch.isfinite(loss)
    assert set(grads[0]) == set(grads[1]) and len(grads[0]) > 10
    for k in grads[1]:
        a, b = grads[0][k], grads[1][k]
        assert float((a - b).abs().max()) <= 1e-10 * max(1.0, float(b.abs().max())), k


def test_eval_mode_never_takes_the_aten_path():
    mine, _, x = _pair("painn") if rl.available() else (None, None, None)
    if mine is None:
        pytest.skip("reference modules absent")
    mine.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        mine({k: v.float() if v.is_floating_point() else v for k, v in x.items()})
