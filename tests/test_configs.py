"""CPU: the shipped Hydra config files (schnetpack_b200/configs, drop-ins for the reference's
configs/model/{nnp.yaml,representation/{painn,schnet}.yaml}) name importable ``_target_``s whose constructors accept exactly
the keys the files give -- checked with a minimal re-implementation of ``hydra.utils.instantiate`` (hydra is not installed
here): ``_target_`` -> import + call, ``defaults`` -> sub-file merge, ``${globals.cutoff}`` -> value."""
import importlib
import os

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "schnetpack_b200", "configs", "model")


def _load(path, base):
    d = yaml.safe_load(open(path))
    for item in d.pop("defaults", []):
        (group, name), = item.items()
        d[group] = _load(os.path.join(base, group, name + ".yaml"), os.path.join(base, group))
    return d


def _instantiate(node, env):
    if isinstance(node, dict):
        kw = {k: _instantiate(v, env) for k, v in node.items() if k != "_target_"}
        if "_target_" not in node:
            return kw
        mod, _, attr = node["_target_"].rpartition(".")
        return getattr(importlib.import_module(mod), attr)(**kw)
    if isinstance(node, list):
        return [_instantiate(v, env) for v in node]
    if isinstance(node, str) and node.startswith("${"):
        return env[node[2:-1]]
    return node


def test_representation_configs_instantiate():
    import schnetpack_b200 as sb

    env = {"globals.cutoff": 5.0}
    for name, cls, n_int in (("painn_b200", sb.representation.PaiNN, 3), ("schnet_b200", sb.representation.SchNet, 6)):
        rep = _instantiate(_load(os.path.join(CFG, "representation", name + ".yaml"), os.path.join(CFG, "representation")), env)
        assert isinstance(rep, cls) and rep.n_atom_basis == 128 and len(rep.interactions) == n_int
        assert rep.radial_basis.n_rbf == 20 and float(rep.cutoff) == 5.0


def test_nnp_config_names_b200_modules():
    d = _load(os.path.join(CFG, "nnp_b200.yaml"), CFG)
    assert d["_target_"] == "schnetpack.model.NeuralNetworkPotential"          # the reference's own orchestration class
    env = {"globals.cutoff": 5.0, "model.representation.n_atom_basis": 128}
    ins = _instantiate(d["input_modules"], env)
    outs = _instantiate(d["output_modules"], env)
    rep = _instantiate(d["representation"], env)
    from schnetpack_b200 import atomistic
    from schnetpack_b200.model import NeuralNetworkPotential

    assert isinstance(ins[0], atomistic.PairwiseDistances) and isinstance(outs[1], atomistic.Forces)
    model = NeuralNetworkPotential(rep, ins, outs)                               # the mirror accepts the same arguments
    assert sorted(model.model_outputs) == ["energy", "forces"] and model.required_derivatives == ["_positions"]
