"""ORACLE / test infrastructure, NOT product code: stub-package loader that imports the UNMODIFIED reference hot-path
modules (plain Python + torch), either from the read-only reference tree ``/root/reference`` (build container) or from
``oracle/_ref`` -- a git-ignored, byte-identical copy of exactly the module files this loader imports, made by the committed
recipe ``oracle/make_ref.py`` (run by ``__graft_entry__.build()``), which travels to the GPU box with the snapshot so that the
reference itself is the checker and the timed baseline there (``cpu_baseline.kind = "reference"``).

Used by ``tests/golden/make_golden*.py`` (fixtures), ``tests/test_reference_live.py`` (direct reference-vs-CUDA parity and
``convert_model``), ``bench.py --impl reference`` and the ``cpu_baseline`` / ``eager_gpu_baseline`` legs of ``bench.py``.
Nothing under ``schnetpack_b200/`` imports it.

Recipe follows SURVEY.md Appendix B: the reference's ``schnetpack/__init__.py`` imports ase / lightning /
hydra (absent here), so a bare parent package is registered and only the torch-only sub-modules on the
hot path are imported from the read-only tree.  Nothing is written to the reference tree.
"""
import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
VENDORED = os.path.join(HERE, "_ref")                      # oracle/_ref/{schnetpack/..., testdata/...}, see make_ref.py


def _resolve():
    if os.path.isdir(REF_ROOT + "/src/schnetpack"):
        return REF_ROOT + "/src/schnetpack", REF_ROOT + "/tests/testdata", "tree"
    if os.path.isdir(os.path.join(VENDORED, "schnetpack")):
        return os.path.join(VENDORED, "schnetpack"), os.path.join(VENDORED, "testdata"), "vendored"
    return None, None, None


REF, TESTDATA, SOURCE = _resolve()


def available() -> bool:
    return REF is not None


def _stubpkg(name, path=None, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    if path:
        m.__path__ = [path]
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """Return the stub ``schnetpack`` package with properties/utils/nn/representation/atomistic/model."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference modules not found (neither /root/reference nor oracle/_ref; run oracle/make_ref.py "
                           "in the build container)")
    sys.dont_write_bytecode = True
    _stubpkg("ase", Atoms=object)
    _stubpkg("ase.data", atomic_masses=np.ones(119))
    _stubpkg("ase.neighborlist", neighbor_list=None)
    _stubpkg("matscipy")
    _stubpkg("matscipy.neighbours", neighbour_list=None)
    _stubpkg("vesin", NeighborList=None)
    _stubpkg("fasteners")
    _stubpkg("dirsync", sync=None)
    spk = _stubpkg("schnetpack", REF, __version__="2.2.0")
    for n in ("properties", "utils", "nn", "representation"):
        setattr(spk, n, importlib.import_module("schnetpack." + n))
    ato = _stubpkg("schnetpack.atomistic", REF + "/atomistic")
    spk.atomistic = ato
    for n in ("atomwise", "response", "distances"):
        m = importlib.import_module("schnetpack.atomistic." + n)
        for k in getattr(m, "__all__", []):
            setattr(ato, k, getattr(m, k))
    ato.PairwiseDistances = sys.modules["schnetpack.atomistic.distances"].PairwiseDistances
    tr = _stubpkg("schnetpack.transform", REF + "/transform")
    spk.transform = tr
    tr.Transform = importlib.import_module("schnetpack.transform.base").Transform
    for n in ("atomistic", "casting", "neighborlist"):
        m = importlib.import_module("schnetpack.transform." + n)
        for k in getattr(m, "__all__", []):
            setattr(tr, k, getattr(m, k))
    mdl = _stubpkg("schnetpack.model", REF + "/model")
    spk.model = mdl
    mb = importlib.import_module("schnetpack.model.base")
    mdl.NeuralNetworkPotential = mb.NeuralNetworkPotential
    mdl.AtomisticModel = mb.AtomisticModel
    _loaded = spk
    return spk


def load_model(path):
    load()
    from schnetpack.utils.compatibility import load_model as _lm

    return _lm(path)


def testdata(name: str) -> str:
    """Path of a file of the reference's tests/testdata (md_ethanol.model, md_ethanol.xyz)."""
    return os.path.join(TESTDATA, name)


def loaded_files():
    """Files of the reference the loader has imported (what make_ref.py copies)."""
    load()
    out = []
    for name, m in sorted(sys.modules.items()):
        f = getattr(m, "__file__", None)
        if name.startswith("schnetpack") and f and f.startswith(REF):
            out.append(f)
    return out


def build_from_spec(spec: dict, params: dict, dtype=None, device=None):
    """The reference's own ``NeuralNetworkPotential([PairwiseDistances], {PaiNN|SchNet}, [Atomwise, Forces])`` for a
    ``schnetpack_b200.synthetic.model_spec`` dictionary, loaded with a flat weight dictionary (reference ``state_dict`` keys)."""
    import re

    import torch

    spk = load()
    nn_ = spk.nn
    if spec["rbf"] == "gaussian":
        rbf = nn_.GaussianRBF(n_rbf=spec["n_rbf"], cutoff=spec["cutoff"])
    else:
        rbf = nn_.BesselRBF(n_rbf=spec["n_rbf"], cutoff=spec["cutoff"])
    cut = nn_.CosineCutoff(spec["cutoff"])
    if spec["kind"] == "painn":
        rep = spk.representation.PaiNN(spec["n_atom_basis"], spec["n_interactions"], rbf, cut,
                                       shared_interactions=spec["shared_interactions"],
                                       shared_filters=spec["shared_filters"], epsilon=spec["epsilon"])
    else:
        rep = spk.representation.SchNet(spec["n_atom_basis"], spec["n_interactions"], rbf, cut,
                                        n_filters=spec["n_filters"], shared_interactions=spec["shared_interactions"])
    outs = [spk.atomistic.Atomwise(n_in=spec["n_atom_basis"], output_key="energy")]
    if spec.get("forces", True):
        outs.append(spk.atomistic.Forces(energy_key="energy", force_key="forces"))
    model = spk.model.NeuralNetworkPotential(rep, input_modules=[spk.atomistic.PairwiseDistances()],
                                             output_modules=outs, postprocessors=[], do_postprocessing=False)
    sd = model.state_dict()
    new = {}
    for k in sd:
        kk = k
        if k not in params and spec["shared_interactions"]:   # shared blocks: one module under every index
            kk = re.sub(r"\.(interactions|mixing)\.\d+\.", r".\1.0.", k)
        new[k] = torch.as_tensor(params[kk]).to(sd[k].dtype)
    model.load_state_dict(new)
    if dtype is not None:
        model = model.to(dtype)
    if device is not None:
        model = model.to(device)
    return model.eval()


def evaluate(model, inputs: dict, dtype=None, device=None, forces: bool = True):
    """Run a reference model on a numpy / tensor batch -> dict of detached tensors (energy, forces, representations).  A
    batch that carries ``_Rij`` (padded neighbour list) bypasses PairwiseDistances like the fixtures do."""
    import torch

    x = {}
    for k, v in inputs.items():
        t = torch.as_tensor(v)
        if t.is_floating_point() and dtype is not None:
            t = t.to(dtype)
        x[k] = t.to(device) if device is not None else t
    if "_Rij" in x:
        x["_Rij"].requires_grad_(forces)
        x = model.representation(x)
        x = model.output_modules[0](x)
        res = {"energy": x["energy"]}
    else:
        res = model(x)
    out = {k: v.detach() for k, v in res.items()}
    out["scalar_representation"] = x["scalar_representation"].detach()
    if "vector_representation" in x:
        out["vector_representation"] = x["vector_representation"].detach()
    return out
