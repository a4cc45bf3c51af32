"""Stub-package loader that imports the UNMODIFIED reference hot-path modules from /root/reference.

Only usable in the build container (``/root/reference`` does not exist on the GPU box).  It is used by
``tests/golden/make_golden.py`` to generate the committed golden fixtures and by the optional
``tests/test_oracle_vs_live_reference.py`` (skipped when the reference tree is absent).

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
REF = REF_ROOT + "/src/schnetpack"


def available() -> bool:
    return os.path.isdir(REF)


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
        raise RuntimeError("reference tree not present at " + REF)
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
