"""Golden fixture for the stress response (row a14, optional part): the UNMODIFIED reference modules
``Strain`` + ``PairwiseDistances`` -> ``PaiNN`` -> ``Atomwise`` + ``Forces(calc_forces=True, calc_stress=True)``
(/root/reference/src/schnetpack/atomistic/response.py:18-92,434-464) on a small periodic box, fp64.
Run in the build container only:  python tests/golden/make_golden_stress.py  ->  tests/golden/painn_box_stress.npz"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader as rl  # noqa: E402
from make_golden import build_reference_model  # noqa: E402
from schnetpack_b200 import synthetic as S  # noqa: E402


def main():
    spk = rl.load()
    spec, data = S.make_config("cfg4", n_atoms_total=96)
    spec = dict(spec, n_atom_basis=64, n_interactions=2, stress=True)
    seed = 31
    params = S.init_params(spec, seed)
    model = build_reference_model(spk, spec, params, torch.float64)
    model.input_modules = torch.nn.ModuleList([spk.atomistic.Strain(), spk.atomistic.PairwiseDistances()])
    model.output_modules = torch.nn.ModuleList([model.output_modules[0],
                                                spk.atomistic.Forces(calc_forces=True, calc_stress=True,
                                                                     energy_key="energy", force_key="forces")])
    model.collect_derivatives()
    model.collect_outputs()
    x = {}
    for k, v in data.items():
        t = torch.as_tensor(v)
        x[k] = t.double() if t.is_floating_point() else t
    x["_cell"] = x["_cell"].reshape(-1, 3, 3)
    out = model(x)
    save = {"spec_json": json.dumps(spec), "seed": np.int64(seed)}
    for k, v in data.items():
        save["in:" + k] = np.asarray(v)
    for k in ("energy", "forces", "stress"):
        save["ref64:" + k] = out[k].detach().numpy()
    np.savez_compressed(os.path.join(HERE, "painn_box_stress.npz"), **save)
    print({k: out[k].shape for k in ("energy", "forces", "stress")}, float(out["stress"].abs().max()))
    # pin the oracle right away
    from oracle import spk_oracle as O

    o = O.energy_forces(spec, params, data, dtype=torch.float64, stress=True)
    for k in ("energy", "forces", "stress"):
        err = float((o[k] - out[k].detach()).abs().max() / out[k].detach().abs().max())
        print(k, "oracle vs reference rel err", err)


if __name__ == "__main__":
    main()
