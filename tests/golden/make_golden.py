"""Generate the committed golden fixtures by running the UNMODIFIED reference modules (build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Each fixture stores the model spec, the weight seed (weights are regenerated with
``schnetpack_b200.synthetic.init_params``; the trained-model fixture stores the weights themselves), the inputs and
the outputs the reference produced: energy [B], forces [N,3] (= -dE/dR through the reference's own
``PairwiseDistances`` -> representation -> ``Atomwise`` -> ``Forces`` modules assembled in a reference
``NeuralNetworkPotential`` with ``do_postprocessing=False``), scalar/vector representation.
fp32 AND fp64 reference outputs are stored: fp64 is the truth the 1e-5 tolerance is judged against, fp32 shows the
reference's own rounding noise.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader as rl  # noqa: E402
from schnetpack_b200 import synthetic as S  # noqa: E402


def build_reference_model(spk, spec, params, dtype):
    return rl.build_from_spec(spec, params, dtype)


def run_reference(spk, model, inputs, dtype, spec):
    x = {}
    for k, v in inputs.items():
        t = torch.as_tensor(v)
        if t.is_floating_point():
            t = t.to(dtype)
        x[k] = t
    direct = S.Rij in x
    if direct:
        # padded list: bypass PairwiseDistances, call representation + Atomwise directly
        x[S.Rij].requires_grad_(spec["forces"])
        x = model.representation(x)
        x = model.output_modules[0](x)
        res = {"energy": x["energy"]}
        x_all = x
    else:
        res = model(x)
        x_all = x
    out = {"energy": res["energy"].detach().numpy()}
    if "forces" in res:
        out["forces"] = res["forces"].detach().numpy()
    out["scalar_representation"] = x_all["scalar_representation"].detach().numpy()
    if "vector_representation" in x_all:
        out["vector_representation"] = x_all["vector_representation"].detach().numpy()
    return out


def save_case(name, spec, seed, inputs, ref32, ref64, params=None):
    blob = {"spec_json": np.array(json.dumps(spec)), "seed": np.array(seed)}
    for k, v in inputs.items():
        blob["in:" + k] = v
    for k, v in ref32.items():
        blob["ref32:" + k] = v
    for k, v in ref64.items():
        blob["ref64:" + k] = v
    if params is not None:
        for k, v in params.items():
            blob["param:" + k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: N={inputs[S.Z].shape[0]} E={inputs[S.idx_i].shape[0]} E0={ref64['energy'][0]:.6f} "
          f"-> {os.path.getsize(path) / 1e3:.0f} kB")


def model_case(spk, name, spec, inputs, seed):
    params = S.init_params(spec, seed)
    outs = {}
    for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
        m = build_reference_model(spk, spec, params, dt)
        outs[tag] = run_reference(spk, m, inputs, dt, spec)
    save_case(name, spec, seed, inputs, outs["32"], outs["64"])


def primitives_case(spk):
    """The reference's own known-answer inputs (tests/nn/test_radial.py, test_cutoff.py, test_activations.py),
    evaluated by the reference functions; stored with the closed-form expectation the reference tests assert."""
    nn_ = spk.nn
    blob = {}
    d1 = torch.tensor([[[1.0]]])
    blob["rbf1_in"] = d1.numpy()
    blob["rbf1_ref"] = nn_.GaussianRBF(n_rbf=6, cutoff=5.0)(d1).numpy()
    blob["rbf1_expt"] = torch.exp(-0.5 * torch.tensor([[[1.0, 0.0, 1.0, 4.0, 9.0, 16.0]]])).numpy()
    d2 = torch.tensor([[[0.0, 1.0, 1.5], [0.5, 1.5, 3.0]]])
    g2 = nn_.GaussianRBF(start=1.0, cutoff=4.0, n_rbf=4)
    blob["rbf2_in"] = d2.numpy()
    blob["rbf2_ref"] = g2(d2).numpy()
    blob["rbf2_offsets"] = g2.offsets.numpy()
    blob["rbf2_widths"] = g2.widths.numpy()
    e2 = torch.tensor([[[[1, 2, 3, 4], [0, 1, 2, 3], [0.5, 0.5, 1.5, 2.5]],
                        [[0.5, 1.5, 2.5, 3.5], [0.5, 0.5, 1.5, 2.5], [2, 1, 0, 1]]]])
    blob["rbf2_expt"] = torch.exp(-0.5 * e2**2).numpy()
    d3 = torch.tensor([[[0.0, 1.0, 1.5, 0.25], [0.5, 1.5, 3.0, 1.0]]])
    g3 = nn_.GaussianRBF(start=1.0, cutoff=4.0, n_rbf=5, trainable=True)
    blob["rbf3_in"] = d3.numpy()
    blob["rbf3_ref"] = g3(d3).detach().numpy()
    blob["rbf3_offsets"] = g3.offsets.detach().numpy()
    blob["rbf3_widths"] = g3.widths.detach().numpy()
    bes = nn_.BesselRBF(n_rbf=8, cutoff=5.0)
    db = torch.tensor([0.0, 0.3, 1.0, 2.5, 4.999, 5.0, 6.0])
    blob["bessel_in"] = db.numpy()
    blob["bessel_ref"] = bes(db).numpy()
    blob["bessel_freqs"] = bes.freqs.numpy()
    torch.manual_seed(42)
    dist = torch.rand((10, 5, 20), dtype=torch.float)
    cut = nn_.CosineCutoff(cutoff=1.8)
    blob["cut_in"] = dist.numpy()
    blob["cut_ref"] = cut(dist).numpy()
    blob["cut_ref35"] = cut(3.5 * dist).numpy()
    v = 0.5 * (1.0 + torch.cos(3.5 * dist * np.pi / 1.8))
    v[3.5 * dist >= 1.8] = 0.0
    blob["cut_expt35"] = v.numpy()
    x = torch.tensor([0.0, 1.0, 0.5, 2.0])
    blob["ssp_in"] = x.numpy()
    blob["ssp_ref"] = nn_.shifted_softplus(x).numpy()
    torch.manual_seed(42)
    xd = torch.randn((10, 5), dtype=torch.double)
    xd2 = 10 * torch.randn((10, 5), dtype=torch.double)
    blob["ssp_in64"] = xd.numpy()
    blob["ssp_ref64"] = nn_.shifted_softplus(xd).numpy()
    blob["ssp_in64b"] = xd2.numpy()
    blob["ssp_ref64b"] = nn_.shifted_softplus(xd2).numpy()
    # scatter_add and collate index shifting (tests/data/test_loader.py:16-29)
    xs = torch.arange(24, dtype=torch.float32).reshape(8, 3)
    idx = torch.tensor([0, 0, 2, 1, 2, 2, 4, 0])
    blob["scatter_in"] = xs.numpy()
    blob["scatter_idx"] = idx.numpy()
    blob["scatter_ref"] = nn_.scatter_add(xs, idx, dim_size=5).numpy()
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **blob)
    print("primitives ok")


def trained_case(spk):
    """Shipped trained PaiNN (tests/testdata/md_ethanol.model) on tests/testdata/md_ethanol.xyz, no postprocessing."""
    m = rl.load_model("/root/reference/tests/testdata/md_ethanol.model")
    m.do_postprocessing = False
    m.eval()
    inputs = S.ethanol_batch(batch=2, jitter=0.03, seed=3)
    # first molecule = exact xyz geometry
    inputs[S.R][:9] = S.ETHANOL_R.astype(np.float32)
    spec = S.model_spec(kind="painn", n_atom_basis=128, n_interactions=3, n_rbf=20, cutoff=5.0)
    params = {k: v.detach().numpy() for k, v in m.state_dict().items()
              if k.startswith("representation.") or k.startswith("output_modules.0.")}
    outs = {}
    for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
        mm = m.to(dt)
        outs[tag] = run_reference(spk, mm, inputs, dt, spec)
    save_case("painn_md_ethanol_trained", spec, -1, inputs, outs["32"], outs["64"], params=params)


def main():
    spk = rl.load()
    torch.set_num_threads(8)
    primitives_case(spk)
    model_case(spk, "painn_aspirin_b4", S.model_spec("painn"), S.aspirin_batch(4, seed=1), seed=11)
    model_case(spk, "schnet_ethanol_b1", S.model_spec("schnet"), S.ethanol_batch(1), seed=12)
    model_case(spk, "schnet_qm9_b8", S.model_spec("schnet", n_interactions=6, forces=False),
               S.qm9like_batch(8, seed=2), seed=13)
    model_case(spk, "schnet_qm9_b8_padded", S.model_spec("schnet", n_interactions=6, forces=False),
               S.qm9like_batch(8, seed=2, padded=True), seed=13)
    model_case(spk, "schnet_qm9_b8_forces", S.model_spec("schnet", n_interactions=6, forces=True),
               S.qm9like_batch(8, seed=2), seed=13)
    model_case(spk, "painn_box_216", S.model_spec("painn"), S.periodic_box(216, seed=4), seed=14)
    model_case(spk, "painn_bessel_shared", S.model_spec("painn", n_atom_basis=64, n_interactions=2, n_rbf=16,
                                                         rbf="bessel", shared_interactions=True, shared_filters=True),
               S.aspirin_batch(3, seed=5), seed=15)
    model_case(spk, "schnet_box_216", S.model_spec("schnet", n_interactions=3), S.periodic_box(216, seed=6), seed=16)
    trained_case(spk)


if __name__ == "__main__":
    main()
