"""The special-function-unit shifted softplus of the fused SchNet block (csrc/common.cuh: spk_ssp_fast) restated in float32
numpy: softplus(x) - ln 2 = max(x, 0) + log1p(t) - ln 2 with t = 2^(-|x| log2 e) and the compensated
log1p(t) = ln(u) - ((u - 1) - t) / u, u = fl(1 + t).  The CUDA function is covered on the GPU by the fused-block tests (3e-6
against fp64); this pins the FORMULA -- including the branch-free behaviour beyond torch's softplus threshold of 20
(nn/activations.py:22 of the reference: there softplus(x) = x exactly; here the difference is log1p(e^-20) = 2e-9) -- with
the approximation error of ex2.approx / lg2.approx (2^-22 relative / 2^-22.6 absolute) injected as worst-case noise."""
import numpy as np

F = np.float32
LOG2E, LN2 = F(1.4426950408889634), F(0.6931471805599453)


def ssp_fast(x, rng=None):
    x = x.astype(F)
    t = np.exp2((-np.abs(x) * LOG2E).astype(F)).astype(F)
    if rng is not None:
        t = (t * (1 + rng.choice([-1.0, 1.0], size=x.shape) * 2.0 ** -22)).astype(F)
    u = (F(1) + t).astype(F)
    lg = np.log2(u.astype(np.float64))
    if rng is not None:
        lg = lg + rng.choice([-1.0, 1.0], size=x.shape) * 2.0 ** -22.6
    ln_u = (lg.astype(F) * LN2).astype(F)
    corr = (((u - F(1)).astype(F) - t).astype(F) / u).astype(F)
    return ((np.maximum(x, F(0)) + (ln_u - corr).astype(F)).astype(F) - LN2).astype(F)


def ssp_ref(x):
    x = x.astype(np.float64)
    return np.logaddexp(0.0, x) - np.log(2.0)


def test_formula_is_fp32_grade_over_the_whole_range():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(400_000) * 3, rng.uniform(-30, 30, 200_000), [0.0, -0.0, 20.0, 20.001, -20.0, 88.0, -88.0]])
    ref = ssp_ref(x)
    scale = np.maximum(np.abs(ref), 1.0)
    exact = ssp_fast(x)
    noisy = ssp_fast(x, rng)
    assert np.max(np.abs(exact - ref) / scale) < 2e-7
    assert np.max(np.abs(noisy - ref) / scale) < 4e-7
    assert np.sqrt(np.mean(((noisy - ref) / scale) ** 2)) < 2e-7


def test_matches_the_thresholded_reference_definition():
    # torch: softplus(x) = x for x > 20 (threshold), log1p(exp(x)) otherwise
    x = np.array([19.9, 20.0, 20.1, 25.0, 40.0], dtype=np.float64)
    torch_like = np.where(x > 20.0, x, np.log1p(np.exp(x))) - np.log(2.0)
    assert np.max(np.abs(ssp_fast(x) - torch_like)) < 4e-6        # one fp32 ulp at |y| ~ 20-40
