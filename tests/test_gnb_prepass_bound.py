"""CPU check of the mathematics behind the GaussianNB fp32 pre-pass (csrc/scorers.cu, gnb_prepass_rows).

The kernel evaluates  T~_c = sum_j fma32(x_j, a~_cj, -b~_cj)^2  (fp32 fma chain, two rows per FFMA2) and
acc~_c = c~_c - T~_c  in fp32 and claims
    |acc~_c - jll_c| <= E_c = 2^-19 (T_c + |c_c| + sum_j b_cj^2),   T_c = c~_c - acc~_c,
where jll_c is the fp64 definition.  A row is certified when the class with the largest upper bound acc~ + E has a
lower bound acc~ - E above every other class's upper bound; then the fp64 argmax must be that class.  This test
emulates the fp32 arithmetic with numpy (products of two float32 are exact in float64, so one float64 operation
followed by a cast is a correctly rounded fp32 fma up to a negligible double-rounding effect) and checks both claims on
models with wildly different per-class variances, rows near and far from the class means, exact ties and huge values."""
import numpy as np
import pytest

EPS = 2.0 ** -19


def _model(rng, C, d, spread):
    theta = rng.uniform(0.0, 10.0 ** rng.uniform(0, spread, (C, 1)), (C, d))
    var = 10.0 ** rng.uniform(-3, 2 * spread, (C, d))
    prior = rng.dirichlet(np.ones(C))
    a = 1.0 / np.sqrt(2.0 * var)                       # the kernel's constants (abi.cu: scorer packing for GaussianNB)
    b = theta * a
    c = np.log(prior) - 0.5 * np.sum(np.log(2.0 * np.pi * var), axis=1)
    return theta, var, a, b, c


def _fp64(x, a, b, c):
    t = x[:, None, :] * a[None] - b[None]
    return c[None] - np.sum(t * t, axis=2)


def _fp32(x, a, b, c):
    af, bf, cf = a.astype(np.float32), b.astype(np.float32), c.astype(np.float32)
    Tacc = np.zeros((len(x), len(c)), np.float32)
    xf = x.astype(np.float32)
    for j in range(x.shape[1]):
        t = (xf[:, None, j].astype(np.float64) * af[None, :, j].astype(np.float64) - bf[None, :, j].astype(np.float64)).astype(np.float32)
        Tacc = (Tacc.astype(np.float64) + t.astype(np.float64) * t.astype(np.float64)).astype(np.float32)
    acc = (cf[None].astype(np.float64) - Tacc.astype(np.float64)).astype(np.float32)
    T = cf[None].astype(np.float64) - acc.astype(np.float64)
    E = EPS * (T + np.abs(c)[None] + np.sum(b * b, axis=1)[None])
    return acc.astype(np.float64), E


@pytest.mark.parametrize("seed,spread", [(0, 1.0), (1, 3.0), (2, 5.0), (3, 2.0)])
def test_error_bound_and_certification(seed, spread):
    rng = np.random.default_rng(seed)
    C, d, n = 6, 8, 20_000
    theta, var, a, b, c = _model(rng, C, d, spread)
    cls = rng.integers(0, C, n)
    x = theta[cls] + rng.normal(0, 1, (n, d)) * np.sqrt(var[cls]) * rng.choice([0.01, 1.0, 30.0], (n, 1))
    x[: n // 10] = rng.uniform(0, 1e6, (n // 10, d))            # far from everything
    x[n // 10: n // 5] = theta[cls[n // 10: n // 5]]             # exactly on a class mean
    x = x.astype(np.float32).astype(np.float64)                  # the pre-pass only runs on float32 rows
    jll = _fp64(x, a, b, c)
    acc, E = _fp32(x, a, b, c)
    finite = np.isfinite(acc) & np.isfinite(E)
    assert finite.mean() > 0.9
    assert np.all(np.abs(acc - jll)[finite] <= E[finite]), "fp32 error exceeds the bound the kernel relies on"
    hi, lo = acc + E, acc - E
    best = np.argmax(hi, axis=1)
    others = np.where(np.arange(C)[None] == best[:, None], -np.inf, hi).max(axis=1)
    sure = lo[np.arange(n), best] > others
    assert sure.mean() > 0.5                                      # the bound is useful, not vacuous
    assert np.array_equal(np.argmax(jll, axis=1)[sure], best[sure])


def test_twin_classes_are_never_certified():
    rng = np.random.default_rng(9)
    theta, var, a, b, c = _model(rng, 3, 8, 2.0)
    theta[1], var[1] = theta[0], var[0]
    a, b = 1.0 / np.sqrt(2 * var), theta / np.sqrt(2 * var)
    c[1] = c[0]
    x = (theta[0] + rng.normal(0, 1, (5000, 8)) * np.sqrt(var[0])).astype(np.float32).astype(np.float64)
    acc, E = _fp32(x, a, b, c)
    hi, lo = acc + E, acc - E
    best = np.argmax(hi, axis=1)
    others = np.where(np.arange(3)[None] == best[:, None], -np.inf, hi).max(axis=1)
    sure = lo[np.arange(len(x)), best] > others
    winners = np.argmax(_fp64(x, a, b, c), axis=1)
    assert not np.any(sure & (winners != 2))                      # wherever a twin wins, the fp64 path must decide
