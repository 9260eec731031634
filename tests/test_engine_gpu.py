"""GPU: the tensor-core (tcgen05) distance engine against the oracle.

KNN through the engine must be bit-identical to the fp64 kernel / oracle / sklearn's sequential heap: the
tensor-core distances only FILTER candidates, every survivor is re-evaluated in fp64 in index order.
SVC through the engine evaluates exp() and the one-vs-one sums in fp32 (tile sums promoted to fp64):
decision values agree with libsvm's fp64 to SVC_ENGINE_TOL absolute; labels agree wherever no pairwise
decision value is closer to zero than that tolerance (such rows are counted and must be rare).
"""
import numpy as np
import pytest

import oracle
from traffic_classifier_sdn_b200 import _lib, from_spec, synth

pytestmark = pytest.mark.gpu

SVC_ENGINE_TOL = 2e-3   # absolute, on decision values of magnitude O(1..100)
KAPPA = 2.0 ** -18      # csrc/dist_engine.cu kKappa


def _force(est, mode):
    est.set_option(_lib.OPT_ENGINE, mode)
    return est


def _knn_spec(nt, k=5, seed=0, d=12):
    X, y = synth.make_flows(nt, seed=seed)
    return dict(kind="knn", fit_X=X[:, :d].copy(), y=y.astype(np.int32), k=k, classes=synth.CLASSES, n_features=d)


@pytest.mark.parametrize("nt,nq,k", [(5000, 6000, 5), (777, 4100, 1), (20000, 9000, 5), (3000, 5000, 17), (64, 4096, 3)])
def test_knn_engine_bit_exact_vs_oracle(nt, nq, k):
    spec = _knn_spec(nt, k=k, seed=nt)
    est = _force(from_spec(spec), 2)
    Xq = synth.make_flows(nq, seed=nq + 1, return_labels=False)
    dup = min(300, nt)
    Xq[:dup] = spec["fit_X"][:dup]           # exact duplicates of training rows (distance 0, ties with their twins)
    idx, pr = est._run(Xq, True)
    ridx, rpr = oracle.knn(spec, Xq)
    st = est.stats()
    assert st[1] == nq and st[2] == 0, "rows did not go through the tensor-core engine"
    assert np.array_equal(pr, rpr) and np.array_equal(idx, ridx)
    # the filter must actually filter: far fewer exact evaluations than pairs
    assert st[3] < 0.05 * nq * nt + 64 * nq, (st[3], nq * nt)
    print(f"knn nt={nt} nq={nq} k={k}: {st[3] / nq:.1f} exact re-evaluations per query of {nt} candidates")
    idx32, pr32 = est._run(Xq.astype(np.float32), True)
    r32 = oracle.knn(spec, Xq.astype(np.float32).astype(np.float64))
    assert np.array_equal(pr32, r32[1]) and np.array_equal(idx32, r32[0])


@pytest.mark.parametrize("case", ["flows", "positive_large", "offset_4e5", "heavy_tail"])
def test_knn_filter_slack_audit_all_pairs(case):
    """audit mode re-evaluates EVERY pair exactly and records max |tensor-core value - exact| / (||x||^2 + max ||t||^2):
    the filter slack kappa = 2^-18 must dominate it with a wide margin on adversarial magnitudes."""
    rng = np.random.default_rng(11)
    nt, nq = 2048, 4096
    if case == "flows":
        tr = synth.make_flows(nt, seed=1, return_labels=False); q = synth.make_flows(nq, seed=2, return_labels=False)
    elif case == "positive_large":   # no sign cancellation inside the dot products, every term near 4e5^2
        tr = rng.uniform(1e5, 4e5, (nt, 12)); q = rng.uniform(1e5, 4e5, (nq, 12))
    elif case == "offset_4e5":
        tr = 4e5 + rng.normal(0, 50.0, (nt, 12)); q = 4e5 + rng.normal(0, 50.0, (nq, 12))
    else:
        tr = rng.standard_cauchy((nt, 12)) * 100; q = rng.standard_cauchy((nq, 12)) * 100
        tr = np.clip(tr, -1e6, 1e6); q = np.clip(q, -1e6, 1e6)
    y = rng.integers(0, 3, nt).astype(np.int32)
    spec = dict(kind="knn", fit_X=tr, y=y, k=5, classes=np.arange(3), n_features=12)
    est = _force(from_spec(spec), 3)
    idx, pr = est._run(q, True)
    st = est.stats()
    assert st[3] == nq * nt, "audit mode must re-evaluate every pair"
    ratio = st[5] / 2.0 ** 40
    print(f"audit[{case}]: max error ratio {ratio:.3e} = 2^{np.log2(ratio):.1f} (kappa = 2^-18)")
    assert 0 < ratio < KAPPA / 4
    ridx, rpr = oracle.knn(spec, q)
    assert np.array_equal(pr, rpr) and np.array_equal(idx, ridx)


def test_knn_engine_lattice_ties_heap_order():
    rng = np.random.default_rng(7)
    tr = np.floor(rng.random((900, 4)) * 3)
    y = rng.integers(0, 5, 900).astype(np.int32)
    q = np.floor(rng.random((5000, 4)) * 3)
    spec = dict(kind="knn", fit_X=tr, y=y, k=5, classes=np.arange(5), n_features=4)
    est = _force(from_spec(spec), 2)
    idx, pr = est._run(q, True)
    ridx, rpr = oracle.knn(spec, q)
    assert np.array_equal(pr, rpr) and np.array_equal(idx, ridx)
    from sklearn.neighbors import KNeighborsClassifier
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=1):
        sk = KNeighborsClassifier(5, algorithm="brute").fit(tr, y)
        assert np.array_equal(pr, sk.predict_proba(q))


def test_knn_engine_equals_fp64_kernel_on_golden(golden, specs):
    X = golden["X"]
    a = _force(from_spec(specs["knn"]), 2)._run(X, True)
    b = _force(from_spec(specs["knn"]), 1)._run(X, True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[0], golden["knn.expected_label"])


def test_knn_engine_huge_magnitudes_and_offsets():
    """features around 4e5 with tiny relative differences: the filter must still keep every true neighbour"""
    rng = np.random.default_rng(3)
    tr = 4e5 + rng.normal(0, 3.0, (4000, 12))
    y = rng.integers(0, 4, 4000).astype(np.int32)
    q = 4e5 + rng.normal(0, 3.0, (4500, 12))
    spec = dict(kind="knn", fit_X=tr, y=y, k=5, classes=np.arange(4), n_features=12)
    est = _force(from_spec(spec), 2)
    idx, pr = est._run(q, True)
    ridx, rpr = oracle.knn(spec, q)
    assert np.array_equal(pr, rpr) and np.array_equal(idx, ridx)


def _svc_spec(nsv, C=6, seed=0, gamma=None):
    rng = np.random.default_rng(seed)
    Xs, ys = synth.make_flows(nsv, seed=seed + 1, class_weights=np.r_[np.ones(C), np.zeros(6 - C)] if C < 6 else None)
    order = np.argsort(ys, kind="stable")
    Xs, ys = Xs[order], ys[order]
    nsup = np.bincount(ys, minlength=C)[:C].astype(np.int32)
    g = gamma if gamma is not None else 1.0 / (12 * Xs.var())
    dual = rng.uniform(-1.0, 1.0, (C - 1, nsv)) * (rng.random((C - 1, nsv)) < 0.6)
    return dict(kind="svc", sv=Xs, dual_coef=dual, intercept=rng.normal(0, 0.5, C * (C - 1) // 2), n_support=nsup,
                gamma=float(g), classes=synth.CLASSES[:C], n_features=12, decision_function_shape="ovr", break_ties=False,
                n_classes=C)


def _check_svc(spec, X, tol=SVC_ENGINE_TOL):
    est = _force(from_spec(spec), 2)
    idx, dec = est._run(X, True)
    ridx, rdec = oracle.svc(spec, X.astype(np.float64))
    st = est.stats()
    assert st[1] == len(X) and st[2] == 0
    err = np.max(np.abs(dec - rdec))
    assert err < tol, f"max |dec - libsvm fp64| = {err:.3e}"
    safe = np.min(np.abs(rdec), axis=1) > tol
    assert np.array_equal(idx[safe], ridx[safe])
    assert (~safe).mean() < 0.01 or len(X) < 100
    return err


def test_svc_engine_golden(golden, specs):
    err = _check_svc(specs["svc"], golden["X"])
    est = from_spec(specs["svc"])            # auto mode: 7 653 rows take the engine
    assert np.array_equal(est.predict_indices(golden["X"]), golden["svc.expected_label"])
    print(f"svc golden: max abs dec error {err:.3e}")


@pytest.mark.parametrize("nsv,C,nq", [(3000, 6, 5000), (513, 3, 4200), (20000, 6, 4608), (130, 2, 4096)])
def test_svc_engine_synthetic(nsv, C, nq):
    spec = _svc_spec(nsv, C, seed=nsv)
    X = synth.make_flows(nq, seed=nsv + 5, return_labels=False)
    err = _check_svc(spec, X)
    err32 = _check_svc(spec, X.astype(np.float32))
    print(f"svc nsv={nsv} C={C}: max abs dec error {err:.3e} (f64 rows) {err32:.3e} (f32 rows)")


def test_engine_ragged_and_auto_dispatch(specs):
    """n not a multiple of 512; small batches stay on the fp64 kernels, large ones take the engine"""
    for kind in ("knn", "svc"):
        est = from_spec(specs[kind])
        for n in (100, 4095, 4096, 4097, 5000):
            X = synth.make_flows(n, seed=n, return_labels=False)
            idx, sc = est._run(X, True)
            st = est.stats()
            assert (st[1] == n) == (n >= 4096) and (st[2] == n) == (n < 4096)
            ridx, rsc = oracle.predict(specs[kind], X)
            if kind == "knn":
                assert np.array_equal(idx, ridx) and np.array_equal(sc, rsc)
            else:
                assert np.max(np.abs(sc - rsc)) < SVC_ENGINE_TOL


def test_engine_nonfinite_rows_raise(specs):
    X = synth.make_flows(5000, seed=1, return_labels=False)
    X[4321, 7] = np.inf
    for kind in ("knn", "svc"):
        with pytest.raises(ValueError, match="NaN|infinity"):
            from_spec(specs[kind]).predict(X)


# ------------------------------------------------------------------ BASELINE sizes: 10M query rows (configs[2], configs[3])
@pytest.mark.parametrize("name", ["knn", "svc"])
def test_engine_full_size_properties(name):
    """The bench workloads at their full size (10M rows x 50k training rows / 20k support vectors), checked through
    properties that do not need a 10M-row oracle: a permutation of the rows permutes the labels, one launch equals
    eight slices, a strided sample equals the CPU oracle and the fp64 CUDA-core kernel (KNN: exactly; SVC: wherever no
    pairwise decision value is within the engine tolerance of zero), and the exact-evaluation counter stays sane."""
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    w = bench.build_workload(name)
    n = w["full_rows"]
    est = from_spec(w["spec"])
    X = bench.synth_rows(n, w["d"], seed=4242, device=torch.device("cuda", 0))
    full = est.predict_indices(X)
    st = est.stats()
    assert st[1] == n and st[2] == 0                       # every row went through the tensor-core engine
    g = torch.Generator(device="cpu").manual_seed(3)
    perm = torch.randperm(n, generator=g).cuda()
    assert torch.equal(full[perm], est.predict_indices(X[perm].contiguous()))
    assert torch.equal(full, torch.cat([est.predict_indices(c.contiguous()) for c in X.chunk(8)]))
    est.sync_check()
    hist = np.bincount(full.cpu().numpy(), minlength=8)
    assert hist.sum() == n and (hist > 0).sum() >= 2       # not a constant answer
    # strided sample against the oracle (fp64 restatement of sklearn) and the fp64 kernel
    s = slice(0, n, 20_011)
    xs = X[s].cpu().numpy().astype(np.float64)
    lab_o, sc_o = oracle.predict(w["spec"], xs, want_scores=True)
    lab_e = full[s].cpu().numpy()
    est64 = _force(from_spec(w["spec"]), 1)
    lab_k = est64.predict_indices(xs)
    assert np.array_equal(lab_k, lab_o)
    if name == "knn":
        assert np.array_equal(lab_e, lab_o)
        evals_per_query = (est.stats()[3]) / (3.0 * n)     # three full passes so far
        assert 5 <= evals_per_query < 400
    else:
        safe = np.abs(sc_o).min(axis=1) > SVC_ENGINE_TOL
        assert np.array_equal(lab_e[safe], lab_o[safe]) and (~safe).mean() < 0.02
