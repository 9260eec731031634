"""GaussianNB.fit and KMeans.fit on the GPU (csrc/fit.cu) against scikit-learn.

The GPU sums are fp64 and deterministic but not in numpy's order, so fitted parameters are compared with a relative
tolerance (FIT_RTOL) instead of bit for bit; labels predicted with the fitted models must agree exactly on the rows
used here (their top-1/top-2 margins are far above that tolerance)."""
import ctypes as C

import numpy as np
import pytest

import traffic_classifier_sdn_b200 as tc
from traffic_classifier_sdn_b200 import _lib, synth

FIT_RTOL = 1e-10


def test_fit_argument_errors_need_no_gpu():
    lib = _lib.load()
    out = np.zeros(8)
    f = _lib._f64p
    assert lib.tcsdn_gnb_fit(None, None, 10, 4, 2, 1, 0, 1e-9, out.ctypes.data_as(f), out.ctypes.data_as(f),
                             out.ctypes.data_as(f), None, None, None) == _lib.EINVAL
    assert lib.tcsdn_kmeans_fit(None, 10, 4, 2, 1, 0, out.ctypes.data_as(f), 10, 1e-4, out.ctypes.data_as(f), None, None,
                                None, None) == _lib.EINVAL
    x = np.zeros((4, 3))
    assert lib.tcsdn_kmeans_fit(x.ctypes.data_as(C.c_void_p), 4, 3, 40, 1, 0, out.ctypes.data_as(f), 10, 1e-4,
                                out.ctypes.data_as(f), None, None, None, None) == _lib.EINVAL   # k > 33


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gnb_fit_matches_sklearn(golden, dtype):
    from sklearn.naive_bayes import GaussianNB as SK
    for X, y in ((golden["X"], golden["y"]), synth.make_flows(150_001, seed=5, d=8)):
        X = np.ascontiguousarray(X, dtype)
        # float32 rows: scikit-learn then does the moments in float32 arithmetic; the GPU accumulates the same float32
        # values in fp64, so the yardstick is scikit-learn on the widened rows
        sk = SK().fit(X.astype(np.float64), y)
        est = tc.GaussianNB().fit(X, y, backend="gpu")
        assert np.array_equal(est.classes_, sk.classes_)
        np.testing.assert_allclose(est.theta_, sk.theta_, rtol=FIT_RTOL, atol=0)
        np.testing.assert_allclose(est.var_, sk.var_, rtol=FIT_RTOL, atol=0)
        np.testing.assert_allclose(est.class_prior_, sk.class_prior_, rtol=1e-15)
        np.testing.assert_allclose(est.epsilon_, sk.epsilon_, rtol=FIT_RTOL)
        assert np.array_equal(est.class_count_, sk.class_count_)
        assert np.array_equal(est.predict(X), sk.predict(X))
        again = tc.GaussianNB().fit(X, y, backend="gpu")            # deterministic: same bits on a second run
        assert np.array_equal(again.theta_, est.theta_) and np.array_equal(again.var_, est.var_)


@pytest.mark.gpu
def test_gnb_fit_device_rows_and_errors():
    import torch
    X, y = synth.make_flows(20_000, seed=9)
    est_h = tc.GaussianNB().fit(X, y, backend="gpu")
    est_d = tc.GaussianNB().fit(torch.from_numpy(X).cuda(), y, backend="gpu")
    assert np.array_equal(est_h.theta_, est_d.theta_) and np.array_equal(est_h.var_, est_d.var_)
    bad = X.copy()
    bad[17, 3] = np.nan
    with pytest.raises(ValueError):
        tc.GaussianNB().fit(bad, y, backend="gpu")
    with pytest.raises(ValueError):
        tc.GaussianNB().fit(X, y[:-1], backend="gpu")
    with pytest.raises(ValueError):
        tc.GaussianNB(priors=[0.5, 0.5]).fit(X, y, backend="gpu")
    assert tc.RandomForestClassifier(n_estimators=3).fit(X[:500], y[:500]).predict(X[:5]).shape == (5,)   # host route


@pytest.mark.gpu
def test_kmeans_fit_from_init_array_matches_sklearn(golden):
    from sklearn.cluster import KMeans as SK
    for X, k in ((golden["X"], 4), (synth.make_flows(60_000, seed=11, return_labels=False), 6)):
        rng = np.random.default_rng(3)
        init = X[rng.choice(len(X), k, replace=False)].copy()
        sk = SK(n_clusters=k, init=init, n_init=1, algorithm="lloyd").fit(X)
        est = tc.KMeans(n_clusters=k, init=init, n_init=1).fit(X, backend="gpu")
        scale = np.abs(sk.cluster_centers_).max()
        np.testing.assert_allclose(est.cluster_centers_, sk.cluster_centers_, rtol=1e-9, atol=1e-9 * scale)
        assert np.array_equal(est.labels_, sk.labels_)
        assert est.n_iter_ == sk.n_iter_
        np.testing.assert_allclose(est.inertia_, sk.inertia_, rtol=1e-9)
        assert np.array_equal(est.predict(X), sk.predict(X))


@pytest.mark.gpu
def test_kmeans_fit_default_seeding_is_sklearns(golden):
    from sklearn.cluster import KMeans as SK
    X = synth.make_flows(40_000, seed=21, return_labels=False)
    sk = SK(n_clusters=5, n_init=1, random_state=7).fit(X)
    est = tc.KMeans(n_clusters=5, n_init=1, random_state=7).fit(X)      # backend auto -> GPU
    scale = np.abs(sk.cluster_centers_).max()
    np.testing.assert_allclose(est.cluster_centers_, sk.cluster_centers_, rtol=1e-9, atol=1e-9 * scale)
    assert np.array_equal(est.labels_, sk.labels_) and est.n_iter_ == sk.n_iter_
    with pytest.raises(ValueError):
        tc.KMeans(n_clusters=5, n_init=10).fit(X, backend="gpu")         # restarts stay with scikit-learn
    assert tc.KMeans(n_clusters=3, n_init=3, random_state=0).fit(X).cluster_centers_.shape == (3, 12)
