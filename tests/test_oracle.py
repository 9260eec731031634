"""CPU: pin the oracle (oracle/tcsdn_oracle.c) against scikit-learn's answers.

(a) golden vectors: the reference's six pickles x its 7 653 bundled rows (tests/golden/bundled.npz);
(b) live scikit-learn on seeded random models, including the edge cases sklearn's own tests cover
    (ties in votes and distances, duplicate training rows, single-row batches).
"""
import numpy as np
import pytest

import oracle
from conftest import KINDS
from sk_rebuild import sklearn_from_spec, quiet


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_golden_labels(golden, specs, kind):
    lab, _ = oracle.predict(specs[kind], golden["X"])
    assert np.array_equal(lab, golden[f"{kind}.expected_label"])


def test_oracle_scores_linear(golden, specs):
    _, s = oracle.linear(specs["linear"], golden["X"])
    assert _rel(s, golden["linear.expected_score"]) < 1e-13


def test_oracle_scores_gnb_bit_exact(golden, specs):
    _, s = oracle.gnb(specs["gnb"], golden["X"])
    assert np.array_equal(s, golden["gnb.expected_score"])  # same association order as numpy


def test_oracle_scores_kmeans(golden, specs):
    X = golden["X"]
    _, s = oracle.kmeans(specs["kmeans"], X)
    d2 = s + np.einsum("ij,ij->i", X, X)[:, None]
    ref = golden["kmeans.expected_score"] ** 2
    assert np.max(np.abs(d2 - ref) / np.maximum(1.0, ref)) < 1e-6  # ||x||^2 cancellation in fp64


def test_oracle_scores_knn(golden, specs):
    _, p = oracle.knn(specs["knn"], golden["X"])
    assert np.array_equal(np.rint(p * specs["knn"]["k"]).astype(np.uint8), golden["knn.expected_score"])


def test_oracle_scores_svc(golden, specs):
    _, dec = oracle.svc(specs["svc"], golden["X"])
    assert np.max(np.abs(dec - golden["svc.expected_score"])) < 1e-11
    ovr = oracle.ovr_from_ovo(dec, 6)
    assert np.max(np.abs(ovr - golden["svc.expected_ovr"])) < 1e-11


def test_oracle_scores_forest_bit_exact(golden, specs):
    _, p = oracle.forest(specs["forest"], golden["X"])
    assert np.array_equal(p, golden["forest.expected_score"])


# ------------------------------------------------------------------ live sklearn, seeded
def _flows(rng, n, d=12):
    """integer-ish, heavy-tailed, many zeros and duplicates -- the bundled data's character"""
    X = np.floor(rng.gamma(0.7, 50.0, size=(n, d)))
    X[rng.random((n, d)) < 0.3] = 0.0
    X[:, 3] = rng.gamma(2.0, 3.0, n)
    return X


def test_knn_heap_ties_vs_sklearn_brute():
    rng = np.random.default_rng(7)
    tr = np.floor(rng.random((600, 4)) * 3)  # tiny lattice: masses of exact distance ties
    y = rng.integers(0, 5, 600)
    q = np.floor(rng.random((3000, 4)) * 3)
    from sklearn.neighbors import KNeighborsClassifier
    from threadpoolctl import threadpool_limits
    spec = dict(kind="knn", fit_X=tr, y=y.astype(np.int32), k=5, classes=np.arange(5), n_features=4)
    lab, p = oracle.knn(spec, q)
    with threadpool_limits(limits=1):  # 3000 > 4*256*1 -> sklearn's sequential `parallel_on_X` reduction
        sk = KNeighborsClassifier(5, algorithm="brute").fit(tr, y)
        assert np.array_equal(p, sk.predict_proba(q))
        assert np.array_equal(lab, sk.predict(q))


@pytest.mark.parametrize("seed", [0, 1])
def test_forest_vs_sklearn_fresh_fit(seed):
    from sklearn.ensemble import RandomForestClassifier
    from traffic_classifier_sdn_b200.modelio import spec_from_estimator
    rng = np.random.default_rng(seed)
    X = _flows(rng, 3000)
    y = (X[:, 0] + X[:, 5] > 60).astype(int) + (X[:, 7] > 20)
    rf = RandomForestClassifier(n_estimators=20, max_depth=9, random_state=seed).fit(X[:1500], y[:1500])
    spec = spec_from_estimator(rf)
    Xq = X[1500:] + (rng.random((1500, 12)) < 0.1) * 1e-4  # values that straddle fp32 rounding
    lab, p = oracle.forest(spec, Xq)
    assert np.array_equal(p, rf.predict_proba(Xq))
    assert np.array_equal(rf.classes_[lab], rf.predict(Xq))


def test_svc_vs_sklearn_fresh_fit_binary_and_multiclass():
    from sklearn.svm import SVC
    from traffic_classifier_sdn_b200.modelio import spec_from_estimator
    rng = np.random.default_rng(3)
    X = _flows(rng, 1200)
    for nc in (2, 4):
        y = rng.integers(0, nc, 1200)
        y[X[:, 0] > 40] = 0
        m = SVC().fit(X[:600], y[:600])
        spec = spec_from_estimator(m)
        lab, dec = oracle.svc(spec, X[600:])
        assert np.array_equal(m.classes_[lab], m.predict(X[600:]))
        ref = m._decision_function(X[600:])
        ref = -ref.reshape(-1, 1) if nc == 2 else ref
        assert np.max(np.abs(dec - ref)) < 1e-11


def test_gnb_linear_kmeans_vs_sklearn_fresh_fit():
    from sklearn.cluster import KMeans
    from sklearn.linear_model import LogisticRegression
    from sklearn.naive_bayes import GaussianNB
    from traffic_classifier_sdn_b200.modelio import spec_from_estimator
    rng = np.random.default_rng(5)
    X = _flows(rng, 2000, d=8)
    y = rng.integers(0, 3, 2000)
    nb = GaussianNB().fit(X, y)
    lab, s = oracle.gnb(spec_from_estimator(nb), X)
    assert np.array_equal(s, nb._joint_log_likelihood(X)) and np.array_equal(nb.classes_[lab], nb.predict(X))
    for yy in (y, (y > 0).astype(int)):  # multinomial and binary
        lr = quiet(LogisticRegression(max_iter=50).fit, X, yy)
        lab, s = oracle.linear(spec_from_estimator(lr), X)
        assert _rel(s.squeeze(), lr.decision_function(X)) < 1e-12
        assert np.array_equal(lr.classes_[lab], lr.predict(X))
    km = KMeans(5, n_init=1, random_state=0).fit(X)
    lab, _ = oracle.kmeans(spec_from_estimator(km), X)
    assert np.array_equal(lab, km.predict(X))


def test_rebuilt_estimators_equal_spec_roundtrip(golden, specs):
    """sklearn_from_spec (used to make the golden file) reproduces the golden labels."""
    X = golden["X"][::7]
    for kind in KINDS:
        sk = sklearn_from_spec(specs[kind])
        pred = quiet(sk.predict, X)
        exp = golden[f"{kind}.expected_label"][::7]
        if kind != "kmeans":
            exp = np.asarray(specs[kind]["classes"])[exp]
        assert np.array_equal(np.asarray(pred), exp)
