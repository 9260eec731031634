"""CPU: host-side logic -- model import, flow table, CLI shim, and the C-ABI library's surface."""
import ctypes
import io
import os
import re

import numpy as np
import pytest

import oracle
from conftest import ROOT, KINDS
from traffic_classifier_sdn_b200 import _lib, cli, flows, modelio


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert lib.tcsdn_version() == 100
    header = open(os.path.join(ROOT, "include", "tcsdn.h")).read()
    declared = set(re.findall(r"\b(tcsdn_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tcsdn.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device create() must fail loudly (there is no CPU path)."""
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    from traffic_classifier_sdn_b200 import from_spec
    spec = dict(kind="kmeans", centers=np.zeros((2, 3)), classes=np.arange(2), n_features=3)
    with pytest.raises(_lib.TcsdnError, match="no CUDA device|CPU"):
        from_spec(spec)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "traffic_classifier_sdn_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "tcsdn_oracle" not in src, f


def test_data_only_unpickler_rejects_code():
    import pickle
    evil = pickle.dumps(os.system)
    with pytest.raises(pickle.UnpicklingError):
        modelio.spec_from_pickle_bytes(evil)


def test_specs_from_golden_are_complete(specs):
    assert set(specs) == set(KINDS)
    assert specs["svc"]["sv"].shape == (2281, 12) and specs["svc"]["n_support"].sum() == 2281
    assert specs["forest"]["tree_offsets"][-1] == 5306 and len(specs["forest"]["tree_offsets"]) == 101
    assert specs["knn"]["fit_X"].shape == (4448, 12) and specs["knn"]["k"] == 5
    v = specs["forest"]["value"]
    assert np.allclose(v.sum(axis=1), 1.0)  # counts of the 1.0.1 pickle were normalised to fractions


def test_spec_from_live_estimators_roundtrip():
    from sklearn.ensemble import RandomForestClassifier
    rng = np.random.default_rng(0)
    X = rng.random((300, 5)); y = rng.integers(0, 3, 300)
    rf = RandomForestClassifier(n_estimators=7, max_depth=5, random_state=0).fit(X, y)
    spec = modelio.spec_from_estimator(rf)
    lab, p = oracle.forest(spec, X)
    assert np.array_equal(p, rf.predict_proba(X))


# ---------------------------------------------------------------- flow table vs the reference's formulas
class _RefFlow:
    """Test-local restatement of reference traffic_classifier.py:29-96 (one direction) for differential tests."""
    def __init__(self, t0, packets, nbytes, active):
        self.t0, self.p, self.b = t0, packets, nbytes
        self.dp = self.db = 0
        self.ipps = self.apps = self.ibps = self.abps = 0.0
        self.last = t0
        self.status = "ACTIVE" if active else "INACTIVE"

    def update(self, packets, nbytes, now):
        self.dp = packets - self.p; self.p = packets
        if now != self.t0: self.apps = packets / float(now - self.t0)
        if now != self.last: self.ipps = self.dp / float(now - self.last)
        self.db = nbytes - self.b; self.b = nbytes
        if now != self.t0: self.abps = nbytes / float(now - self.t0)
        if now != self.last: self.ibps = self.db / float(now - self.last)
        self.last = now
        self.status = "INACTIVE" if (self.db == 0 or self.dp == 0) else "ACTIVE"


def _monitor_log(rng, n_flows=6, polls=25):
    macs = [f"00:00:00:00:00:{i:02x}" for i in range(1, 2 * n_flows + 1)]
    counters = {}
    lines = [b"time\tdatapath\tin-port\teth-src\teth-dst\tout-port\ttotal_packets\ttotal_bytes\n"]
    t = 1_600_000_000
    for _ in range(polls):
        t += int(rng.integers(0, 3))  # repeated timestamps exercise the `!=` guards
        for f in range(n_flows):
            for (src, dst) in ((macs[2 * f], macs[2 * f + 1]), (macs[2 * f + 1], macs[2 * f])):
                if rng.random() < 0.15:
                    continue
                p, b = counters.get((src, dst), (0, 0))
                dp = int(rng.integers(0, 40)) if rng.random() < 0.7 else 0
                p, b = p + dp, b + dp * int(rng.integers(60, 1500))
                counters[(src, dst)] = (p, b)
                lines.append(b"data\t%d\t1\t%x\t%s\t%s\t%x\t%d\t%d\n" % (t, f + 1, src.encode(), dst.encode(), f + 2, p, b))
        lines.append(b"some unrelated ryu log line\n")
    return lines


def test_flow_table_matches_reference_formulas():
    rng = np.random.default_rng(11)
    lines = _monitor_log(rng)
    table = flows.FlowTable()
    ref = {}
    order = []
    for ln in lines:
        rec = flows.parse_monitor_line(ln)
        if rec is None:
            continue
        table.ingest(rec)
        t, dp, _in, src, dst, _out, p, b = rec
        if (dp, src, dst) in ref:
            ref[(dp, src, dst)][0].update(p, b, t)
        elif (dp, dst, src) in ref:
            ref[(dp, dst, src)][1].update(p, b, t)
        else:
            ref[(dp, src, dst)] = (_RefFlow(t, p, b, True), _RefFlow(t, 0, 0, False))
            order.append((dp, src, dst))
    feats = table.features()
    assert feats.shape == (len(order), 12)
    for i, key in enumerate(order):
        fw, rv = ref[key]
        exp = [fw.dp, fw.db, fw.ipps, fw.apps, fw.ibps, fw.abps, rv.dp, rv.db, rv.ipps, rv.apps, rv.ibps, rv.abps]
        assert np.array_equal(feats[i], np.asarray(exp, float)), (i, feats[i], exp)
    for (fid, src, dst, fs, rs), key in zip(table.rows(), order):
        assert (fs, rs) == (ref[key][0].status, ref[key][1].status)
    # oracle's C restatement of the same update (used by the GPU test of tcsdn_flow_update)
    st = np.zeros(9); st[8] = 5.0
    out = oracle.flow_update(st, 5.0, 10, 900, 7.0)
    blk = np.zeros(9); blk[8] = 5.0
    flows.update_direction(blk, 5.0, 10, 900, 7.0)
    assert np.array_equal(out, blk)


def test_training_lines_format():
    table = flows.FlowTable()
    table.ingest((100, "1", "1", "aa", "bb", "2", 3, 300))
    table.ingest((101, "1", "2", "bb", "aa", "1", 7, 760))
    line = next(table.training_lines("dns"))
    assert line == "3\t300\t0\t0\t0.0\t0.0\t0.0\t0.0\t7\t760\t7\t760\t7.0\t7.0\t760.0\t760.0\tdns\n"
    assert flows.TRAINING_HEADER.count("\t") == 16


class _FakeModel:
    classes_ = np.array(["dns", "voice"])
    def predict(self, X):
        X = np.asarray(X)
        assert X.ndim == 2 and X.shape[1] == 12
        return self.classes_.take((X[:, 0] > 0).astype(int))


def test_cli_run_monitor_and_table():
    rng = np.random.default_rng(3)
    stream = io.BytesIO(b"".join(_monitor_log(rng, n_flows=2, polls=12)))
    out = io.StringIO()
    table = cli.run_monitor(stream, model=_FakeModel(), every=10, out=out)
    text = out.getvalue()
    assert len(table) == 2 and "Traffic Type" in text and "Flow ID" in text
    assert text.count("+--") > 0 and ("dns" in text or "voice" in text)
    rows = cli.classify_table(table, type("K", (), {"predict": lambda self, X: np.array([0, 3], np.int32)})())
    assert [r[3] for r in rows] == ["dns", "quake"]  # reference's fixed int map :109-114


def test_cli_words(capsys):
    assert cli.main([]) == 0
    assert "ERROR: Incorrect # of args" in capsys.readouterr().out
    assert cli.main(["nonsense"]) == 0
    assert "Unknown subcommand" in capsys.readouterr().out
    assert cli.main(["train"]) == 0
    assert "specify traffic type" in capsys.readouterr().out
    assert cli.ALIASES["supervised"] == "logistic" and cli.ALIASES["unsupervised"] == "kmeans"
    assert set(modelio.MODEL_FILES) >= {"logistic", "kmeans", "svm", "kneighbors", "Randomforest", "gaussiannb"}


def test_logistic_proba_forms_match_sklearn():
    """multinomial -> softmax, binary -> sigmoid (sk:linear_model/_logistic.py:1620-1625), one-vs-rest models ->
    normalised sigmoids (sk:linear_model/_base.py:429-451); the spec records which one the model uses."""
    from sklearn.linear_model import LogisticRegression
    from sklearn.multiclass import OneVsRestClassifier
    from traffic_classifier_sdn_b200.estimators import lr_proba
    rng = np.random.default_rng(0)
    X = rng.normal(size=(300, 5)); y = rng.integers(0, 4, 300)
    sk = LogisticRegression(max_iter=200).fit(X, y)
    spec = modelio.spec_from_estimator(sk)
    assert spec["ovr"] is False
    assert np.allclose(lr_proba(sk.decision_function(X), False), sk.predict_proba(X), rtol=0, atol=1e-15)
    skb = LogisticRegression().fit(X, y % 2)
    assert np.allclose(lr_proba(skb.decision_function(X)[:, None], False), skb.predict_proba(X), rtol=0, atol=1e-15)
    # a one-vs-rest model: scikit-learn's own formula, reached through LinearClassifierMixin._predict_proba_lr
    ovr = OneVsRestClassifier(LogisticRegression()).fit(X, y)
    dec = np.column_stack([e.decision_function(X) for e in ovr.estimators_])
    sk.coef_ = np.vstack([e.coef_ for e in ovr.estimators_]); sk.intercept_ = np.concatenate([e.intercept_ for e in ovr.estimators_])
    assert np.allclose(lr_proba(dec.copy(), True), sk._predict_proba_lr(X), rtol=0, atol=1e-15)
    sk.multi_class = "ovr"            # attribute of scikit-learn <= 1.6 estimators (the reference's pickles are 1.0.1)
    assert modelio.spec_from_estimator(sk)["ovr"] is True
    assert np.allclose(lr_proba(np.full((2, 3), -1e4), True), 1.0 / 3)   # all-zero sigmoids -> uniform


def test_take_labels_matches_numpy_take():
    """tcsdn_take_labels (the host tail of predict on large batches) == classes_.take(idx) for fixed-width label dtypes;
    out-of-range indices are an error, not a wild read"""
    lib = _lib.load()
    rng = np.random.default_rng(3)
    for cls in (np.array(["dns", "game", "ping", "quake", "telnet", "voice"]), np.arange(7, dtype=np.int64), np.array([b"a", b"bcd"])):
        idx = rng.integers(0, len(cls), 300_001).astype(np.int32)
        out = np.empty(len(idx), dtype=cls.dtype)
        for threads in (1, 5):
            out[:] = cls[0]
            assert lib.tcsdn_take_labels(_lib.ptr(idx), len(idx), _lib.ptr(cls), len(cls), cls.dtype.itemsize, _lib.ptr(out), threads) == 0
            assert np.array_equal(out, cls.take(idx))
        idx[1234] = len(cls)
        assert lib.tcsdn_take_labels(_lib.ptr(idx), len(idx), _lib.ptr(cls), len(cls), cls.dtype.itemsize, _lib.ptr(out), 3) == _lib.EINVAL
    assert lib.tcsdn_take_labels(None, 0, None, 1, 4, None, 1) == 0


def test_synthetic_counters_are_consistent_with_the_closed_form():
    """synth.make_flows(counters=True) exposes the cumulative counters behind the rows (the input of the GPU generator
    make_flows_device): deltas, instantaneous and average rates recomputed from them equal the closed-form rows"""
    from traffic_classifier_sdn_b200 import synth
    X, y = synth.make_flows(5000, seed=3)
    Cn, y2 = synth.make_flows(5000, seed=3, counters=True)
    assert np.array_equal(y, y2) and Cn.shape == (5000, 9)
    age = Cn[:, 0]
    for base, cols in ((1, (0, 1, 2, 3, 4, 5)), (5, (6, 7, 8, 9, 10, 11))):
        dp, db = Cn[:, base + 2] - Cn[:, base], Cn[:, base + 3] - Cn[:, base + 1]
        ref = np.column_stack([dp, db, dp, Cn[:, base + 2] / age, db, Cn[:, base + 3] / age])
        assert np.array_equal(ref, X[:, cols])
