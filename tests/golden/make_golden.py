#!/usr/bin/env python
"""Generate tests/golden/bundled.npz from the reference tree (run in the build container only).

    python tests/golden/make_golden.py [/root/reference]

Inputs (read-only): the reference's five ``datasets/*_training_data.csv`` and its six
``models/*`` pickles.  The rows are assembled with the notebooks' recipe (SURVEY.md 8c:
ping, voice, dns, telnet tab-separated, game comma-separated, concat, dropna, drop the four
cumulative counters).  Expected outputs come from scikit-learn itself -- the library whose
``predict`` the reference calls at traffic_classifier.py:106 -- evaluated on estimators
revived from the pickles: the four that still unpickle are loaded with ``pickle.load`` and
must agree exactly with their rebuilt twins; KNeighbors / RandomForestClassifier are rebuilt
from the data-only spec (tests/sk_rebuild.py).  Nothing under /root/reference is copied as
source; the .npz holds numeric rows, fitted parameters and sklearn's answers.
"""
import os
import pickle
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from traffic_classifier_sdn_b200 import modelio  # noqa: E402
from sk_rebuild import sklearn_from_spec  # noqa: E402

DROP = ["Forward Packets", "Forward Bytes", "Reverse Packets", "Reverse Bytes"]


def load_bundled_rows(ref):
    frames = []
    for name, sep in (("ping", "\t"), ("voice", "\t"), ("dns", "\t"), ("telnet", "\t"), ("game", ",")):
        frames.append(pd.read_csv(os.path.join(ref, "datasets", f"{name}_training_data.csv"), delimiter=sep))
    df = pd.concat(frames, ignore_index=True).dropna()
    df = df.drop(columns=DROP)
    y = df["Traffic Type"].to_numpy().astype(str)
    X = df.drop(columns=["Traffic Type"]).to_numpy(dtype=np.float64)
    return np.ascontiguousarray(X), y


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    warnings.simplefilter("ignore")
    X, y = load_bundled_rows(ref)
    assert X.shape == (7653, 12), X.shape
    out = {"X": X, "y": y}
    for word, fname in modelio.MODEL_FILES.items():
        path = os.path.join(ref, "models", fname)
        spec = modelio.load_reference_pickle(path)
        kind = spec["kind"]
        for k, v in spec.items():
            if k == "kind":
                continue
            a = np.asarray(v)
            out[f"{kind}.{k}"] = a.astype(str) if a.dtype == object else a
        sk = sklearn_from_spec(spec)
        if kind == "knn":
            # one thread => sklearn's thread-count-independent `parallel_on_X` reduction (n > 4*256*1)
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=1)
        direct = None
        try:
            with open(path, "rb") as fh:
                direct = pickle.load(fh)
        except Exception as exc:  # KNeighbors, RandomForestClassifier under sklearn >= 1.3
            print(f"{fname}: pickle.load fails here ({type(exc).__name__}); using the rebuilt estimator")
        pred = sk.predict(X)
        if direct is not None:
            assert np.array_equal(np.asarray(direct.predict(X)), np.asarray(pred)), fname
        classes = np.asarray(spec["classes"])
        if kind == "kmeans":
            lab = np.asarray(pred, np.int32)
        else:
            lab = np.searchsorted(classes, pred).astype(np.int32)
            assert np.array_equal(classes[lab], pred)
        out[f"{kind}.expected_label"] = lab
        if kind == "linear":
            s = sk.decision_function(X)
            if direct is not None:
                assert np.array_equal(direct.decision_function(X), s)
            out["linear.expected_score"] = s
        elif kind == "gnb":
            s = sk._joint_log_likelihood(X)
            if direct is not None:
                assert np.array_equal(direct._joint_log_likelihood(X), s)
            out["gnb.expected_score"] = s
        elif kind == "kmeans":
            out["kmeans.expected_score"] = sk.transform(X)  # euclidean distances to centers
        elif kind == "knn":
            kd = sklearn_from_spec(spec, knn_algorithm="kd_tree")  # what 'auto' picks for 12 features
            assert np.array_equal(kd.predict(X), pred), "kd_tree and brute disagree"
            out["knn.expected_score"] = np.rint(sk.predict_proba(X) * spec["k"]).astype(np.uint8)
        elif kind == "svc":
            sk.decision_function_shape = "ovo"
            s = sk.decision_function(X)
            sk.decision_function_shape = spec["decision_function_shape"]
            if direct is not None:
                direct.decision_function_shape = "ovo"
                assert np.array_equal(direct.decision_function(X), s)
            out["svc.expected_score"] = s
            out["svc.expected_ovr"] = sk.decision_function(X)
        elif kind == "forest":
            out["forest.expected_score"] = sk.predict_proba(X)
        agree = float(np.mean(classes[lab] == y)) if kind != "kmeans" else float("nan")
        print(f"{fname}: kind={kind} labels ok, agreement with CSV labels {agree:.4f}")
    path = os.path.join(HERE, "bundled.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
