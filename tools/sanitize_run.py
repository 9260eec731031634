#!/usr/bin/env python
"""The bundled-row parity run for compute-sanitizer: every kernel family once on the reference's 7 653 rows
(KNeighbors and SVC through the tensor-core engine AND the fp64 kernels), results checked against the golden labels.
usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitize_run.py [kinds...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import KINDS, spec_from_golden
from traffic_classifier_sdn_b200 import _lib, from_spec

z = np.load(os.path.join(ROOT, "tests", "golden", "bundled.npz"), allow_pickle=False)
g = {k: z[k] for k in z.files}
X = np.ascontiguousarray(g["X"])
kinds = sys.argv[1:] or KINDS
for kind in kinds:
    spec = spec_from_golden(g, kind)
    est = from_spec(spec)
    idx, _ = est._run(X, True)                 # labels + scores (fp64 kernels for knn/svc scores)
    assert np.array_equal(idx, g[f"{kind}.expected_label"]), kind
    idx = est.predict_indices(X)               # labels only: engine for knn / svc, fp32 pre-pass for gnb needs float32 rows
    assert np.array_equal(idx, g[f"{kind}.expected_label"]), kind
    idx = est.predict_indices(X.astype(np.float32))
    print(f"{kind}: ok, stats={est.stats().tolist()}", flush=True)
    if kind in ("knn", "svc"):
        est.set_option(_lib.OPT_ENGINE, 1)
        assert np.array_equal(est.predict_indices(X[:600]), g[f"{kind}.expected_label"][:600]), kind
print("sanitize_run done")
