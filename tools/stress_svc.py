#!/usr/bin/env python
"""Repeat small SVC engine runs (few tiles per pass: the regime where stage-release races show) and count bad ones:
the engine's raw decision values (audit option 3) must stay inside the certificate's bounds (option 4) of the fp64 values
on every run -- a stage refilled under a warp's coefficient loads shows up as whole warps far outside them."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracle
from test_engine_gpu import _svc_spec, _force
from traffic_classifier_sdn_b200 import from_spec, synth
nbad = tot = 0
cases = []
for nsv, C, nq in ((600, 4, 4200), (3000, 6, 5000), (513, 3, 4200)):
    spec = _svc_spec(nsv, C, seed=nsv)
    X = synth.make_flows(nq, seed=nsv + 5, return_labels=False)
    ridx, rdec = oracle.svc(spec, X)
    bound = _force(from_spec(spec), 4)._run(X, True)[1]
    cases.append((spec, X, ridx, rdec, bound))
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    for spec, X, ridx, rdec, bound in cases:
        raw = _force(from_spec(spec), 3)._run(X, True)[1]
        lab = _force(from_spec(spec), 2).predict_indices(X)
        nbad += bool((np.abs(raw - rdec) > bound).any()) or not np.array_equal(lab, ridx)
        tot += 1
print("svc stress: runs with bad rows:", nbad, "of", tot)
