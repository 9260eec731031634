#!/usr/bin/env python
"""Repeat small SVC engine runs (few tiles per pass: the regime where stage-release races show) and count bad ones."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, oracle
from test_engine_gpu import _svc_spec, _force
from traffic_classifier_sdn_b200 import from_spec, synth
nbad = tot = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    for nsv, C, nq in ((600, 4, 4200), (3000, 6, 5000), (513, 3, 4200)):
        spec = _svc_spec(nsv, C, seed=nsv)
        X = synth.make_flows(nq, seed=nsv + 5, return_labels=False)
        idx, dec = _force(from_spec(spec), 2)._run(X, True)
        rdec = oracle.svc(spec, X)[1]
        nbad += bool((np.abs(dec - rdec).max(axis=1) > 2e-3).any()); tot += 1
print("svc stress: runs with bad rows:", nbad, "of", tot)
