#!/usr/bin/env python
"""Device time of the KNN bench workload (10M x 50k, k = 5) with pruning on and off, and the engine's counters."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from traffic_classifier_sdn_b200 import from_spec, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
w = bench.build_workload("knn")
X = bench.synth_rows(n, w["d"], seed=3, device=torch.device("cuda", 0))
ref = None
for prune_off in (0, 1):
    est = from_spec(w["spec"])
    est.set_option(_lib.OPT_KNN_PRUNE, prune_off)
    out = est.predict_indices(X)
    torch.cuda.synchronize()
    s0 = est.stats()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(2):
        out = est.predict_indices(X)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 2
    st = est.stats()
    print(f"prune_off={prune_off}: {ms:.2f} ms  {n / ms * 1e3:.3e} rows/s  tiles/pass {st[4] / 1000:.1f}  "
          f"evals/query {st[3] / (3.0 * n):.1f}  tie rows {st[7]}  stats {st}")
    if ref is None:
        ref = out.clone()
    else:
        print("labels equal:", bool(torch.equal(ref, out)))
