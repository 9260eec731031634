#!/usr/bin/env python
"""GaussianNB: time the tiled kernel with (default) and without (TCSDN_GNB_PREPASS=0) the certified fp32 pre-pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from traffic_classifier_sdn_b200 import from_spec
w = bench.build_workload("gnb")
for rows in (1_000_000, 50_000_000):
    X = bench.synth_rows(rows, w["d"], seed=1000, device=torch.device("cuda", 0))
    out = torch.empty(rows, dtype=torch.int32, device="cuda")
    for mode in (0,):
        est = from_spec(w["spec"])
        for _ in range(3): est.predict_indices(X, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): est.predict_indices(X, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"rows={rows} prepass={os.environ.get('TCSDN_GNB_PREPASS', '1')}: {ms*1e3:.1f} us  {rows/ms*1e3:.3e} rows/s  refined={int(est.stats()[6])} of {13*rows}")
