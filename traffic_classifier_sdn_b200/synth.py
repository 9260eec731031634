"""Synthetic D-ITG-shaped flow tables: rows with the structure the reference's own feature
derivation produces, for benchmarks and large-size tests (no dataset can be downloaded here).

Each row is one flow observed at one 1 s poll.  Per traffic class a forward and a reverse packet process is
drawn (rates and packet sizes follow the profiles of the reference's bundled captures: dns, game, ping,
telnet, voice; `quake` -- referenced by the notebooks but not bundled -- is given a Quake3-like profile from
D-IGT_scripts/quake_script_file), cumulative counters are integrated over the flow's age, and the twelve
features are then computed with the reference's formulas (traffic_classifier.py:63-96): delta = this poll's
increment, inst rate = delta / 1 s, avg rate = cumulative / age.  So the structural identities of real
captures hold: integer-valued delta/inst columns, inst_pps == delta_packets, 35-60 % idle polls for the
bursty classes, all-zero reverse features for `game`.
"""
from __future__ import annotations

import numpy as np

CLASSES = np.array(["dns", "game", "ping", "quake", "telnet", "voice"])
# per class: fwd pkts/s when active, fwd bytes/pkt (lo, hi), rev pkts/s, rev bytes/pkt (lo, hi), P(idle poll)
_PROFILE = {
    "dns":    (0.9,  (60, 90),    1.0,  (160, 360),  0.55),
    "game":   (22.0, (90, 280),   0.0,  (0, 0),      0.40),
    "ping":   (1.0,  (98, 98),    1.0,  (98, 98),    0.02),
    "quake":  (30.0, (60, 110),   20.0, (120, 320),  0.05),
    "telnet": (60.0, (40, 70),    55.0, (40, 80),    0.45),
    "voice":  (0.6,  (60, 90),    49.0, (150, 166),  0.03),
}
FEATURES_12 = ("dFp", "dFb", "FiPps", "FaPps", "FiBps", "FaBps", "dRp", "dRb", "RiPps", "RaPps", "RiBps", "RaBps")
# the 8-feature view used by BASELINE config 2 (inst rates equal the deltas at a 1 s poll, so they are dropped)
COLUMNS_8 = (0, 1, 3, 5, 6, 7, 9, 11)


def _direction(rng, n, pps, bpp, idle, age, burst, raw=False):
    """-> (delta_p, delta_b, avg_pps, avg_bps), or with raw=True the counters behind them: (hist_p, hist_b, cum_p, cum_b)"""
    if pps <= 0.0:
        z = np.zeros(n)
        return z, z, z, z
    active = rng.random(n) >= idle
    delta_p = np.where(active, rng.poisson(pps * burst, n), 0).astype(np.float64)
    lo, hi = bpp
    size = rng.integers(lo, hi + 1, n).astype(np.float64)
    delta_b = delta_p * size
    # history before this poll: the flow was active (1 - idle) of the time at its own mean rate
    hist_p = rng.poisson(np.maximum(pps * burst * (1.0 - idle) * (age - 1.0), 0.0)).astype(np.float64)
    hist_b = hist_p * (0.5 * (lo + hi)) + np.rint(rng.normal(0.0, 1.0, n) * np.sqrt(hist_p + 1.0) * (hi - lo) / 3.46)
    hist_b = np.maximum(hist_b, 0.0)
    cum_p = hist_p + delta_p
    cum_b = hist_b + delta_b
    if raw:
        return hist_p, hist_b, cum_p, cum_b
    return delta_p, delta_b, cum_p / age, cum_b / age


def make_flows(n: int, seed: int = 0, d: int = 12, dtype=np.float64, class_weights=None, return_labels=True, counters=False):
    """n synthetic flow rows -> (X [n,d], y [n] class index into CLASSES).  d is 12 or 8.
    counters=True returns instead the cumulative counters the rows derive from: (C [n, 9], y) with columns
    age, then per direction (forward, reverse): packets and bytes one poll ago, packets and bytes now -- the input of
    `make_flows_device`, which pushes them through the reference's own derivation on the GPU."""
    if d not in (8, 12):
        raise ValueError("d must be 12 (the models' feature count) or 8 (BASELINE config 2)")
    rng = np.random.default_rng(seed)
    w = np.full(len(CLASSES), 1.0 / len(CLASSES)) if class_weights is None else np.asarray(class_weights, float)
    y = rng.choice(len(CLASSES), size=n, p=w / w.sum()).astype(np.int32)
    X = np.empty((n, 9 if counters else 12), np.float64)
    age = rng.integers(1, 900, n).astype(np.float64)          # seconds since the flow appeared (15 min captures)
    burst = np.exp(rng.normal(0.0, 0.35, n))                  # per-flow rate heterogeneity
    for ci, name in enumerate(CLASSES):
        m = y == ci
        k = int(m.sum())
        if k == 0:
            continue
        fp, fb, rp, rb, idle = _PROFILE[name]
        a, b = age[m], burst[m]
        if counters:
            X[m] = np.column_stack((a,) + _direction(rng, k, fp, fb, idle, a, b, raw=True) + _direction(rng, k, rp, rb, idle, a, b, raw=True))
            continue
        dFp, dFb, FaP, FaB = _direction(rng, k, fp, fb, idle, a, b)
        dRp, dRb, RaP, RaB = _direction(rng, k, rp, rb, idle, a, b)
        X[m] = np.column_stack([dFp, dFb, dFp, FaP, dFb, FaB, dRp, dRb, dRp, RaP, dRb, RaB])
    if counters:
        return (X, y) if return_labels else X
    if d == 8:
        X = X[:, COLUMNS_8]
    X = np.ascontiguousarray(X, dtype=dtype)
    return (X, y) if return_labels else X


def make_flows_device(n: int, seed: int = 0, d: int = 12, dtype="float32", device=None, class_weights=None):
    """The same rows as `make_flows`, derived ON THE GPU by the reference's own feature derivation
    (``Flow.updateforward / updatereverse``, traffic_classifier.py:63-96 -> tcsdn_flow_update, csrc/flow.cu) instead of the
    closed form: every flow is created at t = 0 with zero counters, polled at t = age - 1 and again at t = age, once per
    direction.  Bit-identical to `make_flows` (tests/test_parity_gpu.py); returns a CUDA tensor [n, d]."""
    import torch
    from . import _lib
    from .flows import STATE
    C = make_flows(n, seed=seed, class_weights=class_weights, return_labels=False, counters=True)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    Cd = torch.from_numpy(C).to(dev)
    age = Cd[:, 0].contiguous()
    state = torch.zeros((n, STATE), dtype=torch.float64, device=dev)      # Flow.__init__ at t = 0, all counters 0
    fdt = torch.float32 if str(dtype).endswith("32") else torch.float64
    feat = torch.empty((n, 12), dtype=fdt, device=dev)
    lib = _lib.load()
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        for direction, base in ((0, 1), (1, 5)):
            dd = torch.full((n,), direction, dtype=torch.uint8, device=dev)
            for when, cols in ((age - 1.0, (base, base + 1)), (age, (base + 2, base + 3))):
                p, b = Cd[:, cols[0]].contiguous(), Cd[:, cols[1]].contiguous()
                _lib.check(lib.tcsdn_flow_update(state.data_ptr(), p.data_ptr(), b.data_ptr(), when.contiguous().data_ptr(),
                                                 dd.data_ptr(), n, feat.data_ptr(), _lib.F32 if fdt == torch.float32 else _lib.F64, st))
    if d == 8:
        feat = feat[:, list(COLUMNS_8)].contiguous()
    return feat


def _full_tree(rng, depth, d, n_classes, impure, scale, uniform=False):
    """Complete binary tree in preorder, vectorised: node i at level l has left = i+1, right = i + 2^(depth-l)."""
    n = (1 << (depth + 1)) - 1
    level = np.zeros(n, np.int32)
    # preorder levels: recursively  [l] + T(l+1) + T(l+1); build bottom-up
    pat = np.array([depth], np.int32)
    for l in range(depth - 1, -1, -1):
        pat = np.concatenate([[l], pat, pat]).astype(np.int32)
    level[:] = pat
    idx = np.arange(n, dtype=np.int64)
    leaf = level == depth
    left = np.where(leaf, -1, idx + 1).astype(np.int32)
    right = np.where(leaf, -1, idx + (1 << (depth - level).astype(np.int64))).astype(np.int32)
    feat = np.where(leaf, -2, rng.integers(0, d, n)).astype(np.int32)
    if uniform:   # thresholds uniform in [0, 1): rows uniform in [0, 1)^d spread over ALL leaves (cache-hostile by construction)
        thr = rng.random(n)
    else:
        thr = np.abs(rng.normal(0.0, 1.0, n)) * scale[np.maximum(feat, 0)] * 0.7 + np.where(rng.random(n) < 0.5, 0.5, 0.0)
    thr = np.where(leaf, -2.0, thr)
    val = np.full((n, n_classes), 1.0 / n_classes)
    nl = int(leaf.sum())
    lv = np.zeros((nl, n_classes))
    lv[np.arange(nl), rng.integers(0, n_classes, nl)] = 1.0
    imp = rng.random(nl) < impure
    k = int(imp.sum())
    if k:
        cnt = rng.integers(0, 5, (k, n_classes)).astype(float)
        cnt[np.arange(k), rng.integers(0, n_classes, k)] += 1.0
        lv[imp] = cnt / cnt.sum(axis=1, keepdims=True)
    val[leaf] = lv
    return left, right, feat, thr, val


def random_forest_spec(n_trees: int, depth: int, d: int = 12, n_classes: int = 6, seed: int = 0, full: bool = True,
                       impure: float = 0.05, scale=None, uniform: bool = False):
    """A synthetic forest spec (modelio layout) with random splits: `full` = complete binary trees of the given
    depth (the adversarial, cache-hostile forest of SURVEY 8d); otherwise ragged trees that stop early at random.
    `uniform` (full trees): thresholds uniform in [0, 1) -- with rows uniform in [0, 1)^d every leaf is reached, so the
    walk touches the WHOLE node array (flow-shaped rows follow a few paths only and leave most of a random tree cold)."""
    rng = np.random.default_rng(seed)
    scale = np.asarray(scale if scale is not None else [24, 2076, 24, 122, 2069, 9651, 27, 2977, 27, 17, 2974, 2441][:d], float)
    lefts, rights, feats, thrs, vals, offs = [], [], [], [], [], [0]
    for _ in range(n_trees):
        if full:
            l, r, f, t, v = _full_tree(rng, depth, d, n_classes, impure, scale, uniform)
            lefts.append(l); rights.append(r); feats.append(f); thrs.append(t); vals.append(v)
            offs.append(offs[-1] + len(l))
            continue
        left, right, feat, thr, val = [], [], [], [], []
        stack = [(0, -1, 0)]  # (level, parent, side) explicit preorder construction
        while stack:
            level, parent, side = stack.pop()
            i = len(left)
            if parent >= 0:
                (left if side == 0 else right)[parent] = i
            left.append(-1); right.append(-1)
            if level == depth or (level >= 2 and rng.random() < 0.25):
                v = np.zeros(n_classes)
                if rng.random() < impure:
                    cnt = rng.integers(0, 5, n_classes).astype(float)
                    cnt[rng.integers(n_classes)] += 1.0
                    v = cnt / cnt.sum()
                else:
                    v[rng.integers(n_classes)] = 1.0
                feat.append(-2); thr.append(-2.0); val.append(v)
                continue
            f = int(rng.integers(d))
            feat.append(f)
            thr.append(float(np.abs(rng.normal(0.0, 1.0)) * scale[f] * 0.7) + (0.5 if rng.random() < 0.5 else 0.0))
            val.append(np.full(n_classes, 1.0 / n_classes))
            stack.append((level + 1, i, 1))
            stack.append((level + 1, i, 0))
        lefts.append(np.asarray(left, np.int32)); rights.append(np.asarray(right, np.int32))
        feats.append(np.asarray(feat, np.int32)); thrs.append(np.asarray(thr, np.float64))
        vals.append(np.stack(val))
        offs.append(offs[-1] + len(left))
    classes = CLASSES[:n_classes].copy() if n_classes <= len(CLASSES) else np.arange(n_classes)
    return dict(kind="forest", tree_offsets=np.asarray(offs, np.int64), left=np.concatenate(lefts),
                right=np.concatenate(rights), feature=np.concatenate(feats), threshold=np.concatenate(thrs),
                value=np.ascontiguousarray(np.concatenate(vals)), classes=classes, n_features=d)
