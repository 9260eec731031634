#!/usr/bin/env python
"""CPU simulation that sized the KNN engine's pruning: for the bench workload (50k training rows, k = 5) it builds the
kd order with tile-level or 8-row leaves, sorts a sample of queries by leaf, and walks passes of 512 / 256 rows nearest tile
first with the engine's skip rule and a stale H (the producer runs a few tiles ahead of the epilogue).  Prints the tiles a pass
multiplies and the tiles a warp filters.  Pure numpy, a few minutes; usage: tools/knn_prune_sim.py [n_queries]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

N = 64
w = bench.build_workload("knn")
T = np.asarray(w["spec"]["fit_X"], dtype=np.float64)
k, d = w["spec"]["k"], T.shape[1]
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000


def build(idx, leafsize):
    if len(idx) <= leafsize:
        return ("leaf", idx)
    sub = T[idx]
    j = int(np.argmax(sub.max(0) - sub.min(0)))
    mid = ((len(idx) // 2 + N - 1) // N) * N if len(idx) > N else len(idx) // 2
    part = np.argpartition(sub[:, j], mid)
    return ("node", j, sub[part[mid], j], build(idx[part[:mid]], leafsize), build(idx[part[mid:]], leafsize))


def leaves(t, out):
    if t[0] == "leaf":
        out.append(t[1])
    else:
        leaves(t[3], out); leaves(t[4], out)


def descend(t, X, ids, out, counter):
    if t[0] == "leaf":
        out[ids] = counter[0]; counter[0] += 1
        return
    m = X[ids, t[1]] < t[2]
    descend(t[3], X, ids[m], out, counter); descend(t[4], X, ids[~m], out, counter)


X = np.asarray(bench.synth_rows(nq, d, seed=5), dtype=np.float64)
for leafsize in (64, 8):
    tree = build(np.arange(len(T)), leafsize)
    lv = []
    leaves(tree, lv)
    order = np.concatenate(lv)
    nt = (len(order) + N - 1) // N
    tiles = [order[i * N:(i + 1) * N] for i in range(nt)]
    cent = np.array([T[t].mean(0) for t in tiles])
    rad = np.array([np.sqrt(((T[t] - c) ** 2).sum(1).max()) for t, c in zip(tiles, cent)])
    pos = np.empty(len(T), dtype=np.int64)
    pos[order] = np.arange(len(T))
    leaf_tile = np.array([pos[l[0]] // N for l in lv])
    key = np.empty(nq, dtype=np.int64)
    descend(tree, X, np.arange(nq), key, [0])
    qs = np.lexsort((np.random.default_rng(0).random(nq), key))
    Xs, ks = X[qs], key[qs]
    for P in (512, 256):
        def simulate(p, lag=6):
            Q = Xs[p * P:(p + 1) * P]
            x0 = Q[0]
            e = np.sqrt(((Q - x0) ** 2).sum(1))
            rho = e.max()
            home = leaf_tile[ks[p * P]]
            tord = np.argsort(np.sqrt(((cent - cent[home]) ** 2).sum(1)))
            gap = np.sqrt(((cent - x0) ** 2).sum(1)) - rad
            best = np.full((P, k), np.inf)
            hist, pv, wv = [], 0, 0
            for t in tord:
                H = hist[len(hist) - lag] if len(hist) >= lag else np.inf
                g = gap[t] - rho
                if g > 0 and g * g > H:
                    continue
                pv += 1
                m = gap[t] - e
                far = (m > 0) & (m * m > best[:, -1])
                wv += (~far).reshape(-1, 32).any(1).sum()
                d2 = ((Q[:, None, :] - T[tiles[t]][None, :, :]) ** 2).sum(2)
                best = np.sort(np.concatenate([best, d2], 1), 1)[:, :k]
                hist.append(best[:, -1].max())
            return pv, wv
        npass = nq // P
        res = [simulate(p) for p in range(0, npass, max(1, npass // 40))]
        pv = np.array([r[0] for r in res]); wv = np.array([r[1] for r in res])
        print(f"leaves of <= {leafsize} rows, {P}-row passes: {pv.mean():.1f} of {nt} tiles multiplied per pass (median {np.median(pv):.0f}); "
              f"{wv.mean() / (P / 32):.1f} tiles filtered per warp", flush=True)
