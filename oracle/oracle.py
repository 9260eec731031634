"""ctypes front end of oracle/tcsdn_oracle.c (TEST INFRASTRUCTURE, see that file's header).

Every function takes a parameter spec (traffic_classifier_sdn_b200.modelio) plus float64 rows and
returns (label_index[int32], scores[float64]) computed by the C restatement on the host.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


def build(force=False):
    so = os.path.join(_HERE, "libtcsdn_oracle.so")
    src = os.path.join(_HERE, "tcsdn_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.tcsdn_oracle_threads.restype = C.c_int
    return _LIB


def threads():
    return int(lib().tcsdn_oracle_threads())


def set_threads(n):
    lib().tcsdn_oracle_set_threads(C.c_int(int(n)))


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def predict(spec, X, want_scores=True):
    """Dispatch on spec['kind'] -> (labels int32 [n], scores f64 or None)."""
    return {"linear": linear, "gnb": gnb, "kmeans": kmeans, "knn": knn, "svc": svc,
            "forest": forest}[spec["kind"]](spec, X, want_scores)


def linear(spec, X, want_scores=True):
    X = _d(X); n, d = X.shape
    coef, icpt = _d(spec["coef"]), _d(spec["intercept"])
    R = coef.shape[0]
    lab = np.empty(n, np.int32)
    sc = np.empty((n, R)) if want_scores else None
    lib().tcsdn_oracle_linear(_p(X, _f64p), C.c_int64(n), C.c_int32(d), _p(coef, _f64p), _p(icpt, _f64p),
                              C.c_int32(R), _p(sc, _f64p), _p(lab, _i32p))
    return lab, sc


def gnb(spec, X, want_scores=True):
    X = _d(X); n, d = X.shape
    th, var, pr = _d(spec["theta"]), _d(spec["var"]), _d(spec["class_prior"])
    Cn = th.shape[0]
    lab = np.empty(n, np.int32)
    sc = np.empty((n, Cn)) if want_scores else None
    lib().tcsdn_oracle_gnb(_p(X, _f64p), C.c_int64(n), C.c_int32(d), _p(th, _f64p), _p(var, _f64p),
                           _p(pr, _f64p), C.c_int32(Cn), _p(sc, _f64p), _p(lab, _i32p))
    return lab, sc


def kmeans(spec, X, want_scores=True):
    X = _d(X); n, d = X.shape
    ctr = _d(spec["centers"])
    k = ctr.shape[0]
    lab = np.empty(n, np.int32)
    sc = np.empty((n, k)) if want_scores else None
    lib().tcsdn_oracle_kmeans(_p(X, _f64p), C.c_int64(n), C.c_int32(d), _p(ctr, _f64p), C.c_int32(k),
                              _p(sc, _f64p), _p(lab, _i32p))
    return lab, sc


def knn(spec, X, want_scores=True, return_neighbors=False):
    X = _d(X); n, d = X.shape
    fx = _d(spec["fit_X"])
    y = np.ascontiguousarray(spec["y"], dtype=np.int32)
    k = int(spec["k"]); Cn = len(spec["classes"])
    lab = np.empty(n, np.int32)
    cnt = np.empty((n, Cn), np.int32)
    nbr = np.empty((n, k), np.int64) if return_neighbors else None
    lib().tcsdn_oracle_knn(_p(X, _f64p), C.c_int64(n), C.c_int32(d), _p(fx, _f64p), _p(y, _i32p),
                           C.c_int64(fx.shape[0]), C.c_int32(k), C.c_int32(Cn), _p(lab, _i32p),
                           _p(nbr, _i64p), _p(cnt, _i32p))
    sc = cnt.astype(np.float64) / k if want_scores else None
    if return_neighbors:
        return lab, sc, nbr
    return lab, sc


def svc(spec, X, want_scores=True):
    X = _d(X); n, d = X.shape
    sv, dual, icpt = _d(spec["sv"]), _d(spec["dual_coef"]), _d(spec["intercept"])
    nsup = np.ascontiguousarray(spec["n_support"], dtype=np.int32)
    Cn = len(nsup); P = Cn * (Cn - 1) // 2
    lab = np.empty(n, np.int32)
    dec = np.empty((n, P)) if want_scores else None
    lib().tcsdn_oracle_svc(_p(X, _f64p), C.c_int64(n), C.c_int32(d), _p(sv, _f64p), _p(dual, _f64p),
                           _p(icpt, _f64p), _p(nsup, _i32p), C.c_int32(sv.shape[0]), C.c_int32(Cn),
                           C.c_double(float(spec["gamma"])), _p(dec, _f64p), _p(lab, _i32p))
    return lab, dec


def ovr_from_ovo(dec, n_classes):
    dec = _d(dec); n = dec.shape[0]
    out = np.empty((n, n_classes))
    lib().tcsdn_oracle_ovr_from_ovo(_p(dec, _f64p), C.c_int64(n), C.c_int32(n_classes), _p(out, _f64p))
    return out


def forest(spec, X, want_scores=True, return_visits=False):
    X = _d(X); n, d = X.shape
    offs = np.ascontiguousarray(spec["tree_offsets"], dtype=np.int64)
    left = np.ascontiguousarray(spec["left"], np.int32); right = np.ascontiguousarray(spec["right"], np.int32)
    feat = np.ascontiguousarray(spec["feature"], np.int32)
    thr, val = _d(spec["threshold"]), _d(spec["value"])
    Cn = val.shape[1]
    lab = np.empty(n, np.int32)
    pr = np.empty((n, Cn)) if want_scores else None
    vis = C.c_int64(0)
    lib().tcsdn_oracle_forest(_p(X, _f64p), C.c_int64(n), C.c_int32(d), _p(offs, _i64p), _p(left, _i32p),
                              _p(right, _i32p), _p(feat, _i32p), _p(thr, _f64p), _p(val, _f64p),
                              C.c_int32(len(offs) - 1), C.c_int32(Cn), _p(pr, _f64p), _p(lab, _i32p),
                              C.byref(vis))
    if return_visits:
        return lab, pr, int(vis.value)
    return lab, pr


def flow_update(state, time_start, packets, nbytes, curr_time):
    st = _d(state).copy()
    lib().tcsdn_oracle_flow_update(_p(st, _f64p), C.c_double(time_start), C.c_double(packets),
                                   C.c_double(nbytes), C.c_double(curr_time))
    return st
