"""Test infrastructure: rebuild a live scikit-learn estimator from a parameter spec.

scikit-learn (third-party; the reference pins no version, its pickles say 1.0.1, this image
ships 1.9.0) holds all arithmetic of the reference's hot path, so a
live sklearn estimator is the strongest checker we can have on any box.  Two of the
reference's pickles do not unpickle under sklearn 1.9, so they are rebuilt here from the
data-only spec that ``traffic_classifier_sdn_b200.modelio`` extracts (SURVEY.md 8c recipe).
Never imported by the product package.
"""
import warnings

import numpy as np


def sklearn_from_spec(spec, knn_algorithm="brute"):
    kind = spec["kind"]
    d = int(spec["n_features"])
    if kind == "linear":
        from sklearn.linear_model import LogisticRegression
        m = LogisticRegression()
        m.coef_ = spec["coef"].copy()
        m.intercept_ = spec["intercept"].copy()
        m.classes_ = np.asarray(spec["classes"])
        m.n_features_in_ = d
        return m
    if kind == "gnb":
        from sklearn.naive_bayes import GaussianNB
        m = GaussianNB()
        m.theta_ = spec["theta"].copy()
        m.var_ = spec["var"].copy()
        m.class_prior_ = spec["class_prior"].copy()
        m.classes_ = np.asarray(spec["classes"])
        m.n_features_in_ = d
        return m
    if kind == "kmeans":
        from sklearn.cluster import KMeans
        c = spec["centers"]
        m = KMeans(n_clusters=c.shape[0])
        m.cluster_centers_ = c.copy()
        m._n_features_out = c.shape[0]
        m._n_threads = 1
        m.n_features_in_ = d
        m.labels_ = np.zeros(1, np.int32)
        m.inertia_ = 0.0
        m.n_iter_ = 1
        return m
    if kind == "knn":
        from sklearn.neighbors import KNeighborsClassifier
        m = KNeighborsClassifier(n_neighbors=int(spec["k"]), algorithm=knn_algorithm)
        classes = np.asarray(spec["classes"])
        m.fit(spec["fit_X"], classes[spec["y"]])
        return m
    if kind == "svc":
        from sklearn.svm import SVC
        m = SVC(kernel="rbf", gamma=float(spec["gamma"]),
                decision_function_shape=spec.get("decision_function_shape", "ovr"),
                break_ties=bool(spec.get("break_ties", False)))
        nsv = spec["sv"].shape[0]
        C = len(spec["classes"])
        m.support_vectors_ = spec["sv"].copy()
        m.support_ = np.arange(nsv, dtype=np.int32)
        m._n_support = spec["n_support"].astype(np.int32)
        m._dual_coef_ = spec["dual_coef"].copy()
        m._intercept_ = spec["intercept"].copy()
        if C == 2:
            m.dual_coef_ = -m._dual_coef_
            m.intercept_ = -m._intercept_
        else:
            m.dual_coef_ = m._dual_coef_
            m.intercept_ = m._intercept_
        m._gamma = float(spec["gamma"])
        m._sparse = False
        m.classes_ = np.asarray(spec["classes"])
        m.class_weight_ = np.ones(C)
        m._probA = np.empty(0)
        m._probB = np.empty(0)
        m.fit_status_ = 0
        m.shape_fit_ = (nsv, d)
        m.n_features_in_ = d
        m._num_iter = np.zeros(C * (C - 1) // 2, np.int32)
        return m
    if kind == "forest":
        from sklearn.ensemble import RandomForestClassifier
        from sklearn.tree import DecisionTreeClassifier
        from sklearn.tree._tree import Tree, NODE_DTYPE
        classes = np.asarray(spec["classes"])
        C = len(classes)
        offs = spec["tree_offsets"]
        ests = []
        for t in range(len(offs) - 1):
            a, b = int(offs[t]), int(offs[t + 1])
            n = b - a
            nodes = np.zeros(n, dtype=NODE_DTYPE)
            nodes["left_child"] = spec["left"][a:b]
            nodes["right_child"] = spec["right"][a:b]
            nodes["feature"] = spec["feature"][a:b]
            nodes["threshold"] = spec["threshold"][a:b]
            nodes["n_node_samples"] = 1
            nodes["weighted_n_node_samples"] = 1.0
            values = np.ascontiguousarray(spec["value"][a:b].reshape(n, 1, C))
            tree = Tree(d, np.array([C], dtype=np.intp), 1)
            depth = _depth(spec["left"][a:b], spec["right"][a:b])
            tree.__setstate__({"max_depth": depth, "node_count": n, "nodes": nodes, "values": values})
            dt = DecisionTreeClassifier()
            dt.tree_ = tree
            dt.n_outputs_ = 1
            dt.n_classes_ = C
            dt.classes_ = np.arange(C, dtype=np.float64)
            dt.n_features_in_ = d
            dt.max_features_ = d
            ests.append(dt)
        m = RandomForestClassifier(n_estimators=len(ests))
        m.estimators_ = ests
        m.classes_ = classes
        m.n_classes_ = C
        m.n_outputs_ = 1
        m.n_features_in_ = d
        return m
    raise ValueError(kind)


def _depth(left, right):
    depth = np.zeros(len(left), np.int64)
    best = 0
    for i in range(len(left)):  # preorder or not, parents precede children in sklearn's builders
        if left[i] >= 0:
            depth[left[i]] = depth[i] + 1
            depth[right[i]] = depth[i] + 1
            best = max(best, int(depth[i]) + 1)
    return best


def quiet(fn, *a, **k):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)
