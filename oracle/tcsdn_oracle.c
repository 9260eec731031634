/*
 * tcsdn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the six scikit-learn predict paths that the reference
 * calls at traffic_classifier.py:106 (`label = model.predict(features.tolist())`).  It is
 * the checker for the CUDA library: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product package never does.
 *
 * The arithmetic of this path lives in a third-party dependency that is NOT under
 * /root/reference: scikit-learn (reference pins nothing; its pickles record 1.0.1; this image
 * ships 1.9.0).  Every function cites the sklearn source it follows as
 *   sk:<path>:<lines>  ==  site-packages/sklearn/<path>.
 * Pinning: tests/test_oracle.py checks every function here against (a) tests/golden/
 * bundled.npz -- sklearn's own answers for the reference's six pickles on the reference's
 * 7 653 bundled CSV rows, produced by tests/golden/make_golden.py -- and (b) live sklearn on
 * seeded random models.  The reference itself holds no tests for this path.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 * -ffp-contract=off matters: sklearn's wheels are baseline x86-64 (no FMA contraction).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int tcsdn_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void tcsdn_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* numpy's pairwise summation for a contiguous double vector
 * (numpy/_core/src/umath/loops_utils.h.src: @TYPE@_pairwise_sum, PW_BLOCKSIZE = 128).
 * GaussianNB reduces with np.sum(axis=1) over C-contiguous rows, which lands here. */
static double np_pairwise_sum(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        /* numpy starts from -0.0 to preserve -0 sums; irrelevant for value equality */
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

/* ------------------------------------------------------------------------------------
 * a1 LogisticRegression.predict
 *   sk:linear_model/_base.py:366-394 decision_function: scores = X @ coef_.T + intercept_
 *   sk:linear_model/_base.py:398-427 predict: 1 row of coef -> (score > 0), else argmax (first max)
 * The BLAS summation order inside dgemm is unspecified; this restatement sums j ascending.
 * scores: [n,R]; labels: class index.
 * ---------------------------------------------------------------------------------- */
void tcsdn_oracle_linear(const double *X, int64_t n, int32_t d, const double *coef,
                         const double *intercept, int32_t R, double *scores, int32_t *labels) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const double *x = X + i * d;
        double best = 0.0;
        int32_t arg = 0;
        for (int32_t c = 0; c < R; c++) {
            double s = 0.0;
            for (int32_t j = 0; j < d; j++) s += x[j] * coef[(int64_t)c * d + j];
            s += intercept[c];
            if (scores) scores[i * R + c] = s;
            if (c == 0 || s > best) { best = s; arg = c; }
        }
        if (R == 1) arg = best > 0.0 ? 1 : 0;
        if (labels) labels[i] = arg;
    }
}

/* ------------------------------------------------------------------------------------
 * a2 GaussianNB.predict
 *   sk:naive_bayes.py:533-545 _joint_log_likelihood, association order kept exactly:
 *     jointi = log(prior_i)
 *     n_ij   = -0.5 * sum_j log(2*pi*var_ij)
 *     n_ij   = n_ij - 0.5 * sum_j ((x_j - theta_ij)**2 / var_ij)      (np.sum -> pairwise)
 *     jll_i  = jointi + n_ij
 *   sk:naive_bayes.py:96-117 predict: argmax (first max).
 * ---------------------------------------------------------------------------------- */
void tcsdn_oracle_gnb(const double *X, int64_t n, int32_t d, const double *theta, const double *var,
                      const double *prior, int32_t C, double *jll, int32_t *labels) {
    double *cst = (double *)malloc(sizeof(double) * (size_t)C);
    double *logp = (double *)malloc(sizeof(double) * (size_t)C);
    double *tmp0 = (double *)malloc(sizeof(double) * (size_t)d);
    for (int32_t c = 0; c < C; c++) {
        for (int32_t j = 0; j < d; j++) tmp0[j] = log(2.0 * M_PI * var[(int64_t)c * d + j]);
        cst[c] = -0.5 * np_pairwise_sum(tmp0, d);
        logp[c] = log(prior[c]);
    }
    free(tmp0);
#pragma omp parallel
    {
        double *tmp = (double *)malloc(sizeof(double) * (size_t)d);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; i++) {
            const double *x = X + i * d;
            double best = 0.0;
            int32_t arg = 0;
            for (int32_t c = 0; c < C; c++) {
                for (int32_t j = 0; j < d; j++) {
                    double df = x[j] - theta[(int64_t)c * d + j];
                    tmp[j] = (df * df) / var[(int64_t)c * d + j];
                }
                double nij = cst[c] - 0.5 * np_pairwise_sum(tmp, d);
                double s = logp[c] + nij;
                if (jll) jll[i * C + c] = s;
                if (c == 0 || s > best) { best = s; arg = c; }
            }
            if (labels) labels[i] = arg;
        }
        free(tmp);
    }
    free(cst);
    free(logp);
}

/* ------------------------------------------------------------------------------------
 * a3 KMeans.predict
 *   sk:cluster/_kmeans.py:1075-1107 -> _labels_inertia :761 -> lloyd_iter_chunked_dense
 *   sk:cluster/_k_means_lloyd.pyx:168-213 _update_chunk_dense:
 *     pd[i,j] = ||c_j||^2 ; pd += -2 * X.C^T (dgemm) ; argmin with strict '<' (first min)
 *   ||c_j||^2 = row_norms(centers, squared=True) = einsum('ij,ij->i') (sk:utils/extmath.py).
 * scores (optional): [n,k] the partial squared distance ||c||^2 - 2 x.c .
 * ---------------------------------------------------------------------------------- */
void tcsdn_oracle_kmeans(const double *X, int64_t n, int32_t d, const double *centers, int32_t k,
                         double *scores, int32_t *labels) {
    double *cn = (double *)malloc(sizeof(double) * (size_t)k);
    for (int32_t c = 0; c < k; c++) {
        double s = 0.0;
        for (int32_t j = 0; j < d; j++) s += centers[(int64_t)c * d + j] * centers[(int64_t)c * d + j];
        cn[c] = s;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const double *x = X + i * d;
        double best = 0.0;
        int32_t arg = 0;
        for (int32_t c = 0; c < k; c++) {
            double dot = 0.0;
            for (int32_t j = 0; j < d; j++) dot += x[j] * centers[(int64_t)c * d + j];
            double s = cn[c] + (-2.0) * dot;
            if (scores) scores[i * k + c] = s;
            if (c == 0 || s < best) { best = s; arg = c; }
        }
        if (labels) labels[i] = arg;
    }
    free(cn);
}

/* ------------------------------------------------------------------------------------
 * a4 KNeighborsClassifier.predict, algorithm='brute', weights='uniform', euclidean, k neighbours
 *   sk:neighbors/_classification.py:245-312 predict.  For the euclidean metric ArgKminClassMode is
 *   excluded (sk:metrics/_pairwise_distances_reduction/_dispatcher.py valid_metrics), so predict goes
 *   kneighbors (sk:neighbors/_base.py:820-880) -> ArgKmin.compute -> the heap reduction
 *   sk:metrics/_pairwise_distances_reduction/_argkmin.pyx.tp:143-169: for every query i and every
 *   train row j ascending: heap_push(heap_i, dist(x_i, y_j), j), heaps initialised to DBL_MAX;
 *   then mode of the k neighbour classes (scipy.stats.mode: smallest class on a vote tie) :302.
 *   sk:utils/_heap.pyx:6-88 heap_push on a k-slot max-heap (rejects val >= root; sift-down prefers
 *     the left child on ties) -- restated verbatim in behaviour, because WHICH of several
 *     equidistant rows survives depends on it.
 *   Strategy: this is sklearn's `parallel_on_X` schedule (chosen when n_queries > 4*256*n_threads,
 *     sk:metrics/_pairwise_distances_reduction/_base.pyx.tp), the only one whose result does not
 *     depend on the thread count; `parallel_on_Y` merges per-thread heaps and resolves boundary
 *     ties differently for every thread count.
 *   Distance: sklearn's float64 euclidean specialisation evaluates ||x||^2 - 2 x.y + ||y||^2 with a
 *     dgemm middle term (_middle_term_computer.pyx.tp), whose rounding depends on the BLAS build.
 *     The restatement uses the definition it approximates -- sum_j (x_j - y_j)^2 in fp64, j ascending
 *     (sk:metrics/_dist_metrics.pxd.tp:39-49 euclidean_rdist) -- identical whenever the arithmetic
 *     is exact (integer-valued features) and otherwise equal up to dgemm cancellation error, which
 *     can only reorder neighbours whose squared distances agree to ~1e-16 * (||x||^2 + ||y||^2).
 * nbr_idx (optional) [n,k]: heap content as left by the pushes (heap order, not sorted).
 * counts (optional) [n,C]: neighbour class histogram (proba * k).
 * ---------------------------------------------------------------------------------- */
static inline void heap_push(double *values, int64_t *indices, int64_t size, double val, int64_t val_idx) {
    int64_t current_idx, left_child_idx, right_child_idx, swap_idx;
    if (val >= values[0]) return;
    values[0] = val;
    indices[0] = val_idx;
    current_idx = 0;
    for (;;) {
        left_child_idx = 2 * current_idx + 1;
        right_child_idx = left_child_idx + 1;
        if (left_child_idx >= size) {
            break;
        } else if (right_child_idx >= size) {
            if (values[left_child_idx] > val) swap_idx = left_child_idx; else break;
        } else if (values[left_child_idx] >= values[right_child_idx]) {
            if (val < values[left_child_idx]) swap_idx = left_child_idx; else break;
        } else {
            if (val < values[right_child_idx]) swap_idx = right_child_idx; else break;
        }
        values[current_idx] = values[swap_idx];
        indices[current_idx] = indices[swap_idx];
        current_idx = swap_idx;
    }
    values[current_idx] = val;
    indices[current_idx] = val_idx;
}

void tcsdn_oracle_knn(const double *X, int64_t n, int32_t d, const double *fitX, const int32_t *y,
                      int64_t n_train, int32_t k, int32_t C, int32_t *labels, int64_t *nbr_idx,
                      int32_t *counts) {
#pragma omp parallel
    {
        double *hv = (double *)malloc(sizeof(double) * (size_t)k);
        int64_t *hi = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
        int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * (size_t)C);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; i++) {
            const double *x = X + i * d;
            for (int32_t s = 0; s < k; s++) { hv[s] = DBL_MAX; hi[s] = 0; }
            for (int64_t t = 0; t < n_train; t++) {
                const double *yv = fitX + t * d;
                double dist = 0.0;
                for (int32_t j = 0; j < d; j++) {
                    double tmp = x[j] - yv[j];
                    dist += tmp * tmp;
                }
                heap_push(hv, hi, k, dist, t);
            }
            memset(cnt, 0, sizeof(int32_t) * (size_t)C);
            for (int32_t s = 0; s < k; s++) cnt[y[hi[s]]] += 1;
            int32_t arg = 0;
            for (int32_t c = 1; c < C; c++) if (cnt[c] > cnt[arg]) arg = c;
            if (labels) labels[i] = arg;
            if (nbr_idx) for (int32_t s = 0; s < k; s++) nbr_idx[i * k + s] = hi[s];
            if (counts) for (int32_t c = 0; c < C; c++) counts[i * C + c] = cnt[c];
        }
        free(hv); free(hi); free(cnt);
    }
}

/* ------------------------------------------------------------------------------------
 * a5 SVC.predict / a5' SVC.decision_function (kernel='rbf', dense)
 *   sk:svm/_base.py:830-861 predict -> :458-510 _dense_predict -> libsvm.predict
 *   sk:svm/src/libsvm/libsvm_helper.c:315-332 copy_predict: serial loop over rows
 *   sk:svm/src/libsvm/libsvm_helper.c:171  rho[i] = -intercept[i]
 *   sk:svm/src/libsvm/svm.cpp:461-478 k_function RBF: m = x - sv ; sum = dot(m, m) ; exp(-gamma*sum)
 *     (BLAS ddot order unspecified; j ascending here)
 *   sk:svm/src/libsvm/svm.cpp:2846-2904 predict_values: per pair (i<j), p++:
 *       sum = sum_{s in class i} coef[j-1][s] K_s + sum_{s in class j} coef[i][s] K_s ; sum -= rho[p]
 *       dec[p] = sum ; (dec[p] > 0) ? ++vote[i] : ++vote[j] ; winner = first max vote (strict >)
 * dec (optional) [n,P] with libsvm's sign (what SVC._decision_function returns for C > 2);
 * labels: class index.  For C == 2 sklearn negates dec (sk:svm/_base.py:535-538); callers do that.
 * ---------------------------------------------------------------------------------- */
void tcsdn_oracle_svc(const double *X, int64_t n, int32_t d, const double *sv, const double *dual_coef,
                      const double *intercept, const int32_t *n_support, int32_t n_sv, int32_t C,
                      double gamma, double *dec, int32_t *labels) {
    int32_t P = C * (C - 1) / 2;
    int32_t *start = (int32_t *)malloc(sizeof(int32_t) * (size_t)C);
    start[0] = 0;
    for (int32_t i = 1; i < C; i++) start[i] = start[i - 1] + n_support[i - 1];
#pragma omp parallel
    {
        double *kv = (double *)malloc(sizeof(double) * (size_t)n_sv);
        int32_t *vote = (int32_t *)malloc(sizeof(int32_t) * (size_t)C);
#pragma omp for schedule(dynamic, 16)
        for (int64_t r = 0; r < n; r++) {
            const double *x = X + r * d;
            for (int32_t s = 0; s < n_sv; s++) {
                const double *v = sv + (int64_t)s * d;
                double sum = 0.0;
                for (int32_t j = 0; j < d; j++) {
                    double m = x[j] - v[j];
                    sum += m * m;
                }
                kv[s] = exp(-gamma * sum);
            }
            for (int32_t i = 0; i < C; i++) vote[i] = 0;
            int32_t p = 0;
            for (int32_t i = 0; i < C; i++)
                for (int32_t j = i + 1; j < C; j++) {
                    double sum = 0.0;
                    int32_t si = start[i], sj = start[j], ci = n_support[i], cj = n_support[j];
                    const double *coef1 = dual_coef + (int64_t)(j - 1) * n_sv;
                    const double *coef2 = dual_coef + (int64_t)i * n_sv;
                    for (int32_t k = 0; k < ci; k++) sum += coef1[si + k] * kv[si + k];
                    for (int32_t k = 0; k < cj; k++) sum += coef2[sj + k] * kv[sj + k];
                    sum -= -intercept[p];
                    if (dec) dec[r * P + p] = sum;
                    if (sum > 0) ++vote[i]; else ++vote[j];
                    p++;
                }
            int32_t arg = 0;
            for (int32_t i = 1; i < C; i++) if (vote[i] > vote[arg]) arg = i;
            if (labels) labels[r] = arg;
        }
        free(kv); free(vote);
    }
    free(start);
}

/* sk:utils/multiclass.py:557-599 _ovr_decision_function(dec < 0, -dec, C), as called from
 * sk:svm/_base.py:798-828 SVC.decision_function when decision_function_shape == 'ovr' and C > 2. */
void tcsdn_oracle_ovr_from_ovo(const double *dec, int64_t n, int32_t C, double *out) {
    int32_t P = C * (C - 1) / 2;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; r++) {
        double votes[64], conf[64];
        for (int32_t c = 0; c < C; c++) { votes[c] = 0.0; conf[c] = 0.0; }
        int32_t k = 0;
        for (int32_t i = 0; i < C; i++)
            for (int32_t j = i + 1; j < C; j++) {
                double cf = -dec[r * P + k];
                conf[i] -= cf;
                conf[j] += cf;
                if (dec[r * P + k] < 0) votes[j] += 1.0; else votes[i] += 1.0;
                k++;
            }
        for (int32_t c = 0; c < C; c++) out[r * C + c] = votes[c] + conf[c] / (3.0 * (fabs(conf[c]) + 1.0));
    }
}

/* ------------------------------------------------------------------------------------
 * a6 RandomForestClassifier.predict
 *   sk:ensemble/_forest.py:606-624 _validate_X_predict: X cast to float32
 *   sk:tree/_tree.pyx:954-996 _apply_dense: while node.left_child != LEAF:
 *        (float32 x promoted) x <= (float64) threshold ? left : right
 *   sk:tree/_classes.py:1026-1061 predict_proba: value[leaf, :C] (fractions as stored)
 *   sk:ensemble/_forest.py:704-716,952-962: out(f64) += proba tree by tree in estimator order
 *        (n_jobs=None -> sequential), then out /= n_estimators
 *   sk:ensemble/_forest.py:882-919 predict: argmax (first max)
 * X here is float64 and is rounded to float32 first, exactly as np.asarray(X, float32) does.
 * ---------------------------------------------------------------------------------- */
void tcsdn_oracle_forest(const double *X, int64_t n, int32_t d, const int64_t *tree_offsets,
                         const int32_t *left, const int32_t *right, const int32_t *feature,
                         const double *threshold, const double *value, int32_t n_trees, int32_t C,
                         double *proba, int32_t *labels, int64_t *visits) {
    int64_t total_visits = 0;
#pragma omp parallel reduction(+ : total_visits)
    {
        float *xf = (float *)malloc(sizeof(float) * (size_t)d);
        double *acc = (double *)malloc(sizeof(double) * (size_t)C);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; i++) {
            for (int32_t j = 0; j < d; j++) xf[j] = (float)X[i * d + j];
            for (int32_t c = 0; c < C; c++) acc[c] = 0.0;
            for (int32_t t = 0; t < n_trees; t++) {
                int64_t base = tree_offsets[t];
                int64_t node = 0;
                while (left[base + node] != -1) {
                    total_visits++;
                    if ((double)xf[feature[base + node]] <= threshold[base + node]) node = left[base + node];
                    else node = right[base + node];
                }
                const double *v = value + (base + node) * C;
                for (int32_t c = 0; c < C; c++) acc[c] += v[c];
            }
            int32_t arg = 0;
            for (int32_t c = 0; c < C; c++) {
                acc[c] /= (double)n_trees;
                if (proba) proba[i * C + c] = acc[c];
            }
            for (int32_t c = 1; c < C; c++) if (acc[c] > acc[arg]) arg = c;
            if (labels) labels[i] = arg;
        }
        free(xf); free(acc);
    }
    if (visits) *visits = total_visits;
}

/* ------------------------------------------------------------------------------------
 * N1 feature derivation: restates Flow.updateforward / updatereverse
 * (reference traffic_classifier.py:63-96) for one direction of one flow.
 *   state: [packets, bytes, delta_packets, delta_bytes, inst_pps, avg_pps, inst_bps, avg_bps,
 *           last_time]  (all double; counters are exact integers below 2^53)
 * ---------------------------------------------------------------------------------- */
void tcsdn_oracle_flow_update(double *st, double time_start, double packets, double bytes, double curr_time) {
    st[2] = packets - st[0];
    st[0] = packets;
    if (curr_time != time_start) st[5] = packets / (curr_time - time_start);
    if (curr_time != st[8]) st[4] = st[2] / (curr_time - st[8]);
    st[3] = bytes - st[1];
    st[1] = bytes;
    if (curr_time != time_start) st[7] = bytes / (curr_time - time_start);
    if (curr_time != st[8]) st[6] = st[3] / (curr_time - st[8]);
    st[8] = curr_time;
}
