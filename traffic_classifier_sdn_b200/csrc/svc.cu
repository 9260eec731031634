// svc.cu -- SVC(kernel='rbf') decision values and vote, fp64 CUDA-core kernel  (SURVEY 8a rows a5, a5').
//
//   sk:svm/src/libsvm/svm.cpp:461-478   K_s = exp(-gamma * sum_j (x_j - sv_sj)^2)
//   sk:svm/src/libsvm/svm.cpp:2866-2891 pair (i<j): dec_p = sum_{s in i} coef[j-1][s] K_s
//                                                        + sum_{s in j} coef[i][s] K_s - rho_p
//                                       dec_p > 0 ? ++vote[i] : ++vote[j]
//   sk:svm/src/libsvm/svm.cpp:2893-2896 winner = first class with the maximum vote
//   sk:svm/src/libsvm/libsvm_helper.c:171  rho = -intercept
//   sk:utils/multiclass.py:557-599      one-vs-rest transform of the OvO values (decision_function)
//
// fp64 reference kernel of the path: small batches, and the cross-check of the tensor-core engine.
// One thread owns one row; support vectors and their dual coefficients stream through a shared-memory
// tile (every lane reads the same vector: broadcast).  Support vectors are grouped by class, so the
// C-1 running sums of the class being scanned stay in registers and are flushed when the class ends.
#include "common.h"

namespace tcsdn {

constexpr int kSvcThreads = 128;
constexpr int kSvcTile = 64;
constexpr int kSvcMaxC = 16;

template <typename T>
__global__ void __launch_bounds__(kSvcThreads) svc_exact_kernel(const T *__restrict__ X, int64_t n, int d,
                                                                const double *__restrict__ sv,
                                                                const double *__restrict__ coef,
                                                                const double *__restrict__ rho,
                                                                const int32_t *__restrict__ start, int n_sv, int C,
                                                                double gamma, int32_t *__restrict__ labels,
                                                                double *__restrict__ dec_out, int32_t *flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *xs = reinterpret_cast<double *>(smem_raw);       // [d][kSvcThreads]
    double *ss = xs + (size_t)d * kSvcThreads;               // [kSvcTile][d]
    double *cs = ss + (size_t)kSvcTile * d;                  // [C-1][kSvcTile]
    const int tid = threadIdx.x;
    const int Cm1 = C - 1;
    float nf = 0.f;
    for (int64_t r0 = (int64_t)blockIdx.x * kSvcThreads; r0 < n; r0 += (int64_t)gridDim.x * kSvcThreads) {
        const int64_t row = r0 + tid;
        const bool live = row < n;
        __syncthreads();
        if (live) {
            for (int j = 0; j < d; ++j) {
                T v = X[row * d + j];
                nf += static_cast<float>(v * static_cast<T>(0));
                xs[j * kSvcThreads + tid] = static_cast<double>(v);
            }
        }
        double S[kSvcMaxC * (kSvcMaxC - 1)];   // S[i][m] = sum_{s in class i} coef[m][s] K_s
        double acc[kSvcMaxC - 1];
#pragma unroll
        for (int mm = 0; mm < kSvcMaxC - 1; ++mm) acc[mm] = 0.0;
        int ci = 0;
        for (int s0 = 0; s0 < n_sv; s0 += kSvcTile) {
            const int tn = (n_sv - s0) < kSvcTile ? (n_sv - s0) : kSvcTile;
            __syncthreads();
            for (int e = tid; e < tn * d; e += kSvcThreads) ss[e] = sv[(size_t)s0 * d + e];
            for (int e = tid; e < Cm1 * tn; e += kSvcThreads) {
                int mm = e / tn, tt = e - mm * tn;
                cs[mm * kSvcTile + tt] = coef[(size_t)mm * n_sv + s0 + tt];
            }
            __syncthreads();
            if (live) {
                for (int tt = 0; tt < tn; ++tt) {
                    const int s = s0 + tt;
                    while (s >= start[ci + 1]) {  // class boundary (uniform across the CTA)
#pragma unroll
                        for (int mm = 0; mm < kSvcMaxC - 1; ++mm)
                            if (mm < Cm1) { S[ci * Cm1 + mm] = acc[mm]; acc[mm] = 0.0; }
                        ++ci;
                    }
                    double sum = 0.0;
                    for (int j = 0; j < d; ++j) {
                        double df = xs[j * kSvcThreads + tid] - ss[tt * d + j];
                        sum = fma(df, df, sum);
                    }
                    const double kv = exp(-gamma * sum);
#pragma unroll
                    for (int mm = 0; mm < kSvcMaxC - 1; ++mm)
                        if (mm < Cm1) acc[mm] = fma(cs[mm * kSvcTile + tt], kv, acc[mm]);
                }
            }
        }
        if (live) {
            while (ci < C) {
#pragma unroll
                for (int mm = 0; mm < kSvcMaxC - 1; ++mm)
                    if (mm < Cm1) { S[ci * Cm1 + mm] = acc[mm]; acc[mm] = 0.0; }
                ++ci;
            }
            int vote[kSvcMaxC];
            for (int c = 0; c < C; ++c) vote[c] = 0;
            int p = 0;
            const int P = C * (C - 1) / 2;
            for (int i = 0; i < C; ++i)
                for (int j = i + 1; j < C; ++j) {
                    double dv = (S[i * Cm1 + (j - 1)] + S[j * Cm1 + i]) - rho[p];
                    if (dec_out) dec_out[row * P + p] = dv;
                    if (dv > 0) ++vote[i]; else ++vote[j];
                    ++p;
                }
            int arg = 0;
            for (int c = 1; c < C; ++c)
                if (vote[c] > vote[arg]) arg = c;
            labels[row] = arg;
        }
    }
    if (flag && nf != nf) atomicOr(flag, 1);
}

__global__ void ovr_from_ovo_kernel(const double *__restrict__ dec, int64_t n, int C, double *__restrict__ out) {
    const int P = C * (C - 1) / 2;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        double votes[kMaxClasses], conf[kMaxClasses];
        for (int c = 0; c < C; ++c) { votes[c] = 0.0; conf[c] = 0.0; }
        int k = 0;
        for (int i = 0; i < C; ++i)
            for (int j = i + 1; j < C; ++j) {
                const double dv = dec[r * P + k];
                const double cf = -dv;
                conf[i] -= cf;
                conf[j] += cf;
                if (dv < 0) votes[j] += 1.0; else votes[i] += 1.0;
                ++k;
            }
        for (int c = 0; c < C; ++c) out[r * C + c] = votes[c] + conf[c] / (3.0 * (fabs(conf[c]) + 1.0));
    }
}

int launch_ovr_from_ovo(const double *dec, int64_t n, int C, double *out, cudaStream_t st) {
    if (n == 0) return TCSDN_OK;
    if (C < 2 || C > kMaxClasses) { set_error("ovr_from_ovo: n_classes out of range"); return TCSDN_EINVAL; }
    int64_t blocks = (n + 127) / 128;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ovr_from_ovo_kernel<<<(unsigned)blocks, 128, 0, st>>>(dec, n, C, out);
    TCSDN_CUDA(cudaGetLastError());
    return TCSDN_OK;
}

int launch_svc_exact(tcsdn_model *m, const void *x, int64_t n, int dtype, int32_t *labels, double *scores,
                     cudaStream_t st) {
    if (n == 0) return TCSDN_OK;
    if (m->n_classes > kSvcMaxC) { set_error("svc: more than %d classes", kSvcMaxC); return TCSDN_EINVAL; }
    const size_t smem = ((size_t)m->d * kSvcThreads + (size_t)kSvcTile * m->d + (size_t)(m->n_classes - 1) * kSvcTile) *
                        sizeof(double);
    int64_t blocks = (n + kSvcThreads - 1) / kSvcThreads;
    int64_t cap = (int64_t)m->sm_count * 8;
    if (blocks > cap) blocks = cap;
    int32_t *flag = m->opt_check_finite ? m->d_flag : nullptr;
    m->stats[0] += 1;
    m->stats[2] += n;
    if (dtype == TCSDN_F32) {
        auto kern = svc_exact_kernel<float>;
        TCSDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, kSvcThreads, smem, st>>>(static_cast<const float *>(x), n, m->d, m->d_sv, m->d_coef,
                                                          m->d_rho, m->d_start, m->n_sv, m->n_classes, m->gamma,
                                                          labels, scores, flag);
    } else {
        auto kern = svc_exact_kernel<double>;
        TCSDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, kSvcThreads, smem, st>>>(static_cast<const double *>(x), n, m->d, m->d_sv, m->d_coef,
                                                          m->d_rho, m->d_start, m->n_sv, m->n_classes, m->gamma,
                                                          labels, scores, flag);
    }
    TCSDN_CUDA(cudaGetLastError());
    return TCSDN_OK;
}

}  // namespace tcsdn
