// umma_probe.cu -- standalone bring-up test for the tcgen05 building blocks of csrc/dist_engine.cu:
// TMEM alloc, no-swizzle K-major smem descriptors, kind::f16 (bf16 x bf16 -> fp32) MMA with M=128, N=64,
// K=16 x KSTEPS, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b.x32, bulk async copy of a pre-tiled B image.
// Every wait is bounded (a stuck barrier sets an error code instead of hanging the GPU).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 64, KSTEPS = 5, K = 16 * KSTEPS;  // K = 80
constexpr int kChunkBytes = 128;                             // one 8x8 bf16 core matrix
constexpr int kSBO = (K / 8) * kChunkBytes;                  // stride between 8-row groups

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity, int *err, int code) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22); ++spin) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return true;
    }
    atomicExch(err, code);
    return false;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
    // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
    return d;
}

__global__ void __launch_bounds__(160, 1) probe_kernel(const float *A, const uint8_t *Bimg, float *D, int *err) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem;                       // M x K bf16, core-matrix layout
    unsigned char *sB = smem + M * K * 2;           // N x K bf16, same layout (copied as is)
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + M * K * 2 + N * K * 2);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));  // B landed
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));  // MMA done
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {  // one warp allocates 64 TMEM columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // A: thread r (warps 0-3) packs row r (fp32 -> bf16) into the canonical K-major layout
    if (tid < M) {
        const int r = tid;
#pragma unroll
        for (int c = 0; c < K / 8; ++c) {
            __align__(16) __nv_bfloat16 v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(A[r * K + c * 8 + e]);
            *reinterpret_cast<uint4 *>(sA + (r / 8) * kSBO + c * kChunkBytes + (r % 8) * 16) = *reinterpret_cast<uint4 *>(v);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4 && lane == 0) {
        // B: one bulk async copy of the pre-tiled image
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(N * K * 2) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sB)),
                     "l"(Bimg), "r"(N * K * 2), "r"(smem_u32(&bars[0]))
                     : "memory");
        if (mbar_wait_bounded(&bars[0], 0, err, 1)) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // instruction descriptor: D=F32 (1<<4), A=BF16 (1<<7), B=BF16 (1<<10), K-major both, N>>3 @17, M>>4 @24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
                const uint64_t da = make_desc(smem_u32(sA) + k * 2 * kChunkBytes, kChunkBytes, kSBO);
                const uint64_t db = make_desc(smem_u32(sB) + k * 2 * kChunkBytes, kChunkBytes, kSBO);
                const uint32_t accum = k > 0;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_base),
                    "l"(da), "l"(db), "r"(idesc), "r"(accum)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
        }
    }
    if (warp < 4) {  // epilogue: warp w reads TMEM lanes 32w..32w+31, thread = row
        if (mbar_wait_bounded(&bars[1], 0, err, 2)) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[64];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[h * 32 + 0]), "=r"(v[h * 32 + 1]), "=r"(v[h * 32 + 2]), "=r"(v[h * 32 + 3]), "=r"(v[h * 32 + 4]),
                      "=r"(v[h * 32 + 5]), "=r"(v[h * 32 + 6]), "=r"(v[h * 32 + 7]), "=r"(v[h * 32 + 8]), "=r"(v[h * 32 + 9]),
                      "=r"(v[h * 32 + 10]), "=r"(v[h * 32 + 11]), "=r"(v[h * 32 + 12]), "=r"(v[h * 32 + 13]), "=r"(v[h * 32 + 14]),
                      "=r"(v[h * 32 + 15]), "=r"(v[h * 32 + 16]), "=r"(v[h * 32 + 17]), "=r"(v[h * 32 + 18]), "=r"(v[h * 32 + 19]),
                      "=r"(v[h * 32 + 20]), "=r"(v[h * 32 + 21]), "=r"(v[h * 32 + 22]), "=r"(v[h * 32 + 23]), "=r"(v[h * 32 + 24]),
                      "=r"(v[h * 32 + 25]), "=r"(v[h * 32 + 26]), "=r"(v[h * 32 + 27]), "=r"(v[h * 32 + 28]), "=r"(v[h * 32 + 29]),
                      "=r"(v[h * 32 + 30]), "=r"(v[h * 32 + 31])
                    : "r"(taddr + h * 32));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = warp * 32 + lane;
#pragma unroll
            for (int c = 0; c < N; ++c) D[row * N + c] = __uint_as_float(v[c]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64));
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N, -1.f), ref(M * N);
    srand(1);
    for (auto &v : A) v = bf16_round((float)(rand() % 2001 - 1000) / 64.f);
    for (auto &v : B) v = bf16_round((float)(rand() % 2001 - 1000) / 32.f);
    std::vector<uint8_t> Bimg(N * K * 2);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            __nv_bfloat16 h = __float2bfloat16(B[n * K + k]);
            size_t off = (size_t)(n / 8) * kSBO + (k / 8) * kChunkBytes + (n % 8) * 16 + (k % 8) * 2;
            memcpy(&Bimg[off], &h, 2);
        }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
            ref[m * N + n] = (float)s;
        }
    float *dA, *dD; uint8_t *dB; int *derr;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dB, Bimg.size()); cudaMalloc(&derr, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bimg.data(), Bimg.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(derr, 0, 4);
    size_t smem = M * K * 2 + N * K * 2 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 160, smem>>>(dA, dB, dD, derr);
    cudaError_t e = cudaDeviceSynchronize();
    int herr = 0;
    cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    int bad = 0;
    for (int i = 0; i < M * N; ++i) {
        double er = fabs((double)D[i] - ref[i]);
        maxerr = fmax(maxerr, er); maxref = fmax(maxref, fabs((double)ref[i]));
        if (er > 1e-3 * (1 + fabs(ref[i])) && bad < 8) { printf("mismatch [%d,%d] got %g want %g\n", i / N, i % N, D[i], ref[i]); ++bad; }
    }
    printf("umma_probe: cuda=%s err_code=%d max_abs_err=%g (max |ref| %g) -> %s\n", cudaGetErrorString(e), herr, maxerr, maxref,
           (e == cudaSuccess && herr == 0 && maxerr <= 1e-3 * (1 + maxref)) ? "PASS" : "FAIL");
    return 0;
}
