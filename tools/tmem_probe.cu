// tmem_probe.cu -- two per-SM throughput numbers the distance engine's epilogues are bounded by:
//   (1) tcgen05.ld (TMEM -> registers) bytes/clk/SM, for x16 / x32 shapes and 4 / 8 / 16 warps;
//   (2) broadcast LDS.128 / LDS.64 / LDS.32 (all lanes read the same address) instructions/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tmem_probe tools/tmem_probe.cu && tools/tmem_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>   // 16 or 32 columns per load
__global__ void __launch_bounds__(512, 1) ldtm_kernel(int iters, long long *cycles, float *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const uint32_t addr = base + (uint32_t)((i * SHAPE + (warp >> 2) * 64) & 511 & ~(SHAPE - 1));
        if (SHAPE == 32) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(r[0] ^ r[31]);
        } else {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(r[0] ^ r[15]);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 1.2345f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <int W>   // 4, 2, 1 words per lane, same address in every lane (broadcast)
__global__ void __launch_bounds__(512, 1) lds_kernel(int iters, long long *cycles, float *sink) {
    __shared__ __align__(16) float buf[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = (float)i;
    __syncthreads();
    float acc = 0.f;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int off = ((i * 8 + u) * 4) & 4092;
            if (W == 4) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(buf + off))); acc += v.x + v.w; }
            if (W == 2) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem_u32(buf + off))); acc += v.x + v.y; }
            if (W == 1) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(buf + off))); acc += v; }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 1.2345f) sink[0] = acc;
}

// MUFU.EX2 / FFMA2 / FFMA issue rates: 8 independent chains per thread, `warps` warps on one SM
template <int OP>
__global__ void __launch_bounds__(1024, 1) alu_kernel(int iters, long long *cycles, float *sink) {
    float a[8];
    float2 b[8];
    for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); b[i] = make_float2(a[i], a[i] * 0.5f); }
    const float2 c2 = make_float2(0.999f, 1.001f);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[u]));
            if (OP == 1) { unsigned long long d = *reinterpret_cast<unsigned long long *>(&b[u]); const unsigned long long m = *reinterpret_cast<const unsigned long long *>(&c2);
                           asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d) : "l"(m)); b[u] = *reinterpret_cast<float2 *>(&d); }
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[u]) : "f"(0.999f));
            if (OP == 3) { double dd; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(dd) : "f"(a[u])); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(a[u]) : "d"(dd)); }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float acc = 0.f;
    for (int i = 0; i < 8; ++i) acc += a[i] + b[i].x + b[i].y;
    if (acc == 1.2345f) sink[0] = acc;
}

// the SVC epilogue's instruction mix: per group 1 MUFU.EX2 + NF FFMA2 (+ optionally one broadcast LDS.128), independent chains
template <int NF, bool LDS>
__global__ void __launch_bounds__(1024, 1) mix_kernel(int iters, long long *cycles, float *sink) {
    __shared__ __align__(16) float buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = 1.0f + 1e-6f * i;
    float a[8];
    float2 b[8][3];
    for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); for (int k = 0; k < 3; ++k) b[i][k] = make_float2(a[i], a[i] * 0.5f + k); }
    float4 cf = make_float4(0.999f, 1.001f, 0.998f, 1.002f);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[u]));
            if (LDS) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cf.x), "=f"(cf.y), "=f"(cf.z), "=f"(cf.w) : "r"(smem_u32(buf + ((i * 8 + u) * 4 & 1020))));
#pragma unroll
            for (int k = 0; k < NF; ++k) {
                unsigned long long d = *reinterpret_cast<unsigned long long *>(&b[u][k % 3]);
                const float2 c2 = (k & 1) ? make_float2(cf.x, cf.y) : make_float2(cf.z, cf.w);
                const unsigned long long m = *reinterpret_cast<const unsigned long long *>(&c2);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d) : "l"(m));
                b[u][k % 3] = *reinterpret_cast<float2 *>(&d);
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float acc = 0.f;
    for (int i = 0; i < 8; ++i) acc += a[i] + b[i][0].x + b[i][1].y + b[i][2].x;
    if (acc == 1.2345f) sink[0] = acc;
}

// closer to the SVC epilogue: VAR 0: 4 FFMA2 make 8 exponents, 8 MUFU.EX2 (out of place), 4 FFMA2 consume the pairs
//                            VAR 1: the same with scalar FFMA for the exponents;  VAR 2: MUFU in place on the FFMA2 results
template <int VAR>
__global__ void __launch_bounds__(1024, 1) epi_kernel(int iters, long long *cycles, float *sink) {
    float2 v[4], acc[4];
    for (int i = 0; i < 4; ++i) { v[i] = make_float2(-0.001f * (threadIdx.x + i), -0.002f * i); acc[i] = make_float2(0.f, 0.f); }
    const float2 g = make_float2(0.999f, 0.999f), bias = make_float2(-0.5f, -0.5f), cf = make_float2(0.25f, -0.25f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float2 e[4], k[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (VAR == 1) { e[i].x = fmaf(v[i].x, g.x, bias.x); e[i].y = fmaf(v[i].y, g.y, bias.y); }
            else { unsigned long long d = *reinterpret_cast<const unsigned long long *>(&bias);
                   asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&v[i])), "l"(*reinterpret_cast<const unsigned long long *>(&g)));
                   e[i] = *reinterpret_cast<float2 *>(&d); }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (VAR == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[i].x)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[i].y)); k[i] = e[i]; }
            else { asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(k[i].x) : "f"(e[i].x)); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(k[i].y) : "f"(e[i].y)); }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned long long d = *reinterpret_cast<unsigned long long *>(&acc[i]);
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&k[i])), "l"(*reinterpret_cast<const unsigned long long *>(&cf)));
            acc[i] = *reinterpret_cast<float2 *>(&d);
            v[i].x = acc[i].y * 1e-3f; v[i].y = acc[i].x * 1e-3f;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc[0].x + acc[1].y + acc[2].x + acc[3].y == 1.2345f) sink[0] = acc[0].x;
}

int main() {
    long long *cyc; float *sink;
    cudaMalloc(&cyc, 1024 * sizeof(long long)); cudaMalloc(&sink, 4);
    const int iters = 20000;
    long long h = 0;
    for (int warps : {4, 8, 16}) {
        ldtm_kernel<16><<<1, warps * 32>>>(iters, cyc, sink); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDTM x16 %2d warps: %.1f clk per load per warp, %.1f B/clk/SM  (%s)\n", warps, (double)h / iters, warps * 2048.0 * iters / h, cudaGetErrorString(cudaGetLastError()));
        ldtm_kernel<32><<<1, warps * 32>>>(iters, cyc, sink); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDTM x32 %2d warps: %.1f clk per load per warp, %.1f B/clk/SM  (%s)\n", warps, (double)h / iters, warps * 4096.0 * iters / h, cudaGetErrorString(cudaGetLastError()));
    }
    for (int warps : {4, 8, 16}) {
        lds_kernel<4><<<1, warps * 32>>>(iters, cyc, sink); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.128 broadcast %2d warps: %.2f clk per instruction per SM\n", warps, (double)h / (iters * 8.0 * warps));
        lds_kernel<2><<<1, warps * 32>>>(iters, cyc, sink); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.64  broadcast %2d warps: %.2f clk per instruction per SM\n", warps, (double)h / (iters * 8.0 * warps));
        lds_kernel<1><<<1, warps * 32>>>(iters, cyc, sink); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.32  broadcast %2d warps: %.2f clk per instruction per SM\n", warps, (double)h / (iters * 8.0 * warps));
    }
    const char *names[4] = {"MUFU.EX2", "FFMA2", "FFMA", "F2F f32->f64->f32 (2 instr)"};
    for (int op = 0; op < 4; ++op)
        for (int warps : {4, 8, 16, 32}) {
            if (op == 0) alu_kernel<0><<<1, warps * 32>>>(iters, cyc, sink);
            if (op == 1) alu_kernel<1><<<1, warps * 32>>>(iters, cyc, sink);
            if (op == 2) alu_kernel<2><<<1, warps * 32>>>(iters, cyc, sink);
            if (op == 3) alu_kernel<3><<<1, warps * 32>>>(iters, cyc, sink);
            cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-28s %2d warps: %.2f clk per warp-instruction per SM (%.1f lanes/clk/SM)\n", names[op], warps,
                   (double)h / (iters * 8.0 * warps), 32.0 * iters * 8.0 * warps / h);
        }
    for (int warps : {4, 8, 16}) {
        epi_kernel<0><<<1, warps * 32>>>(iters, cyc, sink); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("epilogue-like (FFMA2 args, 8 MUFU out of place) %2d warps: %.2f clk per MUFU per SM\n", warps, (double)h / (iters * 8.0 * warps));
        epi_kernel<1><<<1, warps * 32>>>(iters, cyc, sink); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("epilogue-like (scalar FFMA args)               %2d warps: %.2f clk per MUFU per SM\n", warps, (double)h / (iters * 8.0 * warps));
        epi_kernel<2><<<1, warps * 32>>>(iters, cyc, sink); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("epilogue-like (MUFU in place)                  %2d warps: %.2f clk per MUFU per SM\n", warps, (double)h / (iters * 8.0 * warps));
    }
    for (int warps : {4, 8, 16}) {
        mix_kernel<3, false><<<1, warps * 32>>>(iters, cyc, sink); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mix 1 MUFU + 3 FFMA2         %2d warps: %.2f clk per group per SM   (MUFU alone 2.0, 3 FFMA2 alone 1.56)\n", warps, (double)h / (iters * 8.0 * warps));
        mix_kernel<3, true><<<1, warps * 32>>>(iters, cyc, sink); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mix 1 MUFU + 3 FFMA2 + LDS.128 %2d warps: %.2f clk per group per SM\n", warps, (double)h / (iters * 8.0 * warps));
        mix_kernel<6, false><<<1, warps * 32>>>(iters, cyc, sink); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mix 1 MUFU + 6 FFMA2         %2d warps: %.2f clk per group per SM   (6 FFMA2 alone 3.12)\n", warps, (double)h / (iters * 8.0 * warps));
    }
    return 0;
}
