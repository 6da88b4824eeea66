// Probe of the tcgen05 pieces used by the batched analysis kernel: TMEM allocation, K-major no-swizzle shared-memory
// descriptors for TF32 operands, tcgen05.mma kind::tf32 (M=128, N=64, K=8), commit -> mbarrier, tcgen05.ld, and the
// rounding behaviour of the accumulator (long accumulation chains vs float64).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // layout type 0 (no swizzle), base offset 0
}

__device__ __forceinline__ bool mbar_wait(uint32_t mbar, uint32_t parity) {
    for (long long it = 0; it < (1ll << 24); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// A: [128 rows][8 k], B: [64 cols][8 k] (row-major inputs in global); out D [128][64]; reps accumulation steps with the
// same operands; mode 0: D = sum over reps of A.B accumulated in TMEM;  split: 0 = plain tf32, 1 = 3xTF32
__global__ void __launch_bounds__(128, 1) probe(const float *A, const float *B, float *D, int reps, int split, int *status) {
    __shared__ __align__(128) float sAhi[2 * 128 * 4], sAlo[2 * 128 * 4], sBhi[2 * 64 * 4], sBlo[2 * 64 * 4];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    // operands -> K-major core-matrix layout: element (r, k) at [(k/4)][r][k%4]
    for (int i = tid; i < 128 * 8; i += 128) {
        const int r = i / 8, k = i % 8;
        const float x = A[i];
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
        float hi = __uint_as_float(hb);
        if (!split) hi = x;
        sAhi[((k / 4) * 128 + r) * 4 + (k % 4)] = hi;
        sAlo[((k / 4) * 128 + r) * 4 + (k % 4)] = x - hi;
    }
    for (int i = tid; i < 64 * 8; i += 128) {
        const int r = i / 8, k = i % 8;
        const float x = B[i];
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
        float hi = __uint_as_float(hb);
        if (!split) hi = x;
        sBhi[((k / 4) * 64 + r) * 4 + (k % 4)] = hi;
        sBlo[((k / 4) * 64 + r) * 4 + (k % 4)] = x - hi;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base;
    // instruction descriptor: D f32, A/B tf32, K-major both, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    if (tid == 0) {
        const uint64_t dAhi = make_desc(smem_u32(sAhi), 128 * 16, 128), dAlo = make_desc(smem_u32(sAlo), 128 * 16, 128);
        const uint64_t dBhi = make_desc(smem_u32(sBhi), 64 * 16, 128), dBlo = make_desc(smem_u32(sBlo), 64 * 16, 128);
        for (int r = 0; r < reps; ++r) {
            const uint32_t accum = r > 0;
            if (split) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tbase), "l"(dAlo), "l"(dBhi), "r"(idesc), "r"(accum) : "memory");
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tbase), "l"(dAhi), "l"(dBlo), "r"(idesc), "r"(1u) : "memory");
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tbase), "l"(dAhi), "l"(dBhi), "r"(idesc), "r"(1u) : "memory");
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tbase), "l"(dAhi), "l"(dBhi), "r"(idesc), "r"(accum) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    const bool ok = mbar_wait(smem_u32(&mbar), 0);
    if (!ok) {
        if (tid == 0) *status = 1;
    } else {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[64];
        const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
            "%25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
            "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
              "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
              "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
              "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
              "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 64; ++j) D[tid * 64 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase));
}


// timing: `tiles` x 3 MMAs (M=128, N=ncol, K=8) issued back to back by one thread, clocks from first issue to mbarrier completion
__global__ void __launch_bounds__(128, 1) mma_time(int tiles, int ncol, int reps, long long *clk) {
    extern __shared__ __align__(128) float sm[];
    float *sA = sm, *sB = sm + 2 * 8 * 2 * 128 * 4;   // A: up to 8 tiles hi/lo ... reuse same data
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 2 * 8 * 2 * 128 * 4 + 2 * 8 * 2 * 256 * 4; i += 128) sm[i] = 1.0f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)ncol >> 3) << 17) | ((128u >> 4) << 24);
    long long best = 1ll << 60;
    for (int r = 0; r < reps; ++r) {
        long long t0 = 0;
        if (tid == 0) {
            t0 = clock64();
            for (int t = 0; t < tiles; ++t) {
                const uint64_t dA = make_desc(smem_u32(sA + t * 128 * 4), 8 * 128 * 16, 128);
                const uint64_t dA2 = make_desc(smem_u32(sA + 8 * 2 * 128 * 4 + t * 128 * 4), 8 * 128 * 16, 128);
                const uint64_t dB = make_desc(smem_u32(sB), ncol * 16, 128), dB2 = make_desc(smem_u32(sB + 8 * 2 * 256 * 4), ncol * 16, 128);
                const uint32_t d = tbase + (t * ncol) % 512;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(dA2), "l"(dB), "r"(idesc), "r"(0u) : "memory");
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(dA), "l"(dB2), "r"(idesc), "r"(1u) : "memory");
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(dA), "l"(dB), "r"(idesc), "r"(1u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        mbar_wait(smem_u32(&mbar), r & 1);
        if (tid == 0) { long long dt = clock64() - t0; if (dt < best) best = dt; }
        __syncthreads();
    }
    if (tid == 0) *clk = best;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

int main() {
    {
        long long *dclk; cudaMalloc(&dclk, 8);
        const int smem = (2 * 8 * 2 * 128 * 4 + 2 * 8 * 2 * 256 * 4) * 4;
        cudaFuncSetAttribute(mma_time, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int ncol : {64, 128, 256}) for (int tiles : {1, 2, 4, 8}) {
            mma_time<<<1, 128, smem>>>(tiles, ncol, 5, dclk);
            cudaError_t e = cudaDeviceSynchronize(); long long c = 0; cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost);
            printf("mma_time N=%d tiles=%d (%d MMAs): %lld clk  (%s)\n", ncol, tiles, 3 * tiles, c, cudaGetErrorString(e));
        }
    }
    std::vector<float> A(128 * 8), B(64 * 8), D(128 * 64);
    srand(1);
    for (auto &x : A) x = (float)rand() / RAND_MAX + 0.25f;   // positive: accumulation grows monotonically (worst case for RZ)
    for (auto &x : B) x = (float)rand() / RAND_MAX + 0.25f;
    float *dA, *dB, *dD;
    int *dS;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    for (int split = 0; split < 2; ++split)
        for (int reps : {1, 16, 64, 256, 1024, 4096}) {
            cudaMemset(dS, 0, 4);
            cudaMemset(dD, 0, D.size() * 4);
            probe<<<1, 128>>>(dA, dB, dD, reps, split, dS);
            cudaError_t e = cudaDeviceSynchronize();
            int st = 0;
            cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double num = 0, den = 0, bias = 0;
            for (int i = 0; i < 128; ++i)
                for (int j = 0; j < 64; ++j) {
                    double ref = 0;
                    for (int k = 0; k < 8; ++k) ref += (double)A[i * 8 + k] * (double)B[j * 8 + k];
                    ref *= reps;
                    const double d = (double)D[i * 64 + j] - ref;
                    num += d * d; den += ref * ref; bias += d / ref;
                }
            printf("split=%d reps=%5d: cuda=%s status=%d rel-L2 err %.3e mean signed rel err %.3e  D[0]=%g\n", split, reps,
                   cudaGetErrorString(e), st, sqrt(num / den), bias / (128 * 64), D[0]);
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
