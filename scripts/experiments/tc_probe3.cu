// Probe 3 (round 2, derived from tc_probe2.cu): tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (written by tcgen05.st) against the
// shared-memory-descriptor form, M = 128, K = 8, N = 16..256:
//   (1) correctness of the assumed A layout in tensor memory (lane = row, 32-bit column = k) with a 3xTF32 product;
//   (2) clocks of back-to-back MMAs, SS form vs TS form, per N;
//   (3) the same while other warps of the CTA stream shared memory (LDS.128 + STS.128), to see whether the tensor core's
//       operand fetch and the LSU share one shared-memory bandwidth.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe2 tc_probe2.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
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
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void split(float x, float &hi, float &lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    lo = x - hi;
}

constexpr int ACOL = 448;  // tensor-memory columns of the A operand: hi at ACOL..+7, lo at ACOL+8..+15

// (1) D[128][N] = A[128][8] . B[N][8]^T with 3xTF32, A in tensor memory
__global__ void __launch_bounds__(128, 1) ts_check(const float *A, const float *B, float *D, int N, int *status) {
    extern __shared__ __align__(128) float sm[];
    float *sBhi = sm, *sBlo = sm + 2 * 256 * 4;
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < N * 8; i += 128) {
        const int r = i / 8, k = i % 8;
        float hi, lo;
        split(B[i], hi, lo);
        sBhi[((k / 4) * N + r) * 4 + (k % 4)] = hi;
        sBlo[((k / 4) * N + r) * 4 + (k % 4)] = lo;
    }
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
    // A row `tid` -> lane tid of tensor memory (warp w writes the lane quarter w)
    float hi[8], lo[8];
    for (int k = 0; k < 8; ++k) split(A[tid * 8 + k], hi[k], lo[k]);
    const uint32_t arow = tbase + ((uint32_t)(warp * 32) << 16) + ACOL;
    tmem_st8(arow, hi);
    tmem_st8(arow + 8, lo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    if (tid == 0) {
        const uint64_t dBhi = make_desc(smem_u32(sBhi), N * 16, 128), dBlo = make_desc(smem_u32(sBlo), N * 16, 128);
        mma_ts(tbase, tbase + ACOL + 8, dBhi, idesc, 0u);   // a_lo . B_hi
        mma_ts(tbase, tbase + ACOL, dBlo, idesc, 1u);       // a_hi . B_lo
        mma_ts(tbase, tbase + ACOL, dBhi, idesc, 1u);       // a_hi . B_hi
        commit(smem_u32(&mbar));
    }
    const bool ok = mbar_wait(smem_u32(&mbar), 0);
    if (!ok) {
        if (tid == 0) *status = 1;
    } else {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
            for (int j = 0; j < 32 && c0 + j < N; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}


// Issue-rate probe: `nissue` warps (warps 0..nissue-1, one per scheduler) each issue `nmma` x 3 MMAs (TS or SS form) into their
// own accumulator columns; `dep` = 1: the three MMAs of a group accumulate into the same tile (as 3xTF32 does), 0: every MMA
// its own tile.  Hammer warps: htype 1 = FFMA loop (issue slots only), 2 = volatile LDS.128 + STS.128 stream (shared-memory
// bandwidth), placed on all schedulers (hsched = 0) or only on schedulers 1..3 (hsched = 1: never the scheduler of warp 0).
__global__ void __launch_bounds__(1024, 1) issue_time(int mode, int ncol, int nmma, int nissue, int dep, int htype, int hwarps,
                                                      int hsched, int hiters, long long *clk) {
    extern __shared__ __align__(128) float sm[];
    float *sA = sm;                       // [hi|lo][2][128][4] = 2048 floats
    float *sB = sm + 2048;                // [hi|lo][2][256][4] = 4096 floats
    float *sH = sm + 2048 + 4096;         // hammer region: 32 warps x 32 lanes x 8 floats = 8192 floats
    __shared__ __align__(8) unsigned long long mbar[4];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 2048 + 4096 + 8192; i += blockDim.x) sm[i] = 1.0f;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[i])));
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
    if (warp < 4) {
        float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + ACOL, one);
        tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + ACOL + 8, one);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)ncol >> 3) << 17) | ((128u >> 4) << 24);
    long long best = 1ll << 60, ibest = 1ll << 60;
    const int ntile = max(1, (448 / nissue) / ncol);   // accumulator tiles of this issuing warp
    bool is_hammer = false;
    if (warp >= 4) {
        const int h = warp - 4;
        if (hsched == 0) is_hammer = h < hwarps;
        else is_hammer = (warp & 3) != 0 && (h - (h + 3) / 4) < hwarps;
    }
    for (int r = 0; r < 5; ++r) {
        __syncthreads();
        const long long t0 = clock64();
        if (warp < nissue) {
            if (lane == 0) {
                const uint64_t dAhi = make_desc(smem_u32(sA), 128 * 16, 128), dAlo = make_desc(smem_u32(sA + 1024), 128 * 16, 128);
                const uint64_t dBhi = make_desc(smem_u32(sB), ncol * 16, 128), dBlo = make_desc(smem_u32(sB + 2048), ncol * 16, 128);
                const uint32_t d0 = tbase + (uint32_t)(warp * (448 / nissue));
                int tile = 0;
                for (int t = 0; t < nmma; ++t) {
                    uint32_t d[3];
                    for (int q = 0; q < 3; ++q) {
                        d[q] = d0 + (uint32_t)(tile * ncol);
                        if (!dep || q == 2) tile = tile + 1 == ntile ? 0 : tile + 1;
                    }
                    if (mode == 0) {
                        mma_ss(d[0], dAlo, dBhi, idesc, 1u);
                        mma_ss(d[1], dAhi, dBlo, idesc, 1u);
                        mma_ss(d[2], dAhi, dBhi, idesc, 1u);
                    } else {
                        mma_ts(d[0], tbase + ACOL + 8, dBhi, idesc, 1u);
                        mma_ts(d[1], tbase + ACOL, dBlo, idesc, 1u);
                        mma_ts(d[2], tbase + ACOL, dBhi, idesc, 1u);
                    }
                }
                commit(smem_u32(&mbar[warp]));
                const long long ti = clock64() - t0;
                if (warp == 0 && ti < ibest) ibest = ti;
            }
            __syncwarp();
            mbar_wait(smem_u32(&mbar[warp]), r & 1);
            if (lane == 0 && warp == 0) { const long long dt = clock64() - t0; if (dt < best) best = dt; }
        } else if (is_hammer) {
            if (htype == 1) {
                float a = lane * 1e-3f, b = a + 1.f, c = a + 2.f, e = a + 3.f;
                for (int i = 0; i < hiters * 4; ++i) {
                    a = fmaf(a, 0.999f, 1e-4f); b = fmaf(b, 0.999f, 1e-4f); c = fmaf(c, 0.999f, 1e-4f); e = fmaf(e, 0.999f, 1e-4f);
                }
                if (a + b + c + e == 12345.f) sH[0] = a;
            } else {
                volatile float4 *p = reinterpret_cast<volatile float4 *>(sH + (warp * 32 + lane) * 8);
                for (int i = 0; i < hiters; ++i) {
                    float4 v;
                    v.x = p[i & 1].x; v.y = p[i & 1].y; v.z = p[i & 1].z; v.w = p[i & 1].w;
                    p[(i & 1) ^ 1].x = v.x; p[(i & 1) ^ 1].y = v.y; p[(i & 1) ^ 1].z = v.z; p[(i & 1) ^ 1].w = v.w;
                }
            }
        }
    }
    if (tid == 0) { clk[0] = best; clk[1] = ibest; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

int main() {
    long long *dclk;
    cudaMalloc(&dclk, 16);
    const int smem = (2048 + 4096 + 8192) * 4;
    cudaFuncSetAttribute(issue_time, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    auto run = [&](int mode, int ncol, int nissue, int dep, int htype, int hwarps, int hsched) {
        const int nmma = 16;
        issue_time<<<1, 1024, smem>>>(mode, ncol, nmma, nissue, dep, htype, hwarps, hsched, 300, dclk);
        cudaError_t e = cudaDeviceSynchronize();
        long long c[2] = {0, 0};
        cudaMemcpy(c, dclk, 16, cudaMemcpyDeviceToHost);
        const int total = 3 * nmma * nissue;
        printf("%s N=%3d issuers=%d dep=%d hammer(type %d, %2d warps, sched %d): %3d MMAs done in %6lld clk = %6.1f clk/MMA overall "
               "(ideal N/2 = %3d), issue loop of warp 0: %6lld clk = %5.1f per MMA  (%s)\n",
               mode ? "TS" : "SS", ncol, nissue, dep, htype, hwarps, hsched, total, c[0], (double)c[0] / total, ncol / 2, c[1],
               (double)c[1] / (3 * nmma), cudaGetErrorString(e));
        if (e != cudaSuccess) exit(1);
    };
    for (int mode = 0; mode < 2; ++mode)
        for (int ncol : {32, 128, 256})
            for (int nissue : {1, 2, 4})
                for (int dep = 0; dep < 2; ++dep) if (nissue * ncol <= 448) run(mode, ncol, nissue, dep, 0, 0, 0);
    for (int mode = 0; mode < 2; ++mode)
        for (int ncol : {128, 256})
            for (int htype : {1, 2})
                for (int hwarps : {6, 12, 18})
                    for (int hsched = 0; hsched < 2; ++hsched) run(mode, ncol, 1, 1, htype, hwarps, hsched);
    return 0;
}
