// Probe 5 (round 2, derived from tc_probe4.cu): what does tcgen05.commit cost in a steady issue loop?  One converged warp issues NSEG x 3
// MMAs per chunk (elect.sync) and commits onto a ring of mbarriers every CEVERY chunks (0 = only once at the end); nobody waits on the ring.
// Reports the time until the last MMA retired and the issue loop's own time.
// (header of probe 4 follows)
// Probe 4 (round 2, derived from tc_probe2.cu): how fast can ONE warp issue tcgen05.mma: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (written by tcgen05.st) against the
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


__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(p));
    return p != 0;
}
// TS-form MMA with an always-accumulate predicate folded into the instruction (no setp from a register)
__device__ __forceinline__ void mma_ts_acc(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
}


template <int NSEG>
__global__ void __launch_bounds__(128, 1) commit_loop(int ncol, int nchunk, int nslot, int cevery, long long *clk) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long mbar, ring[8];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8 * 4096; i += blockDim.x) sm[i] = 1.0f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ring[i])));
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
    {
        float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + ACOL, one);
        tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + ACOL + 8, one);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)ncol >> 3) << 17) | ((128u >> 4) << 24);
    long long best = 1ll << 60;
    if (warp == 0) {
        for (int r = 0; r < 5; ++r) {
            const long long t0 = clock64();
            const uint64_t dB0 = make_desc(smem_u32(sm), 512 * 16, 128);
            const uint64_t lo_off = (uint64_t)(8192 >> 4), slot_step = (uint64_t)(16384 >> 4);
            uint64_t dB = dB0;
            int slot = 0;
            long long tcommit = 0;
            for (int c = 0; c < nchunk; ++c) {
                if (elect_one()) {
#pragma unroll
                    for (int g = 0; g < NSEG; ++g) {
                        const uint64_t b = dB + (uint64_t)(g * ncol);
                        const uint32_t d = tbase + (uint32_t)(g * ncol), a = tbase + ACOL;
                        mma_ts_acc(d, a, b + lo_off, idesc);
                        mma_ts_acc(d, a + 8, b, idesc);
                        mma_ts_acc(d, a, b, idesc);
                    }
                    if (cevery > 0 && (c + 1) % cevery == 0) {
                        const long long c0 = clock64();
                        commit(smem_u32(&ring[c & 7]));
                        tcommit += clock64() - c0;
                    }
                }
                __syncwarp();
                dB += slot_step;
                if (++slot == nslot) { slot = 0; dB = dB0; }
            }
            if (elect_one()) commit(smem_u32(&mbar));
            __syncwarp();
            const long long ti = clock64() - t0;
            mbar_wait(smem_u32(&mbar), r & 1);
            if (lane == 0) {
                const long long dt = clock64() - t0;
                if (dt < best) { best = dt; clk[1] = ti; }
            }
            tcommit = __shfl_sync(0xffffffffu, tcommit, 0) ;
            long long tmax = tcommit;
            for (int o = 16; o; o >>= 1) { long long v = __shfl_xor_sync(0xffffffffu, tmax, o); tmax = v > tmax ? v : tmax; }
            if (lane == 0) clk[2] = tmax;
        }
    }
    if (tid == 0) clk[0] = best;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

template <int NSEG>
void run(int ncol, int cevery, long long *dclk) {
    const int smem = 8 * 4096 * 4, nchunk = 64;
    cudaFuncSetAttribute(commit_loop<NSEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    commit_loop<NSEG><<<1, 128, smem>>>(ncol, nchunk, 6, cevery, dclk);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[3] = {0, 0, 0};
    cudaMemcpy(c, dclk, 24, cudaMemcpyDeviceToHost);
    const int total = 3 * NSEG * nchunk;
    printf("NSEG=%d N=%3d commit every %d chunks: %3d MMAs retired in %6lld clk = %6.1f per chunk (pipe ideal %4d), issue loop %6lld clk = %6.1f per chunk, "
           "inside commit instructions %6lld clk  (%s)\n", NSEG, ncol, cevery, total, c[0], (double)c[0] / nchunk, 3 * NSEG * ncol / 2, c[1],
           (double)c[1] / nchunk, c[2], cudaGetErrorString(e));
    if (e != cudaSuccess) exit(1);
}

int main() {
    long long *dclk;
    cudaMalloc(&dclk, 24);
    for (int ce : {0, 1, 2, 4}) {
        run<1>(32, ce, dclk);
        run<3>(128, ce, dclk);
        run<4>(96, ce, dclk);
        run<1>(256, ce, dclk);
    }
    return 0;
}
