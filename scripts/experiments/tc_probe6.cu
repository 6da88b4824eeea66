// Probe 6 (round 2): which shared-memory float does a TS-form tcgen05.mma kind::tf32 read for B element (n, k) with an MN-MAJOR descriptor?
// A[row][k] = (k == row % 8), so D[row][n] = B(n, row % 8); shared memory holds its own float index.
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


constexpr int ACOL = 448;
__global__ void __launch_bounds__(128, 1) mn_probe(float *D, int N, int lbo16, int sbo16, unsigned extra, int *status, unsigned long long layout_type) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8192; i += 128) sm[i] = (float)(i % 2048);
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
    float a[8], z[32];
    for (int k = 0; k < 8; ++k) a[k] = (k == tid % 8) ? 1.0f : 0.0f;
    for (int k = 0; k < 32; ++k) z[k] = -7.0f;
    const uint32_t lane0 = tbase + ((uint32_t)(warp * 32) << 16);
    tmem_st8(lane0 + ACOL, a);
    for (int c = 0; c < 64; c += 8) tmem_st8(lane0 + c, z);   // D pre-filled with -7: an MMA that does not write shows
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | extra | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    if (tid == 0) {
        const uint64_t dB = make_desc(smem_u32(sm), lbo16 * 16, sbo16 * 16) | (layout_type << 61);
        mma_ts(tbase, tbase + ACOL, dB, idesc, 0u);
        commit(smem_u32(&mbar));
    }
    const bool ok = mbar_wait(smem_u32(&mbar), 0);
    if (!ok) {
        if (tid == 0) *status = 1;
    } else {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[32];
        tmem_ld32(lane0, v);
        for (int j = 0; j < 32; ++j) D[tid * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

int main() {
    float *dD; int *dS;
    cudaMalloc(&dD, 128 * 32 * 4); cudaMalloc(&dS, 4);
    cudaFuncSetAttribute(mn_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 4);
    struct Case { int N, lbo, sbo; unsigned extra; const char *name; unsigned long long lt; };
    const Case cases[] = {{16, 64, 8, 0u, "K-major no swizzle LBO 64 SBO 8", 0},
                          {16, 64, 128, 1u << 16, "MN-major SW32  LBO 64 SBO 128", 6}, {16, 128, 64, 1u << 16, "MN-major SW32  LBO 128 SBO 64", 6},
                          {16, 64, 128, 1u << 16, "MN-major SW64  LBO 64 SBO 128", 4}, {16, 64, 128, 1u << 16, "MN-major SW128 LBO 64 SBO 128", 2},
                          {32, 64, 128, 1u << 16, "MN-major SW128 N 32 LBO 64 SBO 128", 2}, {16, 64, 128, 1u << 16, "MN-major SW128_BASE32B", 1},
                          {16, 64, 128, 0u, "K-major SW32 LBO 64 SBO 128", 6}};
    for (const Case &c : cases) {
        cudaMemset(dS, 0, 4);
        mn_probe<<<1, 128, 8192 * 4>>>(dD, c.N, c.lbo, c.sbo, c.extra, dS, c.lt);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> h(128 * 32);
        cudaMemcpy(h.data(), dD, h.size() * 4, cudaMemcpyDeviceToHost);
        printf("== %s: %s\n", c.name, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        for (int row = 0; row < 8; ++row) {   // row = k
            printf("  k=%d:", row);
            for (int n = 0; n < 20; ++n) printf(" %5.0f", h[row * 32 + n]);
            printf("\n");
        }
    }
    return 0;
}
