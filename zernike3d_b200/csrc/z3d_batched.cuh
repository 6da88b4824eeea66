// zernike3d_b200 - batched analysis on the 5th-generation tensor cores: tiles of up to 32 volumes.
//
// For ONE volume the per-l contraction needs the per-volume panel a*P (a = folded volume x row trig), so nothing but the
// R panel is shared between volumes (k_decompose).  For a BATCH the volume-independent product
//     T[(n,l,m)][k] = R_nl(r_k) * Q_lm(theta_k)          (Q = P_lm / kappa_lm, one value per moment and folded voxel)
// is generated once and the analysis becomes, per (m, z-parity pz = (l+m)&1), a real GEMM over the voxel axis
//     D[(n,l)][(re/im, b)] += T[(n,l,m)][k] * a[(m,pz)][k][(re/im, b)],      a = F_A,b trig_A + F_B,b trig_B,
// i.e. exactly "[n x V] . [V x m], batched across volumes" with the batch as the GEMM's N dimension (64 columns for 32
// volumes).  The GEMM runs on tcgen05.mma (kind::tf32, M = 128 rows of one (m, pz), N = 64, K = 8 voxels = one chunk)
// with the accumulators in tensor memory.  TF32 alone is far too coarse for the 1e-5 parity bar, so both operands are
// split into a TF32-exact high part and a float32 remainder and three products are issued per tile (lo.hi, hi.lo, hi.hi:
// "3xTF32", relative product error 2^-21).  The tensor core adds into the accumulator with truncation (measured with
// scripts/experiments/tc_probe.cu: about -3e-8 relative per accumulating instruction), so every BFLUSH_CHUNKS chunks the
// accumulators are read back (tcgen05.ld) and added, round-to-nearest, into the CTA's float32 partial buffer (L2
// resident), the next product overwrites instead of accumulating, and the a panels are pre-scaled by the inverse of the
// expected truncation loss their products will suffer until the read-back (BTRUNC_PER_MMA).
//
// Three kernels:
//   k_build_basis  (once per plan - this is what is left of the reference's CalculatePolynomials, zm:2237-2323): the
//                  recurrences of every chunk of 8 folded voxels -> R[chunk][(n,l)][8], Q[chunk][(l,m)][8] and the row
//                  trig tables; 16-fold mirror / transpose folding makes these 16x smaller than the voxel grids;
//   k_fold_tile    (per tile of volumes): the 16 mirror / transpose images of every chunk -> F[chunk][16][volume][8];
//   k_decompose_batched: a CTA owns a row group of <= 8 tiles of 128 rows (all 512 tensor-memory columns) and one split
//                  of the chunk list.  Per chunk one thread issues three bulk copies (cp.async.bulk + mbarrier: F, R and
//                  the row group's slice of Q) two chunks ahead; 31 producer warps turn them into the T and a operand
//                  panels (K-major core-matrix layout of the shared-memory descriptors); the same thread issues the MMAs.
#pragma once
#if defined(Z3D_SYNTH_DEBUG) || defined(Z3D_WAIT_DIAG)
#include <cstdio>
#endif
#include "z3d_kernels.cuh"

namespace z3d {

constexpr int BNT = 1024;           // threads per CTA of the contraction kernel (31 producer warps + 1 issuing warp; accumulators in TMEM)
constexpr int BNWARP = BNT / 32;
constexpr int BNPROD = BNT - 32;    // producer threads: the last warp only issues the MMAs and the bulk copies
constexpr int BTM = 128;            // rows of an MMA tile (UMMA M)
constexpr int BTILES_MAX = 8;       // tiles per row group at most (32-volume tiles)
constexpr int BQ_MAX = 192;         // Q values per voxel in a row group's slice
constexpr int BSLOT_WORDS = 512 * BTM;  // partial sums of one CTA: [tensor-memory column = tile * COL + col][row]
constexpr int BFLUSH_CHUNKS = 32;   // accumulator read-back interval (see above)
// Every accumulating MMA rounds the running sum toward zero: an expected relative loss of c = 3.1e-8 per instruction
// (scripts/experiments/tc_probe.cu: -8e-8 per chunk of 3 on monotone sums; scripts/bias_check.py: a uniform shrink of
// -4.7e-8 x chunks of the moments against the float64 oracle, the same for 5, 11 and 32 chunk intervals).  The products
// of chunk k of an interval of n chunks are followed by 3 (n - 1 - k) + 1 truncations, i.e. arrive at the read-back
// multiplied by (1 - c)^(3 (n - 1 - k) + 1): the a panels of that chunk are pre-scaled by the inverse, which removes the
// bias for any distribution of the volume's weight over the chunks (a single factor at the read-back only does on average).
constexpr float BTRUNC_PER_MMA = 3.13e-8f;

// compile-time geometry of the two tile widths: NV = 32 volumes (N = 64 MMAs, 8 tiles of 128 rows per row group) or
// NV = 64 (N = 128 MMAs, which run at the full tensor-pipe rate and fetch the T operand once per 128 columns; 4 tiles)
template <int NV>
struct BCfg {
    static constexpr int COL = 2 * NV;                 // GEMM columns: [re | im][volume]
    static constexpr int TILES = 512 / COL;            // tiles per row group: all of tensor memory
    static constexpr int ROWS = TILES * BTM;
    static constexpr int F_BYTES = 16 * NV * KC * 4;   // folds of one chunk: [set 2][combo 8][volume][k], k halves swizzled
    static constexpr int A_WORDS = TILES * 2 * COL * 4;  // one a panel set (hi or lo): [panel][2 k-halves][COL][4]
    static constexpr int OFF_THI = 0;                              // T panels, TF32 high part [2 k-halves][ROWS][4]
    static constexpr int OFF_TLO = OFF_THI + 2 * ROWS * 16;        //           float32 remainder
    static constexpr int OFF_ROWTAB = OFF_TLO + 2 * ROWS * 16;     // (R index, Q index in the slice) per row [ROWS] ushort2
    static constexpr int OFF_TABS = OFF_ROWTAB + ROWS * 4;         // mpz, tile_mpz [2][BTILES_MAX]
    static constexpr int OFF_MBAR = OFF_TABS + 2 * BTILES_MAX * 4; // mbarriers: full[2], mma; tensor-memory base address
    static constexpr int OFF_A = (OFF_MBAR + 32 + 127) / 128 * 128;  // a panels [1 or 2 buffers][hi | lo], then 2 copy stages
    static_assert(OFF_TLO % 128 == 0 && OFF_MBAR % 8 == 0 && (A_WORDS * 4) % 128 == 0, "smem alignment");
    // stage: F, then R with extra all-zero entries (padding rows point at the first), then the Q slice
    __host__ __device__ static int stage_bytes(int nrad) { return F_BYTES + (nrad + 4) * 32 + BQ_MAX * 32; }
    __host__ __device__ static int smem_bytes(int nrad, bool adouble) {
        return OFF_A + 2 * A_WORDS * 4 * (adouble ? 2 : 1) + 2 * stage_bytes(nrad);
    }
};

struct DevPlanB {
    int order, dim, H;
    int ngroups;             // row groups
    int gpr, nrounds;        // row groups per round, rounds: one launch of the contraction kernel per round (see z3d_api.cu)
    int nrad, nq;            // entries per chunk of the R and Q tables
    int adouble;             // 1: two a-panel buffers fit in shared memory (the a panels of a chunk are then produced while the
                             //    previous chunk's MMAs run)
    const int2 *rowtab;      // [ngroups][ROWS]   (R index, Q index relative to the group's slice); padding -> zero entry
    const int *tile_mpz;     // [ngroups][BTILES_MAX]     a-panel index of each tile
    const int *mpz;          // [ngroups][BTILES_MAX]    m | pz << 16 of each a panel, -1 padded
    const int *nmpz;         // [ngroups]
    const int *ntile;        // [ngroups]
    const int *qlo, *qlen;   // [ngroups]              the row group's slice of the Q table
    // basis tables (k_build_basis)
    const float *Rt;         // [chunks][nrad][8]      k halves swizzled by (index >> 2) & 1
    const float *Qt;         // [chunks][nq][8]
    const float *trig;       // [items][4][LMAX + 1]   cos / sin(m phi) of the row group and of its transpose
};

// ---- tcgen05 / mbarrier wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor, K-major, no swizzle: 8 x 16-byte core matrices; LBO = byte distance between the two
// 16-byte K halves of a row, SBO = byte distance between groups of 8 rows.  Element (row, k) of an operand lives at
//   (k / 4) * LBO + (row / 8) * SBO + (row % 8) * 16 + (k % 4) * 4.
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((saddr >> 4) & 0x3fff) | ((unsigned long long)((lbo >> 4) & 0x3fff) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);  // bits 46-47: descriptor version 1
}
__device__ __forceinline__ void umma_tf32(unsigned d_tmem, unsigned long long a, unsigned long long b, unsigned idesc,
                                          unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned mbar) {  // the mbarrier completes a phase when all MMAs issued so far retire
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
#pragma unroll 1
    for (int it = 0; it < (1 << 26); ++it) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (ok) return;
    }
#if defined(Z3D_SYNTH_DEBUG) || defined(Z3D_WAIT_DIAG)
    if ((threadIdx.x & 31) == 0) printf("wait timed out: block (%d,%d) warp %d barrier @%u parity %u\n", blockIdx.x, blockIdx.y, threadIdx.x >> 5, mbar, parity);
#endif
    __trap();  // never hang the device: a lost completion is a bug, fail the launch instead
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// this warp's 32 lanes x 32 consecutive columns of tensor memory -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&v)[32]) {
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
// x = hi + lo with hi exactly representable in TF32 (round to nearest); the tensor core truncates lo to TF32 itself
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    unsigned h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    lo = x - hi;
}

// bulk asynchronous copy global -> shared (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_copy(void *dst, const void *src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}

// position in the chunk list: row-group item (ip, jp, paired) and chunk kc inside it
struct ChunkPos {
    int item, kc, ip, jp;
    bool paired;
};
__device__ __forceinline__ void chunk_pos(long long cc, int CPR, int H, ChunkPos &p) {
    p.item = (int)(cc / CPR);
    p.kc = (int)(cc % CPR);
    decode_item(p.item, 0, H, H, p.ip, p.jp, p.paired);
}

// ---------------------------------------------------------------------------------------------------------------------
// Basis tables: one CTA per chunk, one thread per (chain, voxel).  Same float32 recurrences as the per-volume kernel
// (generate_panels): R_n = (K1 r^2 + K2) R_{n-2} + K3 R_{n-4} (zm:1991-2000) and the monic parity-decoupled Legendre
// chains Q_{l+2} = (4c^2 - gamma) Q_l - delta Q_{l-2} (zm:2006-2026 / kappa_lm); geometry in float64 (zm:1886-1890).
// ---------------------------------------------------------------------------------------------------------------------
struct BasisArgs {
    int order, dim, H, nrad, nq;
    const double *side;
    const float4 *rk;      // [nrad + pad]  (K1, K2, K3, 0), l-major
    const int *rbase;      // [L+1]
    const float4 *gd;      // (gamma_E, -delta_E, gamma_O, -delta_O) per (m, step), all m
    const int *goff;       // [L+1]
    const int *qoff;       // [L+1]  index of Q_mm in the Q table (Q_lm at qoff[m] + l - m)
    float *Rt, *Qt, *trig;
};
__global__ void __launch_bounds__(1024) k_build_basis(const BasisArgs A) {
    const int L = A.order, H = A.H, S = L + 1;
    const int CPR = (H + KC - 1) / KC;
    const long long chunk = blockIdx.x;
    ChunkPos p;
    chunk_pos(chunk, CPR, H, p);
    const double y0 = A.side[p.ip], x0 = A.side[p.jp];
    const double rho2 = x0 * x0 + y0 * y0, rho = sqrt(rho2);
    const int k = threadIdx.x % KC, kk = p.kc * KC + k;
    const double z = (kk < H) ? A.side[kk] : 0.0;
    const double r2d = rho2 + z * z, rd = sqrt(r2d);
    double ctd = 1.0, std_ = 0.0;  // r == 0: theta is NaN in the reference and nan_to_num zeroes Y_lm, l > 0 (zm:2298)
    if (rd > 0.0) {
        ctd = z / rd;
        std_ = rho / rd;
    }
    const float c2 = (float)(2.0 * ctd), c2sq = (float)(4.0 * ctd * ctd), st = (float)std_, st2 = (float)(std_ * std_);
    const float r = (float)rd, r2 = (float)r2d;
    for (int ch = threadIdx.x / KC; ch < 2 * S; ch += blockDim.x / KC) {
        if (ch < S) {  // radial chain of l
            const int l = ch, nn = (L - l) / 2 + 1, base = A.rbase[l];
            float R1 = 0.0f, R2 = powi2(r, r2, l);
            float *dst = A.Rt + ((size_t)chunk * A.nrad + base) * KC;
            for (int ni = 0; ni < nn; ++ni) {
                const float4 K = A.rk[base + ni];
                const float R = fmaf(fmaf(K.x, r2, K.y), R1, K.z * R2);
                dst[ni * KC + (k ^ ((((base + ni) >> 2) & 1) << 2))] = R;
                R2 = R1;
                R1 = R;
            }
        } else {  // Legendre chains of m: E (l = m, m+2, ..) and O (l = m+1, m+3, ..)
            const int m = ch - S, nE = (L - m) / 2 + 1, nO = (m < L) ? (L - m - 1) / 2 + 1 : 0, q0 = A.qoff[m];
            float qE1 = powi2(st, st2, m), qE2 = 0.f, qO1 = qE1 * c2, qO2 = 0.f;
            const float4 *t0 = A.gd + A.goff[m];
            float *dst = A.Qt + (size_t)chunk * A.nq * KC;
            for (int i = 0; i < nE; ++i) {
                const float4 g = t0[i];
                const int qe = q0 + 2 * i, qo = qe + 1;
                dst[qe * KC + (k ^ (((qe >> 2) & 1) << 2))] = qE1;
                dst[qo * KC + (k ^ (((qo >> 2) & 1) << 2))] = (i < nO) ? qO1 : 0.0f;
                const float nE1 = fmaf(c2sq - g.x, qE1, g.y * qE2), nO1 = fmaf(c2sq - g.z, qO1, g.w * qO2);
                qE2 = qE1; qE1 = nE1; qO2 = qO1; qO1 = nO1;
            }
        }
    }
    // cos / sin(m phi) of the row group and of its transpose, float64 complex powering (exp(1j m phi), zm:2297)
    if (p.kc == 0 && threadIdx.x <= L) {
        const int m = threadIdx.x;
        double br = 1.0, bi = 0.0;
        if (rho > 0.0) {
            br = x0 / rho;
            bi = y0 / rho;
        }
        double cr = 1.0, ci = 0.0;
        for (int e = m; e; e >>= 1) {
            if (e & 1) {
                const double t = cr * br - ci * bi;
                ci = cr * bi + ci * br;
                cr = t;
            }
            const double t = br * br - bi * bi;
            bi = 2.0 * br * bi;
            br = t;
        }
        float *tr = A.trig + (size_t)p.item * 4 * (LMAX + 1);
        const int q = m & 3;  // transpose: phi' = atan2(x0, y0) = pi/2 - phi (mod 2 pi)
        tr[m] = (float)cr;
        tr[(LMAX + 1) + m] = (float)ci;
        tr[2 * (LMAX + 1) + m] = (float)((q == 0) ? cr : (q == 1) ? ci : (q == 2) ? -cr : -ci);
        tr[3 * (LMAX + 1) + m] = (float)((q == 0) ? -ci : (q == 1) ? cr : (q == 2) ? ci : -cr);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Folds of a tile of volumes: F[chunk][set][combo][volume][k] = 8-point Walsh-Hadamard transform of the mirror images of
// voxel k of the chunk's row group (set 0) and of its transpose (set 1; zero on the diagonal).  One CTA per chunk.
// ---------------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(2 * NV * KC) k_fold_tile(const float *__restrict__ vol, int nvol, int N, int H,
                                                            float *__restrict__ Ft) {
    const int CPR = (H + KC - 1) / KC;
    const long long chunk = blockIdx.x;
    ChunkPos p;
    chunk_pos(chunk, CPR, H, p);
    const int tid = threadIdx.x;
    const int k = tid % KC, b = (tid / KC) % NV, set = tid / (KC * NV);
    const int kk = p.kc * KC + k, k2 = N - 1 - kk;
    const int ri = set ? p.jp : p.ip, rj = set ? p.ip : p.jp;
    const int i2 = N - 1 - ri, j2 = N - 1 - rj;
    const bool live = b < nvol && kk < H && (set == 0 || p.paired);
    const float *vb = vol + (size_t)b * N * N * N;
    float v[8];
#pragma unroll
    for (int img = 0; img < 8; ++img) {
        const bool sy = img & 4, sx = img & 2, sz = img & 1;
        const bool dup = (sy && i2 == ri) || (sx && j2 == rj) || (sz && k2 == kk);
        v[img] = (live && !dup) ? __ldg(vb + ((size_t)(sy ? i2 : ri) * N + (sx ? j2 : rj)) * N + (sz ? k2 : kk)) : 0.0f;
    }
    wht8(v);
    float *dst = Ft + (size_t)chunk * (BCfg<NV>::F_BYTES / 4);
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) dst[((set * 8 + c8) * NV + b) * KC + (k ^ (((b >> 2) & 1) << 2))] = v[c8];
}

// ---------------------------------------------------------------------------------------------------------------------
// (a) T rows = R * Q;  (b) a panels: a = F_A trig_A + F_B trig_B per (m, pz), re / im and volume, 8 voxels at a time
//     (re: even in y, (-1)^m in x; im: odd in y, -(-1)^m in x, negated; z parity pz).  Both are split into TF32 high
//     parts and remainders and written in the K-major core-matrix layout of umma_desc: [k half][row | column][4].
// ---------------------------------------------------------------------------------------------------------------------
struct CtxB {
    float *Thi, *Tlo, *Ahi, *Alo;
    const ushort2 *rowtab;
    const int *mpz, *tile_mpz;
};
// T row = R * Q of this thread's row: loads, products and the TF32 split (registers only) ...
struct TRow {
    float4 t0, t1;  // the eight float32 products; the TF32 split happens at the store: 8 live registers, not 16 (with 12
};                  // or 16 the compiler spills them across the MMA wait, and the reload sits on the critical path)
template <int F_BYTES>
__device__ __forceinline__ TRow t_compute(const CtxB &c, const float *stage, int nrad, int qlo, int row) {
    const float *Rs = stage + F_BYTES / 4, *Qs = Rs + (nrad + 4) * KC;
    const ushort2 rq = c.rowtab[row];
    const float *rp = Rs + rq.x * KC, *qp = Qs + rq.y * KC;
    const int sr = ((rq.x >> 2) & 1) << 2, sq = (((rq.y + qlo) >> 2) & 1) << 2;  // the tables' k-half swizzle
    const float4 r0 = *reinterpret_cast<const float4 *>(rp + sr), r1 = *reinterpret_cast<const float4 *>(rp + (4 ^ sr));
    const float4 q0 = *reinterpret_cast<const float4 *>(qp + sq), q1 = *reinterpret_cast<const float4 *>(qp + (4 ^ sq));
    TRow t;
    t.t0 = make_float4(r0.x * q0.x, r0.y * q0.y, r0.z * q0.z, r0.w * q0.w);
    t.t1 = make_float4(r1.x * q1.x, r1.y * q1.y, r1.z * q1.z, r1.w * q1.w);
    return t;
}
__device__ __forceinline__ void split4(const float4 &x, float4 &hi, float4 &lo) {
    split_tf32(x.x, hi.x, lo.x);
    split_tf32(x.y, hi.y, lo.y);
    split_tf32(x.z, hi.z, lo.z);
    split_tf32(x.w, hi.w, lo.w);
}
// ... and its split and four 16-byte stores into the operand panels, once the previous chunk's MMAs have released them
template <int ROWS>
__device__ __forceinline__ void t_store(const CtxB &c, const TRow &t, int row) {
    float4 hi, lo;
    split4(t.t0, hi, lo);
    *reinterpret_cast<float4 *>(c.Thi + row * 4) = hi;
    *reinterpret_cast<float4 *>(c.Tlo + row * 4) = lo;
    split4(t.t1, hi, lo);
    *reinterpret_cast<float4 *>(c.Thi + (ROWS + row) * 4) = hi;
    *reinterpret_cast<float4 *>(c.Tlo + (ROWS + row) * 4) = lo;
}
// the same for one k half (4 voxels) of a row: the 64-volume kernel has two threads per row
struct THalf {
    float4 t;
};
template <int F_BYTES>
__device__ __forceinline__ THalf t_compute_half(const CtxB &c, const float *stage, int nrad, int qlo, int row, int kh) {
    const float *Rs = stage + F_BYTES / 4, *Qs = Rs + (nrad + 4) * KC;
    const ushort2 rq = c.rowtab[row];
    const int sr = ((rq.x >> 2) & 1) << 2, sq = (((rq.y + qlo) >> 2) & 1) << 2;  // the tables' k-half swizzle
    const float4 r = *reinterpret_cast<const float4 *>(Rs + rq.x * KC + ((4 * kh) ^ sr));
    const float4 q = *reinterpret_cast<const float4 *>(Qs + rq.y * KC + ((4 * kh) ^ sq));
    THalf t;
    t.t = make_float4(r.x * q.x, r.y * q.y, r.z * q.z, r.w * q.w);
    return t;
}
template <int ROWS>
__device__ __forceinline__ void t_store_half(const CtxB &c, const THalf &t, int row, int kh) {
    float4 hi, lo;
    split4(t.t, hi, lo);
    *reinterpret_cast<float4 *>(c.Thi + (kh * ROWS + row) * 4) = hi;
    *reinterpret_cast<float4 *>(c.Tlo + (kh * ROWS + row) * 4) = lo;
}
template <int NV>
__device__ __forceinline__ void a_products(const CtxB &c, const float *stage, int nmpz, const float *__restrict__ trig, float comp) {
    constexpr int BNV = NV, BCOL = 2 * NV;
    const int tid = threadIdx.x;
    const float *Fs = stage;
    for (int it = tid; it < nmpz * BCOL; it += BNPROD) {  // (panel, re / im, volume); producer threads only
        const int b = it % NV, reim = (it / NV) & 1, j = it / BCOL;
        const int code = c.mpz[j], m = code & 0xffff, pz = code >> 16;
        const int comb = (reim << 2) | (((m & 1) ^ reim) << 1) | pz;
        float tA = __ldg(trig + reim * (LMAX + 1) + m), tB = __ldg(trig + (2 + reim) * (LMAX + 1) + m);
        const float sg = reim ? -comp : comp;  // comp: truncation pre-compensation of this chunk (see BTRUNC_PER_MMA)
        tA *= sg;
        tB *= sg;
        const int sw = ((b >> 2) & 1) << 2;  // k_fold_tile's k-half swizzle: 8 consecutive volumes cover all banks
        const float *fa = Fs + (comb * BNV + b) * KC, *fb = fa + 8 * BNV * KC;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
            const float4 a = *reinterpret_cast<const float4 *>(fa + ((4 * kh) ^ sw)), bq = *reinterpret_cast<const float4 *>(fb + ((4 * kh) ^ sw));
            float4 hi, lo;
            split_tf32(fmaf(a.x, tA, bq.x * tB), hi.x, lo.x);
            split_tf32(fmaf(a.y, tA, bq.y * tB), hi.y, lo.y);
            split_tf32(fmaf(a.z, tA, bq.z * tB), hi.z, lo.z);
            split_tf32(fmaf(a.w, tA, bq.w * tB), hi.w, lo.w);
            const int o = ((j * 2 + kh) * BCOL + reim * BNV + b) * 4;
            *reinterpret_cast<float4 *>(c.Ahi + o) = hi;
            *reinterpret_cast<float4 *>(c.Alo + o) = lo;
        }
    }
}

// one thread: the three TF32 products of every tile of the chunk, then a commit that completes the mbarrier phase when
// they (and their shared-memory reads) have retired
template <int NV>
__device__ __forceinline__ void issue_mma(const CtxB &c, unsigned tmem, int ntile, bool accumulate, unsigned mbar) {
    constexpr int BCOL = BCfg<NV>::COL, BROWS_MAX = BCfg<NV>::ROWS;
    // instruction descriptor: D float32, A and B TF32, both K-major, N = COL, M = 128
    constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(BCOL >> 3) << 17) | ((unsigned)(BTM >> 4) << 24);
    const unsigned long long dThi = umma_desc(smem_u32(c.Thi), BROWS_MAX * 16, 128), dTlo = umma_desc(smem_u32(c.Tlo), BROWS_MAX * 16, 128);
    const unsigned long long dAhi = umma_desc(smem_u32(c.Ahi), BCOL * 16, 128), dAlo = umma_desc(smem_u32(c.Alo), BCOL * 16, 128);
    for (int t = 0; t < ntile; ++t) {
        const unsigned long long ta = (unsigned long long)(t * BTM * 16 >> 4), aa = (unsigned long long)(c.tile_mpz[t] * 2 * BCOL * 16 >> 4);
        const unsigned d = tmem + t * BCOL;
        umma_tf32(d, dTlo + ta, dAhi + aa, idesc, accumulate ? 1u : 0u);  // small terms first
        umma_tf32(d, dThi + ta, dAlo + aa, idesc, 1u);
        umma_tf32(d, dThi + ta, dAhi + aa, idesc, 1u);
    }
    umma_commit(mbar);
}

// accumulators -> this CTA's partial buffer [tensor-memory column][row] (float32, round to nearest); warp w reads the lane
// quarter w % 4 (the only one it can address) of the 64 columns 64 (w / 4) .. 64 (w / 4) + 63
__device__ __forceinline__ void flush_accumulators(unsigned tmem, int ncols, float *out, bool add) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3;
#pragma unroll 1
    for (int c64 = 64 * (warp >> 2); c64 < ncols; c64 += 64 * (BNWARP / 4)) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            unsigned v[32];
            tmem_ld32(tmem + ((unsigned)(32 * q) << 16) + c64 + 32 * half, v);
            float *p = out + ((size_t)c64 + 32 * half) * BTM + 32 * q + lane;
            if (add) {
#pragma unroll
                for (int col = 0; col < 32; ++col) p[col * BTM] += __uint_as_float(v[col]);
            } else {
#pragma unroll
                for (int col = 0; col < 32; ++col) p[col * BTM] = __uint_as_float(v[col]);
            }
        }
    }
}

// partial layout: [round][CTA of the round][BSLOT_WORDS]
//
// Per chunk t (stage s = t & 1):  wait full[s] (bulk copies of chunk t)  |  wait for the MMAs of chunk t-1 (panels free)
//   | products  |  barrier  |  one thread: MMAs of chunk t (asynchronous) + the bulk copies of chunk t+2 into stage s.
// One launch per ROUND of row groups: CTAs [0, groups x cpg) - group g0 + blockIdx / cpg, split blockIdx % cpg of the
// launch's chunk range.  All CTAs of a launch sweep the chunk list in lock step, so the R / Q / F data of a chunk is
// fetched from DRAM once per round and served to the round's row groups from L2 (see z3d_api.cu for the choice of rounds).
template <int NV>
__global__ void __launch_bounds__(BNT, 1)
k_decompose_batched(const DevPlanB P, const float *__restrict__ Ft, int part, int nparts, int g0, int cpg,
                    float *__restrict__ partial /* slots of this round */) {
    using G = BCfg<NV>;
    constexpr int BROWS_MAX = G::ROWS, BF_BYTES = G::F_BYTES, BA_WORDS = G::A_WORDS;
    constexpr int BOFF_THI = G::OFF_THI, BOFF_TLO = G::OFF_TLO, BOFF_ROWTAB = G::OFF_ROWTAB, BOFF_TABS = G::OFF_TABS;
    constexpr int BOFF_MBAR = G::OFF_MBAR, BOFF_A = G::OFF_A;
    extern __shared__ __align__(16) unsigned char smem[];
    const int H = P.H, nrad = P.nrad;
    const int tid = threadIdx.x;
    const int grp = g0 + blockIdx.x / cpg, split = blockIdx.x % cpg;
    CtxB c;
    c.Thi = reinterpret_cast<float *>(smem + BOFF_THI);
    c.Tlo = reinterpret_cast<float *>(smem + BOFF_TLO);
    ushort2 *rowtab = reinterpret_cast<ushort2 *>(smem + BOFF_ROWTAB);
    int *mpz = reinterpret_cast<int *>(smem + BOFF_TABS), *tile_mpz = mpz + BTILES_MAX;
    c.rowtab = rowtab;
    c.mpz = mpz;
    c.tile_mpz = tile_mpz;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + BOFF_MBAR);
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem + BOFF_MBAR + 24);
    const int stage_bytes = G::stage_bytes(nrad);
    const int adouble = P.adouble;
    float *abase = reinterpret_cast<float *>(smem + BOFF_A);
    unsigned char *stage0 = smem + BOFF_A + 2 * BA_WORDS * 4 * (adouble ? 2 : 1);
    // ---- tables of this row group -> smem; zero the panels and the stages (padding rows read the zero R entry)
    for (int i = tid; i < BTILES_MAX; i += BNT) {
        mpz[i] = P.mpz[grp * BTILES_MAX + i];
        tile_mpz[i] = P.tile_mpz[grp * BTILES_MAX + i];
    }
    for (int i = tid; i < BROWS_MAX; i += BNT) {
        const int2 rq = P.rowtab[(size_t)grp * BROWS_MAX + i];
        rowtab[i] = make_ushort2((unsigned short)rq.x, (unsigned short)rq.y);
    }
    for (int i = tid; i < BOFF_ROWTAB / 4; i += BNT) reinterpret_cast<float *>(smem)[i] = 0.0f;
    for (int i = tid; i < (int)(stage0 + 2 * stage_bytes - smem - BOFF_A) / 4; i += BNT) abase[i] = 0.0f;
    const int nmpz = P.nmpz[grp], ntile = P.ntile[grp], qlo = P.qlo[grp], qlen = P.qlen[grp];
    const unsigned full0 = smem_u32(bars), mbar = smem_u32(bars + 2);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full0 + 8));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {  // warp 0 owns the tensor-memory allocation: all 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_async_smem();  // the zeroed stages are written by the async proxy next
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    const int CPR = (H + KC - 1) / KC;
    const long long total = (long long)analysis_items(0, H, H) * CPR;
    const long long gbeg = total * part / nparts, glen = total * (part + 1) / nparts - gbeg;
    const long long cbeg = gbeg + glen * split / cpg, cend = gbeg + glen * (split + 1) / cpg;
    float *out = partial + (size_t)blockIdx.x * BSLOT_WORDS;
    const unsigned rbytes = (unsigned)nrad * 32u, qbytes = (unsigned)qlen * 32u;

    // one thread: F, R and the Q slice of chunk cc -> stage s, completion on full[s]
    auto issue_copies = [&](long long cc, int s) {
        unsigned char *st = stage0 + (size_t)s * stage_bytes;
        const unsigned bar = full0 + 8 * s;
        mbar_expect_tx(bar, BF_BYTES + rbytes + qbytes);
        bulk_copy(st, Ft + (size_t)cc * (BF_BYTES / 4), BF_BYTES, bar);
        bulk_copy(st + BF_BYTES, P.Rt + (size_t)cc * nrad * KC, rbytes, bar);
        bulk_copy(st + BF_BYTES + (nrad + 4) * 32, P.Qt + ((size_t)cc * P.nq + qlo) * KC, qbytes, bar);
    };
    if (tid == BNPROD) {
        if (cbeg < cend) issue_copies(cbeg, 0);
        if (cbeg + 1 < cend) issue_copies(cbeg + 1, 1);
    }
    int since_flush = 0;  // chunks accumulated in tensor memory
    unsigned t = 0;       // chunk counter: stage t & 1, full[] parity (t >> 1) & 1, MMA commit t has parity t & 1
    bool flushed = false;
    for (long long cc = cbeg; cc < cend; ++cc, ++t) {
        const int s = t & 1;
        const float *stage = reinterpret_cast<const float *>(stage0 + (size_t)s * stage_bytes);
        mbar_wait(full0 + 8 * s, (t >> 1) & 1);  // this chunk's F, R, Q have landed
        c.Ahi = abase + (adouble ? (int)(t & 1) * 2 * BA_WORDS : 0);
        c.Alo = c.Ahi + BA_WORDS;
        const float *trig = P.trig + (size_t)(cc / CPR) * 4 * (LMAX + 1);
        // position of this chunk in its read-back interval and the interval's length (the last one may be shorter)
        const int pos = since_flush == BFLUSH_CHUNKS ? 0 : since_flush;
        const int ilen = (int)min((long long)BFLUSH_CHUNKS, (cend - cc) + pos);
        const float comp = 1.0f + BTRUNC_PER_MMA * (float)(3 * (ilen - 1 - pos) + 1);
        // everything that does not touch the panels the previous chunk's MMAs are still reading comes first: the a panels
        // (other buffer) and the T products, which stay in registers until the MMAs have retired.  Warps 0-30 produce
        // (one row per thread, warp 0 also the last 32 rows); warp 31 only issues, so that the ~24 MMA issues and 3 copy
        // descriptors per chunk do not delay a producer warp the whole block would then wait for at the barrier.
        // 64-volume tiles have 512 rows: two threads per row, lanes 0-15 of a warp the first k half of 16 consecutive rows,
        // lanes 16-31 the second (conflict-free loads and stores); warp 0 also does the 16 rows of the issuing warp.
        const bool producer = tid < BNPROD, two_rows = tid < BNT - BNPROD;
        const int hrow = (tid >> 5) * 16 + (tid & 15), hkh = (tid >> 4) & 1;
        TRow trow;
        THalf thalf;
        if (producer) {
            if (adouble) a_products<NV>(c, stage, nmpz, trig, comp);
            if constexpr (NV == 32) trow = t_compute<BF_BYTES>(c, stage, nrad, qlo, tid);
            else thalf = t_compute_half<BF_BYTES>(c, stage, nrad, qlo, hrow, hkh);
        }
        if (t) {  // the previous chunk's MMAs have retired: the operand panels are free, the accumulators current
            mbar_wait(mbar, (t - 1) & 1);
            tc_fence_after();
        }
        if (since_flush == BFLUSH_CHUNKS) {
            flush_accumulators(tmem, ntile * G::COL, out, flushed);
            flushed = true;
            since_flush = 0;
            tc_fence_before();
        }
        if (producer) {
            if (!adouble) a_products<NV>(c, stage, nmpz, trig, comp);
            if constexpr (NV == 32) {
                t_store<BROWS_MAX>(c, trow, tid);
                if (two_rows) t_store<BROWS_MAX>(c, t_compute<BF_BYTES>(c, stage, nrad, qlo, BNPROD + tid), BNPROD + tid);
            } else {
                t_store_half<BROWS_MAX>(c, thalf, hrow, hkh);
                if (two_rows) {
                    const int v = BNPROD + tid, row = (v >> 5) * 16 + (v & 15), kh = (v >> 4) & 1;
                    t_store_half<BROWS_MAX>(c, t_compute_half<BF_BYTES>(c, stage, nrad, qlo, row, kh), row, kh);
                }
            }
        }
        fence_async_smem();  // generic-proxy panel writes -> visible to the tensor core's async-proxy reads
        __syncthreads();     // panels complete; everybody is done reading stage s
        if (tid == BNPROD) {
            tc_fence_after();
            issue_mma<NV>(c, tmem, ntile, since_flush > 0, mbar);
            if (cc + 2 < cend) issue_copies(cc + 2, s);
        }
        ++since_flush;
    }
    if (t) {
        mbar_wait(mbar, (t - 1) & 1);
        tc_fence_after();
        flush_accumulators(tmem, ntile * G::COL, out, flushed);
    } else {
        for (int i = tid; i < BSLOT_WORDS; i += BNT) out[i] = 0.0f;  // the reduction reads every split
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// moments[b][mu][reim] = scale[mu] * sum over splits of the owning accumulator (fixed order, float64)
template <int NV>
__global__ void k_reduce_batched(const float *__restrict__ partial, const int2 *__restrict__ rowmap /* [nmom] (group, row) */,
                                 const double *__restrict__ scale, int nmom, int ngroups, int gpr, int nctas, int nvol,
                                 float *__restrict__ moments) {
    constexpr int BNV = NV, BCOL = 2 * NV;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over nmom * BCOL
    if (idx >= nmom * BCOL) return;
    const int mu = idx / BCOL, col = idx % BCOL;
    const int reim = col / BNV, b = col % BNV;
    if (b >= nvol) return;
    const int2 gr = rowmap[mu];
    // slots are [round][CTA of the round's launch]: nctas per round, the group's CTAs are contiguous in split order
    const int round = gr.x / gpr, first = round * gpr, grn = min(gpr, ngroups - first), cpg = nctas / grn;
    const float *p = partial + ((size_t)round * nctas + (size_t)(gr.x - first) * cpg) * BSLOT_WORDS +
                     ((size_t)(gr.y / BTM) * BCOL + col) * BTM + gr.y % BTM;
    double sum = 0.0;
    for (int s = 0; s < cpg; ++s) sum += (double)p[(size_t)s * BSLOT_WORDS];
    moments[((size_t)b * nmom + mu) * 2 + reim] = (float)(sum * scale[mu]);
}

}  // namespace z3d
