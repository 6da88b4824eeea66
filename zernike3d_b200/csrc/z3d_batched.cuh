// zernike3d_b200 - batched analysis kernel on the 5th-generation tensor cores: a tile of up to 32 volumes per launch.
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
// resident) and the next product overwrites instead of accumulating.
//
// Work decomposition: the rows of every (m, pz) are padded to tiles of 128; a CTA owns a row group of <= 8 tiles (all 512
// tensor-memory columns) and one split of the chunk list (row-group items x 8-voxel chunks, 16-fold mirror / transpose
// folding as in k_decompose).  All 16 warps are producers: they stage the voxels (cp.async, double buffered), fold them,
// run the recurrences (R for all l >= the group's smallest m, Q for the group's m) and write the T and a operand panels
// in the K-major core-matrix layout the shared-memory descriptors describe; one thread issues the MMAs, whose execution
// overlaps the next chunk's folds and recurrences.
#pragma once
#include "z3d_kernels.cuh"

namespace z3d {

constexpr int BNT = 1024;           // threads per CTA (32 warps of producers, 64 registers each; accumulators in TMEM)
constexpr int BNWARP = BNT / 32;
constexpr int BNV = 32;             // volumes per batch tile
constexpr int BCOL = 2 * BNV;       // GEMM columns: [re | im][volume]
constexpr int BTM = 128;            // rows of an MMA tile (UMMA M)
constexpr int BTILES = 8;           // tiles per row group: 8 x 64 columns = all of tensor memory
constexpr int BROWS_MAX = BTILES * BTM;
constexpr int BMPZ_MAX = BTILES;    // (m, pz) combinations per row group (a tile belongs to exactly one)
constexpr int BQ_MAX = 192;         // Q values per voxel per row group
constexpr int BQS = BQ_MAX + 4;     // voxel stride of the Q panel: = 4 (mod 32) so the 8 voxel lanes of a chain hit 8 banks
constexpr int BR_MAX = 1024;        // R values per voxel (all (n, l): 676 at order 50, 961 at order 60); last slot stays 0
constexpr int BRS = BR_MAX + 4;     // voxel stride of the R panel (bank spread, as BQS)
constexpr int BGD_MAX = 512;        // float4 Legendre coefficient entries per row group
constexpr int BMAXTASK = 4;         // generation tasks per warp
constexpr int BRAW = 2 * 4 * BNV * 16;  // floats of one raw staging buffer: [set][row (sy,sx)][volume][direct 8 | mirrored 8]
constexpr int BSLOT_WORDS = BTILES * BCOL * BTM;  // partial sums of one CTA: [tile][column][row]
constexpr int BFLUSH_CHUNKS = 32;   // accumulator read-back interval (see above)

constexpr int BOFF_THI = 0;                                      // T panels, TF32 high part   [2 k-halves][BROWS_MAX][4]
constexpr int BOFF_TLO = BOFF_THI + 2 * BROWS_MAX * 16;          //           float32 remainder
constexpr int BOFF_AHI = BOFF_TLO + 2 * BROWS_MAX * 16;          // a panels  [BMPZ_MAX][2 k-halves][BCOL][4]
constexpr int BOFF_ALO = BOFF_AHI + BMPZ_MAX * 2 * BCOL * 16;
constexpr int BOFF_R = BOFF_ALO + BMPZ_MAX * 2 * BCOL * 16;      // R panel    [KC][BRS]
constexpr int BOFF_Q = BOFF_R + KC * BRS * 4;                    // Q panel    [KC][BQS]
constexpr int BOFF_F = BOFF_Q + KC * BQS * 4;                    // folds      [2 sets][8 combos][BNV][KC], k halves swizzled
constexpr int BOFF_RAW = BOFF_F + 2 * 8 * BNV * KC * 4;          // raw voxel segments, double buffered (cp.async)
constexpr int BOFF_ROWTAB = BOFF_RAW + 2 * BRAW * 4;             // (R index, Q index) per row   [BROWS_MAX] int2
constexpr int BOFF_GD = BOFF_ROWTAB + BROWS_MAX * 8;             // Legendre coefficients of the group's m
constexpr int BOFF_RK = BOFF_GD + BGD_MAX * 16;                  // radial coefficients (all l)
constexpr int BOFF_TABS = BOFF_RK + BR_MAX * 16;                 // rbase, goff, qoff [3][LMAX+1]; mpz, tile_mpz [2][BTILES]
constexpr int BOFF_TASKS = BOFF_TABS + 3 * (LMAX + 1) * 4 + 2 * BTILES * 4;  // generation tasks [BNWARP][BMAXTASK]
constexpr int BOFF_TRIG = BOFF_TASKS + BNWARP * BMAXTASK * 4;
constexpr int BOFF_MBAR = BOFF_TRIG + 4 * (LMAX + 1) * 4;        // mbarrier (8 B) + tensor-memory base address (4 B)
constexpr int BOFF_GEO = BOFF_MBAR + 16;                         // per-voxel geometry of one row group, 32 B each
static_assert(BOFF_TLO % 128 == 0 && BOFF_AHI % 128 == 0 && BOFF_ALO % 128 == 0 && BOFF_R % 16 == 0 && BOFF_Q % 16 == 0 &&
              BOFF_F % 16 == 0 && BOFF_RAW % 16 == 0 && BOFF_ROWTAB % 16 == 0 && BOFF_GD % 16 == 0 && BOFF_RK % 16 == 0 &&
              BOFF_MBAR % 8 == 0 && BOFF_GEO % 16 == 0, "smem alignment");
__host__ __device__ inline int batched_smem_bytes(int H) { return BOFF_GEO + (H + KC) * 32; }

#ifdef Z3D_DEBUG_TIMING
// per-warp phase totals are kept in shared memory (global atomics per phase would serialise the whole GPU)
#define DBGB_ADD(slot, t0, t1) do { if ((threadIdx.x & 31) == 0) s_dbg[threadIdx.x >> 5][slot] += (t1) - (t0); } while (0)
#else
#define DBGB_ADD(slot, t0, t1)
#endif

struct DevPlanB {
    int order, dim, H;
    int ngroups, nsplit;
    const double *side;
    const int2 *rowtab;      // [ngroups][BROWS_MAX]   (index into the R panel, index into the Q panel); padding -> zero slot
    const int *tile_mpz;     // [ngroups][BTILES]      a-panel index of each tile
    const int *mpz;          // [ngroups][BMPZ_MAX]    m | pz << 16 of each a panel, -1 padded
    const int *nmpz;         // [ngroups]
    const int *ntile;        // [ngroups]
    const int *mlo, *mhi;    // [ngroups]              range of m whose rows the group owns
    const float4 *gd;        // [ngroups][BGD_MAX]     (gamma_E, -delta_E, gamma_O, -delta_O) per (m, step)
    const int *goff;         // [ngroups][L+1]         offset of m's steps in the group's gd table
    const int *qoff;         // [ngroups][L+1]         offset of Q_{m,m} in the group's Q panel (Q_lm at qoff + l - m)
    const float4 *rk;        // [BR_MAX]               (K1, K2, K3, 0), l-major
    const int *rbase;        // [L+1]                  offset of (l, n = l) in rk == index in the R panel
    const int *gen_tasks;    // [ngroups][BNWARP][BMAXTASK]   (type << 16 | base): 1 = Legendre, 2 = radial
};

struct CtxB {
    float *Thi, *Tlo, *Ahi, *Alo, *Rp, *Qp, *Fp, *raw, *cm, *sm, *cm2, *sm2;
    int2 *rowtab;
    int *tasks;
    float4 *gd, *rk;
    int *rbase, *goff, *qoff, *mpz, *tile_mpz;
    unsigned long long *mbar;
    unsigned *tmem;
    Geo *geo;
};
__device__ __forceinline__ CtxB make_ctx_b(unsigned char *smem) {
    CtxB c;
    c.Thi = reinterpret_cast<float *>(smem + BOFF_THI);
    c.Tlo = reinterpret_cast<float *>(smem + BOFF_TLO);
    c.Ahi = reinterpret_cast<float *>(smem + BOFF_AHI);
    c.Alo = reinterpret_cast<float *>(smem + BOFF_ALO);
    c.Rp = reinterpret_cast<float *>(smem + BOFF_R);
    c.Qp = reinterpret_cast<float *>(smem + BOFF_Q);
    c.Fp = reinterpret_cast<float *>(smem + BOFF_F);
    c.raw = reinterpret_cast<float *>(smem + BOFF_RAW);
    c.tasks = reinterpret_cast<int *>(smem + BOFF_TASKS);
    c.rowtab = reinterpret_cast<int2 *>(smem + BOFF_ROWTAB);
    c.gd = reinterpret_cast<float4 *>(smem + BOFF_GD);
    c.rk = reinterpret_cast<float4 *>(smem + BOFF_RK);
    int *tabs = reinterpret_cast<int *>(smem + BOFF_TABS);
    c.rbase = tabs;
    c.goff = tabs + (LMAX + 1);
    c.qoff = tabs + 2 * (LMAX + 1);
    c.mpz = tabs + 3 * (LMAX + 1);
    c.tile_mpz = c.mpz + BTILES;
    float *tr = reinterpret_cast<float *>(smem + BOFF_TRIG);
    c.cm = tr;
    c.sm = tr + (LMAX + 1);
    c.cm2 = tr + 2 * (LMAX + 1);
    c.sm2 = tr + 3 * (LMAX + 1);
    c.mbar = reinterpret_cast<unsigned long long *>(smem + BOFF_MBAR);
    c.tmem = reinterpret_cast<unsigned *>(smem + BOFF_MBAR + 8);
    c.geo = reinterpret_cast<Geo *>(smem + BOFF_GEO);
    return c;
}

// ---- tcgen05 / mbarrier wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor, K-major, no swizzle: 8 x 16-byte core matrices; LBO = byte distance between the two
// 16-byte K halves of a row, SBO = byte distance between groups of 8 rows.  Element (row, k) of an operand lives at
//   (k / 4) * LBO + (row / 8) * SBO + (row % 8) * 16 + (k % 4) * 4.
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((saddr >> 4) & 0x3fff) | ((unsigned long long)((lbo >> 4) & 0x3fff) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);  // bits 46-47: descriptor version 1
}
// instruction descriptor: D float32, A and B TF32, both K-major, N = 64, M = 128
constexpr unsigned UMMA_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(BCOL >> 3) << 17) | ((unsigned)(BTM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(unsigned d_tmem, unsigned long long a, unsigned long long b, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(UMMA_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned mbar) {  // the mbarrier completes a phase when all MMAs issued so far retire
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    for (int it = 0; it < (1 << 26); ++it) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (ok) return;
    }
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

// cos / sin(m phi) of a row group and of its transpose for m <= L, float64 complex powering (see row_trig)
__device__ __forceinline__ void row_trig_b(int L, const CtxB &c, double x0, double y0, double rho) {
    const int m = threadIdx.x;
    if (m <= L) {
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
        c.cm[m] = (float)cr;
        c.sm[m] = (float)ci;
        // transpose: phi' = atan2(x0, y0) = pi/2 - phi (mod 2 pi)
        const int q = m & 3;
        c.cm2[m] = (float)((q == 0) ? cr : (q == 1) ? ci : (q == 2) ? -cr : -ci);
        c.sm2[m] = (float)((q == 0) ? -ci : (q == 1) ? cr : (q == 2) ? ci : -cr);
    }
}

// 16-byte / 4-byte asynchronous global -> shared copies (LDGSTS); the staging of chunk t+1 overlaps the T / a products
// and the contraction of chunk t
__device__ __forceinline__ void cp_async16(float *dst, const float *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float *dst, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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

// issue the copies of chunk `p` for all volumes of the tile: per (set, row, volume) the 8 voxels k0..k0+7 of the row and
// their 8 z-mirrors, 32 contiguous bytes each, one 16-byte piece per thread (consecutive threads fill consecutive shared
// memory).  raw[((set*4 + row)*BNV + b)*16 + (half ^ swz(b))*8 + j], swz spreads the fold's reads over the banks.
__device__ __forceinline__ void stage_chunk(const float *__restrict__ vol, float *raw, const ChunkPos &p, int N, size_t N3,
                                            int nvol, bool vec4) {
    const int tid = threadIdx.x;
    static_assert(BNT == 2 * 4 * BNV * 4, "one 16-byte piece per thread");
    const int q = tid & 1, hs = (tid >> 1) & 1, b = (tid >> 2) % BNV, row = (tid >> 7) & 3, set = tid >> 9;
    const int half = hs ^ ((b >> 1) & 1);
    if (b < nvol && (set == 0 || p.paired)) {
        const int ri = set ? p.jp : p.ip, rj = set ? p.ip : p.jp;
        const int ii = (row & 2) ? N - 1 - ri : ri, jj = (row & 1) ? N - 1 - rj : rj;
        const int k0 = p.kc * KC;
        const float *src = vol + (size_t)b * N3 + ((size_t)ii * N + jj) * N + (half ? N - KC - k0 : k0) + 4 * q;
        float *dst = raw + tid * 4;
        if (vec4) {
            cp_async16(dst, src);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) cp_async4(dst + j, src + j);
        }
    }
    cp_async_commit();
}

// per row group: geometry of its folded voxels, float64 like the reference (zm:1886-1890, 2240-2242)
__device__ __forceinline__ void item_geometry(const DevPlanB &P, const CtxB &c, const ChunkPos &p) {
    const double y0 = P.side[p.ip], x0 = P.side[p.jp];
    const double rho2 = x0 * x0 + y0 * y0, rho = sqrt(rho2);
    for (int k = threadIdx.x; k < P.H + KC; k += BNT) {
        const double z = (k < P.H) ? P.side[k] : 0.0;
        const double r2d = rho2 + z * z;
        const double r = sqrt(r2d);
        double ct = 1.0, st = 0.0;  // r == 0: theta is NaN in the reference and nan_to_num zeroes Y_lm, l > 0 (zm:2298)
        if (r > 0.0) {
            ct = z / r;
            st = rho / r;
        }
        float4 *g = reinterpret_cast<float4 *>(c.geo + k);
        g[0] = make_float4((float)(2.0 * ct), (float)(4.0 * ct * ct), (float)st, (float)(st * st));
        g[1] = make_float4((float)r, (float)r2d, 0.f, 0.f);
    }
}
__device__ __forceinline__ void item_trig(const DevPlanB &P, const CtxB &c, const ChunkPos &p) {
    const double y0 = P.side[p.ip], x0 = P.side[p.jp];
    row_trig_b(P.order, c, x0, y0, sqrt(x0 * x0 + y0 * y0));
}

// (1a) fold the 16 mirror / transpose images of the chunk's voxels for every volume: raw -> F[set][combo][volume][k]
__device__ __forceinline__ void fold_chunk(const DevPlanB &P, const CtxB &c, const ChunkPos &p, const float *raw, int nvol) {
    const int N = P.dim, H = P.H;
    const int tid = threadIdx.x;
    const int k0 = p.kc * KC;
    if (tid < 2 * BNV * KC) {
        const int k = tid % KC, b = (tid / KC) % BNV, set = tid / (KC * BNV);
        const int kk = k0 + k, k2 = N - 1 - kk;
        const int ri = set ? p.jp : p.ip, rj = set ? p.ip : p.jp;
        const bool live = b < nvol && kk < H && (set == 0 || p.paired);
        const bool dy = (N - 1 - ri == ri), dx = (N - 1 - rj == rj), dz = (k2 == kk);
        const float *rw = raw + (set * 4 * BNV + b) * 16;
        const int sw = ((b >> 1) & 1) * 8;
        float v[8];
#pragma unroll
        for (int img = 0; img < 8; ++img) {
            const bool sy = img & 4, sx = img & 2, sz = img & 1;
            const bool dup = (sy && dy) || (sx && dx) || (sz && dz);
            const float x = rw[(img >> 1) * (BNV * 16) + ((sz ? 8 : 0) ^ sw) + (sz ? 7 - k : k)];
            v[img] = (live && !dup) ? x : 0.0f;
        }
        wht8(v);
#pragma unroll
        for (int cc8 = 0; cc8 < 8; ++cc8) c.Fp[((set * 8 + cc8) * BNV + b) * KC + (k ^ (((b >> 2) & 1) << 2))] = v[cc8];
    }
}

// (1b) recurrences of the chunk: R for all l >= mlo, Q (both parities) for the row group's m, one task per warp
__device__ __forceinline__ void recur_chunk(const DevPlanB &P, const CtxB &c, const ChunkPos &p, int mhi) {
    const int L = P.order;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k0 = p.kc * KC;
    const int gk = lane % KC, gsub = lane / KC, kk = k0 + gk;
    const float4 g0 = reinterpret_cast<const float4 *>(c.geo + kk)[0], g1 = reinterpret_cast<const float4 *>(c.geo + kk)[1];
    const float c2 = g0.x, c2sq = g0.y, st = g0.z, st2 = g0.w, r = g1.x, r2 = g1.y;
#pragma unroll 1
    for (int t = 0; t < BMAXTASK; ++t) {
        const int task = c.tasks[warp * BMAXTASK + t];
        const int type = task >> 16, base = task & 0xffff;
        if (type == 0) break;
        if (type == 1) {  // Legendre: lane m = base + gsub, chains E (l = m, m+2, ..) and O (l = m+1, m+3, ..)
            const int m = base + gsub;
            const bool lanev = m <= mhi;
            const int mm = lanev ? m : mhi;
            const int nE = lanev ? (L - mm) / 2 + 1 : 0;  // chain O has nE or nE - 1 levels; its last store may land on the
            const int nmax = (L - base) / 2 + 1;          // pair's padding slot (Q panel offsets are even-sized per m)
            float qE1 = powi2(st, st2, mm), qE2 = 0.f, qO1 = qE1 * c2, qO2 = 0.f;
            const float4 *t0 = c.gd + c.goff[mm];
            float2 *d = reinterpret_cast<float2 *>(c.Qp + gk * BQS + c.qoff[mm]);
            float4 g = t0[0];
#pragma unroll 2
            for (int i = 0; i < nmax; ++i) {
                const float4 gn = t0[i + 1];
                if (i < nE) d[i] = make_float2(qE1, qO1);
                const float nE1 = fmaf(c2sq - g.x, qE1, g.y * qE2), nO1 = fmaf(c2sq - g.z, qO1, g.w * qO2);
                qE2 = qE1; qE1 = nE1; qO2 = qO1; qO1 = nO1;
                g = gn;
            }
        } else {  // radial: lane l = base + gsub;  R_n = (K1 r^2 + K2) R_{n-2} + K3 R_{n-4}  (zm:1991-2000)
            const int l = base + gsub;
            const bool active = l <= L;
            const int ll = active ? l : L;
            const int nn = active ? (L - ll) / 2 + 1 : 0;
            const int nmax = (L - base) / 2 + 1;
            float R1 = 0.0f, R2 = powi2(r, r2, ll);
            float *dst = c.Rp + gk * BRS + c.rbase[ll];
            const float4 *kp = c.rk + c.rbase[ll];
            float4 Kc = kp[0];
#pragma unroll 4
            for (int ni = 0; ni < nmax; ++ni) {
                const float4 Kn = kp[ni + 1];
                const float R = fmaf(fmaf(Kc.x, r2, Kc.y), R1, Kc.z * R2);
                if (ni < nn) dst[ni] = R;
                R2 = R1;
                R1 = R;
                Kc = Kn;
            }
        }
    }
}

// (2a) T rows = R * Q;  (2b) a panels: a = F_A trig_A + F_B trig_B per (m, pz), re / im, volume and 4 voxels at a time
//      (re: even in y, (-1)^m in x; im: odd in y, -(-1)^m in x, negated; z parity pz).  Both are split into TF32 high
//      parts and remainders and written in the K-major core-matrix layout of umma_desc: [k half][row | column][4].
__device__ __forceinline__ void products(const CtxB &c, int nmpz) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int h = 0; h < BROWS_MAX / BNT; ++h) {
        const int row = tid + h * BNT;
        const int2 rq = c.rowtab[row];
        float hi[KC], lo[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) split_tf32(c.Rp[k * BRS + rq.x] * c.Qp[k * BQS + rq.y], hi[k], lo[k]);
        *reinterpret_cast<float4 *>(c.Thi + row * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4 *>(c.Thi + (BROWS_MAX + row) * 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<float4 *>(c.Tlo + row * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4 *>(c.Tlo + (BROWS_MAX + row) * 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
    }
    for (int it = tid; it < nmpz * (2 * BNV); it += BNT) {  // (panel, re / im, volume): all 8 voxels
        const int b = it & (BNV - 1), reim = (it >> 5) & 1, j = it >> 6;
        const int code = c.mpz[j], m = code & 0xffff, pz = code >> 16;
        const int comb = (reim << 2) | (((m & 1) ^ reim) << 1) | pz;
        float tA = c.cm[reim * (LMAX + 1) + m], tB = c.cm2[reim * (LMAX + 1) + m];  // cm | sm and cm2 | sm2 are adjacent
        if (reim) {
            tA = -tA;
            tB = -tB;
        }
        const int sw = ((b >> 2) & 1) << 2;  // fold_chunk's k-half swizzle: 8 consecutive volumes cover all banks
        const float *fa = c.Fp + (comb * BNV + b) * KC, *fb = fa + 8 * BNV * KC;
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
__device__ __forceinline__ void issue_mma(const CtxB &c, unsigned tmem, int ntile, bool accumulate) {
    const unsigned long long dThi = umma_desc(smem_u32(c.Thi), BROWS_MAX * 16, 128), dTlo = umma_desc(smem_u32(c.Tlo), BROWS_MAX * 16, 128);
    const unsigned long long dAhi = umma_desc(smem_u32(c.Ahi), BCOL * 16, 128), dAlo = umma_desc(smem_u32(c.Alo), BCOL * 16, 128);
    for (int t = 0; t < ntile; ++t) {
        const unsigned long long ta = (unsigned long long)(t * BTM * 16 >> 4), aa = (unsigned long long)(c.tile_mpz[t] * 2 * BCOL * 16 >> 4);
        const unsigned d = tmem + t * BCOL;
        umma_tf32(d, dTlo + ta, dAhi + aa, accumulate ? 1u : 0u);  // small terms first
        umma_tf32(d, dThi + ta, dAlo + aa, 1u);
        umma_tf32(d, dThi + ta, dAhi + aa, 1u);
    }
    umma_commit(smem_u32(c.mbar));
}

// accumulators -> this CTA's partial buffer [tile][column][row] (float32, round to nearest); warp w reads the lane
// quarter w % 4 (the only one it can address) of tile w / 4
__device__ __forceinline__ void flush_accumulators(unsigned tmem, int ntile, float *out, bool add) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3;
#pragma unroll 1
    for (int t = warp >> 2; t < ntile; t += BNWARP / 4) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            unsigned v[32];
            tmem_ld32(tmem + ((unsigned)(32 * q) << 16) + t * BCOL + 32 * half, v);
            float *p = out + ((size_t)t * BCOL + 32 * half) * BTM + 32 * q + lane;
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

__device__ __forceinline__ bool advance_pos(ChunkPos &p, long long cc, int CPR, int H) {  // true when a new row group starts
    if (++p.kc < CPR) return false;
    chunk_pos(cc, CPR, H, p);
    return true;
}

// partial layout: [ngroups][nsplit][BSLOT_WORDS]
//
// Software pipeline over the chunk list, two barriers per chunk:
//   X(t): wait for the MMAs of chunk t-1 (panels free) | T and a panels of chunk t from R, Q, F, trig | geometry of t+1
//   ----- barrier; one thread issues the MMAs of chunk t (asynchronous)
//   Y(t): copies of chunk t+2 issued | fold + recurrences of chunk t+1 (R, Q, F) | trig of chunk t+1's row group
//   ----- barrier
__global__ void __launch_bounds__(BNT, 1)
k_decompose_batched(const DevPlanB P, const float *__restrict__ vol, int nvol, int part, int nparts,
                    float *__restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int L = P.order, N = P.dim, H = P.H;
    const int tid = threadIdx.x;
    const int grp = blockIdx.x % P.ngroups, split = blockIdx.x / P.ngroups;
    const CtxB c = make_ctx_b(smem);
#ifdef Z3D_DEBUG_TIMING
    __shared__ long long s_dbg[BNWARP][8];
    if ((tid & 31) == 0)
        for (int i = 0; i < 8; ++i) s_dbg[tid >> 5][i] = 0;
#endif
    // ---- tables of this row group -> smem; clear the panels (padding rows / unused a panels must stay finite)
    for (int i = tid; i <= L; i += BNT) {
        c.rbase[i] = P.rbase[i];
        c.goff[i] = P.goff[grp * (L + 1) + i];
        c.qoff[i] = P.qoff[grp * (L + 1) + i];
    }
    for (int i = tid; i < BTILES; i += BNT) {
        c.mpz[i] = P.mpz[grp * BMPZ_MAX + i];
        c.tile_mpz[i] = P.tile_mpz[grp * BTILES + i];
    }
    for (int i = tid; i < BNWARP * BMAXTASK; i += BNT) c.tasks[i] = P.gen_tasks[grp * BNWARP * BMAXTASK + i];
    for (int i = tid; i < BROWS_MAX; i += BNT) c.rowtab[i] = P.rowtab[grp * BROWS_MAX + i];
    for (int i = tid; i < BGD_MAX; i += BNT) c.gd[i] = P.gd[grp * BGD_MAX + i];
    for (int i = tid; i < BR_MAX; i += BNT) c.rk[i] = P.rk[i];
    for (int i = tid; i < (BOFF_ROWTAB - BOFF_THI) / 4; i += BNT) c.Thi[i] = 0.0f;  // all panels and staging buffers
    const int nmpz = P.nmpz[grp], ntile = P.ntile[grp], mhi = P.mhi[grp];
    const unsigned mbar = smem_u32(c.mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {  // warp 0 owns the tensor-memory allocation: all 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(c.tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();  // tables, cleared panels, mbarrier and the tensor-memory address are visible
    tc_fence_after();
    const unsigned tmem = *c.tmem;

    const int CPR = (H + KC - 1) / KC;
    const long long total = (long long)analysis_items(0, H, H) * CPR;
    const long long gbeg = total * part / nparts, glen = total * (part + 1) / nparts - gbeg;
    const long long cbeg = gbeg + glen * split / P.nsplit, cend = gbeg + glen * (split + 1) / P.nsplit;
    const size_t N3 = (size_t)N * N * N;
    const bool vec4 = (N & 3) == 0;
    float *out = partial + ((size_t)(grp * P.nsplit + split)) * BSLOT_WORDS;
    int since_flush = 0;  // chunks accumulated in tensor memory
    unsigned issued = 0;  // commits so far (mbarrier phase of commit i has parity i & 1)
    bool flushed = false;

    ChunkPos cur, nxt;
    bool have_next = false, next_new = false;
    int buf = 1;  // staging buffer that holds (or receives) chunk cc + 1
    if (cbeg < cend) {  // prologue: chunk cbeg's operands, chunk cbeg + 1's copies
        chunk_pos(cbeg, CPR, H, cur);
        stage_chunk(vol, c.raw, cur, N, N3, nvol, vec4);
        item_geometry(P, c, cur);
        item_trig(P, c, cur);
        cp_async_wait_all();
        __syncthreads();
        nxt = cur;
        have_next = cbeg + 1 < cend;
        if (have_next) {
            next_new = advance_pos(nxt, cbeg + 1, CPR, H);
            stage_chunk(vol, c.raw + BRAW, nxt, N, N3, nvol, vec4);
        }
        fold_chunk(P, c, cur, c.raw, nvol);
        recur_chunk(P, c, cur, mhi);
    }
    for (long long cc = cbeg; cc < cend; ++cc) {
        DBG_T(t0);
        __syncthreads();  // R, Q, F of chunk cc are complete
        DBG_T(t1);
        DBGB_ADD(0, t0, t1);
        // ---- X(cc)
        if (have_next && next_new) item_geometry(P, c, nxt);  // geometry is only read by the recurrences (phase Y)
        if (issued) {  // the previous chunk's MMAs have retired: the operand panels are free, the accumulators current
            mbar_wait(mbar, (issued - 1) & 1);
            tc_fence_after();
        }
        DBG_T(t2);
        DBGB_ADD(1, t1, t2);
#ifdef Z3D_ABL_NOFLUSH
        if (false) {
#else
        if (since_flush == BFLUSH_CHUNKS) {
#endif
            flush_accumulators(tmem, ntile, out, flushed);
            flushed = true;
            since_flush = 0;
            tc_fence_before();
        }
        DBG_T(t3);
        DBGB_ADD(2, t2, t3);
#ifndef Z3D_ABL_NOPROD
        products(c, nmpz);
#endif
        fence_async_smem();   // generic-proxy panel writes -> visible to the tensor core's async-proxy reads
        cp_async_wait_all();  // this thread's copies of chunk cc + 1
        DBG_T(t4);
        DBGB_ADD(3, t3, t4);
        __syncthreads();  // panels complete; R, Q, F free; raw[buf] visible
        DBG_T(t5);
        DBGB_ADD(4, t4, t5);
        if (tid == 0) {
            tc_fence_after();
#ifdef Z3D_ABL_NOMMA
            umma_commit(mbar);
#else
            issue_mma(c, tmem, ntile, since_flush > 0);
#endif
        }
        ++since_flush;
        ++issued;
        // ---- Y(cc)
        ChunkPos nn2 = nxt;
        bool have2 = false, new2 = false;
        if (have_next) {
            have2 = cc + 2 < cend;
            if (have2) {
                new2 = advance_pos(nn2, cc + 2, CPR, H);
#ifndef Z3D_ABL_NOSTAGE
                stage_chunk(vol, c.raw + (buf ^ 1) * BRAW, nn2, N, N3, nvol, vec4);
#endif
            }
        }
        DBG_T(t6);
        DBGB_ADD(5, t5, t6);
#ifndef Z3D_ABL_NOFOLD
        if (have_next) fold_chunk(P, c, nxt, c.raw + buf * BRAW, nvol);
#endif
        DBG_T(t7);
        DBGB_ADD(6, t6, t7);
        if (have_next) {
#ifndef Z3D_ABL_NORECUR
            recur_chunk(P, c, nxt, mhi);
#endif
            if (next_new) item_trig(P, c, nxt);  // trig is only read by the a panels (phase X)
        }
        DBG_T(t8);
        DBGB_ADD(7, t7, t8);
        cur = nxt;
        nxt = nn2;
        have_next = have2;
        next_new = new2;
        buf ^= 1;
    }
    if (issued) {
        mbar_wait(mbar, (issued - 1) & 1);
        tc_fence_after();
        flush_accumulators(tmem, ntile, out, flushed);
    } else {
        for (int i = tid; i < BSLOT_WORDS; i += BNT) out[i] = 0.0f;  // the reduction reads every split
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
#ifdef Z3D_DEBUG_TIMING
    if ((tid & 31) == 0)
        for (int i = 0; i < 8; ++i) atomicAdd((unsigned long long *)&g_dbg[i], (unsigned long long)s_dbg[tid >> 5][i]);
#endif
}

// moments[b][mu][reim] = scale[mu] * sum over splits of the owning accumulator (fixed order, float64)
__global__ void k_reduce_batched(const float *__restrict__ partial, const int2 *__restrict__ rowmap /* [nmom] (group, row) */,
                                 const double *__restrict__ scale, int nmom, int nsplit, int nvol,
                                 float *__restrict__ moments) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over nmom * BCOL
    if (idx >= nmom * BCOL) return;
    const int mu = idx / BCOL, col = idx % BCOL;
    const int reim = col / BNV, b = col % BNV;
    if (b >= nvol) return;
    const int2 gr = rowmap[mu];
    const float *p = partial + (size_t)gr.x * nsplit * BSLOT_WORDS + ((size_t)(gr.y / BTM) * BCOL + col) * BTM + gr.y % BTM;
    double sum = 0.0;
    for (int s = 0; s < nsplit; ++s) sum += (double)p[(size_t)s * BSLOT_WORDS];
    moments[((size_t)b * nmom + mu) * 2 + reim] = (float)(sum * scale[mu]);
}

}  // namespace z3d
