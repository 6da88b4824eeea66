// zernike3d_b200 - batched analysis, TABLE-FREE tensor-core variant (the default for batches): nothing volume-independent is
// read from memory - the operand T = R_nl * Q_lm is generated inside the kernel by the reference's recurrences.
//
// Same contraction as z3d_stream.cuh, per (m, pz = (l + m) & 1) a real GEMM over the folded voxel axis,
//     D[(re/im, b)][(n,l,m)] += a[(m,pz)][(re/im, b)][k] * T[(n,l,m)][k],   a = F_A,b trig_A + F_B,b trig_B,
// on tcgen05.mma kind::tf32 (M = 128 = re/im x 64 volumes, N = moment rows of one (m, pz) padded to 16, K = 8 folded voxels =
// one chunk), 3xTF32 (hi.lo + lo.hi + hi.hi), accumulators in tensor memory.  What is new:
//   * T is produced by the CTA itself, in registers, straight into the K-major core-matrix layout the shared-memory
//     descriptors describe.  A chain (l, m) of T over n obeys the radial recurrence of zm:1991-2000 because that recurrence is
//     linear in R:  T_n = (K1 r^2 + K2) T_{n-2} + K3 T_{n-4},  seeded with  S_lm = r^l Q_lm.  The seeds obey the Legendre
//     recurrence of zm:2021-2026 (monic, parity-decoupled, times r^l):  S_{l+2} = (4 z^2 - gamma_l r^2) S_l - delta_l r^4 S_{l-2},
//     S_mm = rho^m, S_{m+1,m} = 2 z rho^m - polynomial in the voxel coordinates, no division and no square root per voxel.
//     No R / Q / T table exists anywhere (the streamed-T kernel reads 13.7 GB per tile at 128^3 / order 50).
//   * the a panels (A operand) live in TENSOR memory: the a-panel teams write them with tcgen05.st and the MMAs take A from
//     tensor memory, so neither the producers' stores nor the tensor core's A fetch touch shared memory, which is left to the
//     T ring (64 B written + 96 B fetched per moment row and chunk).
//   * the folds come straight from global memory into registers (prefetched one round ahead): no fold staging.
//   * MMAs are issued by ONE converged warp through elect.sync (back-to-back UTCHMMA on uniform registers: 23 clk or the tensor
//     pipe's N / 2 per instruction; an `if (lane == 0)` region costs 64-72 clk per MMA - scripts/experiments/tc_probe4.cu).
//
// Roles (28 warps = 896 threads, no block barrier in the loop):
//   warp 0        issues the MMAs of chunk t when its T block (tfull) and a panels (afull) are ready; ONE commit per chunk -> tfree (T slot and a-panel slot)
//   warps 1-3     seed warps (chunks t = w-1 mod 3): per chunk the float64 geometry (z, rho^2), the row trig of the set's m
//                 (float64 complex powering like exp(1j m phi), zm:2297) and the seeds S_lm of every (panel, l, voxel) -> G ring
//   warps 4-11    two a-panel teams (chunk t -> team t & 1 -> tensor-memory slot t & 1); a thread owns one row (re/im, volume)
//   warps 12-27   four T teams (chunk t -> team t & 3); a thread owns up to two (chain, k half) tasks, 4 voxels each
// All 24 producer warps (4-27) read the accumulators back every BFLUSH_CHUNKS chunks (see z3d_batched.cuh for why).
// (With G_NISS_ > 1 or G_NALT_ = 2 - compile-time variants that were measured and lost - there are more issuing warps, the producers
// start at warp 8 and the CTA has 32 warps.)
#pragma once
#include "z3d_stream.cuh"

namespace z3d {

constexpr int G_PANELS = 8;        // a panels ((m, pz) pieces) per column set at most
constexpr int G_SEGS = 8;          // MMA segments (<= 256 columns each) per column set at most
constexpr int G_TASKS = 256;       // T chain tasks (chain, k half) per column set at most: two passes of a 128-thread team
#ifndef G_NSEEDW_
#define G_NSEEDW_ 3
#endif
#ifndef G_DG_
#define G_DG_ 12
#endif
constexpr int G_NSEEDW = G_NSEEDW_;        // seed warps
constexpr int G_NAT = 2;           // a-panel teams
#ifndef G_ASLOTS_
#define G_ASLOTS_ 2
#endif
constexpr int G_ASLOTS = G_ASLOTS_;  // a-panel slots (chunks in flight) in tensor memory with one tile per T block (two tiles: 2), a multiple of G_NAT
__host__ __device__ constexpr int g_aslots(int ntile) { return ntile == 1 ? G_ASLOTS : 2; }
static_assert(G_ASLOTS % G_NAT == 0 && G_ASLOTS <= 4, "a-panel slots");
constexpr int G_NTT = 4;           // T teams
constexpr int G_DG = G_DG_;            // depth of the geometry / seed ring
constexpr int G_NS_MAX = 8;        // depth of the T ring at most
constexpr int G_FD = 4;            // depth of every a-panel warp's private fold ring (steps of 2 KB: 32 rows x (set A | set B) x 8 voxels)
constexpr int G_FOLD_BYTES = 4 * G_NAT * G_FD * 2048 + 4 * G_NAT * G_FD * 8;   // the rings of the 8 a-panel warps + one mbarrier per slot
// A parity wait is only meaningful when the waiter knows the barrier is at most one phase behind: the previous user of a ring
// slot must be the SAME agent (warp / team), whose own progress then proves the earlier phase complete.  Chunk t is handled by
// seed warp t % G_NSEEDW, a-panel team t % G_NAT and T team t % G_NTT and uses geometry slot t % G_DG, so G_DG must be a common
// multiple of the three (with G_DG = 6 a T team could pass gfull one phase early whenever the seed warps were the bottleneck:
// sporadic wrong moments, lost arrivals).  The T ring is safe for any depth >= G_NTT because the tensor core retires chunks in order.
static_assert(G_DG % G_NSEEDW == 0 && G_DG % G_NAT == 0 && G_DG % G_NTT == 0, "geometry ring depth vs. agent strides");
#ifndef G_NISS_
#define G_NISS_ 1
#endif
constexpr int G_NISS = G_NISS_;           // issuing warps by SEGMENT, one per scheduler (SM sub-partition = warp % 4): warp w issues the segments g = w mod G_NISS
#ifndef G_NALT_
#define G_NALT_ 1
#endif
constexpr int G_NALT = G_NALT_;           // issuing warps by CHUNK: warp a issues the chunks t = a mod G_NALT (2: measured slower, see the issuing role below)
static_assert(G_NISS == 1 || G_NALT == 1, "split the issue work by segment or by chunk, not both");
static_assert(G_NALT == 1 || G_NALT == 2, "alternating issuers");
constexpr int G_NISSW = G_NISS * G_NALT;  // issuing warps in all
constexpr int G_WARP_SEED0 = G_NISSW;
constexpr int G_WARP_A0 = (G_NISSW + G_NSEEDW + 3) / 4 * 4;
constexpr int G_WARP_T0 = G_WARP_A0 + 4 * G_NAT;
constexpr int G_NWARP = G_WARP_T0 + 4 * G_NTT;     // 28 (32 with more than one issuing warp)
constexpr int GNT = G_NWARP * 32;                  // 896 threads
constexpr int G_SEG_PER_ISS = G_SEGS / G_NISS;
static_assert(G_NWARP - G_WARP_A0 == S_NPW && G_WARP_A0 % 4 == 0 && G_SEGS % G_NISS == 0, "s_flush deals the read-back units over the 24 producer warps");
constexpr int G_OFF_TABS = 0;      // GenTabs
constexpr int G_OFF_BAR = 512;     // mbarriers, tensor-memory base
constexpr int G_OFF_GD = 1024;     // seed recurrence coefficients of the set's panels, then the radial table, the G ring, the T ring

struct GenTabs {                   // per column set
    int ncols, nseg, npanel, kind;
    int ntask, nseedrows, acol0, lpar;      // acol0: first tensor-memory column of the a-panel slots; lpar: parity of the set's l
    int seg_cols[G_SEGS], seg_off[G_SEGS], seg_panel[G_SEGS];
    int panel_m[G_PANELS], panel_pz[G_PANELS], panel_nl[G_PANELS], panel_soff[G_PANELS], panel_gd[G_PANELS];
};
static_assert(sizeof(GenTabs) <= G_OFF_BAR, "tables");

// What the issuing warp needs per column set, passed in the KERNEL PARAMETERS (constant bank): the values arrive in uniform
// registers through ULDC, where UTCHMMA wants them.  Read from shared memory they sat in vector registers and every MMA cost a
// chain of R2UR moves (19 instructions per segment, ~130 clk per segment on a scheduler shared with six producer warps - the
// issuing warp's serial path was the whole kernel's period, whatever the producers and the tensor pipe had to do).
constexpr int G_MAXSETS = 152;
struct GenIssue {
    int4 seg[G_MAXSETS][G_SEGS];   // instruction descriptor, a-panel column offset, T row offset (16 B units) = accumulator column, valid ? 0 : -1
    int4 hdr[G_MAXSETS];           // ncols, nseg, npanel, acol0
};
static_assert(sizeof(GenIssue) < 24 * 1024, "kernel parameter space");

struct DevPlanG {
    int order, dim, H;
    int nsets, nsplits, nslots;    // T ring depth
    int tslot_bytes;               // T block of the widest set: 64 B per column
    int gslot_bytes;               // seeds of the set with the most (panel, l) rows + r^2 + trig
    int gd_bytes, rk_bytes;        // shared-memory copies of the coefficient tables
    int off_rk, off_g, off_t;      // byte offsets in dynamic shared memory
    int off_f;                     // the a-panel warps' fold rings (G_FOLD_BYTES)
    int dbg;                       // developer switches (-DZ3D_GEN_DEBUG builds only)
    const GenTabs *tabs;           // [nsets]
    const int4 *tasks;             // [nsets][G_TASKS]  (first column, length, index in the radial table, seed row | k half << 16)
    const float2 *gd;              // [nsets][gd_bytes / 8]   (gamma_l, -delta_l) per panel and step
    const float4 *rk;              // [2][rk_bytes / 16]      (K1, K2, K3, 0) of the l of one parity, l-major
    const double *side;            // [dim]
};

// Producer-side wait.  A failed mbarrier.try_wait comes back after ~100 clk whatever suspend-time hint it is given, so a waiting
// warp executes a 7-instruction loop ten times per microsecond; with six of the seven warps of a scheduler waiting most of the time
// (ncu: nine of ten executed instructions of the kernel were these loops) the warp that has work - above all the single issuing
// warp - gets a fraction of its scheduler's issue slots.  Producers therefore SLEEP between polls (G_SLEEP_NS; their rings are
// deep enough to absorb the wake-up latency); only the issuing warp spins.
#ifndef G_SLEEP_NS
#define G_SLEEP_NS 128
#endif
__device__ __forceinline__ void g_wait_relaxed(unsigned mbar, unsigned parity) {
#pragma unroll 1
    for (int it = 0; it < (1 << 24); ++it) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (ok) return;
#if G_SLEEP_NS > 0
        __nanosleep(G_SLEEP_NS);
#endif
    }
#if defined(Z3D_SYNTH_DEBUG) || defined(Z3D_WAIT_DIAG)
    if ((threadIdx.x & 31) == 0) printf("relaxed wait timed out: block (%d,%d) warp %d barrier @%u parity %u\n", blockIdx.x, blockIdx.y, threadIdx.x >> 5, mbar, parity);
#endif
    __trap();
}

#ifdef Z3D_GEN_CTACLK
// developer instrumentation, cheap enough not to disturb the timing: per CTA of the last launch, in SM clocks, [0] the CTA's
// duration and what the issuing warp spent waiting for [1] read-backs (flushed), [2] T blocks (tfull), [3] a panels (afull),
// [4] inside the elected MMA / commit block (z3d_debug_gen_cta) - what the planner's cost model of a column set is fitted against
__device__ long long g_cta_clk[5][1024];
#define GC_T(var) const long long var = clock64()
#define GC_ADD(acc, t0, t1) acc += (t1) - (t0)
#else
#define GC_T(var)
#define GC_ADD(acc, t0, t1)
#endif
#ifdef Z3D_GEN_TIMING
// developer instrumentation: clocks spent per role and wait reason, CTA 0 only (z3d_debug_gen)
__device__ long long g_gdbg[32];
#define GT_T(var) const long long var = clock64()
#define GT_ADD(slot, t0, t1) do { if (blockIdx.y * gridDim.x + blockIdx.x == GT_BLOCK && (threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)&g_gdbg[slot], (unsigned long long)((t1) - (t0))); } while (0)
#ifndef GT_BLOCK
#define GT_BLOCK P.dbg   // the CTA to time: Z3D_GEN_DBG=<set * splits + split>
#endif
#else
#define GT_T(var)
#define GT_ADD(slot, t0, t1)
#endif

#if defined(Z3D_GEN_DEBUG) && !defined(Z3D_GEN_ABL)
// developer aid: hazard probes.  GD_SLEEP(bit) delays one role at one point when the bit is set in P.dbg (a race that
// disappears under a delay names its hazard); bounded waits report the barrier that never completed through mapped host memory.
__device__ int *g_gen_trap;
#define GD_SLEEP(bit) do { if (P.dbg & (bit)) __nanosleep(3000); } while (0)
__device__ __forceinline__ void g_wait_dbg(unsigned mbar, unsigned parity, int id, int t) {
#pragma unroll 1
    for (int it = 0; it < (1 << 21); ++it) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (ok) return;
    }
    if ((threadIdx.x & 31) == 0 && g_gen_trap && atomicAdd(g_gen_trap, 1) < 60) {
        int *r = g_gen_trap + 8 + 8 * (atomicAdd(g_gen_trap + 1, 1) % 60);
        r[0] = id; r[1] = blockIdx.y * gridDim.x + blockIdx.x; r[2] = threadIdx.x >> 5; r[3] = (int)parity; r[4] = t;
        __threadfence_system();
    }
    __nanosleep(1000000);
    __trap();
}
#define GD_FLUSHSLEEP ((P.dbg & 512) != 0)
#define GW(bar, par, id, t) g_wait_dbg(bar, par, id, t)
#define GWR(bar, par, id, t) g_wait_dbg(bar, par, id, t)
#else
// (-DZ3D_GEN_DEBUG -DZ3D_GEN_ABL: only the work-skipping switches of P.dbg, with the product's waits - timing ablations)
#define GD_SLEEP(bit)
#define GD_FLUSHSLEEP false
#define GW(bar, par, id, t) mbar_wait(bar, par)
#define GWR(bar, par, id, t) g_wait_relaxed(bar, par)
#endif

// ---- tensor-memory helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_tf32_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long b, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
// x = hi + lo, hi = x rounded to TF32 (nearest, ties away: what cvt.rna.tf32.f32 computes, which ptxas expands into four
// instructions per value - FSETP / VIADD / SEL / LOP3 - for the infinity case; the operands here are finite): add half an ulp of
// the 13 dropped bits to the magnitude, clear them.  (Plain truncation - one instruction less - biases hi: moments 2.3e-6 from the
// float64 oracle instead of 6.5e-7, for 0.7 % of the kernel's time.)
__device__ __forceinline__ void split_fast(float x, float &hi, float &lo) {
#ifdef G_OLDSPLIT
    split_tf32(x, hi, lo);
#else
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = x - hi;
#endif
}
__device__ __forceinline__ void split4_fast(const float4 &x, float4 &hi, float4 &lo) {
    split_fast(x.x, hi.x, lo.x);
    split_fast(x.y, hi.y, lo.y);
    split_fast(x.z, hi.z, lo.z);
    split_fast(x.w, hi.w, lo.w);
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Folds of a tile by kind, as k_fold_tile_kind but without the k-half swizzle (the a-panel threads read them from global memory):
// F[chunk][kind][set A|B][re|im][volume][k]
__global__ void __launch_bounds__(2 * SNV * KC) k_fold_tile_gen(const float *__restrict__ vol, int nvol, int N, int H,
                                                                float *__restrict__ Ft) {
    const int CPR = (H + KC - 1) / KC;
    const long long chunk = blockIdx.x;
    ChunkPos p;
    chunk_pos(chunk, CPR, H, p);
    const int tid = threadIdx.x;
    const int k = tid % KC, b = (tid / KC) % SNV, set = tid / (KC * SNV);
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
    float *dst = Ft + (size_t)chunk * (4 * S_FK_BYTES / 4);
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {  // class c8 = (y parity = re/im) << 2 | (x parity) << 1 | z parity
        const int reim = c8 >> 2, kind = ((((c8 >> 1) & 1) ^ reim) << 1) | (c8 & 1);
        dst[(((kind * 2 + set) * 2 + reim) * SNV + b) * KC + k] = v[c8];
    }
}

__device__ __forceinline__ float4 ldg_stream4(const float *p) {  // read-once data: do not allocate in L1
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// one read-back event of a producer warp (see s_flush)
__device__ __forceinline__ void g_flush_event(unsigned tmem, int ncols, float *out, bool &had, unsigned drained, unsigned flushed, int &nev,
                                              bool last, bool dbg_sleep = false) {
    g_wait_relaxed(drained, nev & 1);   // (s_flush's own wait then succeeds at once)
    s_flush<G_WARP_A0>(tmem, ncols, out, had, drained, nev & 1, dbg_sleep);
    had = true;
    if (!last) {
        tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(flushed);
    }
    ++nev;
}

// NTILE = 2: two tiles of 64 volumes per generated T block.  Tile 1's accumulators follow tile 0's in tensor memory (columns
// [ncols, 2 ncols)), every a-panel slot holds both tiles' panels (slot (chunk parity, tile) at acol0 + (parity * NTILE + tile) *
// 16 npanel), tile 1's folds lie `tile_stride` floats behind tile 0's.  T, the seeds and the geometry - everything that does not
// depend on the volumes - is produced once per chunk and feeds 6 MMAs per segment instead of 3.
template <int NTILE>
__global__ void __launch_bounds__(GNT, 1)
k_decompose_gen(const __grid_constant__ DevPlanG P, const __grid_constant__ GenIssue I, const float *__restrict__ Ft, size_t tile_stride,
                int part, int nparts, float *__restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int set = blockIdx.y, split = blockIdx.x, bid = blockIdx.y * gridDim.x + blockIdx.x;   // grid = (splits, sets)
#ifdef Z3D_GEN_CTACLK
    const long long cta_t0 = clock64();
#endif
    GenTabs &T = *reinterpret_cast<GenTabs *>(smem + G_OFF_TABS);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + G_OFF_BAR);
    float2 *gds = reinterpret_cast<float2 *>(smem + G_OFF_GD);
    float4 *rks = reinterpret_cast<float4 *>(smem + P.off_rk);
    unsigned char *gring = smem + P.off_g, *tring = smem + P.off_t;
    const int NS = P.nslots;
    // barriers: gfull[DG] gfree[DG] tfull[NS_MAX] tfree[NS_MAX] afull[2][NTILE] afree[2][NTILE] drained flushed
    // (the a-panel slots are handed over per TILE: tile 0's MMAs run while the team still builds tile 1's panels, and tile 0's slot
    // is free again as soon as its own MMAs have retired)
    const unsigned gfull0 = smem_u32(bars), gfree0 = gfull0 + 8 * G_DG, tfull0 = gfree0 + 8 * G_DG, tfree0 = tfull0 + 8 * G_NS_MAX,
                   afull0 = tfree0 + 8 * G_NS_MAX, afree0 = afull0 + 8 * G_NAT * 2, drained = afree0 + 8 * G_NAT * 2, flushed = drained + 8;
    const unsigned issued0 = flushed + 8;   // issued[a]: issuing warp a has handed its chunk's MMAs to the tensor core (chunk order between the warps)
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem + G_OFF_BAR + 8 * (2 * G_DG + 2 * G_NS_MAX + 4 * G_NAT + 2 + 2));
    static_assert(8 * (2 * G_DG + 2 * G_NS_MAX + 4 * G_NAT + 2 + 2) + 4 <= G_OFF_GD - G_OFF_BAR, "barrier block");
    for (int i = tid; i < (int)(sizeof(GenTabs) / 4); i += GNT)
        reinterpret_cast<int *>(smem + G_OFF_TABS)[i] = reinterpret_cast<const int *>(P.tabs + set)[i];
    for (int i = tid; i < P.gd_bytes / 8; i += GNT) gds[i] = P.gd[(size_t)set * (P.gd_bytes / 8) + i];
    __syncthreads();
    const int ncols = T.ncols, nseg = T.nseg, npanel = T.npanel;
    constexpr int NCS = g_aslots(NTILE);   // chunk t uses a-panel slot t % NCS
    const int nactive = min(nseg, G_NISS);   // issuing warps that own a segment
    for (int i = tid; i < P.rk_bytes / 16; i += GNT) rks[i] = P.rk[(size_t)T.lpar * (P.rk_bytes / 16) + i];
    if (tid == 0) {
        for (int i = 0; i < G_DG; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gfull0 + 8 * i));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(gfree0 + 8 * i));   // 4 warps of a T team + 4 of an a-panel team
        }
        for (int i = 0; i < G_NS_MAX; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(tfull0 + 8 * i));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tfree0 + 8 * i), "r"(nactive));
        }
        for (int i = 0; i < G_NAT * 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(afull0 + 8 * i));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(afree0 + 8 * i), "r"(nactive));
        }
        for (int i = 0; i < 4 * G_NAT * G_FD; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(smem + P.off_f + 4 * G_NAT * G_FD * 2048) + 8 * i));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(drained), "r"(nactive * G_NALT));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(flushed), "r"(S_NPW));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(issued0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(issued0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    const int CPR = (P.H + KC - 1) / KC;
    const long long total = (long long)analysis_items(0, P.H, P.H) * CPR;
    const long long gbeg = total * part / nparts, glen = total * (part + 1) / nparts - gbeg;
    const long long cbeg = gbeg + glen * split / P.nsplits, cend = gbeg + glen * (split + 1) / P.nsplits;
    const int nchunk = (int)(cend - cbeg);
    float *out = partial + (size_t)bid * BSLOT_WORDS;
    // first read-back event (chunks): staggered over the CTAs, which sweep the chunk list in lock step (see s_first_event)
#ifdef Z3D_GEN_DEBUG
    const int ev0 = (P.dbg & 1) ? (1 << 28) : (bid % 8 + 1) * (BFLUSH_CHUNKS / 8);
#else
    const int ev0 = (bid % 8 + 1) * (BFLUSH_CHUNKS / 8);
#endif
    const int nregular = nchunk > ev0 ? (nchunk - 1 - ev0) / BFLUSH_CHUNKS + 1 : 0;  // events in front of chunks ev0 + k BFLUSH_CHUNKS < nchunk
    const int np16 = npanel * 16;

    if (nchunk == 0) {
        for (int i = tid; i < NTILE * ncols * BTM; i += GNT) out[i] = 0.0f;  // the reduction reads every split
    } else if (warp < G_NISSW) {
        // ---- issuing warp: ONE converged warp issues all MMAs of a chunk (elect.sync), then one commit.
        // Tried (compile with -DG_NALT_=2): two warps that ALTERNATE over the chunks - warp a issues the chunks t = a mod 2 and does
        // its waits, tcgen05.fence and commit while the other warp's MMAs execute; the chunks enter the tensor core in order through
        // the hand-over barriers `issued[a]` (a warp arrives after its chunk's last MMA instruction and waits for the other's before
        // its next chunk's first).  Bit-identical results, but 4.24 ms against 3.73 ms per 64 x 128^3 tile: what the single issuing
        // warp loses is mostly starvation (read-back bubbles, T / a-panel hand-overs), not its own bookkeeping, and the second
        // spinning warp costs the producers issue slots.
        // (G_NISS > 1, the earlier scheme: warp w owns the segments g = w, w + G_NISS, .. - different accumulator columns, no order.)
        const int wi = warp / G_NALT, alt = warp % G_NALT;
        if (wi < nactive) {
            int4 sg[G_SEG_PER_ISS];
#pragma unroll
            for (int q = 0; q < G_SEG_PER_ISS; ++q) sg[q] = I.seg[set][wi + q * G_NISS];
            const unsigned dlo0 = (smem_u32(smem) + (unsigned)P.off_t) >> 4;          // T ring, 16-byte units (descriptor start-address field)
            const unsigned long long dhi = ((unsigned long long)(128 >> 4) << 32) | (1ull << 46) | ((unsigned long long)ncols << 16);  // SBO 128 B, LBO = ncols * 16 B, version 1
            const unsigned lo_t = (unsigned)(2 * ncols), step = (unsigned)(P.tslot_bytes >> 4);
            const unsigned a0 = tmem + (unsigned)T.acol0;
            const unsigned my_issued = issued0 + 8 * alt, other_issued = issued0 + 8 * (alt ^ 1);
            int ts = alt % NS;
            unsigned tuse = 0, nev = 0;
            int next_ev = ev0;        // the next read-back event lies in front of this chunk
            unsigned mine = 0;        // chunks this warp has issued
#ifdef Z3D_GEN_CTACLK
            long long gc_flushed = 0, gc_tfull = 0, gc_afull = 0, gc_issue = 0;
#endif
            auto cross_event = [&](bool wait) {   // pass the event in front of chunk next_ev (or the final one, next_ev >= nchunk)
                const int before = min(next_ev, nchunk) - 1;   // the last chunk in front of the event
                if (G_NALT > 1 && before % G_NALT != alt) {    // not mine: my own MMAs so far must have retired as well
                    if (elect_one()) umma_commit(drained);
                    __syncwarp();
                }
                if (wait) {
                    GC_T(g0);
                    GW(flushed, nev & 1, 1, next_ev);
                    GD_SLEEP(2);
                    GC_T(g1);
                    GC_ADD(gc_flushed, g0, g1);
                }
                ++nev;
                next_ev += BFLUSH_CHUNKS;
            };
            for (int t = alt; t < nchunk; t += G_NALT) {
                GT_T(i0);
                bool first = t == 0;
                if (next_ev <= t) {  // the accumulators have been read back -> overwrite
                    first = next_ev == t;
                    cross_event(true);
                }
                GT_T(i1);
                const int as = t % NCS;
                const unsigned dlo = dlo0 + (unsigned)ts * step;
                GC_T(g2);
                GW(tfull0 + 8 * ts, tuse & 1, 2, t);
                GC_T(g3);
                GC_ADD(gc_tfull, g2, g3);
                GT_T(i2);
                GT_T(i3);
                if (warp == 0) { GT_ADD(0, i0, i1); GT_ADD(1, i1, i2); }
#pragma unroll
                for (int tile = 0; tile < NTILE; ++tile) {
                    GT_T(j0);
                    GC_T(g4);
                    GW(afull0 + 8 * (as * NTILE + tile), (unsigned)(t / NCS) & 1, 3, t);
                    GC_T(g5);
                    GC_ADD(gc_afull, g4, g5);
                    GD_SLEEP(4);
                    GT_T(j1);
                    if (warp == 0) GT_ADD(2, j0, j1);
                    tc_fence_after();
                    if (G_NALT > 1 && tile == 0 && t > 0) GW(other_issued, (alt ? mine : mine - 1) & 1, 10, t);   // chunk t - 1 is in the queue
                    GT_T(k0);
                    GC_T(g6);
                    if (elect_one()) {
                        const unsigned a = a0 + (unsigned)((as * NTILE + tile) * np16);
#pragma unroll
                        for (int q = 0; q < G_SEG_PER_ISS; ++q) {
#ifdef Z3D_GEN_DEBUG
                            if (P.dbg & 16384) continue;
#endif
                            const int4 sq = sg[q];
                            if (sq.w >= 0) {
                                const unsigned long long b = dhi | (unsigned long long)(dlo + (unsigned)sq.z);
                                const unsigned d = tmem + (unsigned)(tile * ncols) + (unsigned)sq.z, ah = a + (unsigned)sq.y;
#ifdef Z3D_GEN_DEBUG
                                if (P.dbg & 2048) { umma_tf32_ts(d, ah, b, (unsigned)sq.x, first ? 0u : 1u); continue; }
#endif
                                umma_tf32_ts(d, ah, b + lo_t, (unsigned)sq.x, first ? 0u : 1u);  // small terms first
                                umma_tf32_ts(d, ah + 8, b, (unsigned)sq.x, 1u);
                                umma_tf32_ts(d, ah, b, (unsigned)sq.x, 1u);
                            }
                        }
                        // ONE commit per chunk: tcgen05.commit holds the issuing thread for ~420 clk (measured: the kernel stripped of
                        // all work spent 845 clk per chunk in this block with two commits, 1 260 with three), so the T slot and the
                        // a-panel slot are released by the same mbarrier - the a-panel team waits on the T ring's barrier of its
                        // slot's previous chunk
                        GT_T(k1);
                        if (tile == NTILE - 1) {
                            if (G_NALT > 1) mbar_arrive(my_issued);   // the other warp may issue chunk t + 1 while this one commits
                            umma_commit(tfree0 + 8 * ts);
                            if (t + 1 == next_ev || t == nchunk - 1) umma_commit(drained);  // a read-back follows this chunk
                        }
                        GT_T(k2);
#ifdef Z3D_GEN_TIMING
                        if (warp == 0 && bid == GT_BLOCK) { atomicAdd((unsigned long long *)&g_gdbg[18], (unsigned long long)(k1 - k0)); atomicAdd((unsigned long long *)&g_gdbg[19], (unsigned long long)(k2 - k1)); }
#endif
                    }
                    __syncwarp();
                    GC_T(g7);
                    GC_ADD(gc_issue, g6, g7);
                    GT_T(k3);
                    if (warp == 0) { GT_ADD(17, j1, k0); GT_ADD(20, k0, k3); }
                }
                GT_T(i4);
                if (warp == 0) GT_ADD(3, i3, i4);
                ++mine;
                ts += G_NALT;
                if (ts >= NS) { ts -= NS; ++tuse; }
            }
#ifdef Z3D_GEN_CTACLK
            if (lane == 0 && alt == 0 && bid < 1024) {
                g_cta_clk[1][bid] = gc_flushed; g_cta_clk[2][bid] = gc_tfull; g_cta_clk[3][bid] = gc_afull; g_cta_clk[4][bid] = gc_issue;
            }
#endif
            if (G_NALT > 1) {  // events behind this warp's last chunk: every issuing warp arrives on `drained` at every event
                while (next_ev < nchunk) cross_event(true);
                if ((nchunk - 1) % G_NALT != alt) cross_event(false);
            }
        }
    } else if (warp < G_WARP_SEED0 + G_NSEEDW) {
        // ---- seed warps: lane = (panel j = lane / 8 (+ 4 in the second pass), voxel k = lane % 8)
        const int sw = warp - G_WARP_SEED0, k = lane & 7;
        int pm[2], ppz[2], pnl[2], psoff[2], pgd[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = (lane >> 3) + 4 * q;
            const bool v = j < npanel;
            pm[q] = v ? T.panel_m[j] : 0;
            ppz[q] = v ? T.panel_pz[j] : 0;
            pnl[q] = v ? T.panel_nl[j] : 0;
            psoff[q] = v ? T.panel_soff[j] : 0;
            pgd[q] = v ? T.panel_gd[j] : 0;
        }
        const int npass = npanel > 4 ? 2 : 1;
        int maxnl = 0;
        for (int j = 0; j < npanel; ++j) maxnl = max(maxnl, T.panel_nl[j]);
        int item_have = -1;
        double rho2 = 0.0, rhom[2] = {0.0, 0.0};
        float tr[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (int t = sw; t < nchunk; t += G_NSEEDW) {
            const long long cc = cbeg + t;
            const int item = (int)(cc / CPR), kc = (int)(cc % CPR);
            if (item != item_have) {  // row group changed: rho, rho^m and the row trig of this lane's panels (float64)
                item_have = item;
                int ip, jp;
                bool paired;
                decode_item(item, 0, P.H, P.H, ip, jp, paired);
                const double y0 = P.side[ip], x0 = P.side[jp];
                rho2 = x0 * x0 + y0 * y0;
                const double rho = sqrt(rho2);
                double ur = 1.0, ui = 0.0;
                if (rho > 0.0) {
                    ur = x0 / rho;
                    ui = y0 / rho;
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int m = pm[q];
                    double br = ur, bi = ui, cr = 1.0, ci = 0.0, pw = 1.0, pb = rho;
                    for (int e = m; e; e >>= 1) {  // (cos, sin)(m phi) by complex powering, rho^m by real powering
                        if (e & 1) {
                            const double tt = cr * br - ci * bi;
                            ci = cr * bi + ci * br;
                            cr = tt;
                            pw *= pb;
                        }
                        const double tt = br * br - bi * bi;
                        bi = 2.0 * br * bi;
                        br = tt;
                        pb *= pb;
                    }
                    rhom[q] = pw;
                    const int mq = m & 3;  // transpose: phi' = atan2(x0, y0) = pi/2 - phi (mod 2 pi)
                    tr[q][0] = (float)cr;
                    tr[q][1] = (float)ci;
                    tr[q][2] = (float)((mq == 0) ? cr : (mq == 1) ? ci : (mq == 2) ? -cr : -ci);
                    tr[q][3] = (float)((mq == 0) ? -ci : (mq == 1) ? cr : (mq == 2) ? ci : -cr);
                }
            }
            const int kk = kc * KC + k;
            const double z = (kk < P.H) ? P.side[kk] : 0.0;
            const double zz = z * z;
            const float r2 = (float)(rho2 + zz), z4 = (float)(4.0 * zz), r4 = r2 * r2;
            const int gs = t % G_DG;
            GT_T(s0);
            if (t >= G_DG) GWR(gfree0 + 8 * gs, (unsigned)(t / G_DG - 1) & 1, 4, t);
            GD_SLEEP(8);
            GT_T(s1);
            float *slot = reinterpret_cast<float *>(gring + (size_t)gs * P.gslot_bytes);
#ifdef Z3D_GEN_DEBUG
            if (P.dbg & 32768) { __syncwarp(); if (lane == 0) mbar_arrive(gfull0 + 8 * gs); continue; }
#endif
            float *seeds = slot + 40;  // [0..7] r^2, [8..39] trig of the panels, then the seeds [row][8]
            if (lane < 8) slot[lane] = r2;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q < npass) {
                    const int j = (lane >> 3) + 4 * q;
                    const float tv = k == 0 ? tr[q][0] : k == 1 ? tr[q][1] : k == 2 ? tr[q][2] : tr[q][3];
                    if (k < 4 && j < npanel) slot[8 + j * 4 + k] = tv;
                    // S_{l0}: l0 = m (pz = 0): rho^m;  l0 = m + 1: 2 z rho^m   (Q_mm = sin^m, Q_{m+1,m} = 2c Q_mm, zm:2012-2020 / kappa)
                    float S1 = (float)(ppz[q] ? 2.0 * z * rhom[q] : rhom[q]), S2 = 0.0f;
                    float *sp = seeds + psoff[q] * 8 + k;
                    const float2 *gp = gds + pgd[q];
                    const int nl = pnl[q];
                    if (nl > 0) sp[0] = S1;
                    // coefficients in blocks of 8, loaded before the block's stores (loads cannot be hoisted over shared-memory
                    // stores by the compiler; one load per step in front of a dependent store is 50 clocks per step)
#ifdef G_OLDSEED
                    for (int i = 1; i < maxnl; ++i) {
                        const float2 g = gp[i - 1];
                        const float S = fmaf(fmaf(-g.x, r2, z4), S1, (g.y * r4) * S2);
                        S2 = S1;
                        S1 = S;
                        if (i < nl) sp[i * 8] = S;
                    }
#else
                    for (int i0 = 1; i0 < maxnl; i0 += 8) {
                        float2 g[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) g[e] = gp[i0 - 1 + e];  // tables are padded: reads past a chain are throw-away
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float S = fmaf(fmaf(-g[e].x, r2, z4), S1, (g[e].y * r4) * S2);
                            S2 = S1;
                            S1 = S;
                            if (i0 + e < nl) sp[(i0 + e) * 8] = S;
                        }
                    }
#endif
                }
            }
            GD_SLEEP(16);
            __syncwarp();
            if (lane == 0) mbar_arrive(gfull0 + 8 * gs);
            GT_T(s2);
            if (sw == 0) { GT_ADD(4, s0, s1); GT_ADD(5, s1, s2); }
        }
    } else if (warp < G_WARP_A0) {
        // (a special warp without a role)
    } else if (warp < G_WARP_T0) {
        // ---- a-panel teams: thread = row (re/im, volume b) of the MMA's M dimension = tensor-memory lane
        const int team = (warp - G_WARP_A0) >> 2, quarter = warp & 3;
        const int row = quarter * 32 + lane, reim = row >> 6, b = row & 63;
        const float sign = reim ? -1.0f : 1.0f;  // im rows are negated
        const unsigned arow0 = tmem + ((unsigned)(32 * quarter) << 16) + (unsigned)T.acol0;
        bool had = false;
        int nev = 0;
        // The folds of a warp's 32 rows (2 x 1 KB per (chunk, tile) step) come through the warp's PRIVATE ring of G_FD slots in shared
        // memory: lane 0 issues two bulk copies per step, G_FD steps ahead, completion counted on the slot's mbarrier; nobody else
        // touches the ring, so there is no consumer barrier - the warp re-arms a slot itself once its lanes have read it.  (Loaded
        // into registers one step ahead, a DRAM round trip sat on the team's critical path at every step - a tile's folds are
        // 0.5 GB - and two steps ahead the registers spilled.)
        const int aw = warp - G_WARP_A0;
        const float *fring = reinterpret_cast<const float *>(smem + P.off_f + (size_t)aw * (G_FD * 2048));
        const unsigned fring_s = smem_u32(fring), fbar0 = smem_u32(smem + P.off_f + 4 * G_NAT * G_FD * 2048) + (unsigned)(aw * G_FD * 8);
        const float *fwarp = Ft + (size_t)T.kind * (S_FK_BYTES / 4) + (size_t)(reim * SNV + (b & 32)) * KC;   // the warp's 32 rows: 1 KB
        const int npos = nchunk > team ? ((nchunk - team + G_NAT - 1) / G_NAT) * NTILE : 0;
        auto fissue = [&](int p) {   // lane 0 only
            const int t = team + G_NAT * (p / NTILE), tile = p % NTILE, sl = p % G_FD;
            const float *q = fwarp + (size_t)tile * tile_stride + (size_t)(cbeg + t) * (4 * S_FK_BYTES / 4);
            mbar_expect_tx(fbar0 + 8 * sl, 2048);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 1024, [%2];"
                         ::"r"(fring_s + (unsigned)(sl * 2048)), "l"(q), "r"(fbar0 + 8 * sl) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 1024, [%2];"
                         ::"r"(fring_s + (unsigned)(sl * 2048 + 1024)), "l"(q + 2 * SNV * KC), "r"(fbar0 + 8 * sl) : "memory");
        };
        if (lane == 0)
            for (int p = 0; p < G_FD && p < npos; ++p) fissue(p);
        int p = 0;
        for (int t = team; t < nchunk; t += G_NAT) {
            int pos, len;
            s_interval(t, ev0, nchunk, pos, len);
            const float sg = sign * (1.0f + BTRUNC_PER_MMA * (float)(3 * (len - 1 - pos) + 1));  // truncation pre-compensation
            const int gs = t % G_DG;
            const float *slot = reinterpret_cast<const float *>(gring + (size_t)gs * P.gslot_bytes);
            const int cs = t % NCS;
            const unsigned arow = arow0 + (unsigned)(cs * NTILE * np16);
            GT_T(a0);
            GWR(gfull0 + 8 * gs, (unsigned)(t / G_DG) & 1, 5, t);
            GT_T(a1);
#pragma unroll
            for (int tile = 0; tile < NTILE; ++tile, ++p) {
                GT_T(b0);
#ifdef Z3D_GEN_ABL
                if (!(P.dbg & 65536))   // timing ablation: an a-panel ring of unbounded depth
#endif
                if (t >= NCS && tile == 0) GWR(tfree0 + 8 * ((t - NCS) % NS), (unsigned)((t - NCS) / NS) & 1, 6, t);  // the slot's previous MMAs (chunk t - NCS) retired
                GD_SLEEP(32);
                const int sl = p % G_FD;
                GWR(fbar0 + 8 * sl, (unsigned)(p / G_FD) & 1, 9, t);   // this step's folds have landed
                GT_T(b1);
                if (warp == G_WARP_A0) GT_ADD(7, b0, b1);
                tc_fence_after();
                const float4 *fq = reinterpret_cast<const float4 *>(fring + sl * 512) + lane * 2;
                const float4 fa0 = fq[0], fa1 = fq[1], fb0 = fq[64], fb1 = fq[65];
                const float fa[8] = {fa0.x, fa0.y, fa0.z, fa0.w, fa1.x, fa1.y, fa1.z, fa1.w};
                const float fb[8] = {fb0.x, fb0.y, fb0.z, fb0.w, fb1.x, fb1.y, fb1.z, fb1.w};
#ifdef Z3D_GEN_DEBUG
                if (!(P.dbg & 4096))
#endif
                for (int j = 0; j < npanel; ++j) {
                    const float tA = slot[8 + j * 4 + reim] * sg, tB = slot[8 + j * 4 + 2 + reim] * sg;
                    float hi[8], lo[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) split_fast(fmaf(fa[q], tA, fb[q] * tB), hi[q], lo[q]);
                    tmem_st8(arow + (unsigned)(tile * np16 + j * 16), hi);
                    tmem_st8(arow + (unsigned)(tile * np16 + j * 16 + 8), lo);
                }
                GT_T(c0);
                tmem_st_wait();
                GT_T(c1);
                GD_SLEEP(64);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(afull0 + 8 * (cs * NTILE + tile));
                    if (tile == NTILE - 1) mbar_arrive(gfree0 + 8 * gs);
                    // refill the fold slot only now, off the path between "slot free" and "panels ready" (the slot round trip through
                    // the issuing warp paces the kernel; 3.72 -> 3.62 ms per 64 x 128^3 tile).  Every lane consumed its folds in the
                    // panel loop above, so the loads have completed before the bulk copies may overwrite the slot.
                    if (p + G_FD < npos) {
                        fence_async_smem();
                        fissue(p + G_FD);
                    }
                }
                GT_T(c2);
                if (warp == G_WARP_A0) { GT_ADD(14, b1, c0); GT_ADD(15, c0, c1); GT_ADD(16, c1, c2); }
            }
            GT_T(a3);
            while (nev < nregular && ev0 + nev * BFLUSH_CHUNKS <= t) g_flush_event(tmem, NTILE * ncols, out, had, drained, flushed, nev, false, GD_FLUSHSLEEP);
            GT_T(a4);
            if (warp == G_WARP_A0) { GT_ADD(6, a0, a1); GT_ADD(8, a1, a3); GT_ADD(9, a3, a4); }
        }
        while (nev < nregular) g_flush_event(tmem, NTILE * ncols, out, had, drained, flushed, nev, false, GD_FLUSHSLEEP);
        g_flush_event(tmem, NTILE * ncols, out, had, drained, flushed, nev, true, GD_FLUSHSLEEP);
    } else {
        // ---- T teams: thread u of a team owns tasks u and u + 128: (chain (l, m), k half) -> 4 voxels of every n of the chain
        // the tasks are sorted by chain length, so a team's first logical warp is its busiest: rotate the logical warps over the
        // physical ones by team, or the four teams' busiest warps all sit on scheduler 0 next to the issuing warp
#if defined(G_NOROT)
        const int team = (warp - G_WARP_T0) >> 2, u = ((warp - G_WARP_T0) & 3) * 32 + lane;
#elif defined(G_LIGHT0)
        // the issuing warp's scheduler (warps = 0 mod 4) gets every team's LIGHTEST logical warp; the other three rotate by team
        const int team = (warp - G_WARP_T0) >> 2, pq = warp & 3, u = (pq == 0 ? 3 : (pq - 1 + team) % 3) * 32 + lane;
#else
        const int team = (warp - G_WARP_T0) >> 2, u = ((warp - team + 1) & 3) * 32 + lane;
#endif
        int4 task[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = u + 128 * q;
            task[q] = i < T.ntask ? P.tasks[(size_t)set * G_TASKS + i] : make_int4(0, 0, 0, 0);
        }
        bool had = false;
        int nev = 0;
        for (int t = team; t < nchunk; t += G_NTT) {
            const int ts = t % NS;
            const unsigned tuse = (unsigned)(t / NS);  // how often the slot has been filled before
            const int gs = t % G_DG;
            const float *slot = reinterpret_cast<const float *>(gring + (size_t)gs * P.gslot_bytes);
            float *tb = reinterpret_cast<float *>(tring + (size_t)ts * P.tslot_bytes);
            GT_T(t0);
            GWR(gfull0 + 8 * gs, (unsigned)(t / G_DG) & 1, 7, t);
            GT_T(t1);
            if (tuse) GWR(tfree0 + 8 * ts, (tuse - 1) & 1, 8, t);  // the slot's previous MMAs retired
            GD_SLEEP(128);
            GT_T(t2);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#ifdef Z3D_GEN_DEBUG
                const int len = (P.dbg & 8192) ? 0 : task[q].y;
#else
                const int len = task[q].y;
#endif
                if (len > 0) {
                    const int kh = task[q].w >> 16, srow = task[q].w & 0xffff;
                    const float4 s4 = *reinterpret_cast<const float4 *>(slot + 40 + srow * 8 + 4 * kh);
                    const float4 r2 = *reinterpret_cast<const float4 *>(slot + 4 * kh);
                    const float4 *K = rks + task[q].z;
                    float *hi = tb + (size_t)(kh * ncols + task[q].x) * 4, *lo = hi + (size_t)2 * ncols * 4;
                    float4 T1 = make_float4(0.f, 0.f, 0.f, 0.f), T2 = s4;
                    float4 kc4 = K[0];
#pragma unroll 2
                    for (int i = 0; i < len; ++i) {
                        const float4 kn = K[i + 1];  // prefetch (the table is padded)
                        float4 Tn, h, l;
                        Tn.x = fmaf(fmaf(kc4.x, r2.x, kc4.y), T1.x, kc4.z * T2.x);
                        Tn.y = fmaf(fmaf(kc4.x, r2.y, kc4.y), T1.y, kc4.z * T2.y);
                        Tn.z = fmaf(fmaf(kc4.x, r2.z, kc4.y), T1.z, kc4.z * T2.z);
                        Tn.w = fmaf(fmaf(kc4.x, r2.w, kc4.y), T1.w, kc4.z * T2.w);
                        split4_fast(Tn, h, l);
#ifdef Z3D_GEN_DEBUG
                        if (P.dbg & 1024) { if (h.x == 123.456f) *reinterpret_cast<float4 *>(hi + i * 4) = l; } else
#endif
                        {
                        *reinterpret_cast<float4 *>(hi + i * 4) = h;
                        *reinterpret_cast<float4 *>(lo + i * 4) = l;
                        }
                        T2 = T1;
                        T1 = Tn;
                        kc4 = kn;
                    }
                }
            }
            fence_async_smem();  // generic-proxy T writes -> visible to the tensor core's async-proxy reads
            GD_SLEEP(256);
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(tfull0 + 8 * ts);
                mbar_arrive(gfree0 + 8 * gs);
            }
            GT_T(t3);
            while (nev < nregular && ev0 + nev * BFLUSH_CHUNKS <= t) g_flush_event(tmem, NTILE * ncols, out, had, drained, flushed, nev, false, GD_FLUSHSLEEP);
            GT_T(t4);
            if (u < 32 && team == 0) { GT_ADD(10, t0, t1); GT_ADD(11, t1, t2); GT_ADD(12, t2, t3); GT_ADD(13, t3, t4); }
        }
        while (nev < nregular) g_flush_event(tmem, NTILE * ncols, out, had, drained, flushed, nev, false, GD_FLUSHSLEEP);
        g_flush_event(tmem, NTILE * ncols, out, had, drained, flushed, nev, true, GD_FLUSHSLEEP);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
#ifdef Z3D_GEN_CTACLK
    if (tid == 0 && bid < 1024) g_cta_clk[0][bid] = clock64() - cta_t0;
#endif
}


// ---------------------------------------------------------------------------------------------------------------------
// Measured tensor-pipe peak for the roofline (bench.py): back-to-back tcgen05.mma kind::tf32, M 128, N 256, K 8, shared-memory
// operands (zeros), one converged warp issuing through elect.sync, one CTA per SM.  What the pipe sustains when nothing else
// happens on the SM: N / 2 clk per instruction = 4096 flop / clk / SM.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_tf32_peak(int iters) {
    __shared__ __align__(128) float sm[2 * 256 * 8 + 2 * 128 * 8];   // B: 256 rows x 8 (k halves apart), A: 128 rows x 8
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 2 * 256 * 8 + 2 * 128 * 8; i += 128) sm[i] = 0.0f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = tmem_slot;
    if (warp == 0) {
        const unsigned long long dB = umma_desc(smem_u32(sm), 256 * 16, 128), dA = umma_desc(smem_u32(sm + 2 * 256 * 8), 128 * 16, 128);
        const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 8; ++j) umma_tf32(tmem + (unsigned)((j & 1) * 256), dA, dB, idesc, 1u);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

}  // namespace z3d
