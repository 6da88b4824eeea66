// zernike3d_b200 - batched analysis, streamed-T variant: tiles of up to 64 volumes, operand T read straight from HBM.
//
// Same contraction as z3d_batched.cuh,
//     D[(n,l,m)][(re/im, b)] += T[(n,l,m)][k] * a[(m,pz)][k][(re/im, b)],   T = R_nl * Q_lm,  a = F_A,b trig_A + F_B,b trig_B,
// but with the GEMM transposed and the volume-independent operand precomputed:
//   * the MMA's M = 128 rows are the 2 x 64 (re/im, volume) columns of a tile - always full - and its N dimension are the
//     moment rows of one (m, pz), padded to 16 instead of 128 (fill 94 % instead of 61 % at order 50);
//   * T, already split into its TF32 high part and float32 remainder and already in the K-major core-matrix layout of
//     the shared-memory descriptors, is a plan-owned table in HBM (k_build_T, once per plan: 13.7 GB at 128^3 / order 50,
//     64 B per moment row and chunk) that the TMA engine (cp.async.bulk) streams into a ring of stages; the tensor core
//     reads it where it lands.  No thread ever touches T: what the 31 producer warps of z3d_batched.cuh spent most of
//     their shared-memory bandwidth on is gone, and so is the R / Q staging.
//   * the producers only build the a panels (one per (m, pz) of the CTA's column set, 128 x 8 values) from the folds.
// A CTA owns a *column set* - up to 512 accumulator columns (all of tensor memory) made of segments of <= 256 moment rows
// of <= S_PANELS different (m, pz) - and one split of the chunk list.  Sets are packed by the planner (z3d_api.cu) to
// equal cost; a (m, pz) may be cut between two sets.  The folds are stored by *kind* = (m parity, pz): a panel needs only
// the 4 fold classes of its kind (8 KB per chunk), a set holds one kind.
//
// Designed floors per tile: the HBM stream of T (every tile streams the whole table once, 64 B x rows16 x chunks) and
// 3 x 2 x 128 x 8 flops per row and chunk on the tensor pipe, within 20 % of each other; DESIGN.md section 3 has what the kernel
// measures against them and why.
#pragma once
#include "z3d_batched.cuh"

namespace z3d {

constexpr int SNV = 64;               // volumes per tile
constexpr int SM_ROWS = 2 * SNV;      // MMA M: [re | im][volume]
constexpr int S_PANELS = 4;           // a panels ((m, pz) pieces) per column set at most
constexpr int S_SEGS = 8;             // MMA segments (<= 256 columns each) per column set at most
constexpr int S_KINDS = 1;            // fold kinds staged per chunk: the planner closes a column set at a kind boundary
constexpr int S_FK_BYTES = 4 * SNV * KC * 4;       // folds of one kind and chunk: [set A|B][re|im class][volume][k]
constexpr int S_PANEL_BYTES = 2 * SM_ROWS * KC * 4;  // one a panel: [hi | lo][2 k halves][128 rows][4]
constexpr int S_MAX_STAGES = 8;
constexpr int S_OFF_TABS = 0;                       // int tables, see StreamTabs
constexpr int S_OFF_MBAR = 256;                     // fready[3], ready[3], done[3], drained, flushed, tensor-memory base
constexpr int S_OFF_SEG = 512;                      // per-segment descriptor parts, int4 [S_SEGS]
constexpr int S_OFF_BLK = 640;                      // read-back list: int nblk, pad, int blk[32] (16-column blocks of the accumulators)
constexpr int S_OFF_A = 1024;                       // the three slots: [T block][folds][a panels]
constexpr int S_ABUF_BYTES = S_PANELS * S_PANEL_BYTES;

struct StreamTabs {          // per column set, copied to shared memory
    int ncols, nseg, npanel, nkind;
    int kind[S_KINDS];
    int seg_cols[S_SEGS], seg_off[S_SEGS], seg_panel[S_SEGS];
    int panel_code[S_PANELS];   // m | pz << 16 | kind slot << 24
    long long toff;             // offset (floats) of the set's block inside a chunk record of the T table
};
static_assert(sizeof(StreamTabs) <= S_OFF_MBAR, "tables");

struct DevPlanS {
    int order, dim, H;
    int nsets, nsplits, nstages;
    int stage_bytes;           // bytes of a slot: T block of the widest set + the folds of one kind + the a panels
    int t_bytes_max;
    long long trec;            // floats per chunk record of the T table (16 x all columns)
    const StreamTabs *tabs;    // [nsets]
    const float *Tt;           // [chunks][trec]: per set [hi | lo][2 k halves][ncols][4]
    const float *trig;         // [items][4][LMAX + 1]
};

// ---------------------------------------------------------------------------------------------------------------------
// T table: one CTA per chunk; a thread per (column, k half): product of the R and Q table entries (k_build_basis), TF32
// split, two 16-byte stores.  colsrc[j] = (R index, Q index, float offset of (column, k half 0, hi) in the record, set's
// column count); padding columns have R index -1 and are written as zeros.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_build_T(const float *__restrict__ Rt, const float *__restrict__ Qt, int nrad, int nq,
                                                  const int4 *__restrict__ colsrc, int ncols_all, long long trec,
                                                  float *__restrict__ Tt) {
    const long long chunk = blockIdx.x;
    const float *Rs = Rt + (size_t)chunk * nrad * KC, *Qs = Qt + (size_t)chunk * nq * KC;
    float *rec = Tt + (size_t)chunk * trec;
    for (int it = threadIdx.x; it < 2 * ncols_all; it += blockDim.x) {
        const int kh = it & 1, j = it >> 1;
        const int4 cs = colsrc[j];
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cs.x >= 0) {
            const int sr = ((cs.x >> 2) & 1) << 2, sq = ((cs.y >> 2) & 1) << 2;  // the tables' k-half swizzle
            const float4 r = *reinterpret_cast<const float4 *>(Rs + cs.x * KC + ((4 * kh) ^ sr));
            const float4 q = *reinterpret_cast<const float4 *>(Qs + cs.y * KC + ((4 * kh) ^ sq));
            t = make_float4(r.x * q.x, r.y * q.y, r.z * q.z, r.w * q.w);
        }
        float4 hi, lo;
        split4(t, hi, lo);
        float *p = rec + cs.z + (size_t)kh * cs.w * 4;
        *reinterpret_cast<float4 *>(p) = hi;
        *reinterpret_cast<float4 *>(p + (size_t)2 * cs.w * 4) = lo;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Folds of a tile, by kind: F[chunk][kind = (m parity) << 1 | pz][set A|B][re|im][volume][k].  The real part of a moment
// with m parity mp needs the class (y even, x parity mp, z parity pz), the imaginary part (y odd, x parity mp ^ 1, pz).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(2 * SNV * KC) k_fold_tile_kind(const float *__restrict__ vol, int nvol, int N, int H,
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
        dst[(((kind * 2 + set) * 2 + reim) * SNV + b) * KC + (k ^ (((b >> 2) & 1) << 2))] = v[c8];
    }
}

// ---- warp roles -----------------------------------------------------------------------------------------------------
// Warps 0-3 are the special warps, one per scheduler (SM sub-partition = warp % 4), so that none of them competes with
// another for issue slots: warp 0 issues the bulk copies, warps 1-3 issue the MMAs of segments g = w-1, w+2, w+5 (a single
// thread needs several hundred clocks per segment for descriptor arithmetic and the uniform-register traffic of three
// UTCHMMA; one issuing warp was the critical path at 2 900 clk per chunk).  Warps 4-27 are three producer TEAMS of 8 warps:
// team i builds the a panels of the chunks t = i, i + 3, .. into a buffer i (a thread owns one (row, k half) of every panel
// of the set and loads its folds once per chunk), so a team has three chunk periods for the load -> split -> store ->
// proxy fence -> arrive latency chain; all 24 warps read the accumulators back.
constexpr int S_NMMA = 3;                 // issuing warps
constexpr int S_WARP_TMA = 0;
constexpr int S_WARP_PROD0 = 1 + S_NMMA;  // first producer warp
constexpr int S_NTEAM = 3;                // producer teams = a-panel buffers
constexpr int S_TEAM_WARPS = 8;           // 256 threads = 128 rows x 2 k halves
constexpr int S_NPW = S_NTEAM * S_TEAM_WARPS;
constexpr int S_NPROD = S_NPW * 32;
constexpr int SNT = (S_WARP_PROD0 + S_NPW) * 32;   // 896 threads
constexpr int S_SEG_PER_WARP = (S_SEGS + S_NMMA - 1) / S_NMMA;
// first read-back event of a CTA (chunks): CTAs sweep the chunk list in lock step, staggering spreads the L2 traffic of the
// read-backs (360 KB per CTA and event) over the interval instead of bursting it from all 148 CTAs at once
__device__ __forceinline__ int s_first_event() { return (int)(blockIdx.x % 8 + 1) * (BFLUSH_CHUNKS / 8); }

__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
    unsigned p;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
// wait with a suspend-time hint: the producers run chunks ahead of the tensor core, their wake-up latency is hidden
__device__ __forceinline__ void mbar_wait_relaxed(unsigned mbar, unsigned parity) {
#pragma unroll 1
    for (int it = 0; it < (1 << 24); ++it) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity), "r"(2000u) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_normal() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_copy_hint(void *dst, const void *src, unsigned bytes, unsigned mbar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(mbar), "l"(pol) : "memory");
}

__device__ __forceinline__ float ld_hint(const float *p, unsigned long long pol) {
    float v;
    asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void st_hint(float *p, float v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

// Accumulator read-back: accumulators -> added into the CTA's partial buffer [column][row].  Producer warp w can address lane
// quarter w % 4; the 6 warps of a quarter take the 32-column units iq, iq + 6, ..  A unit is one dependent round trip to L2
// (tcgen05.ld, 32 loads, add, 32 stores) of 2 000 - 3 500 clk while the T stream saturates the memory system (measured with
// clock64 stamps), and the tensor core idles until the slowest warp is through: the units are dealt out evenly (at most 3 per
// warp for 512 columns) and the partial sums are kept L2-resident (evict-last; the folds, 537 MB per tile, are evict-normal).
template <int WARP0 = S_WARP_PROD0>
__device__ __forceinline__ void s_flush(unsigned tmem, int ncols, float *out, bool add, unsigned drained_bar, unsigned parity, bool dbg_sleep = false) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3;
    const unsigned long long pol = l2_policy_evict_last();
    mbar_wait(drained_bar, parity);  // every MMA issued so far has retired
    if (dbg_sleep) __nanosleep(3000);
    tc_fence_after();
#pragma unroll 1
    for (int c32 = 32 * ((warp - WARP0) >> 2); c32 < ncols; c32 += 32 * (S_NPW / 4)) {
        unsigned v[32];
        tmem_ld32(tmem + ((unsigned)(32 * q) << 16) + c32, v);
        float *p = out + (size_t)c32 * BTM + 32 * q + lane;
        if (add) {
            float old[32];
#pragma unroll
            for (int col = 0; col < 32; ++col) old[col] = ld_hint(p + col * BTM, pol);
#pragma unroll
            for (int col = 0; col < 32; ++col) st_hint(p + col * BTM, old[col] + __uint_as_float(v[col]), pol);
        } else {
#pragma unroll
            for (int col = 0; col < 32; ++col) st_hint(p + col * BTM, __uint_as_float(v[col]), pol);
        }
    }
}

// read-back interval containing chunk t: the accumulators are read back in front of the chunks ev0 + k * BFLUSH_CHUNKS and
// after the last chunk
__device__ __forceinline__ void s_interval(int t, int ev0, int nchunk, int &pos, int &len) {
    const int og = BFLUSH_CHUNKS - ev0;
    int start = ((t + og) / BFLUSH_CHUNKS) * BFLUSH_CHUNKS - og;
    const int next = start + BFLUSH_CHUNKS;
    start = max(start, 0);
    len = min(next, nchunk) - start;
    pos = t - start;
}

// Warp-specialised, no block barrier in the loop.  ONE ring of S_NTEAM = 3 slots; slot i = chunk t % 3 holds the chunk's T
// block, its folds and its a panels and belongs to producer team i.  Three mbarriers per slot:
//   fready[i]: the folds have landed (tx bytes)                                  copy warp  -> team i
//   ready[i] : the T block has landed (tx bytes) AND team i's 8 warps have arrived  copy warp + team i -> issuing warps
//   done[i]  : the MMAs of the slot's chunk have retired (tcgen05.commit of every active issuing warp)
//                                                                             issuing warps -> copy warp + team i
// so an issuing warp waits once and commits once per chunk (its serial per-chunk work is the kernel's critical path).
// Accumulator read-back every BFLUSH_CHUNKS chunks (see z3d_batched.cuh for why), the first one after s_first_event()
// chunks: the issuing warps commit the last chunk in front of an event onto `drained`; every producer warp, after the
// panels of its first chunk behind the boundary (they do not depend on the accumulators), waits for it, adds its 16-column
// blocks into the CTA's partial buffer and arrives on `flushed`, which the issuing warps wait for before that chunk's
// (overwriting) MMAs.
__global__ void __launch_bounds__(SNT, 1)
k_decompose_stream(const DevPlanS P, const float *__restrict__ Ft, int part, int nparts, float *__restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int set = blockIdx.x / P.nsplits, split = blockIdx.x % P.nsplits;
    StreamTabs &T = *reinterpret_cast<StreamTabs *>(smem + S_OFF_TABS);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + S_OFF_MBAR);
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem + S_OFF_MBAR + 8 * (3 * S_NTEAM + 2));
    int4 *seginfo = reinterpret_cast<int4 *>(smem + S_OFF_SEG);
    int *nblk = reinterpret_cast<int *>(smem + S_OFF_BLK), *blk = nblk + 2;
    // slot: [T block (t_bytes_max)][folds of the set's kind (S_FK_BYTES)][a panels (S_ABUF_BYTES)]
    unsigned char *slot0 = smem + S_OFF_A;
    const int slot_bytes = P.stage_bytes, t_bytes_max = P.t_bytes_max;
    for (int i = tid; i < (int)(sizeof(StreamTabs) / 4); i += SNT)
        reinterpret_cast<int *>(smem + S_OFF_TABS)[i] = reinterpret_cast<const int *>(P.tabs + set)[i];
    __syncthreads();
    const int ncols = T.ncols, nseg = T.nseg;
    const int nactive = min(nseg, S_NMMA);  // issuing warps that own a segment
    const unsigned fready0 = smem_u32(bars), ready0 = fready0 + 8 * S_NTEAM, done0 = ready0 + 8 * S_NTEAM,
                   drained = done0 + 8 * S_NTEAM, flushed = drained + 8;
    if (tid == 0) {
        for (int i = 0; i < S_NTEAM; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fready0 + 8 * i));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ready0 + 8 * i), "r"(1 + S_TEAM_WARPS));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(done0 + 8 * i), "r"(nactive));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(drained), "r"(nactive));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(flushed), "r"(S_NPW));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid < nseg) {  // per segment: instruction descriptor, a panel and T offsets in descriptor units (16 B), accumulator column
        const int n = T.seg_cols[tid];
        seginfo[tid] = make_int4((int)((1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(SM_ROWS >> 4) << 24)),
                                 T.seg_panel[tid] * (S_PANEL_BYTES >> 4), T.seg_off[tid], T.seg_off[tid]);
    }
    if (tid == 32) {   // read-back list
        int n0 = 0;
        for (int g = 0; g < nseg; ++g)
            for (int c = 0; c < T.seg_cols[g]; c += 16) blk[n0++] = T.seg_off[g] + c;
        nblk[0] = n0;
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
    float *out = partial + (size_t)blockIdx.x * BSLOT_WORDS;

    if (warp == S_WARP_TMA) {
        // ---- copy warp: T block (evict-first: streamed once) and the folds of the set's kind (evict-normal: every set of the
        //      kind reads them, within a short window)
        const unsigned t_bytes = (unsigned)ncols * 64u;
        const unsigned long long pol_t = l2_policy_evict_first(), pol_f = l2_policy_evict_normal();
        const float *tsrc = P.Tt + (size_t)cbeg * P.trec + T.toff;
        const float *fsrc = Ft + ((size_t)cbeg * 4 + T.kind[0]) * (S_FK_BYTES / 4);
        int s = 0;
        unsigned use = 0;  // how often slot s has been filled before
        for (int t = 0; t < nchunk; ++t) {
            if (use) mbar_wait(done0 + 8 * s, (use - 1) & 1);
            if (lane == 0) {
                unsigned char *st = slot0 + (size_t)s * slot_bytes;
                mbar_expect_tx(fready0 + 8 * s, S_FK_BYTES);
                bulk_copy_hint(st + t_bytes_max, fsrc, S_FK_BYTES, fready0 + 8 * s, pol_f);
                mbar_expect_tx(ready0 + 8 * s, t_bytes);
                bulk_copy_hint(st, tsrc, t_bytes, ready0 + 8 * s, pol_t);
            }
            __syncwarp();
            tsrc += P.trec;
            fsrc += 4 * S_FK_BYTES / 4;
            if (++s == S_NTEAM) { s = 0; ++use; }
        }
    } else if (warp < S_WARP_PROD0) {
        // ---- issuing warps: segments wi, wi + S_NMMA, ..; a warp without segments has nothing to do
        const int wi = warp - 1;
        if (wi < nactive) {
            int4 sg[S_SEG_PER_WARP];
#pragma unroll
            for (int q = 0; q < S_SEG_PER_WARP; ++q) {
                const int g = wi + q * S_NMMA;
                sg[q] = g < nseg ? seginfo[g] : make_int4(0, 0, 0, -1);
            }
            int s = 0, since = 0;   // since: chunks accumulated since the last read-back
            unsigned use = 0, nev = 0;
            // shared-memory descriptors advance with the ring: only the 14-bit start address field changes
            const unsigned long long dT0 = umma_desc(smem_u32(slot0), ncols * 16, 128);
            const unsigned long long dA0 = umma_desc(smem_u32(slot0 + t_bytes_max + S_FK_BYTES), SM_ROWS * 16, 128);
            unsigned long long dA = dA0, dT = dT0;
            const unsigned long long lo_a = (unsigned long long)(S_PANEL_BYTES / 2 >> 4), lo_t = (unsigned long long)(2 * ncols);
            const unsigned long long step = (unsigned long long)(slot_bytes >> 4);
            int to_event = s_first_event();  // chunks until the next read-back event
            for (int t = 0; t < nchunk; ++t) {
                if (to_event == 0) {  // the accumulators have been read back -> overwrite
                    mbar_wait(flushed, nev & 1);
                    since = 0;
                    ++nev;
                    to_event = BFLUSH_CHUNKS;
                }
                mbar_wait(ready0 + 8 * s, use & 1);
                tc_fence_after();
                // converged warp + elect.sync: the MMAs of the elected lane compile to back-to-back UTCHMMA on uniform registers
                // (an `if (lane == 0)` region costs an ELECT / R2UR / BRA.U.ANY loop per instruction: 64-72 clk per MMA for
                // any N against 23 clk or the tensor pipe's N / 2, scripts/experiments/tc_probe4.cu)
                if (elect_one()) {
#pragma unroll
                    for (int q = 0; q < S_SEG_PER_WARP; ++q) {
                        if (sg[q].w >= 0) {
                            const unsigned long long a = dA + (unsigned long long)sg[q].y, b = dT + (unsigned long long)sg[q].z;
                            const unsigned d = tmem + (unsigned)sg[q].w;
                            umma_tf32(d, a, b + lo_t, (unsigned)sg[q].x, since > 0 ? 1u : 0u);  // small terms first
                            umma_tf32(d, a + lo_a, b, (unsigned)sg[q].x, 1u);
                            umma_tf32(d, a, b, (unsigned)sg[q].x, 1u);
                        }
                    }
                    umma_commit(done0 + 8 * s);
                    if (to_event == 1 || t == nchunk - 1) umma_commit(drained);  // a read-back follows this chunk
                }
                __syncwarp();
                ++since;
                --to_event;
                dT += step;
                dA += step;
                if (++s == S_NTEAM) { s = 0; ++use; dT = dT0; dA = dA0; }
            }
        }
    } else {
        // ---- producer teams
        const int pw = warp - S_WARP_PROD0, team = pw / S_TEAM_WARPS;
        const int r = (pw % S_TEAM_WARPS) * 32 + lane, kh = r >> 7, row = r & 127, reim = row >> 6, b = row & 63;
        const int fofs = (reim * SNV + b) * KC + ((4 * kh) ^ (((b >> 2) & 1) << 2));  // the thread's 4 fold values (set A)
        const int oofs = (kh * SM_ROWS + row) * 4;
        const float sign = reim ? -1.0f : 1.0f;  // im rows are negated
        const int npanel = T.npanel;
        int tofs[S_PANELS];
#pragma unroll
        for (int j = 0; j < S_PANELS; ++j) tofs[j] = reim * (LMAX + 1) + (T.panel_code[j < npanel ? j : 0] & 0xffff);
        const float *Fs = reinterpret_cast<const float *>(slot0 + (size_t)team * slot_bytes + t_bytes_max);
        float *abuf = reinterpret_cast<float *>(slot0 + (size_t)team * slot_bytes + t_bytes_max + S_FK_BYTES);
        const unsigned my_fready = fready0 + 8 * team, my_ready = ready0 + 8 * team;
        unsigned use = 0;
        int kc = (int)((cbeg + team) % CPR);
        const float *trig = P.trig + (size_t)((cbeg + team) / CPR) * 4 * (LMAX + 1);
        bool had = false;  // the accumulators have been read back before: add to the partial sums
        const int ev0 = s_first_event();
        const int nregular = nchunk > ev0 ? (nchunk - 1 - ev0) / BFLUSH_CHUNKS + 1 : 0;  // events in front of chunks ev0 + k BFLUSH_CHUNKS < nchunk
        int nev = 0;
        for (int t = team; t < nchunk; t += S_NTEAM) {
            // truncation pre-compensation (BTRUNC_PER_MMA): the products of position pos of an interval of len chunks are
            // followed by 3 (len - 1 - pos) + 1 truncating accumulations
            int pos, len;
            s_interval(t, ev0, nchunk, pos, len);
            const float sg = sign * (1.0f + BTRUNC_PER_MMA * (float)(3 * (len - 1 - pos) + 1));
            float tA[S_PANELS], tB[S_PANELS];   // row trig of every panel, fetched before the waits
#pragma unroll
            for (int j = 0; j < S_PANELS; ++j) {
                tA[j] = __ldg(trig + tofs[j]) * sg;
                tB[j] = __ldg(trig + tofs[j] + 2 * (LMAX + 1)) * sg;
            }
            mbar_wait_relaxed(my_fready, use & 1);                    // the folds have landed
            const float4 fa = *reinterpret_cast<const float4 *>(Fs + fofs), fb = *reinterpret_cast<const float4 *>(Fs + fofs + 2 * SNV * KC);
            // (the slot's previous MMAs retired before the copy warp refilled it: the a panels are free)
#pragma unroll
            for (int j = 0; j < S_PANELS; ++j)
                if (j < npanel) {
                    float4 hi, lo;
                    split_tf32(fmaf(fa.x, tA[j], fb.x * tB[j]), hi.x, lo.x);
                    split_tf32(fmaf(fa.y, tA[j], fb.y * tB[j]), hi.y, lo.y);
                    split_tf32(fmaf(fa.z, tA[j], fb.z * tB[j]), hi.z, lo.z);
                    split_tf32(fmaf(fa.w, tA[j], fb.w * tB[j]), hi.w, lo.w);
                    float *o = abuf + oofs + j * (S_PANEL_BYTES / 4);
                    *reinterpret_cast<float4 *>(o) = hi;
                    *reinterpret_cast<float4 *>(o + 2 * SM_ROWS * 4) = lo;
                }
            fence_async_smem();  // generic-proxy panel writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(my_ready);
            // read-back events in front of this chunk come AFTER its panels: the a panels do not depend on the accumulators, and
            // the issuing warps find them ready when the read-back is done
            while (nev < nregular && ev0 + nev * BFLUSH_CHUNKS <= t) {
                s_flush(tmem, ncols, out, had, drained, nev & 1);
                had = true;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(flushed);
                ++nev;
            }
            ++use;
            kc += S_NTEAM;
            while (kc >= CPR) { kc -= CPR; trig += 4 * (LMAX + 1); }
        }
        if (nchunk) {  // remaining events, then the read-back after the final chunk
            while (nev < nregular) {
                s_flush(tmem, ncols, out, had, drained, nev & 1);
                had = true;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(flushed);
                ++nev;
            }
            s_flush(tmem, ncols, out, had, drained, nev & 1);
        } else {
            for (int i = tid - S_WARP_PROD0 * 32; i < ncols * BTM; i += S_NPROD) out[i] = 0.0f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// moments[b][mu][reim] = scale[mu] * sum over the splits of the owning accumulator (fixed order, float64)
__global__ void k_reduce_stream(const float *__restrict__ partial, const int2 *__restrict__ colmap /* [nmom] (set, column) */,
                                const double *__restrict__ scale, int nmom, int nsplits, int nvol, float *__restrict__ moments) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over nmom * 128
    if (idx >= nmom * SM_ROWS) return;
    const int mu = idx / SM_ROWS, row = idx % SM_ROWS, reim = row / SNV, b = row % SNV;
    if (b >= nvol) return;
    const int2 sc = colmap[mu];
    const float *p = partial + (size_t)sc.x * nsplits * BSLOT_WORDS + (size_t)sc.y * BTM + row;
    double sum = 0.0;
    for (int s = 0; s < nsplits; ++s) sum += (double)p[(size_t)s * BSLOT_WORDS];
    moments[((size_t)b * nmom + mu) * 2 + reim] = (float)(sum * scale[mu]);
}

}  // namespace z3d
