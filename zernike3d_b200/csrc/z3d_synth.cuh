// zernike3d_b200 - batched SYNTHESIS on the tensor cores (maps from moments for tiles of 64 moment sets), table-free.
//
// The exact adjoint of z3d_gen.cuh.  The reference's sum (zm:2596-2614)
//     vol(v) = sum_{n,l,m>=0} w_m kappa_lm T_nlm(v) (Re O_nlm cos m phi_v - Im O_nlm sin m phi_v),   T = R_nl Q_lm,  w_0 = 1, w_m = 2,
// is evaluated on the 1/16 of the voxels the grid's mirror / transpose symmetries leave: for a folded voxel k of a row group
//     G_X[kind][re|im](k) = sum over the (m, pz) panels j of the kind of   trig_X,j[re|im] * out_j[re|im](k),     X = row group A or its transpose B,
//     out_j[(re|im, b)](k) = sum over the rows of panel j of   O'[(re|im, b)][row] * T[row](k),     O'[re] = w kappa Re O,  O'[im] = - w kappa Im O,
// and the 8 mirror images of k are the 8-point Walsh-Hadamard transform of the 8 class values (kind, re|im) - k_fold_tile_gen run backwards.
// out_j is a GEMM over the moment rows: tcgen05.mma kind::tf32 with M = 128 = (re|im) x 64 maps, K = 8 rows per instruction, B = the T block in
// shared memory, MN-major.  A pipeline step is a PAIR of chunks: a T row of 128 bytes holds [T_hi A | T_hi B | T_lo A | T_lo B] (8 voxels each),
// so that one N = 32 MMA with A = O'_hi yields O'_hi T_hi and O'_hi T_lo of both chunks and one N = 16 MMA with A = O'_lo adds O'_lo T_hi
// (3xTF32: the three are summed in the epilogue).
//   * O' of the CTA's column set stays in TENSOR memory for the whole kernel (columns [0, ncols) hi, [ncols, 2 ncols) lo; the MMA's A operand),
//   * T is generated exactly as in k_decompose_gen (seed warps, T teams, the same plan tables: the two-tile layout of the analysis kernel,
//     whose sets hold <= 224 columns),
//   * the accumulators out_j (48 columns per panel, one bank, released to the issuing warp panel by panel) are read by two epilogue teams
//     (thread = (re|im, map) = tensor-memory lane), multiplied by the row trig of the chunk and summed over the set's panels,
//   * CTAs run in thread-block clusters of two column sets: where both are of the same kind the odd CTA sends its class partials to the even
//     one through distributed shared memory and only the sum is stored; the stored partials of one kind add up in k_unfold_tile, which also
//     applies the Walsh-Hadamard transform and writes the 16 images.  No read-back interval: an accumulator lives for one chunk pair.
// Roles (28 warps): warp 0 issues, warps 1-3 seeds, warps 4-11 two epilogue teams, warps 12-27 four T teams.
#pragma once
#include "z3d_gen.cuh"

namespace z3d {

__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float (&v)[8]) {
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// O' of a tile of maps in the layout the synthesis kernel loads into tensor memory: Op[set][row = (re|im, map)][hi: maxcols | lo: maxcols]
// (zero for padding columns and for maps beyond nvol)
__global__ void k_prepare_synth(const float *__restrict__ moments, const int *__restrict__ colsrc /* [nsets][maxcols] -> mu or -1 */,
                                const double *__restrict__ ofac, int maxcols, int nmom, int nvol, float *__restrict__ Op) {
    const int set = blockIdx.x, row = blockIdx.y, reim = row / SNV, b = row % SNV;
    float *dst = Op + ((size_t)set * SM_ROWS + row) * (2 * maxcols);
    for (int c = threadIdx.x; c < maxcols; c += blockDim.x) {
        const int mu = colsrc[(size_t)set * maxcols + c];
        float v = 0.0f;
        if (mu >= 0 && b < nvol) {
            const double x = (double)moments[((size_t)b * nmom + mu) * 2 + reim] * ofac[mu];
            v = (float)(reim ? -x : x);
        }
        float hi, lo;
        split_fast(v, hi, lo);
        dst[c] = hi;
        dst[maxcols + c] = lo;
    }
}


// ---- thread-block cluster helpers (pairs of CTAs that hold two column sets of the same kind add their class partials through
// distributed shared memory before they are stored)
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned mapa_shared(unsigned addr, unsigned rank) {   // the same shared-memory offset in CTA `rank` of the cluster
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(unsigned raddr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(unsigned raddr) {   // arrive on an mbarrier of another CTA of the cluster, releasing this thread's writes
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned mbar, unsigned parity) {   // local mbarrier, acquire at cluster scope (remote writes become visible)
#pragma unroll 1
    for (int it = 0; it < (1 << 24); ++it) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (ok) return;
        __nanosleep(64);
    }
#if defined(Z3D_SYNTH_DEBUG) || defined(Z3D_WAIT_DIAG)
    if ((threadIdx.x & 31) == 0) printf("cluster wait timed out: block (%d,%d) warp %d barrier @%u parity %u\n", blockIdx.x, blockIdx.y, threadIdx.x >> 5, mbar, parity);
#endif
    __trap();
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr int SY_XBUF_BYTES = 2 * SM_ROWS * 32 * 4;   // exchange buffers of the two epilogue teams: [team][8 x float4][row]

// Partial folds of the launch's chunk range: Fs[set][chunk - gbeg][row group A|B][row = (re|im, map)][k]
//
// A pipeline STEP is a PAIR of chunks (A = 2s, B = 2s + 1 of the CTA's chunk range): a T row of 128 bytes holds [T_hi A | T_hi B | T_lo A |
// T_lo B] (8 voxels each), so that per K-step of 8 moment rows ONE N = 32 MMA with O'_hi and ONE N = 16 MMA with O'_lo (-> T_hi A, T_hi B)
// serve both chunks: half the MMAs and half the hand-overs per voxel of a chunk-per-step pipeline (the MMAs are issue bound at N <= 32).
// Accumulators: 48 tensor-memory columns per panel, ONE bank - the issuing warp waits per panel until the epilogue team of the previous
// pair has read that panel (`dfree[j]`), so it re-fills panel 0 while the later panels are still being read.
constexpr int SY_DG = 24;          // depth of the geometry / seed ring: a common multiple of the agents' chunk strides (3 seed warps, 4 T teams x 2
                                   // chunks, 2 epilogue teams x 2 chunks), see the note on parity waits in z3d_gen.cuh
__host__ __device__ constexpr int sy_gring_bytes(int gslot_bytes) { return SY_DG * gslot_bytes; }
__global__ void __launch_bounds__(GNT, 1)
k_synth_gen(const __grid_constant__ DevPlanG P, const float *__restrict__ Op, int maxcols, int off_ts, int NS, int part, int nparts,
            const int *__restrict__ pairinfo /* [sets]: slot << 1 | exchange with the cluster partner */, float *__restrict__ Fs) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int set = blockIdx.y, split = blockIdx.x;
    if (set >= P.nsets) return;   // (padding CTA of an odd number of sets: its cluster partner does not exchange)
    const int pinfo = pairinfo[set], slot_id = pinfo >> 1;
    const bool exch = pinfo & 1;
    const unsigned crank = cluster_ctarank();   // 0: adds the partner's class partials and stores the sum; 1: sends its own
    GenTabs &T = *reinterpret_cast<GenTabs *>(smem + G_OFF_TABS);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + P.off_g + sy_gring_bytes(P.gslot_bytes));   // 1 KB behind the seed ring
    float *xbuf = reinterpret_cast<float *>(smem + P.off_g + sy_gring_bytes(P.gslot_bytes) + 1024);   // exchange buffers, SY_XBUF_BYTES
    float2 *gds = reinterpret_cast<float2 *>(smem + G_OFF_GD);
    float4 *rks = reinterpret_cast<float4 *>(smem + P.off_rk);
    // (the analysis kernel's fold rings are not needed: the seed ring is twice as deep and the T ring follows it, on a 1024-byte boundary
    // of the shared-memory WINDOW - the swizzle of the MN-major layout works on absolute address bits)
    off_ts = (int)(((smem_u32(smem) + (unsigned)off_ts + 1023u) & ~1023u) - smem_u32(smem));
    unsigned char *gring = smem + P.off_g, *tring = smem + off_ts;
    const int slot_bytes = 2 * P.tslot_bytes;   // 128 B per row
    // barriers: gfull[SY_DG] gfree[SY_DG] tfull[NS_MAX] tfree[NS_MAX] dfree[G_PANELS] dfull[2] oready  (dfull per epilogue team: a parity wait
    // is only meaningful on a barrier whose every phase the waiter consumes)
    const unsigned gfull0 = smem_u32(bars), gfree0 = gfull0 + 8 * SY_DG, tfull0 = gfree0 + 8 * SY_DG, tfree0 = tfull0 + 8 * G_NS_MAX,
                   dfree0 = tfree0 + 8 * G_NS_MAX, dfull0 = dfree0 + 8 * G_PANELS, oready = dfull0 + 16, xfull0 = oready + 8, xfree0 = xfull0 + 16;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(bars + (2 * SY_DG + 2 * G_NS_MAX + G_PANELS + 7));
    static_assert(8 * (2 * SY_DG + 2 * G_NS_MAX + G_PANELS + 7) + 4 <= 1024, "barrier block");
    for (int i = tid; i < (int)(sizeof(GenTabs) / 4); i += GNT)
        reinterpret_cast<int *>(smem + G_OFF_TABS)[i] = reinterpret_cast<const int *>(P.tabs + set)[i];
    for (int i = tid; i < P.gd_bytes / 8; i += GNT) gds[i] = P.gd[(size_t)set * (P.gd_bytes / 8) + i];
    // rows of a panel are padded to 16: the padding rows of the T ring are never written by a chain and multiply O' = 0 - they must be finite
    for (int i = tid; i < NS * slot_bytes / 16; i += GNT) reinterpret_cast<float4 *>(tring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int ncols = T.ncols, nseg = T.nseg, npanel = T.npanel;
#if defined(Z3D_SYNTH_DEBUG) || defined(Z3D_WAIT_DIAG)
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0)
        printf("synth: bars @%u gfree0 %u tfull0 %u tfree0 %u dfree0 %u dfull0 %u oready %u | ncols %d nseg %d npanel %d NS %d\n", gfull0, gfree0, tfull0, tfree0, dfree0,
               dfull0, oready, T.ncols, T.nseg, T.npanel, NS);
#endif
    for (int i = tid; i < P.rk_bytes / 16; i += GNT) rks[i] = P.rk[(size_t)T.lpar * (P.rk_bytes / 16) + i];
    if (tid == 0) {
        for (int i = 0; i < SY_DG; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gfull0 + 8 * i));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(gfree0 + 8 * i));   // 4 warps of a T team + 4 of an epilogue team
        }
        for (int i = 0; i < G_NS_MAX; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(tfull0 + 8 * i));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tfree0 + 8 * i));
        }
        for (int i = 0; i < G_PANELS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(dfree0 + 8 * i));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dfull0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dfull0 + 8));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(oready));
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(xfull0 + 8 * i));   // the partner's epilogue team (4 warps) has sent a pair
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(xfree0 + 8 * i));   // the partner's epilogue team has read the buffer
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_async_smem();   // the zeroed T ring is read by the tensor core's async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (exch) cluster_sync_all();   // the partner's barriers exist before anything is sent (both CTAs of the pair take this branch)
    const unsigned tmem = *tmem_slot;
    const unsigned dcol0 = 512u - 48u * (unsigned)npanel;   // accumulators of panel j: dcol0 + 48 j (+ 32: the O'_lo product, 16 columns)

    const int CPR = (P.H + KC - 1) / KC;
    const long long total = (long long)analysis_items(0, P.H, P.H) * CPR;
    const long long gbeg = total * part / nparts, glen = total * (part + 1) / nparts - gbeg;
    const long long cbeg = gbeg + glen * split / P.nsplits, cend = gbeg + glen * (split + 1) / P.nsplits;
    const int nchunk = (int)(cend - cbeg);
    const int npair = (nchunk + 1) / 2;

    if (nchunk == 0) {
        // nothing to do
    } else if (warp == 0) {
        // ---- issuing warp
        const unsigned dlo0 = (smem_u32(smem) + (unsigned)off_ts) >> 4;
        // MN-major B for 32-bit operands exists in ONE shared-memory layout (scripts/experiments/tc_probe6.cu: every other layout type
        // makes the tensor core read zeros): SWIZZLE_128B_BASE32B - rows of 128 B (32 elements along MN), 4 rows = one 512-byte atom,
        // the 32-byte unit c of row k stored at unit c ^ (k % 4), atoms along K `SBO` apart; an MMA reads 2 atoms = 8 rows.
        const unsigned long long dhi = ((unsigned long long)(512 >> 4) << 32) | (1ull << 46) | ((unsigned long long)(512 >> 4) << 16) | (1ull << 61);
        const unsigned idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((unsigned)(SM_ROWS >> 4) << 24);
        const unsigned idesc32 = idesc0 | ((32u >> 3) << 17), idesc16 = idesc0 | ((16u >> 3) << 17);
        const unsigned step = (unsigned)(slot_bytes >> 4);
        int ts = 0;
        unsigned tuse = 0;
        mbar_wait(oready, 0);
        for (int s = 0; s < npair; ++s) {
            mbar_wait(tfull0 + 8 * ts, tuse & 1);
            tc_fence_after();
            if (elect_one()) {
                const unsigned dl = dlo0 + (unsigned)ts * step;
                int prev_panel = -1;
#ifdef Z3D_SYNTH_DEBUG
                if (!(P.dbg & 1))
#endif
                for (int g = 0; g < nseg; ++g) {
                    const int j = T.seg_panel[g], off = T.seg_off[g], cols = T.seg_cols[g];
                    const unsigned d = tmem + dcol0 + (unsigned)(j * 48);
                    const bool fresh = j != prev_panel;
                    prev_panel = j;
                    if (fresh && s > 0) {   // the epilogue team of pair s - 1 has read this panel
                        mbar_wait(dfree0 + 8 * j, (unsigned)(s - 1) & 1);
                        tc_fence_after();
                    }
                    for (int ks = 0; ks < cols; ks += 8) {
                        const unsigned r = (unsigned)(off + ks);
                        const unsigned long long b = dhi | (unsigned long long)(dl + r * 8u);   // row r -> atom r / 4 -> 512 B each
                        const unsigned acc = (fresh && ks == 0) ? 0u : 1u;
                        umma_tf32_ts(d, tmem + r, b, idesc32, acc);                              // O'_hi x [T_hi A | T_hi B | T_lo A | T_lo B]
                        umma_tf32_ts(d + 32, tmem + (unsigned)ncols + r, b, idesc16, acc);       // O'_lo x [T_hi A | T_hi B]
                    }
                }
#ifdef Z3D_SYNTH_DEBUG
                if ((P.dbg & 1) && s > 0) for (int j = 0; j < npanel; ++j) mbar_wait(dfree0 + 8 * j, (unsigned)(s - 1) & 1);
#endif
                umma_commit(tfree0 + 8 * ts);
                umma_commit(dfull0 + 8 * (s & 1));
            }
            __syncwarp();
            if (++ts == NS) { ts = 0; ++tuse; }
        }
    } else if (warp < G_WARP_SEED0 + G_NSEEDW) {
        // ---- seed warps: as in k_decompose_gen (lane = (panel j = lane / 8 (+ 4 in the second pass), voxel k = lane % 8)), one chunk at a time
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
            if (item != item_have) {
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
                    for (int e = m; e; e >>= 1) {
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
                    const int mq = m & 3;
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
            const int gs = t % SY_DG;
            if (t >= SY_DG) g_wait_relaxed(gfree0 + 8 * gs, (unsigned)(t / SY_DG - 1) & 1);
            float *slot = reinterpret_cast<float *>(gring + (size_t)gs * P.gslot_bytes);
            float *seeds = slot + 40;
#ifdef Z3D_SYNTH_DEBUG
            if (P.dbg & 8) { __syncwarp(); if (lane == 0) mbar_arrive(gfull0 + 8 * gs); continue; }
#endif
            if (lane < 8) slot[lane] = r2;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q < npass) {
                    const int j = (lane >> 3) + 4 * q;
                    const float tv = k == 0 ? tr[q][0] : k == 1 ? tr[q][1] : k == 2 ? tr[q][2] : tr[q][3];
                    if (k < 4 && j < npanel) slot[8 + j * 4 + k] = tv;
                    float S1 = (float)(ppz[q] ? 2.0 * z * rhom[q] : rhom[q]), S2 = 0.0f;
                    float *sp = seeds + psoff[q] * 8 + k;
                    const float2 *gp = gds + pgd[q];
                    const int nl = pnl[q];
                    if (nl > 0) sp[0] = S1;
                    for (int i0 = 1; i0 < maxnl; i0 += 8) {
                        float2 g[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) g[e] = gp[i0 - 1 + e];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float S = fmaf(fmaf(-g[e].x, r2, z4), S1, (g[e].y * r4) * S2);
                            S2 = S1;
                            S1 = S;
                            if (i0 + e < nl) sp[(i0 + e) * 8] = S;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(gfull0 + 8 * gs);
        }
    } else if (warp < G_WARP_A0) {
        // (no role)
    } else if (warp < G_WARP_T0) {
        // ---- epilogue teams: thread = row (re|im, map b) = tensor-memory lane; pair s -> team s & 1
        const int team = (warp - G_WARP_A0) >> 2, quarter = warp & 3;
        const int row = quarter * 32 + lane, reim = row >> 6;
        const unsigned lane0 = tmem + ((unsigned)(32 * quarter) << 16);
        if (team == 0) {   // O' of the set -> tensor memory, once
            const float *src = Op + ((size_t)set * SM_ROWS + row) * (2 * maxcols);
            for (int c = 0; c < ncols; c += 8) {
                float hi[8], lo[8];
                const float4 h0 = *reinterpret_cast<const float4 *>(src + c), h1 = *reinterpret_cast<const float4 *>(src + c + 4);
                const float4 l0 = *reinterpret_cast<const float4 *>(src + maxcols + c), l1 = *reinterpret_cast<const float4 *>(src + maxcols + c + 4);
                hi[0] = h0.x; hi[1] = h0.y; hi[2] = h0.z; hi[3] = h0.w; hi[4] = h1.x; hi[5] = h1.y; hi[6] = h1.z; hi[7] = h1.w;
                lo[0] = l0.x; lo[1] = l0.y; lo[2] = l0.z; lo[3] = l0.w; lo[4] = l1.x; lo[5] = l1.y; lo[6] = l1.z; lo[7] = l1.w;
                tmem_st8(lane0 + (unsigned)c, hi);
                tmem_st8(lane0 + (unsigned)(ncols + c), lo);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(oready);
        }
        for (int s = team; s < npair; s += 2) {
            const int t0 = 2 * s;
            const bool hasB = t0 + 1 < nchunk;
            g_wait_relaxed(dfull0 + 8 * team, (unsigned)(s >> 1) & 1);
            tc_fence_after();
            // panel by panel, both chunks of the pair: a panel's accumulators are released to the issuing warp as soon as they have
            // been read (the next pair's MMAs of that panel wait for exactly this), the class sums of the two chunks stay in registers
            const int gsA = t0 % SY_DG, gsB = (t0 + 1) % SY_DG;
            const float *slotA = reinterpret_cast<const float *>(gring + (size_t)gsA * P.gslot_bytes);
            const float *slotB = reinterpret_cast<const float *>(gring + (size_t)gsB * P.gslot_bytes);
            float FA0[8], FB0[8], FA1[8], FB1[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) FA0[v] = FB0[v] = FA1[v] = FB1[v] = 0.0f;
#ifdef Z3D_SYNTH_DEBUG
            if (!(P.dbg & 4))
#endif
            for (int j = 0; j < npanel; ++j) {
                const unsigned ta = lane0 + dcol0 + (unsigned)(j * 48);
                float dh[8], dl[8], e[8];
                tmem_ld8(ta, dh);            // O'_hi T_hi, chunk A
                tmem_ld8(ta + 16, dl);       // O'_hi T_lo
                tmem_ld8(ta + 32, e);        // O'_lo T_hi
                tmem_ld_wait();
                {
                    const float tA = slotA[8 + j * 4 + reim], tB = slotA[8 + j * 4 + 2 + reim];
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const float o = (dl[v] + e[v]) + dh[v];   // small terms first
                        FA0[v] = fmaf(tA, o, FA0[v]);
                        FB0[v] = fmaf(tB, o, FB0[v]);
                    }
                }
                if (hasB) {
                    tmem_ld8(ta + 8, dh);        // chunk B
                    tmem_ld8(ta + 24, dl);
                    tmem_ld8(ta + 40, e);
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dfree0 + 8 * j);
                if (hasB) {
                    const float tA = slotB[8 + j * 4 + reim], tB = slotB[8 + j * 4 + 2 + reim];
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const float o = (dl[v] + e[v]) + dh[v];
                        FA1[v] = fmaf(tA, o, FA1[v]);
                        FB1[v] = fmaf(tB, o, FB1[v]);
                    }
                }
            }
#ifdef Z3D_SYNTH_DEBUG
            if (P.dbg & 4) { tc_fence_before(); __syncwarp(); if (lane == 0) for (int j = 0; j < npanel; ++j) mbar_arrive(dfree0 + 8 * j); }
#endif
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(gfree0 + 8 * gsA);
                if (hasB) mbar_arrive(gfree0 + 8 * gsB);
            }
            if (exch) {
                const int i = s >> 1;   // this team's i-th pair
                float *xb = xbuf + (size_t)team * (SM_ROWS * 32) + (size_t)row * 4;   // [8 groups of 4 floats][row]
                if (crank == 1) {   // send: wait until the partner has read this team's previous pair, write into ITS buffer, arrive on ITS barrier
                    if (i > 0) mbar_wait_cluster(xfree0 + 8 * team, (unsigned)(i - 1) & 1);
                    const unsigned rb = mapa_shared(smem_u32(xb), 0);
                    st_cluster_v4(rb + 0 * SM_ROWS * 16, FA0[0], FA0[1], FA0[2], FA0[3]);
                    st_cluster_v4(rb + 1 * SM_ROWS * 16, FA0[4], FA0[5], FA0[6], FA0[7]);
                    st_cluster_v4(rb + 2 * SM_ROWS * 16, FB0[0], FB0[1], FB0[2], FB0[3]);
                    st_cluster_v4(rb + 3 * SM_ROWS * 16, FB0[4], FB0[5], FB0[6], FB0[7]);
                    st_cluster_v4(rb + 4 * SM_ROWS * 16, FA1[0], FA1[1], FA1[2], FA1[3]);
                    st_cluster_v4(rb + 5 * SM_ROWS * 16, FA1[4], FA1[5], FA1[6], FA1[7]);
                    st_cluster_v4(rb + 6 * SM_ROWS * 16, FB1[0], FB1[1], FB1[2], FB1[3]);
                    st_cluster_v4(rb + 7 * SM_ROWS * 16, FB1[4], FB1[5], FB1[6], FB1[7]);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_shared(xfull0 + 8 * team, 0));
                    continue;   // nothing to store: the partner stores the sum
                }
                mbar_wait_cluster(xfull0 + 8 * team, (unsigned)i & 1);
                {
                    const float4 a0 = *reinterpret_cast<const float4 *>(xb + 0 * SM_ROWS * 4), a1 = *reinterpret_cast<const float4 *>(xb + 1 * SM_ROWS * 4),
                                 b0 = *reinterpret_cast<const float4 *>(xb + 2 * SM_ROWS * 4), b1 = *reinterpret_cast<const float4 *>(xb + 3 * SM_ROWS * 4);
                    FA0[0] += a0.x; FA0[1] += a0.y; FA0[2] += a0.z; FA0[3] += a0.w; FA0[4] += a1.x; FA0[5] += a1.y; FA0[6] += a1.z; FA0[7] += a1.w;
                    FB0[0] += b0.x; FB0[1] += b0.y; FB0[2] += b0.z; FB0[3] += b0.w; FB0[4] += b1.x; FB0[5] += b1.y; FB0[6] += b1.z; FB0[7] += b1.w;
                }
                {
                    const float4 a0 = *reinterpret_cast<const float4 *>(xb + 4 * SM_ROWS * 4), a1 = *reinterpret_cast<const float4 *>(xb + 5 * SM_ROWS * 4),
                                 b0 = *reinterpret_cast<const float4 *>(xb + 6 * SM_ROWS * 4), b1 = *reinterpret_cast<const float4 *>(xb + 7 * SM_ROWS * 4);
                    FA1[0] += a0.x; FA1[1] += a0.y; FA1[2] += a0.z; FA1[3] += a0.w; FA1[4] += a1.x; FA1[5] += a1.y; FA1[6] += a1.z; FA1[7] += a1.w;
                    FB1[0] += b0.x; FB1[1] += b0.y; FB1[2] += b0.z; FB1[3] += b0.w; FB1[4] += b1.x; FB1[5] += b1.y; FB1[6] += b1.z; FB1[7] += b1.w;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_shared(xfree0 + 8 * team, 1));   // the buffer may be refilled
            }
            float *dst = Fs + (((size_t)slot_id * glen + (size_t)(cbeg + t0 - gbeg)) * 2) * (SM_ROWS * KC) + (size_t)row * KC;
            *reinterpret_cast<float4 *>(dst) = make_float4(FA0[0], FA0[1], FA0[2], FA0[3]);
            *reinterpret_cast<float4 *>(dst + 4) = make_float4(FA0[4], FA0[5], FA0[6], FA0[7]);
            *reinterpret_cast<float4 *>(dst + SM_ROWS * KC) = make_float4(FB0[0], FB0[1], FB0[2], FB0[3]);
            *reinterpret_cast<float4 *>(dst + SM_ROWS * KC + 4) = make_float4(FB0[4], FB0[5], FB0[6], FB0[7]);
            if (hasB) {
                dst += 2 * SM_ROWS * KC;   // the next chunk
                *reinterpret_cast<float4 *>(dst) = make_float4(FA1[0], FA1[1], FA1[2], FA1[3]);
                *reinterpret_cast<float4 *>(dst + 4) = make_float4(FA1[4], FA1[5], FA1[6], FA1[7]);
                *reinterpret_cast<float4 *>(dst + SM_ROWS * KC) = make_float4(FB1[0], FB1[1], FB1[2], FB1[3]);
                *reinterpret_cast<float4 *>(dst + SM_ROWS * KC + 4) = make_float4(FB1[4], FB1[5], FB1[6], FB1[7]);
            }
        }
    } else {
        // ---- T teams: the recurrences of k_decompose_gen, for the two chunks of a pair; pair s -> team s & 3 -> T slot s % NS
        const int team = (warp - G_WARP_T0) >> 2, u = ((warp - team + 1) & 3) * 32 + lane;
        int4 task[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = u + 128 * q;
            task[q] = i < T.ntask ? P.tasks[(size_t)set * G_TASKS + i] : make_int4(0, 0, 0, 0);
        }
        for (int s = team; s < npair; s += G_NTT) {
            const int ts = s % NS;
            const unsigned tuse = (unsigned)(s / NS);
            const int t0 = 2 * s;
            const bool hasB = t0 + 1 < nchunk;
            unsigned char *tb = tring + (size_t)ts * slot_bytes;
            g_wait_relaxed(gfull0 + 8 * (t0 % SY_DG), (unsigned)(t0 / SY_DG) & 1);
            if (hasB) g_wait_relaxed(gfull0 + 8 * ((t0 + 1) % SY_DG), (unsigned)((t0 + 1) / SY_DG) & 1);
            if (tuse) g_wait_relaxed(tfree0 + 8 * ts, (tuse - 1) & 1);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                if (half == 1 && !hasB) break;   // (the B units keep an earlier pair's finite values; the epilogue ignores those columns)
                const float *slot = reinterpret_cast<const float *>(gring + (size_t)((t0 + half) % SY_DG) * P.gslot_bytes);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
#ifdef Z3D_SYNTH_DEBUG
                    const int len = (P.dbg & 2) ? 0 : task[q].y;
#else
                    const int len = task[q].y;
#endif
                    if (len > 0) {
                        const int kh = task[q].w >> 16, srow = task[q].w & 0xffff;
                        const float4 s4 = *reinterpret_cast<const float4 *>(slot + 40 + srow * 8 + 4 * kh);
                        const float4 r2 = *reinterpret_cast<const float4 *>(slot + 4 * kh);
                        const float4 *K = rks + task[q].z;
                        const int row0 = task[q].x;
                        float4 T1 = make_float4(0.f, 0.f, 0.f, 0.f), T2 = s4;
                        float4 kc4 = K[0];
#pragma unroll 2
                        for (int i = 0; i < len; ++i) {
                            const float4 kn = K[i + 1];
                            float4 Tn, h, l;
                            Tn.x = fmaf(fmaf(kc4.x, r2.x, kc4.y), T1.x, kc4.z * T2.x);
                            Tn.y = fmaf(fmaf(kc4.x, r2.y, kc4.y), T1.y, kc4.z * T2.y);
                            Tn.z = fmaf(fmaf(kc4.x, r2.z, kc4.y), T1.z, kc4.z * T2.z);
                            Tn.w = fmaf(fmaf(kc4.x, r2.w, kc4.y), T1.w, kc4.z * T2.w);
                            split4_fast(Tn, h, l);
                            const int r = row0 + i, k4 = r & 3;
                            unsigned char *rowp = tb + (size_t)r * 128 + kh * 16;
                            *reinterpret_cast<float4 *>(rowp + ((half ^ k4) << 5)) = h;          // unit (T_hi A | T_hi B) ^ k4
                            *reinterpret_cast<float4 *>(rowp + (((2 + half) ^ k4) << 5)) = l;    // unit (T_lo A | T_lo B) ^ k4
                            T2 = T1;
                            T1 = Tn;
                            kc4 = kn;
                        }
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(tfull0 + 8 * ts);
                mbar_arrive(gfree0 + 8 * (t0 % SY_DG));
                if (hasB) mbar_arrive(gfree0 + 8 * ((t0 + 1) % SY_DG));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    if (exch) cluster_sync_all();   // neither CTA of the pair leaves while the other may still write into its shared memory
}

// Sum of the partial folds over the column sets of every kind, 8-point Walsh-Hadamard transform of the 8 class values, the 8 mirror
// images of both row groups -> the maps.  One CTA per chunk of the launch's range; thread = (row group A|B, map, k).
struct KindSets { int first[5]; };
__global__ void __launch_bounds__(2 * SNV * KC) k_unfold_tile(const float *__restrict__ Fs, long long gbeg, long long glen, KindSets ks,
                                                              int nvol, int N, int H, float *__restrict__ vol) {
    const int CPR = (H + KC - 1) / KC;
    const long long rel = blockIdx.x, chunk = gbeg + rel;
    ChunkPos p;
    chunk_pos(chunk, CPR, H, p);
    const int tid = threadIdx.x;
    const int k = tid % KC, b = (tid / KC) % SNV, set = tid / (KC * SNV);
    const int kk = p.kc * KC + k, k2 = N - 1 - kk;
    const int ri = set ? p.jp : p.ip, rj = set ? p.ip : p.jp;
    const int i2 = N - 1 - ri, j2 = N - 1 - rj;
    const bool live = b < nvol && kk < H && (set == 0 || p.paired);
    if (!live) return;
    float v[8];
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
        const int reim = c8 >> 2, kind = ((((c8 >> 1) & 1) ^ reim) << 1) | (c8 & 1);
        float s = 0.0f;
        for (int q = ks.first[kind]; q < ks.first[kind + 1]; ++q)
            s += Fs[((((size_t)q * glen + rel) * 2 + set) * SM_ROWS + (reim * SNV + b)) * KC + k];
        v[c8] = s;
    }
    wht8(v);
    float *vb = vol + (size_t)b * N * N * N;
#pragma unroll
    for (int img = 0; img < 8; ++img) {
        const bool sy = img & 4, sx = img & 2, sz = img & 1;
        const bool dup = (sy && i2 == ri) || (sx && j2 == rj) || (sz && k2 == kk);
        if (!dup) vb[((size_t)(sy ? i2 : ri) * N + (sx ? j2 : rj)) * N + (sz ? k2 : kk)] = v[img];
    }
}

}  // namespace z3d
