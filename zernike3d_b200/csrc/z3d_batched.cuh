// zernike3d_b200 - batched analysis kernel: a tile of up to 32 volumes per launch.
//
// For ONE volume the per-l contraction needs the per-volume panel a*P (a = folded volume x row trig), so nothing but the
// R panel is shared between volumes (k_decompose).  For a BATCH the volume-independent product
//     T[(n,l,m)][k] = R_nl(r_k) * Q_lm(theta_k)          (Q = P_lm / kappa_lm, one value per moment and folded voxel)
// is generated once and the analysis becomes, per (m, z-parity pz = (l+m)&1), a real GEMM over the voxel axis
//     D[(n,l)][(re/im, b)] += T[(n,l,m)][k] * a[(m,pz)][k][(re/im, b)],      a = F_A,b trig_A + F_B,b trig_B,
// i.e. exactly "[n x V] . [V x m], batched across volumes" with the batch as the GEMM's N dimension (64 columns for 32
// volumes).  Rows of one (m, pz) are padded to 8 (97 % efficiency instead of the 65 % of the ragged per-l tiles), and the
// recurrences, the geometry and the T products are amortised over the 32 volumes.
//
// Work decomposition: the 12 051 rows (order 50) are cut into `ngroups` row groups of <= 78 blocks of 8 rows; a CTA owns one
// row group (its 624 threads = 78 row blocks x 8 column blocks hold an 8x8 accumulator tile each, all register resident)
// and one split of the chunk list (row-group items x 8-voxel chunks, 16-fold mirror/transpose folding as in k_decompose).
// A CTA regenerates R for all l >= its smallest m (the only redundancy between row groups) and Q only for its own m.
#pragma once
#include "z3d_kernels.cuh"

namespace z3d {

constexpr int BNT = 640;            // threads per CTA (20 warps, 96 registers)
constexpr int BNWARP = BNT / 32;
constexpr int BNV = 32;             // volumes per batch tile
constexpr int BCOL = 2 * BNV;       // GEMM columns: [re | im][volume]
constexpr int BBLK_MAX = BNT / 8;   // row blocks per group (80)
constexpr int BROWS_MAX = BBLK_MAX * 8;
constexpr int BMPZ_MAX = 32;        // (m, pz) combinations per row group
constexpr int BQ_MAX = 192;         // Q values per voxel per row group
constexpr int BR_MAX = 1024;        // R values per voxel (all (n, l): 676 at order 50, 961 at order 60)
constexpr int BGD_MAX = 512;        // float4 Legendre coefficient entries per row group
constexpr int BMAXTASK = 4;         // generation tasks per warp
constexpr int BFS = BNV + 4;        // padded volume stride of the fold buffer (16-byte aligned, bank spread)
constexpr int BRAW = 2 * 4 * BNV * 16;  // floats of one raw staging buffer: [set][row (sy,sx)][volume][direct 8 | mirrored 8]
constexpr int BSLOT_WORDS = BNT * TW;
constexpr int BFLUSH_CHUNKS = 256;  // register tiles are added into the CTA's partial buffer every 2048 voxels

constexpr int BOFF_T = 0;                                        // T panel    [KC][BROWS_MAX]
constexpr int BOFF_A = BOFF_T + KC * BROWS_MAX * 4;              // a panels   [BMPZ_MAX][KC][BCOL]
constexpr int BOFF_R = BOFF_A + BMPZ_MAX * KC * BCOL * 4;        // R panel    [KC][BR_MAX]
constexpr int BOFF_Q = BOFF_R + KC * BR_MAX * 4;                 // Q panel    [KC][BQ_MAX]
constexpr int BOFF_F = BOFF_Q + KC * BQ_MAX * 4;                 // folds      [2][8][KC][BFS]
constexpr int BOFF_RAW = BOFF_F + 2 * 8 * KC * BFS * 4;          // raw voxel segments, double buffered (cp.async)
constexpr int BOFF_ROWTAB = BOFF_RAW + 2 * BRAW * 4;             // (R index, Q index) per row   [BROWS_MAX] int2
constexpr int BOFF_GD = BOFF_ROWTAB + BROWS_MAX * 8;             // Legendre coefficients of the group's m
constexpr int BOFF_RK = BOFF_GD + BGD_MAX * 16;                  // radial coefficients (all l)
constexpr int BOFF_TABS = BOFF_RK + BR_MAX * 16;                 // rbase, goff, qoff  [3][LMAX+1]; mpz list [BMPZ_MAX]
constexpr int BOFF_TASKS = BOFF_TABS + 3 * (LMAX + 1) * 4 + BMPZ_MAX * 4;   // generation tasks [BNWARP][BMAXTASK]
constexpr int BOFF_TRIG = BOFF_TASKS + BNWARP * BMAXTASK * 4;
constexpr int BOFF_GEO = BOFF_TRIG + 4 * (LMAX + 1) * 4;         // per-voxel geometry of one row group, 32 B each
static_assert(BOFF_A % 16 == 0 && BOFF_R % 16 == 0 && BOFF_Q % 16 == 0 && BOFF_F % 16 == 0 && BOFF_RAW % 16 == 0 &&
              BOFF_ROWTAB % 16 == 0 && BOFF_GD % 16 == 0 &&
              BOFF_RK % 16 == 0 && BOFF_GEO % 16 == 0, "smem alignment");
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
    const int2 *rowtab;      // [ngroups][BROWS_MAX]   (index into the R panel, index into the Q panel)
    const int *blk_mpz;      // [ngroups][BBLK_MAX]    a-panel index of each row block
    const int *mpz;          // [ngroups][BMPZ_MAX]    m | pz << 16 of each a panel, -1 padded
    const int *nmpz;         // [ngroups]
    const int *mlo, *mhi;    // [ngroups]              range of m whose rows the group owns
    const float4 *gd;        // [ngroups][BGD_MAX]     (gamma_E, -delta_E, gamma_O, -delta_O) per (m, step)
    const int *goff;         // [ngroups][L+1]         offset of m's steps in the group's gd table
    const int *qoff;         // [ngroups][L+1]         offset of Q_{m,m} in the group's Q panel (Q_lm at qoff + l - m)
    const float4 *rk;        // [num_radial + pad]     (K1, K2, K3, 0), l-major
    const int *rbase;        // [L+1]                  offset of (l, n = l) in rk == index in the R panel
    const int *gen_tasks;    // [ngroups][BNWARP][BMAXTASK]   (type << 16 | base): 1 = Legendre, 2 = radial
};

struct CtxB {
    float *Tp, *Ap, *Rp, *Qp, *Fp, *raw, *cm, *sm, *cm2, *sm2;
    int2 *rowtab;
    int *tasks;
    float4 *gd, *rk;
    int *rbase, *goff, *qoff, *mpz;
    Geo *geo;
};
__device__ __forceinline__ CtxB make_ctx_b(unsigned char *smem) {
    CtxB c;
    c.Tp = reinterpret_cast<float *>(smem + BOFF_T);
    c.Ap = reinterpret_cast<float *>(smem + BOFF_A);
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
    float *tr = reinterpret_cast<float *>(smem + BOFF_TRIG);
    c.cm = tr;
    c.sm = tr + (LMAX + 1);
    c.cm2 = tr + 2 * (LMAX + 1);
    c.sm2 = tr + 3 * (LMAX + 1);
    c.geo = reinterpret_cast<Geo *>(smem + BOFF_GEO);
    return c;
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
// their 8 z-mirrors, 32 contiguous bytes each.  raw[((set*4 + row)*BNV + b)*16 + (half ^ swz(b))*8 + j]
__device__ __forceinline__ void stage_chunk(const float *__restrict__ vol, float *raw, const ChunkPos &p, int N, size_t N3,
                                            int nvol, bool vec4) {
    const int tid = threadIdx.x;
    if (tid < 2 * 4 * BNV * 2) {
        const int half = tid & 1, b = (tid >> 1) % BNV, row = (tid >> 6) & 3, set = tid >> 8;
        if (b < nvol && (set == 0 || p.paired)) {
            const int ri = set ? p.jp : p.ip, rj = set ? p.ip : p.jp;
            const int ii = (row & 2) ? N - 1 - ri : ri, jj = (row & 1) ? N - 1 - rj : rj;
            const int k0 = p.kc * KC;
            const float *src = vol + (size_t)b * N3 + ((size_t)ii * N + jj) * N + (half ? N - KC - k0 : k0);
            float *dst = raw + ((set * 4 + row) * BNV + b) * 16 + ((half ^ ((b >> 1) & 1)) * 8);
            if (vec4) {
                cp_async16(dst, src);
                cp_async16(dst + 4, src + 4);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) cp_async4(dst + j, src + j);
            }
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

// (1a) fold the 16 mirror / transpose images of the chunk's voxels for every volume: raw -> F[set][combo][k][volume];
// (1b) recurrences of the chunk: R for all l >= mlo, Q (both parities) for the row group's m, one task per warp
__device__ __forceinline__ void fold_and_recur(const DevPlanB &P, const CtxB &c, const ChunkPos &p, const float *raw, int nvol,
                                               int mhi) {
    const int L = P.order, N = P.dim, H = P.H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
        for (int cc8 = 0; cc8 < 8; ++cc8) c.Fp[((set * 8 + cc8) * KC + k) * BFS + b] = v[cc8];
    }
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
            float2 *d = reinterpret_cast<float2 *>(c.Qp + gk * BQ_MAX + c.qoff[mm]);
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
            float *dst = c.Rp + gk * BR_MAX + c.rbase[ll];
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

// (2a) T rows = R * Q;  (2b) a panels: a = F_A trig_A + F_B trig_B per (m, pz), re / im, voxel and 4 volumes at a time
//      (re: even in y, (-1)^m in x; im: odd in y, -(-1)^m in x, negated; z parity pz)
__device__ __forceinline__ void products(const CtxB &c, int nmpz) {
    const int tid = threadIdx.x;
    {
        const int2 rq = c.rowtab[tid];
        float t[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) t[k] = c.Rp[k * BR_MAX + rq.x] * c.Qp[k * BQ_MAX + rq.y];
#pragma unroll
        for (int k = 0; k < KC; ++k) c.Tp[k * BROWS_MAX + tid] = t[k];
    }
    const int b4 = tid & 7, k = (tid >> 3) & 7, reim = (tid >> 6) & 1;
    const float *fbase = c.Fp + ((reim << 2) * KC + k) * BFS + 4 * b4;
    float *abase = c.Ap + k * BCOL + reim * BNV + 4 * b4;
    const float *trig = c.cm + reim * (LMAX + 1);  // cm | sm | cm2 | sm2 are adjacent arrays of LMAX + 1
    const float sgn = reim ? -1.0f : 1.0f;
#pragma unroll 2
    for (int j = tid >> 7; j < nmpz; j += BNT / 128) {
        const int code = c.mpz[j], m = code & 0xffff, pz = code >> 16;
        const int comb = ((((m & 1) ^ reim) << 1) | pz) * (KC * BFS);
        const float tA = sgn * trig[m], tB = sgn * trig[2 * (LMAX + 1) + m];
        const float4 fa = *reinterpret_cast<const float4 *>(fbase + comb);
        const float4 fb = *reinterpret_cast<const float4 *>(fbase + 8 * KC * BFS + comb);
        float4 a;
        a.x = fmaf(fa.x, tA, fb.x * tB);
        a.y = fmaf(fa.y, tA, fb.y * tB);
        a.z = fmaf(fa.z, tA, fb.z * tB);
        a.w = fmaf(fa.w, tA, fb.w * tB);
        *reinterpret_cast<float4 *>(abase + j * (KC * BCOL)) = a;
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
//   X(t): T and a panels of chunk t from R, Q, F, trig                         | geometry of chunk t+1's row group
//   Y(t): copies of chunk t+2 issued, fold + recurrences of chunk t+1 (R, Q, F) | contraction of chunk t | trig of t+1
// Inside Y there is no barrier between the (latency-bound) recurrences and the (FMA-bound) contraction, so warps that
// carry a recurrence task overlap with warps that are already contracting.
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
    for (int i = tid; i < BMPZ_MAX; i += BNT) c.mpz[i] = P.mpz[grp * BMPZ_MAX + i];
    for (int i = tid; i < BNWARP * BMAXTASK; i += BNT) c.tasks[i] = P.gen_tasks[grp * BNWARP * BMAXTASK + i];
    for (int i = tid; i < BROWS_MAX; i += BNT) c.rowtab[i] = P.rowtab[grp * BROWS_MAX + i];
    for (int i = tid; i < BGD_MAX; i += BNT) c.gd[i] = P.gd[grp * BGD_MAX + i];
    for (int i = tid; i < BR_MAX; i += BNT) c.rk[i] = P.rk[i];
    for (int i = tid; i < (BOFF_ROWTAB - BOFF_T) / 4; i += BNT) c.Tp[i] = 0.0f;  // T, a, R, Q, F, raw are contiguous
    const int nmpz = P.nmpz[grp], mhi = P.mhi[grp];

    // contraction role: row block rb, column block cb
    const int rb = tid >> 3, cb = tid & 7;
    const int fp = (cb >> 2) & 1;  // which half of the 8 columns is loaded first (bank spread, see TileDesc)
    const float *tptr = c.Tp + 8 * rb;
    const float *aptr = c.Ap + (size_t)P.blk_mpz[grp * BBLK_MAX + rb] * (KC * BCOL) + 8 * cb;

    unsigned long long acc[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = 0ull;

    const int CPR = (H + KC - 1) / KC;
    const long long total = (long long)analysis_items(0, H, H) * CPR;
    const long long gbeg = total * part / nparts, glen = total * (part + 1) / nparts - gbeg;
    const long long cbeg = gbeg + glen * split / P.nsplit, cend = gbeg + glen * (split + 1) / P.nsplit;
    const size_t N3 = (size_t)N * N * N;
    const bool vec4 = (N & 3) == 0;
    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(partial + ((size_t)(grp * P.nsplit + split)) * BSLOT_WORDS + (size_t)tid * TW);
    int since_flush = 0;
    bool flushed = false;

    ChunkPos cur, nxt;
    bool have_next = false, next_new = false;
    int buf = 1;  // staging buffer that holds (or receives) chunk cc + 1
    __syncthreads();  // panels cleared before the first copies land
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
        fold_and_recur(P, c, cur, c.raw, nvol, mhi);
    }
    for (long long cc = cbeg; cc < cend; ++cc) {
        DBG_T(t0);
        __syncthreads();  // R, Q, F of chunk cc are complete; the contraction of chunk cc - 1 is done with T and a
        DBG_T(t1);
        DBGB_ADD(0, t0, t1);
        // ---- X(cc)
        if (have_next && next_new) item_geometry(P, c, nxt);  // geometry is only read by the recurrences (phase Y)
        products(c, nmpz);
        cp_async_wait_all();  // this thread's copies of chunk cc + 1
        DBG_T(t2);
        DBGB_ADD(1, t1, t2);
        __syncthreads();  // T, a complete; R, Q, F free; raw[buf] visible
        DBG_T(t3);
        DBGB_ADD(2, t2, t3);
        // ---- Y(cc)
        ChunkPos nn2 = nxt;
        bool have2 = false, new2 = false;
        if (have_next) {
            have2 = cc + 2 < cend;
            if (have2) {
                new2 = advance_pos(nn2, cc + 2, CPR, H);
                stage_chunk(vol, c.raw + (buf ^ 1) * BRAW, nn2, N, N3, nvol, vec4);
            }
            fold_and_recur(P, c, nxt, c.raw + buf * BRAW, nvol, mhi);
            if (next_new) item_trig(P, c, nxt);  // trig is only read by the a panels (phase X)
        }
        DBG_T(t4);
        DBGB_ADD(3, t3, t4);
        // contraction: 8 rows x 8 columns per thread, fma.rn.f32x2 with diagonal pairing (see k_decompose)
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const ulonglong2 ra = *reinterpret_cast<const ulonglong2 *>(tptr + k * BROWS_MAX);
            const ulonglong2 rbv = *reinterpret_cast<const ulonglong2 *>(tptr + k * BROWS_MAX + 4);
            const ulonglong2 pa = *reinterpret_cast<const ulonglong2 *>(aptr + k * BCOL + 4 * fp);
            const ulonglong2 pb = *reinterpret_cast<const ulonglong2 *>(aptr + k * BCOL + 4 * (1 - fp));
            const unsigned long long Rp[4] = {ra.x, ra.y, rbv.x, rbv.y};
            const unsigned long long Pp[4] = {pa.x, pa.y, pb.x, pb.y};
            unsigned long long Px[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) Px[j] = swap2(Pp[j]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ffma2(acc[(i * 4 + j) * 2], Rp[i], Pp[j]);
                    ffma2(acc[(i * 4 + j) * 2 + 1], Rp[i], Px[j]);
                }
        }
        DBG_T(t5);
        DBGB_ADD(4, t4, t5);
        // two-level summation (see k_decompose): bounds the length of any float32 running sum
        if (++since_flush == BFLUSH_CHUNKS && cc + 1 < cend) {
            since_flush = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                ulonglong2 v = make_ulonglong2(acc[2 * q], acc[2 * q + 1]);
                if (flushed) {
                    const ulonglong2 old = o[q];
                    v.x = fadd2(v.x, old.x);
                    v.y = fadd2(v.y, old.y);
                }
                o[q] = v;
                acc[2 * q] = 0ull;
                acc[2 * q + 1] = 0ull;
            }
            flushed = true;
        }
        cur = nxt;
        nxt = nn2;
        have_next = have2;
        next_new = new2;
        buf ^= 1;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        ulonglong2 v = make_ulonglong2(acc[2 * q], acc[2 * q + 1]);
        if (flushed) {
            const ulonglong2 old = o[q];
            v.x = fadd2(v.x, old.x);
            v.y = fadd2(v.y, old.y);
        }
        o[q] = v;
    }
#ifdef Z3D_DEBUG_TIMING
    if ((tid & 31) == 0)
        for (int i = 0; i < 8; ++i) atomicAdd((unsigned long long *)&g_dbg[i], (unsigned long long)s_dbg[tid >> 5][i]);
#endif
}

// moments[b][mu][reim] = scale[mu] * sum over splits of the owning tile element (fixed order, float64)
__global__ void k_reduce_batched(const float *__restrict__ partial, const int2 *__restrict__ rowmap /* [nmom] (group, row) */,
                                 const double *__restrict__ scale, int nmom, int nsplit, int nvol,
                                 float *__restrict__ moments) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over nmom * BCOL
    if (idx >= nmom * BCOL) return;
    const int mu = idx / BCOL, col = idx % BCOL;
    const int reim = col / BNV, b = col % BNV;
    if (b >= nvol) return;
    const int2 gr = rowmap[mu];
    const int r = gr.y, rb = r >> 3, i = r & 7, cb = col >> 3, c8 = col & 7;
    const int fp = (cb >> 2) & 1;
    const int ip = i >> 1, ei = i & 1;
    const int half = c8 >> 2, jp = (half == fp ? 0 : 2) + ((c8 & 3) >> 1), ej = c8 & 1;
    const int word = ((ip * 4 + jp) * 2 + (ei != ej)) * 2 + ei;
    const float *p = partial + (size_t)gr.x * nsplit * BSLOT_WORDS + (size_t)(rb * 8 + cb) * TW + word;
    double sum = 0.0;
    for (int s = 0; s < nsplit; ++s) sum += (double)p[(size_t)s * BSLOT_WORDS];
    moments[((size_t)b * nmom + mu) * 2 + reim] = (float)(sum * scale[mu]);
}

}  // namespace z3d
