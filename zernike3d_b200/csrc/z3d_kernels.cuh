// zernike3d_b200 - fused sm_100a kernels for the spharm 3D-Zernike analysis / synthesis path.
//
// Neither kernel ever materialises an R_nl or Y_lm voxel grid (the reference builds 676 + 1326 of them,
// zm:2281-2323).  Per chunk of KC folded voxels a CTA
//   (0) computes r, cos(theta), sin(theta) once per voxel in float64 (zm:1886-1890) and the per-row
//       cos(m phi), sin(m phi) once per row group,
//   (1) runs the reference's recurrences in registers - R_nl in n at fixed l (zm:1980-2002), the
//       normalised associated Legendre functions in l at fixed m (zm:2006-2026, in monic parity-decoupled
//       form; the normalisation kappa_lm is applied once to the finished sums) - and stages the two operand
//       panels R[n][k] and a*P[(m,re/im)][k] of the per-l contraction in shared memory,
//   (2) contracts them over the voxel axis into register-resident 8x8 output tiles (analysis), or
//       contracts register-resident moment tiles with them into 8 mirror-class sums (synthesis),
//       with packed fma.rn.f32x2.
#pragma once
#include "z3d_plan.h"

namespace z3d {

#ifdef Z3D_DEBUG_TIMING
// developer instrumentation: per-warp cycle totals of the analysis kernel's phases (warp 0 lane 0 of CTA 0..)
__device__ long long g_dbg[8];
#define DBG_T(var) const long long var = clock64()
#define DBG_ADD(slot, t0, t1) do { if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)&g_dbg[slot], (unsigned long long)((t1) - (t0))); } while (0)
#else
#define DBG_T(var)
#define DBG_ADD(slot, t0, t1)
#endif

// ---------------------------------------------------------------------------------------------
// shared-memory carve-up.  Everything the recurrences and the contraction touch sits at a COMPILE-TIME offset
// (fixed-size tables first): with 64 accumulators live the compiler otherwise rematerialises the layout arithmetic
// from kernel parameters inside every recurrence step.  Only the box-size dependent arrays at the end have
// run-time offsets; they are used in the per-row-group setup only.
// ---------------------------------------------------------------------------------------------
constexpr int LMAX = 127;                                   // largest order the fixed tables hold
constexpr int OFF_RS = 0;                                   // R panel        [KC][RS]
constexpr int OFF_PS = OFF_RS + KC * RS * 4;                // a*P panel      [KC][PS]
constexpr int OFF_GD = OFF_PS + KC * PS * 4;                // Legendre coefficients of this pass
constexpr int OFF_RK = OFF_GD + GD_MAX * 8;                 // radial coefficients of this pass
constexpr int OFF_TABS = OFF_RK + RK_MAX * 16;              // poffw, roffw, goff, rbase  [4][LMAX+1]
constexpr int OFF_CM = OFF_TABS + 4 * (LMAX + 1) * 4;       // cos(m phi)
constexpr int OFF_SM = OFF_CM + (LMAX + 1) * 4;             // sin(m phi)
constexpr int OFF_CM2 = OFF_SM + (LMAX + 1) * 4;            // cos / sin(m phi) of the transposed row group (analysis)
constexpr int OFF_SM2 = OFF_CM2 + (LMAX + 1) * 4;
constexpr int OFF_RED = OFF_SM2 + (LMAX + 1) * 4;           // synthesis: cross-warp reduction
constexpr int OFF_GEO = OFF_RED + 4 * NWARP * 16 * 4;       // per-voxel geometry of one row group, 32 B each
struct SmemMap {
    int F, rows, V, total;  // byte offsets of the box-size dependent arrays
};
__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline SmemMap smem_map(int N, int H) {
    SmemMap m;
    int o = OFF_GEO + (H + KC) * 32;
    m.F = o;    o += 2 * align16(8 * H * 4);         // analysis: 8 mirror-folded rows of the row group and of its transpose
    m.rows = o; o += align16(4 * N * 4);             // analysis: 4 staged rows
    m.V = o;    o += align16(8 * H * 4);             // synthesis: class sums of one row group and of its transpose
    m.total = o;
    return m;
}
struct Geo {  // one folded voxel
    float c2, c2sq, st, st2, r, r2, pad0, pad1;
};

// b^e for small non-negative integer e using squares of b2 = b*b (fewer roundings than from b alone)
__device__ __forceinline__ float powi2(float b, float b2, int e) {
    float res = (e & 1) ? b : 1.0f;
    e >>= 1;
    float p = b2;
    while (e) {
        if (e & 1) res *= p;
        p *= p;
        e >>= 1;
    }
    return res;
}

// packed FP32 helpers (sm_100a): two FMAs per issued instruction
__device__ __forceinline__ void ffma2(unsigned long long &d, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ unsigned long long swap2(unsigned long long v) {
    const float2 f = unpack2(v);
    return pack2(f.y, f.x);
}

struct Ctx {  // all members are compile-time offsets from the dynamic shared memory base
    float *Rs, *Ps, *cm, *sm, *cm2, *sm2;
    int *poffw, *roffw, *goff, *rbase;
    float2 *gd;
    float4 *rk;
    Geo *geo;
};
__device__ __forceinline__ Ctx make_ctx(unsigned char *smem) {
    Ctx c;
    c.Rs = reinterpret_cast<float *>(smem + OFF_RS);
    c.Ps = reinterpret_cast<float *>(smem + OFF_PS);
    c.gd = reinterpret_cast<float2 *>(smem + OFF_GD);
    c.rk = reinterpret_cast<float4 *>(smem + OFF_RK);
    int *tabs = reinterpret_cast<int *>(smem + OFF_TABS);
    c.poffw = tabs;
    c.roffw = tabs + (LMAX + 1);
    c.goff = tabs + 2 * (LMAX + 1);
    c.rbase = tabs + 3 * (LMAX + 1);
    c.cm = reinterpret_cast<float *>(smem + OFF_CM);
    c.sm = reinterpret_cast<float *>(smem + OFF_SM);
    c.cm2 = reinterpret_cast<float *>(smem + OFF_CM2);
    c.sm2 = reinterpret_cast<float *>(smem + OFF_SM2);
    c.geo = reinterpret_cast<Geo *>(smem + OFF_GEO);
    return c;
}

// Step (1): every warp advances its share of the recurrence chains for the KC voxels of the chunk.
// lane -> (voxel gk = lane % KC, sub-lane gsub = lane / KC); each lane runs CPL independent chains; tasks are
// warp-uniform.
//
// Legendre chains run in the parity-decoupled monic form: with Q_lm = P_lm / kappa_lm and
// Q_l = 2c Q_{l-1} - beta_l Q_{l-2} (Y_lm_recurrence zm:2021-2026 divided by kappa), eliminating the
// other parity gives  Q_{l+2} = (4c^2 - gamma_l) Q_l - delta_l Q_{l-2},  gamma_l = beta_{l+2} + beta_{l+1},
// delta_l = beta_{l+1} beta_l.  A pass owns one parity of l, so it walks only its own levels, and because
// the recurrence is linear the chain is run on the two products a_re*Q and a_im*Q directly (the z-mirror
// parity (l+m)&1 that selects a_re / a_im is constant along such a chain).
template <int NPH, bool ANALYSIS>  // NPH = npass / 2 when known at compile time (1), 0 = read it from the plan
__device__ __forceinline__ void generate_panels(const DevPlan &P, const Ctx &c, const float *F, int k0, int pass,
                                                int warp, int lane) {
    const int L = P.order, H = P.H;
    const int gk = lane % KC, gsub = lane / KC, kk = k0 + gk;
    const int par = pass & 1, nph = NPH ? NPH : (P.npass >> 1), rho = pass >> 1;
    const int *tasks = P.gen_tasks + (pass * NWARP + warp) * MAXTASK;
    const float4 g0 = reinterpret_cast<const float4 *>(c.geo + kk)[0], g1 = reinterpret_cast<const float4 *>(c.geo + kk)[1];
    const float c2 = g0.x, c2sq = g0.y, st = g0.z, st2 = g0.w, r = g1.x, r2 = g1.y;
    for (int t = 0; t < MAXTASK; ++t) {
        const int task = __ldg(tasks + t);
        const int type = task >> 16, base = task & 0xffff;
        if (type == 0) break;
        if (type == 1) {
            const int lmin = base + ((base ^ par) & 1);          // warp-uniform first level of the group
            const int mlast = min(base + GRP - 1, L);
            const int l0max = min(mlast + ((mlast ^ par) & 1), L);  // all lanes have started after this level
            int l0[CPL];
            const float2 *tp[CPL];
            float *dst[CPL];
            float sa[CPL], sb[CPL], A1[CPL], A2[CPL], B1[CPL], B2[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int m = min(base + q * SUBS + gsub, L);    // lanes past L redo chain L (same stores, same values)
                l0[q] = m + ((m ^ par) & 1);                     // first level of this parity with l >= m
                // a = (folded f) x trig for analysis (re: even in y, (-1)^m in x; im: odd in y, -(-1)^m in x; z parity
                // (l+m)&1 = (par+m)&1), plain trig for synthesis
                float2 a = make_float2(c.cm[m], c.sm[m]);
                if (ANALYSIS) {
                    // row group A = (i, j) and its transpose B = (j, i) share rho, hence R and Q; the moment sum is
                    // linear in a, so both are contracted at once with a = F_A trig_A + F_B trig_B (F_B = 0 on the diagonal)
                    const int px = m & 1, pz = (par + m) & 1;
                    const bool live = kk < H;
                    const int cre = (((px << 1) | pz) * H) + kk, cim = ((4 | ((px ^ 1) << 1) | pz) * H) + kk;
                    const float *FB = F + 8 * H;
                    a.x = live ? fmaf(F[cre], a.x, FB[cre] * c.cm2[m]) : 0.f;
                    a.y = live ? -fmaf(F[cim], a.y, FB[cim] * c.sm2[m]) : 0.f;
                }
                float seed = powi2(st, st2, m);                  // Q_mm = sin^m  (zm:2012-2015 / kappa_mm)
                if (l0[q] != m) seed *= c2;                      // Q_{m+1,m} = 2c Q_mm (zm:2016-2020 / kappa)
                sa[q] = a.x * seed;
                sb[q] = a.y * seed;
                // table pointer such that entry [p-1] advances level lmin + 2(p-1) -> lmin + 2p for this lane
                tp[q] = c.gd + c.goff[m] - ((l0[q] - lmin) >> 1);
                dst[q] = c.Ps + gk * PS + 2 * m;
                A1[q] = A2[q] = B1[q] = B2[q] = 0.f;
            }
            int until = 0;  // positions until the next own level (always 0 when the pass owns its whole parity)
            if (nph > 1) {
                until = (rho - (lmin >> 1)) % nph;
                if (until < 0) until += nph;
            }
            int p = 0, l = lmin;
            // prologue (<= GRP/2 + 1 positions): chains that have not started hold zeros, then take their seed
            for (; l <= l0max; l += 2, ++p) {
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    if (p > 0) {
                        const float2 g = tp[q][p - 1];
                        const float tt = c2sq - g.x;
                        const float An = fmaf(tt, A1[q], g.y * A2[q]), Bn = fmaf(tt, B1[q], g.y * B2[q]);
                        A2[q] = A1[q]; A1[q] = An; B2[q] = B1[q]; B1[q] = Bn;
                    }
                    if (l <= l0[q]) {
                        const bool start = (l == l0[q]);
                        A1[q] = start ? sa[q] : 0.f; B1[q] = start ? sb[q] : 0.f;
                        A2[q] = 0.f; B2[q] = 0.f;
                    }
                }
                if (until == 0) {
                    const int off = c.poffw[l];
#pragma unroll
                    for (int q = 0; q < CPL; ++q)
                        if (l >= l0[q]) *reinterpret_cast<float2 *>(dst[q] + off) = make_float2(A1[q], B1[q]);
                    until = nph;
                }
                --until;
            }
            // main loop: every chain is live.  Coefficients and store offsets are fetched one iteration ahead so the
            // shared-memory latency is off the recurrence's critical path (which is one FFMA per level).
            if (NPH == 1 && CPL == 1) {
                const float2 *t0 = tp[0] + p - 1;
                const int *po = c.poffw + l;
                float *d0 = dst[0];
                float2 g = t0[0];
                int off = po[0];
                float a1 = A1[0], a2 = A2[0], b1 = B1[0], b2 = B2[0];
                const int npos = (l <= L) ? (L - l) / 2 + 1 : 0;
#pragma unroll 4
                for (int i = 0; i < npos; ++i) {
                    const float2 gn = t0[i + 1];       // tables are padded: reads past the chain end are discarded
                    const int offn = po[2 * i + 2];
                    const float tt = c2sq - g.x;
                    const float an = fmaf(tt, a1, g.y * a2), bn = fmaf(tt, b1, g.y * b2);
                    a2 = a1; a1 = an; b2 = b1; b1 = bn;
                    *reinterpret_cast<float2 *>(d0 + off) = make_float2(an, bn);
                    g = gn;
                    off = offn;
                }
            } else {
#pragma unroll 2
                for (; l <= L; l += 2, ++p) {
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const float2 g = tp[q][p - 1];
                        const float tt = c2sq - g.x;
                        const float An = fmaf(tt, A1[q], g.y * A2[q]), Bn = fmaf(tt, B1[q], g.y * B2[q]);
                        A2[q] = A1[q]; A1[q] = An; B2[q] = B1[q]; B1[q] = Bn;
                    }
                    if (until == 0) {
                        const int off = c.poffw[l];
#pragma unroll
                        for (int q = 0; q < CPL; ++q) *reinterpret_cast<float2 *>(dst[q] + off) = make_float2(A1[q], B1[q]);
                        until = nph;
                    }
                    --until;
                }
            }
        } else {
            // Radial chains for own l = ownl[base + q*SUBS + gsub]:  R_n = (K1 r^2 + K2) R_{n-2} + K3 R_{n-4}
            // (R_nl_recurrence zm:1991-2000; the l == n-2 branch zm:1987-1988 is the same formula with K3 = 0 and
            // the seed R_ll = r^l enters through the table entry (0, 0, 1) of n = l).
            const int nown = __ldg(P.nown + pass);
            const int *own = P.ownl + pass * (L + 1);
            const int nmax = (L - __ldg(own + base)) / 2 + 1;
            int nn[CPL];
            float R1[CPL], R2[CPL];
            float *dst[CPL];
            const float4 *kp[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int oi = base + q * SUBS + gsub;
                const bool active = oi < nown;
                const int l = __ldg(own + (active ? oi : base));
                nn[q] = active ? (L - l) / 2 + 1 : 0;
                R1[q] = 0.0f;
                R2[q] = powi2(r, r2, l);
                dst[q] = c.Rs + gk * RS + c.roffw[l];
                kp[q] = c.rk + c.rbase[l];
            }
            // every lane runs the group's longest chain length; shorter chains keep computing past their end (finite
            // throw-away values from the next table entries) and only the store is predicated: no divergence
            float4 Kc[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) Kc[q] = kp[q][0];
#pragma unroll 4
            for (int ni = 0; ni < nmax; ++ni) {
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const float4 Kn = kp[q][ni + 1];  // prefetch (table is padded)
                    const float R = fmaf(fmaf(Kc[q].x, r2, Kc[q].y), R1[q], Kc[q].z * R2[q]);
                    if (ni < nn[q]) dst[q][ni] = R;
                    R2[q] = R1[q];
                    R1[q] = R;
                    Kc[q] = Kn;
                }
            }
        }
    }
}

// Geometry of the H folded voxels of a row group; float64 like the reference (zm:2240-2242, 1886-1890).
__device__ __forceinline__ void row_geometry(const DevPlan &P, const Ctx &c, double rho2, double rho) {
    for (int k = threadIdx.x; k < P.H + KC; k += NT) {
        const double z = (k < P.H) ? P.side[k] : 0.0;
        const double r2d = rho2 + z * z;
        const double r = sqrt(r2d);
        double ct = 1.0, st = 0.0;  // r == 0: theta is NaN in the reference and nan_to_num zeroes Y_lm, l > 0
        if (r > 0.0) {             // (zm:2298); R_nl(0) = 0 for l > 0 gives the same product here.
            ct = z / r;
            st = rho / r;
        }
        float4 *g = reinterpret_cast<float4 *>(c.geo + k);
        g[0] = make_float4((float)(2.0 * ct), (float)(4.0 * ct * ct), (float)st, (float)(st * st));
        g[1] = make_float4((float)r, (float)r2d, 0.f, 0.f);
    }
}

// in-register 8-point Walsh-Hadamard transform: v[c] = sum_img v[img] * (-1)^popcount(c & img)
__device__ __forceinline__ void wht8(float v[8]) {
#pragma unroll
    for (int h = 1; h < 8; h <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (!(i & h)) {
                const float x = v[i], y = v[i | h];
                v[i] = x + y;
                v[i | h] = x - y;
            }
        }
    }
}

// per-pass coefficient and offset tables -> smem (read every chunk by the recurrences)
__device__ __forceinline__ void load_tables(const DevPlan &P, const Ctx &c, int pass) {
    const int L = P.order;
    for (int i = threadIdx.x; i <= L; i += NT) {
        c.poffw[i] = P.poffw[i];
        c.roffw[i] = P.roffw[i];
        c.goff[i] = P.goff[pass * (L + 1) + i];
        c.rbase[i] = P.rbase[i];
    }
    for (int i = threadIdx.x; i < GD_MAX; i += NT) c.gd[i] = P.gd[pass * GD_MAX + i];
    for (int i = threadIdx.x; i < RK_MAX; i += NT) c.rk[i] = P.rk[pass * RK_MAX + i];
    // padding rows / columns of the panels are never written by the recurrences but are multiplied into padding
    // accumulators (analysis) or with zero moment entries (synthesis): they must be finite -> clear both panels once.
    for (int i = threadIdx.x; i < KC * (RS + PS); i += NT) c.Rs[i] = 0.0f;  // Rs and Ps are contiguous
}

// cos(m phi0), sin(m phi0) for a row group, float64 like exp(1j m phi) in zm:2297
__device__ __forceinline__ void row_trig(const DevPlan &P, float *cm, float *sm, double x0, double y0) {
    if (threadIdx.x <= P.order) {
        const double ph = atan2(y0, x0);
        double s, co;
        sincos((double)threadIdx.x * ph, &s, &co);
        cm[threadIdx.x] = (float)co;
        sm[threadIdx.x] = (float)s;
    }
}

// stage the 4 rows (+-y, +-x) of row group (ip, jp) with coalesced float4 loads
__device__ __forceinline__ void stage_rows(const DevPlan &P, const float *__restrict__ vb, float *rows, int ip, int jp) {
    const int N = P.dim, tid = threadIdx.x;
    const int i2 = N - 1 - ip, j2 = N - 1 - jp;
    if (P.dim_vec4) {
        const int nv = N >> 2;
        for (int idx = tid; idx < 4 * nv; idx += NT) {
            const int rw = idx / nv, q = idx - rw * nv;
            const int ii = (rw & 2) ? i2 : ip, jj = (rw & 1) ? j2 : jp;
            const bool dup = ((rw & 2) && i2 == ip) || ((rw & 1) && j2 == jp);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!dup) v = __ldg(reinterpret_cast<const float4 *>(vb + ((size_t)ii * N + jj) * N) + q);
            reinterpret_cast<float4 *>(rows + rw * N)[q] = v;
        }
    } else {
        for (int idx = tid; idx < 4 * N; idx += NT) {
            const int rw = idx / N, q = idx - rw * N;
            const int ii = (rw & 2) ? i2 : ip, jj = (rw & 1) ? j2 : jp;
            const bool dup = ((rw & 2) && i2 == ip) || ((rw & 1) && j2 == jp);
            rows[rw * N + q] = dup ? 0.0f : __ldg(vb + ((size_t)ii * N + jj) * N + q);
        }
    }
}

// fold the 8 mirror images of the staged rows: F[c][k] = sum_img f_img (-1)^popcount(c & img),
// img = sy*4 + sx*2 + sz,  c = py*4 + px*2 + pz
__device__ __forceinline__ void fold_rows(const DevPlan &P, const float *rows, float *F) {
    const int N = P.dim, H = P.H;
    for (int k = threadIdx.x; k < H; k += NT) {
        const int k2 = N - 1 - k;
        float v[8];
#pragma unroll
        for (int img = 0; img < 8; ++img) {
            const float x = rows[(img >> 1) * N + ((img & 1) ? k2 : k)];
            v[img] = ((img & 1) && k2 == k) ? 0.0f : x;
        }
        wht8(v);
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) F[cc * H + k] = v[cc];
    }
}

// Work items of an analysis call over plane pairs [lo, hi): for ip in [lo, hi) the row groups (ip, jp) with jp outside
// the slab (single: the transpose's planes are not available), then jp in [ip, hi) (jp > ip: paired with the transpose
// (jp, ip); jp == ip: diagonal).  Row ip holds H + lo - ip items.
__host__ __device__ inline int analysis_items(int lo, int hi, int H) {
    const int n = hi - lo;
    return n * H - n * (n - 1) / 2;
}
__device__ __forceinline__ void decode_item(int item, int lo, int hi, int H, int &ip, int &jp, bool &paired) {
    // largest t with t*H - t(t-1)/2 <= item
    const double b = 2.0 * H + 1.0;
    int t = (int)((b - sqrt(b * b - 8.0 * (double)item)) * 0.5);
    t = max(0, min(t, hi - lo - 1));
    while (t > 0 && t * H - t * (t - 1) / 2 > item) --t;
    while (t + 1 < hi - lo && (t + 1) * H - (t + 1) * t / 2 <= item) ++t;
    const int rem = item - (t * H - t * (t - 1) / 2);
    ip = lo + t;
    const int nout = lo + (H - hi);
    if (rem < lo) { jp = rem; paired = false; }
    else if (rem < nout) { jp = hi + (rem - lo); paired = false; }
    else { jp = ip + (rem - nout); paired = (jp != ip); }
}

// ---------------------------------------------------------------------------------------------
// Analysis: per-CTA partial sums of  sum_v f R_nl P_lm e^{-i m phi}  for one pass and one split.
// partial layout: [batch][npass][nsplit][SLOT_WORDS]
//
// Register tile (8 rows n x 8 columns (m, re/im)) as 32 packed pairs: with row pairs Rp[ip] = (R[a], R[a+1]) and
// column pairs Pp[jp] = (P[b], P[b+1]) exactly as they come out of LDS.128,
//   D[ip][jp] += Rp[ip] * Pp[jp]        -> (acc[a][b],   acc[a+1][b+1])
//   X[ip][jp] += Rp[ip] * swap(Pp[jp])  -> (acc[a][b+1], acc[a+1][b])
// so no operand has to be duplicated in shared memory or registers.
// ---------------------------------------------------------------------------------------------
template <int NPH>
__global__ void __launch_bounds__(NT, CTAS_PER_SM)
k_decompose(const DevPlan P, const float *__restrict__ vol, int batch, int half_lo, int half_hi, int part, int nparts,
            float *__restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = P.dim, H = P.H;
    const SmemMap mp = smem_map(N, H);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pass = blockIdx.x % P.npass, split = blockIdx.x / P.npass;
    const Ctx c = make_ctx(smem);
    load_tables(P, c, pass);
    float *rows = reinterpret_cast<float *>(smem + mp.rows);
    float *F = reinterpret_cast<float *>(smem + mp.F);
    const int ntile = P.ntiles[pass];

    const float *r0p[TS], *r1p[TS], *p0p[TS], *p1p[TS];
#pragma unroll
    for (int s = 0; s < TS; ++s) {
        const TileDesc t = P.tiles[(pass * TS + s) * NT + tid];
        r0p[s] = c.Rs + t.r0; r1p[s] = c.Rs + t.r1;
        p0p[s] = c.Ps + t.p0; p1p[s] = c.Ps + t.p1;
    }

    const int CPR = (H + KC - 1) / KC;  // chunks per row group
    const long long total = (long long)analysis_items(half_lo, half_hi, H) * CPR;
    // this launch's share of the work list (part / nparts: multi-GPU work sharding), then this CTA's share of that
    const long long gbeg = total * part / nparts, glen = total * (part + 1) / nparts - gbeg;
    const long long cbeg = gbeg + glen * split / P.nsplit, cend = gbeg + glen * (split + 1) / P.nsplit;
    const size_t N3 = (size_t)N * N * N;
    for (int b = 0; b < batch; ++b) {
        unsigned long long acc[TS][32];  // [(ip*4 + jp)*2 + {D, X}]
#pragma unroll
        for (int s = 0; s < TS; ++s)
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[s][e] = 0ull;
        const float *vb = vol + (size_t)b * N3;

        float *out = partial + (((size_t)b * P.npass + pass) * P.nsplit + split) * SLOT_WORDS;
        int since_flush = 0;
        bool flushed = false;
        for (long long cc = cbeg; cc < cend;) {
            const int item = (int)(cc / CPR), kfirst = (int)(cc % CPR);
            int ip, jp;
            bool paired;
            decode_item(item, half_lo, half_hi, H, ip, jp, paired);
            const double y0 = P.side[ip], x0 = P.side[jp];
            const double rho2 = x0 * x0 + y0 * y0, rho = sqrt(rho2);

            // ---- per-item setup: stage + fold the rows of (ip, jp) and of its transpose, trig of both, geometry
            DBG_T(ts0);
            __syncthreads();
            stage_rows(P, vb, rows, ip, jp);
            row_trig(P, c.cm, c.sm, x0, y0);
            row_trig(P, c.cm2, c.sm2, y0, x0);
            row_geometry(P, c, rho2, rho);
            __syncthreads();
            fold_rows(P, rows, F);
            __syncthreads();
            if (paired) {
                stage_rows(P, vb, rows, jp, ip);
                __syncthreads();
                fold_rows(P, rows, F + 8 * H);
            } else {
                for (int k = tid; k < 8 * H; k += NT) F[8 * H + k] = 0.0f;
            }
            __syncthreads();
            DBG_T(ts1);
            DBG_ADD(0, ts0, ts1);

            for (int kc = kfirst; kc < CPR && cc < cend; ++kc, ++cc) {
                const int k0 = kc * KC;
                DBG_T(tw0);
                if (kc != kfirst) __syncthreads();  // the previous chunk's contraction is done reading the panels
                DBG_T(tg0);
                DBG_ADD(4, tw0, tg0);
                // ---- step 1: recurrences -> operand panels in smem
#ifndef Z3D_DEBUG_SKIP_GEN
                generate_panels<NPH, true>(P, c, F, k0, pass, warp, lane);
#endif
                DBG_T(tg1);
                __syncthreads();
                DBG_T(tm0);
                DBG_ADD(1, tg0, tg1);
                DBG_ADD(2, tg1, tm0);
                // ---- step 2: contraction over the chunk's voxels into the register tiles (fma.rn.f32x2)
#ifndef Z3D_DEBUG_SKIP_MAC
#pragma unroll
                for (int k = 0; k < KC; ++k) {
#pragma unroll
                    for (int s = 0; s < TS; ++s) {
                        const ulonglong2 ra = *reinterpret_cast<const ulonglong2 *>(r0p[s] + k * RS);
                        const ulonglong2 rb = *reinterpret_cast<const ulonglong2 *>(r1p[s] + k * RS);
                        const ulonglong2 pa = *reinterpret_cast<const ulonglong2 *>(p0p[s] + k * PS);
                        const ulonglong2 pb = *reinterpret_cast<const ulonglong2 *>(p1p[s] + k * PS);
                        const unsigned long long Rp[4] = {ra.x, ra.y, rb.x, rb.y};
                        const unsigned long long Pp[4] = {pa.x, pa.y, pb.x, pb.y};
                        unsigned long long Px[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) Px[j] = swap2(Pp[j]);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                ffma2(acc[s][(i * 4 + j) * 2], Rp[i], Pp[j]);
                                ffma2(acc[s][(i * 4 + j) * 2 + 1], Rp[i], Px[j]);
                            }
                    }
                }
#endif
                DBG_T(tm1);
                DBG_ADD(3, tm0, tm1);
            }
            // two-level summation: every FLUSH_ITEMS row-group items the register tiles are added into this CTA's
            // partial buffer (owned element-wise by the thread, L2 resident) and cleared, which bounds the length of
            // any float32 running sum to a few thousand terms whatever the box size
            if (++since_flush == FLUSH_ITEMS && cc < cend) {
                since_flush = 0;
#pragma unroll
                for (int s = 0; s < TS; ++s) {
                    if (s * NT + tid < ntile) {
                        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(out + (size_t)(s * NT + tid) * TW);
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            ulonglong2 v = make_ulonglong2(acc[s][2 * q], acc[s][2 * q + 1]);
                            if (flushed) {
                                const ulonglong2 old = o[q];
                                v.x = fadd2(v.x, old.x);
                                v.y = fadd2(v.y, old.y);
                            }
                            o[q] = v;
                            acc[s][2 * q] = 0ull;
                            acc[s][2 * q + 1] = 0ull;
                        }
                    }
                }
                flushed = true;
            }
        }
#pragma unroll
        for (int s = 0; s < TS; ++s) {
            if (s * NT + tid < ntile) {
                ulonglong2 *o = reinterpret_cast<ulonglong2 *>(out + (size_t)(s * NT + tid) * TW);
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    ulonglong2 v = make_ulonglong2(acc[s][2 * q], acc[s][2 * q + 1]);
                    if (flushed) {
                        const ulonglong2 old = o[q];
                        v.x = fadd2(v.x, old.x);
                        v.y = fadd2(v.y, old.y);
                    }
                    o[q] = v;
                }
            }
        }
    }
}

// Deterministic combination of the per-CTA partials (fixed split order, float64), kappa_lm * coeff(n)
// scaling (zm:2198) and scatter into the reference's flat (n, l, m) order.
__global__ void k_reduce_moments(const float *__restrict__ partial, const int *__restrict__ src,
                                 const double *__restrict__ scale, int nmom, int npass, int nsplit, int batch,
                                 float *__restrict__ moments) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (e >= 2 * nmom || b >= batch) return;
    const int s = src[e];
    const int pass = s / SLOT_WORDS, off = s % SLOT_WORDS;
    const float *p = partial + (((size_t)b * npass + pass) * nsplit) * SLOT_WORDS + off;
    double sum = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) sum += (double)p[(size_t)sp * SLOT_WORDS];
    moments[(size_t)b * 2 * nmom + e] = (float)(sum * scale[e >> 1]);
}

// Fused "reduce partials + all-reduce over NVLink peer memory" for the work-sharded analysis of ONE volume spread
// over several GPUs (one process per GPU).  Every rank owns a communication buffer that all ranks have mapped
// (CUDA IPC): [2 slots][MAX_RANKS source ranks][max_elems] floats of data, then [2 slots][MAX_RANKS][P2P_MAX_BLOCKS] epoch flags.
// PUSH, block by block - no grid-wide step and no remote load anywhere:
//   1. each thread reduces this rank's per-CTA partials for its element (same arithmetic as k_reduce_moments) and STORES the
//      value into row `rank` of EVERY rank's buffer (posted NVLink writes, slot = epoch & 1);
//   2. after the block barrier, thread r (< world) publishes `epoch` into flag[slot][rank][block] of rank r (system-scope
//      release, cumulative over the block's stores through the barrier) - the `world` flag stores leave in parallel;
//   3. thread r then waits for rank r's flag of THIS block in the local buffer (acquire; local memory polls), so a block
//      only ever waits for its own counterpart blocks on the peers;
//   4. every thread sums its element over the source rows in rank order from LOCAL memory (L1-bypassing loads) and writes the
//      final moments: one-shot all-gather + ordered sum, bit-identical on all ranks, no NCCL call on the data path.
// Two slots make back-to-back calls safe: a peer can only push epoch e + 2 into a slot after this rank has published all of its
// epoch e + 1 flags, i.e. after its epoch e kernel (the slot's last reader) has completed.
// (Round 1's version published once per rank from the last block to finish and pulled the peers' vectors through remote loads:
// a grid-wide atomic, one thread storing 8 flags in sequence, and 8 dependent NVLink round trips per thread - 0.18-0.45 ms at
// 8 GPUs depending on the node, against 0.145 ms for compute + NCCL.)
constexpr int MAX_RANKS = 16;
constexpr int P2P_MAX_BLOCKS = 1024;
struct PeerBufs {
    float *data[MAX_RANKS];            // base of each rank's buffer (own included), device-accessible from this GPU
};
__host__ __device__ __forceinline__ size_t comm_bytes(size_t max_elems) {
    return 2 * (size_t)MAX_RANKS * max_elems * sizeof(float) + 2 * (size_t)MAX_RANKS * P2P_MAX_BLOCKS * sizeof(unsigned);
}
__device__ __forceinline__ unsigned *comm_flags(float *base, size_t max_elems) {
    return reinterpret_cast<unsigned *>(base + 2 * (size_t)MAX_RANKS * max_elems);
}
__global__ void __launch_bounds__(256)
k_reduce_allreduce_p2p(const float *__restrict__ partial, const int *__restrict__ src,
                       const double *__restrict__ scale, int nmom, int npass, int nsplit, int batch,
                       PeerBufs peers, int rank, int world, size_t max_elems, unsigned epoch,
                       int *error_flag, float *__restrict__ moments) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const int slot = epoch & 1;
    const int block = blockIdx.y * gridDim.x + blockIdx.x;
    const bool live = e < 2 * nmom && b < batch;
    const size_t idx = (size_t)b * 2 * nmom + e;
    const size_t row = ((size_t)slot * MAX_RANKS + rank) * max_elems + idx;
    if (live) {
        const int s = src[e];
        const int pass = s / SLOT_WORDS, off = s % SLOT_WORDS;
        const float *p = partial + (((size_t)b * npass + pass) * nsplit) * SLOT_WORDS + off;
        double sum = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) sum += (double)p[(size_t)sp * SLOT_WORDS];
        const float v = (float)(sum * scale[e >> 1]);
        for (int r = 0; r < world; ++r) peers.data[r][row] = v;
    }
    __syncthreads();
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    if ((int)threadIdx.x < world) {
        const int r = threadIdx.x;
        __threadfence_system();
        unsigned *f = comm_flags(peers.data[r], max_elems) + ((size_t)slot * MAX_RANKS + rank) * P2P_MAX_BLOCKS + block;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
        const unsigned *mine = comm_flags(peers.data[rank], max_elems) + ((size_t)slot * MAX_RANKS + r) * P2P_MAX_BLOCKS + block;
        const long long t0 = clock64();
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if (v != epoch && clock64() - t0 > 4000000000LL) {  // ~2 s: a peer never arrived - fail, do not hang
                atomicExch(&bad, 1);
                atomicExch(error_flag, 1);
                break;
            }
        } while (v != epoch);
    }
    __syncthreads();
    if (live) {
        float acc = 0.0f;
        if (!bad) {
            const float *mine = peers.data[rank] + (size_t)slot * MAX_RANKS * max_elems + idx;
            float v[MAX_RANKS];
#pragma unroll
            for (int r = 0; r < MAX_RANKS; ++r) v[r] = r < world ? __ldcv(mine + (size_t)r * max_elems) : 0.0f;
#pragma unroll
            for (int r = 0; r < MAX_RANKS; ++r)
                if (r < world) acc += v[r];
        } else {
            acc = __int_as_float(0x7fc00000);  // NaN marks the failed exchange
        }
        moments[idx] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// Synthesis
// ---------------------------------------------------------------------------------------------
// moments -> per-pass register-tile layout, pre-multiplied by kappa_lm * w_m (w_0 = 1, w_m = 2: the
// conj term of zm:2612-2613) and with the sign of the -Im(O) sin(m phi) term folded in.
__global__ void k_prepare_tiles(const float *__restrict__ moments, const int *__restrict__ omap,
                                const double *__restrict__ ofac, int nmom, int npass, int batch,
                                float *__restrict__ otiles) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (w >= npass * SLOT_WORDS || b >= batch) return;
    const int e = omap[w];
    float v = 0.0f;
    if (e >= 0) {
        const double f = ofac[e >> 1];
        v = (float)((double)moments[(size_t)b * 2 * nmom + e] * ((e & 1) ? -f : f));
    }
    otiles[(size_t)b * npass * SLOT_WORDS + w] = v;
}

// 16 values per lane -> lanes L and L+16 both end up with the sum over all 32 lanes of v[L & 15]
__device__ __forceinline__ float warp_transpose_reduce16(float v[16], int lane) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
        const bool upper = lane & off;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// pvol layout: [npass][batch][N^3]; only planes of the selected pairs are written.
//
// Per tile and voxel: t[col] = sum_row O[row][col] R[row] (32 packed FMAs on the resident moment tile, in the same
// D / X pair layout as the analysis accumulators), then class[col & 3] += P[col] t[col]: the 8 columns of a tile
// are (m0..m0+3) x (re, im) with m0 a multiple of 4, so column c and c+4 fall in the same mirror class.
template <int NPH>
__global__ void __launch_bounds__(NT, CTAS_PER_SM)
k_reconstruct(const DevPlan P, const float *__restrict__ otiles, int batch, int half_lo, int half_hi,
              float *__restrict__ pvol) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = P.dim, H = P.H;
    const SmemMap mp = smem_map(N, H);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pass = blockIdx.x % P.npass, split = blockIdx.x / P.npass;
    const int par = pass & 1;  // parity of every l in this pass
    const Ctx c = make_ctx(smem);
    load_tables(P, c, pass);
    float *red = reinterpret_cast<float *>(smem + OFF_RED);
    float *V = reinterpret_cast<float *>(smem + mp.V);

    const float *r0p[TS], *r1p[TS], *p0p[TS], *p1p[TS];
#pragma unroll
    for (int s = 0; s < TS; ++s) {
        const TileDesc t = P.tiles[(pass * TS + s) * NT + tid];
        r0p[s] = c.Rs + t.r0; r1p[s] = c.Rs + t.r1;
        p0p[s] = c.Ps + t.p0; p1p[s] = c.Ps + t.p1;
    }

    // a tile whose column halves are loaded in swapped order sees m = 2, 3 where the others see m = 0, 1: for the
    // transposed row group that is one overall sign (see the class accumulation below)
    const float bsign = (p0p[0] > p1p[0]) ? -1.0f : 1.0f;
    const int nG = analysis_items(half_lo, half_hi, H);  // row groups, paired with their transposes inside the slab
    const size_t N3 = (size_t)N * N * N;
    for (int b = 0; b < batch; ++b) {
        unsigned long long om[TS][32];  // moment tiles, same pair layout as the analysis accumulators
        const float *ob = otiles + ((size_t)b * P.npass + pass) * SLOT_WORDS;
#pragma unroll
        for (int s = 0; s < TS; ++s) {
            const ulonglong2 *o = reinterpret_cast<const ulonglong2 *>(ob + (size_t)(s * NT + tid) * TW);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const ulonglong2 x = __ldg(o + q);
                om[s][2 * q] = x.x;
                om[s][2 * q + 1] = x.y;
            }
        }
        float *pv_out = pvol + ((size_t)pass * batch + b) * N3;

        for (int g = split; g < nG; g += P.nsplit) {
            int ip, jp;
            bool paired;
            decode_item(g, half_lo, half_hi, H, ip, jp, paired);
            const double y0 = P.side[ip], x0 = P.side[jp];
            const double rho2 = x0 * x0 + y0 * y0, rho = sqrt(rho2);
            __syncthreads();
            row_trig(P, c.cm, c.sm, x0, y0);
            row_geometry(P, c, rho2, rho);
            __syncthreads();

            for (int k0 = 0; k0 < H; k0 += KC) {
                generate_panels<NPH, false>(P, c, nullptr, k0, pass, warp, lane);
                __syncthreads();
#pragma unroll 1
                for (int kq = 0; kq < KC / 2; ++kq) {  // quarter chunks keep the class accumulators at 12 registers
                    // Per voxel and column pair j = (re, im) of m = m0 + j, m0 = 0 (mod 4):  tt = (t_re, t_im).
                    //   row group A (phi):         (P_re t_re, P_im t_im)                       -> E (m even) / O (m odd)
                    //   transpose  B (pi/2 - phi): cos / sin(m phi') are a signed swap of cos / sin(m phi):
                    //     m = 0: (+E.re, -E.im)   m = 2: (-E.re, +E.im)   m = 1: +swap(P) tt   m = 3: -swap(P) tt
                    // so B costs two extra FFMA2 per voxel (the cross products X of the odd m) instead of a second panel.
                    unsigned long long E0[2], E2[2], O1[2], O3[2], X1[2], X3[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) E0[q] = E2[q] = O1[q] = O3[q] = X1[q] = X3[q] = 0ull;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int k = kq * 2 + kk;
#pragma unroll
                        for (int s = 0; s < TS; ++s) {
                            const ulonglong2 ra = *reinterpret_cast<const ulonglong2 *>(r0p[s] + k * RS);
                            const ulonglong2 rb = *reinterpret_cast<const ulonglong2 *>(r1p[s] + k * RS);
                            const ulonglong2 pa = *reinterpret_cast<const ulonglong2 *>(p0p[s] + k * PS);
                            const ulonglong2 pb = *reinterpret_cast<const ulonglong2 *>(p1p[s] + k * PS);
                            const unsigned long long Rp[4] = {ra.x, ra.y, rb.x, rb.y};
                            const unsigned long long Pp[4] = {pa.x, pa.y, pb.x, pb.y};
                            unsigned long long tt[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                unsigned long long tD = fmul2(Rp[0], om[s][j * 2]);
                                unsigned long long tX = fmul2(Rp[0], om[s][j * 2 + 1]);
#pragma unroll
                                for (int i = 1; i < 4; ++i) {
                                    ffma2(tD, Rp[i], om[s][(i * 4 + j) * 2]);
                                    ffma2(tX, Rp[i], om[s][(i * 4 + j) * 2 + 1]);
                                }
                                tt[j] = fadd2(tD, swap2(tX));
                            }
                            ffma2(E0[kk], Pp[0], tt[0]);
                            ffma2(O1[kk], Pp[1], tt[1]);
                            ffma2(X1[kk], swap2(Pp[1]), tt[1]);
                            ffma2(E2[kk], Pp[2], tt[2]);
                            ffma2(O3[kk], Pp[3], tt[3]);
                            ffma2(X3[kk], swap2(Pp[3]), tt[3]);
                        }
                    }
                    float v[16];  // [k in quarter][A classes j0..j3 | B classes j0..j3]
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const float2 e0 = unpack2(E0[kk]), e2 = unpack2(E2[kk]), o1 = unpack2(O1[kk]), o3 = unpack2(O3[kk]);
                        const float2 x1 = unpack2(X1[kk]), x3 = unpack2(X3[kk]);
                        v[kk * 8 + 0] = e0.x + e2.x;
                        v[kk * 8 + 1] = e0.y + e2.y;
                        v[kk * 8 + 2] = o1.x + o3.x;
                        v[kk * 8 + 3] = o1.y + o3.y;
                        v[kk * 8 + 4] = bsign * (e0.x - e2.x);
                        v[kk * 8 + 5] = bsign * (e2.y - e0.y);
                        v[kk * 8 + 6] = bsign * (x1.x - x3.x);
                        v[kk * 8 + 7] = bsign * (x1.y - x3.y);
                    }
                    const float tot = warp_transpose_reduce16(v, lane);
                    if (lane < 16) red[(kq * NWARP + warp) * 16 + lane] = tot;
                }
                __syncthreads();
                if (tid < 64) {
                    const int kq = tid >> 4, q = tid & 15;
                    float s = 0.0f;
#pragma unroll
                    for (int w = 0; w < NWARP; ++w) s += red[(kq * NWARP + w) * 16 + q];
                    const int k = kq * 2 + (q >> 3), cls = q & 7;
                    if (k0 + k < H) V[cls * H + k0 + k] = s;
                }
            }
            __syncthreads();
            // ---- unfold the 4 class sums of this pass into the 8 mirror voxels of the row group (and of its transpose).
            // mirror class c = py*4 + px*2 + pz of class column j: j0 (m even, re) -> par, j1 (m even, im) -> 6|par,
            // j2 (m odd, re) -> 2|(par^1), j3 (m odd, im) -> 4|(par^1)
            for (int idx = tid; idx < (paired ? 2 : 1) * H; idx += NT) {
                const int set = idx / H, k = idx - set * H;
                const int k2 = N - 1 - k;
                const int ri = set ? jp : ip, rj = set ? ip : jp, ri2 = N - 1 - ri, rj2 = N - 1 - rj;
                const float *Vs = V + set * 4 * H;
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = 0.0f;
                const float V0 = Vs[k], V1 = Vs[H + k], V2 = Vs[2 * H + k], V3 = Vs[3 * H + k];
                if (par == 0) { v[0] = V0; v[6] = V1; v[3] = V2; v[5] = V3; }
                else          { v[1] = V0; v[7] = V1; v[2] = V2; v[4] = V3; }
                wht8(v);
#pragma unroll
                for (int img = 0; img < 8; ++img) {
                    const bool sy = img & 4, sx = img & 2, sz = img & 1;
                    if ((sy && ri2 == ri) || (sx && rj2 == rj) || (sz && k2 == k)) continue;
                    const int ii = sy ? ri2 : ri, jj = sx ? rj2 : rj, kk = sz ? k2 : k;
                    pv_out[((size_t)ii * N + jj) * N + kk] = v[img];
                }
            }
        }
    }
}

// out[b][planes of pairs [half_lo, half_hi)] = sum over passes of the per-pass partial volumes (fixed order)
__global__ void k_sum_passes(const float *__restrict__ pvol, int npass, int batch, int N, int half_lo, int half_hi,
                             float *__restrict__ out) {
    const size_t plane = (size_t)N * N;
    const int np = half_hi - half_lo;
    const size_t per_b = (size_t)2 * np * plane;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_b * batch) return;
    const int b = (int)(idx / per_b);
    const size_t r = idx % per_b;
    const int pl = (int)(r / plane);
    const size_t q = r % plane;
    const int p = half_lo + (pl >> 1);
    const int i = (pl & 1) ? (N - 1 - p) : p;
    if ((pl & 1) && i == p) return;  // odd N: centre plane once
    const size_t o = (size_t)i * plane + q;
    const size_t N3 = plane * N;
    float s = 0.0f;
    for (int ps = 0; ps < npass; ++ps) s += pvol[((size_t)ps * batch + b) * N3 + o];
    out[(size_t)b * N3 + o] = s;
}

// F_nl = sqrt(|O_nl0|^2 + 2 sum_{m>0} |O_nlm|^2)   (Moments.calculate_invariants zm:62-76)
__global__ void k_invariants(const float *__restrict__ moments, const int *__restrict__ nl_base,
                             const int *__restrict__ nl_l, int nrad, int nmom, int batch, float *__restrict__ inv) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (q >= nrad || b >= batch) return;
    const float2 *m = reinterpret_cast<const float2 *>(moments) + (size_t)b * nmom + nl_base[q];
    const int l = nl_l[q];
    double s = 0.0;
    for (int mm = 0; mm <= l; ++mm) {
        const float2 v = m[mm];
        const double a2 = (double)v.x * v.x + (double)v.y * v.y;
        s += (mm == 0) ? a2 : 2.0 * a2;
    }
    inv[(size_t)b * nrad + q] = (float)sqrt(s);
}

// "Zernike FSC" of two moment sets (moment_analysis_3dva.py calc_ZCC, ma:402-422): per order n, over its stored moments (m >= 0),
//   c_n = sum |A conj(B)| / sqrt(sum |A|^2 * sum |B|^2).        One warp per (order, pair); float64 sums; 0 / 0 -> NaN like numpy.
__global__ void k_zcc(const float *__restrict__ ma, const float *__restrict__ mb, int order, int nmom, int batch,
                      float *__restrict__ out) {
    const int n = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
    int base = 0;  // moments of the orders < n: sum over k < n of sum over l = k, k-2, .. of (l + 1)
    for (int k = 0; k < n; ++k)
        for (int l = k & 1; l <= k; l += 2) base += l + 1;
    int cnt = 0;
    for (int l = n & 1; l <= n; l += 2) cnt += l + 1;
    const float2 *A = reinterpret_cast<const float2 *>(ma) + (size_t)b * nmom + base;
    const float2 *B = reinterpret_cast<const float2 *>(mb) + (size_t)b * nmom + base;
    double nom = 0.0, d1 = 0.0, d2 = 0.0;
    for (int i = lane; i < cnt; i += 32) {
        const float2 a = A[i], c = B[i];
        const double a2 = (double)a.x * a.x + (double)a.y * a.y, b2 = (double)c.x * c.x + (double)c.y * c.y;
        nom += sqrt(a2 * b2);  // |A conj(B)| = |A| |B|
        d1 += a2;
        d2 += b2;
    }
    for (int o = 16; o; o >>= 1) {
        nom += __shfl_xor_sync(0xffffffffu, nom, o);
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    if (lane == 0) out[(size_t)b * (order + 1) + n] = (float)(nom / sqrt(d1 * d2));
}

// Pre-processing of an ensemble on the device (ZernikeSpherical.bin_crop_array_center zm:2781-2815, run() zm:2867-2876): centred
// crop, then scipy.ndimage.zoom(v, 1 / bin, order = 0): output index i samples input coordinate x = i (nc - 1) / (n_out - 1)
// (float64), nearest neighbour floor(x + 0.5), and - scipy's mode = 'constant' - 0 when x falls outside [0, nc - 1], which
// rounding makes happen for the last sample of some sizes (e.g. 32 -> 16).  One thread per output voxel.
__device__ __forceinline__ int zoom0_index(int i, int nc, int n_out, bool &inside) {
    if (n_out == nc) {
        inside = true;
        return i;
    }
    const double x = n_out > 1 ? (double)i * ((double)(nc - 1) / (double)(n_out - 1)) : 0.0;
    inside = x >= 0.0 && x <= (double)(nc - 1);
    return min(max((int)floor(x + 0.5), 0), nc - 1);
}
__global__ void k_crop_bin(const float *__restrict__ in, int n_in, int oz, int oy, int ox, int nc, int n_out, float *__restrict__ out) {
    const size_t per = (size_t)n_out * n_out * n_out;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per) return;
    const int b = blockIdx.y;
    const int k = (int)(e % n_out), j = (int)((e / n_out) % n_out), i = (int)(e / ((size_t)n_out * n_out));
    bool a, c, d;
    const int zi = zoom0_index(i, nc, n_out, a), yi = zoom0_index(j, nc, n_out, c), xi = zoom0_index(k, nc, n_out, d);
    const float v = in[(size_t)b * n_in * n_in * n_in + ((size_t)(oz + zi) * n_in + (oy + yi)) * n_in + (ox + xi)];
    out[(size_t)b * per + e] = (a && c && d) ? v : 0.0f;
}

// FP32 FMA throughput probes for the roofline denominator
// 3DVA latent scaling: out[p][e] = dims[0][e] + sum_k lat[p][k] * dims[k+1][e] over the 2 * nmom float components (zm:78-95)
__global__ void k_scale_moments(const float *__restrict__ dims, int ncomp, const float *__restrict__ lat, int npoints, int nelem,
                                float *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (e >= nelem || p >= npoints) return;
    float acc = dims[e];
    for (int k = 0; k < ncomp; ++k) acc = fmaf(__ldg(lat + (size_t)p * ncomp + k), dims[(size_t)(k + 1) * nelem + e], acc);
    out[(size_t)p * nelem + e] = acc;
}

__global__ void k_fma_peak(int iters, float *out, int variant) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float x = 0.999f, y = 1e-4f;
    if (variant == 0) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 = fmaf(a0, x, y); a1 = fmaf(a1, x, y); a2 = fmaf(a2, x, y); a3 = fmaf(a3, x, y);
                a4 = fmaf(a4, x, y); a5 = fmaf(a5, x, y); a6 = fmaf(a6, x, y); a7 = fmaf(a7, x, y);
            }
        }
    } else {
        unsigned long long p0 = pack2(a0, a1), p1 = pack2(a2, a3), p2 = pack2(a4, a5), p3 = pack2(a6, a7);
        const unsigned long long px = pack2(x, x), py = pack2(y, y);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(px), "l"(py));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(px), "l"(py));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(px), "l"(py));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(px), "l"(py));
            }
        }
        const float2 f0 = unpack2(p0), f1 = unpack2(p1), f2 = unpack2(p2), f3 = unpack2(p3);
        a0 = f0.x; a1 = f0.y; a2 = f1.x; a3 = f1.y; a4 = f2.x; a5 = f2.y; a6 = f3.x; a7 = f3.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace z3d
