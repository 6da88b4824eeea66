// zernike3d_b200 - fused sm_100a kernels for the spharm 3D-Zernike analysis / synthesis path.
//
// Neither kernel ever materialises an R_nl or Y_lm voxel grid (the reference builds 676 + 1326 of them,
// zm:2281-2323).  Per chunk of KC folded voxels a CTA
//   (0) computes r, cos(theta), sin(theta) once per voxel in float64 (zm:1886-1890) and the per-row
//       cos(m phi), sin(m phi) once per row group,
//   (1) runs the reference's recurrences in registers - R_nl in n at fixed l (zm:1980-2002), the
//       normalised associated Legendre functions in l at fixed m (zm:2006-2026, in monic form, the
//       normalisation kappa_lm is applied once to the finished sums) - and stages the two operand
//       panels R[n][k] and a*P[(m,re/im)][k] of the per-l contraction in shared memory,
//   (2) contracts them over the voxel axis into register-resident 4x4 output tiles (analysis), or
//       contracts register-resident moment tiles with them into 8 mirror-class sums (synthesis).
#pragma once
#include "z3d_plan.h"

namespace z3d {

// ---------------------------------------------------------------------------------------------
// shared-memory carve-up (identical on host and device)
// ---------------------------------------------------------------------------------------------
struct SmemMap {
    int Rs, Ps, a4, geo, cm, sm, rows, F, tabs, red, V, total;  // byte offsets
};
__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline SmemMap smem_map(int L, int N, int H) {
    SmemMap m;
    int o = 0;
    m.Rs = o;   o += KC * RS * 4;
    m.Ps = o;   o += KC * PS * 4;
    m.a4 = o;   o += align16(KC * (L + 1) * 16);
    m.geo = o;  o += align16(5 * KC * 4);            // c2, st, st2, r, r2
    m.cm = o;   o += align16((L + 1) * 4);
    m.sm = o;   o += align16((L + 1) * 4);
    m.rows = o; o += align16(4 * N * 4);             // analysis: 4 staged rows
    m.F = o;    o += align16(8 * H * 4);             // analysis: 8 mirror-folded rows
    m.tabs = o; o += align16(3 * (L + 1) * 4);       // lpass, poffw, roffw
    m.red = o;  o += NWARP * 32 * 4;                 // synthesis: cross-warp reduction
    m.V = o;    o += align16(4 * H * 4);             // synthesis: class sums of one row group
    m.total = o;
    return m;
}

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

struct Ctx {
    float *Rs, *Ps;
    float4 *a4;
    float *c2, *st, *st2, *r, *r2, *cm, *sm;
    int *lpass, *poffw, *roffw;
};

// Step (1): every warp advances its share of the recurrence chains for the KC voxels of the chunk.
// lane -> (voxel gk = lane % KC, chain gsub = lane / KC); tasks are warp-uniform.
__device__ __forceinline__ void generate_panels(const DevPlan &P, const Ctx &c, int pass, int warp, int lane) {
    const int L = P.order;
    const int gk = lane % KC, gsub = lane / KC;
    const int *tasks = P.gen_tasks + (pass * NWARP + warp) * MAXTASK;
    for (int t = 0; t < MAXTASK; ++t) {
        const int task = __ldg(tasks + t);
        const int type = task >> 16, base = task & 0xffff;
        if (type == 0) break;
        if (type == 1) {
            // Legendre chains m = base + gsub, l = m..L.  Monic form: Q_l = 2c Q_{l-1} - beta_l Q_{l-2},
            // Q_mm = sin^m; P_lm = kappa_lm Q_lm (Y_lm_recurrence zm:2012-2026 divided by kappa).
            const int m = base + gsub;
            const bool active = (m <= L);
            const int mm = active ? m : L;
            const float c2 = c.c2[gk];
            const float seed = powi2(c.st[gk], c.st2[gk], mm);
            const float4 a = c.a4[gk * (L + 1) + mm];
            const float *bp = P.beta + mm * (L + 1);
            float *dst = c.Ps + gk * PS + 2 * mm;
            float q1 = 0.0f, q2 = 0.0f;
            for (int l = base; l <= L; ++l) {
                const float b = __ldg(bp + l);
                float q = fmaf(c2, q1, -b * q2);
                q = (l == mm) ? seed : q;
                q2 = q1;
                q1 = q;
                if (c.lpass[l] == pass && active && l >= m) {
                    const bool odd = (l + m) & 1;
                    const float are = odd ? a.z : a.x, aim = odd ? a.w : a.y;
                    *reinterpret_cast<float2 *>(dst + c.poffw[l]) = make_float2(are * q, aim * q);
                }
            }
        } else {
            // Radial chains for own l = ownl[base + gsub]: R_ll = r^l, then
            // R_n = (K1 r^2 + K2) R_{n-2} + K3 R_{n-4}   (R_nl_recurrence zm:1983-2000; the l == n-2
            // branch is the same formula with K3 = 0).
            const int oi = base + gsub;
            const int nown = __ldg(P.nown + pass);
            const bool active = oi < nown;
            const int l = __ldg(P.ownl + pass * (L + 1) + (active ? oi : base));
            const int l0 = __ldg(P.ownl + pass * (L + 1) + base);
            const int nmax = (L - l0) / 2 + 1;
            const int nn = active ? (L - l) / 2 + 1 : 0;
            const float r2 = c.r2[gk];
            float R1 = powi2(c.r[gk], r2, l), R2 = 0.0f;
            float *dst = c.Rs + gk * RS + c.roffw[l];
            const float4 *kp = P.rk + __ldg(P.rbase + l);
            if (active) dst[0] = R1;
            for (int ni = 1; ni < nmax; ++ni) {
                if (ni < nn) {
                    const float4 K = __ldg(kp + ni);
                    const float R = fmaf(fmaf(K.x, r2, K.y), R1, K.z * R2);
                    dst[ni] = R;
                    R2 = R1;
                    R1 = R;
                }
            }
        }
    }
}

// Step (0) geometry for the KC voxels of a chunk; float64 like the reference (zm:2240-2242, 1886-1890).
__device__ __forceinline__ void chunk_geometry(const DevPlan &P, const Ctx &c, int k, int kk, double rho2, double rho) {
    const double z = (kk < P.H) ? P.side[kk] : 0.0;
    const double r2d = rho2 + z * z;
    const double r = sqrt(r2d);
    double ct = 1.0, st = 0.0;  // r == 0: theta is NaN in the reference and nan_to_num zeroes Y_lm, l > 0
    if (r > 0.0) {             // (zm:2298); R_nl(0) = 0 for l > 0 gives the same product here.
        ct = z / r;
        st = rho / r;
    }
    c.c2[k] = (float)(2.0 * ct);
    c.st[k] = (float)st;
    c.st2[k] = (float)(st * st);
    c.r[k] = (float)r;
    c.r2[k] = (float)r2d;
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

__device__ __forceinline__ Ctx make_ctx(unsigned char *smem, const SmemMap &mp) {
    Ctx c;
    c.Rs = reinterpret_cast<float *>(smem + mp.Rs);
    c.Ps = reinterpret_cast<float *>(smem + mp.Ps);
    c.a4 = reinterpret_cast<float4 *>(smem + mp.a4);
    float *geo = reinterpret_cast<float *>(smem + mp.geo);
    c.c2 = geo;
    c.st = geo + KC;
    c.st2 = geo + 2 * KC;
    c.r = geo + 3 * KC;
    c.r2 = geo + 4 * KC;
    c.cm = reinterpret_cast<float *>(smem + mp.cm);
    c.sm = reinterpret_cast<float *>(smem + mp.sm);
    return c;
}

__device__ __forceinline__ void load_small_tables(const DevPlan &P, Ctx &c, unsigned char *smem, const SmemMap &mp) {
    int *tabs = reinterpret_cast<int *>(smem + mp.tabs);
    const int L = P.order;
    c.lpass = tabs;
    c.poffw = tabs + (L + 1);
    c.roffw = tabs + 2 * (L + 1);
    for (int i = threadIdx.x; i <= L; i += NT) {
        c.lpass[i] = P.lpass[i];
        c.poffw[i] = P.poffw[i];
        c.roffw[i] = P.roffw[i];
    }
}

// cos(m phi0), sin(m phi0) for the row group, float64 like exp(1j m phi) in zm:2297
__device__ __forceinline__ void row_trig(const DevPlan &P, const Ctx &c, double x0, double y0) {
    if (threadIdx.x <= P.order) {
        const double ph = atan2(y0, x0);
        double s, co;
        sincos((double)threadIdx.x * ph, &s, &co);
        c.cm[threadIdx.x] = (float)co;
        c.sm[threadIdx.x] = (float)s;
    }
}

// ---------------------------------------------------------------------------------------------
// Analysis: per-CTA partial sums of  sum_v f R_nl P_lm e^{-i m phi}  for one pass and one split.
// partial layout: [batch][npass][nsplit][SLOT_WORDS]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, CTAS_PER_SM)
k_decompose(const DevPlan P, const float *__restrict__ vol, int batch, int half_lo, int half_hi,
            float *__restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int L = P.order, N = P.dim, H = P.H;
    const SmemMap mp = smem_map(L, N, H);
    Ctx c = make_ctx(smem, mp);
    load_small_tables(P, c, smem, mp);
    float *rows = reinterpret_cast<float *>(smem + mp.rows);
    float *F = reinterpret_cast<float *>(smem + mp.F);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pass = blockIdx.x % P.npass, split = blockIdx.x / P.npass;
    const int ntile = P.ntiles[pass];

    const float *rptr[TS], *pptr[TS];
#pragma unroll
    for (int s = 0; s < TS; ++s) {
        const int2 t = P.tiles[(pass * TS + s) * NT + tid];
        rptr[s] = c.Rs + t.x;
        pptr[s] = c.Ps + t.y;
    }

    const int nG = (half_hi - half_lo) * H;
    const size_t N3 = (size_t)N * N * N;
    for (int b = 0; b < batch; ++b) {
        float acc[TS][16];
#pragma unroll
        for (int s = 0; s < TS; ++s)
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[s][e] = 0.0f;
        const float *vb = vol + (size_t)b * N3;

        for (int g = split; g < nG; g += P.nsplit) {
            const int ip = half_lo + g / H, jp = g % H;
            const int i2 = N - 1 - ip, j2 = N - 1 - jp;
            const double y0 = P.side[ip], x0 = P.side[jp];
            const double rho2 = x0 * x0 + y0 * y0, rho = sqrt(rho2);

            // ---- stage the 4 rows (+-y, +-x) of this row group: coalesced float4 loads -> smem
            __syncthreads();
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
            row_trig(P, c, x0, y0);
            __syncthreads();
            // ---- fold the 8 mirror images: F[c][k] = sum_img f_img (-1)^popcount(c & img),
            //      img = sy*4 + sx*2 + sz,  c = py*4 + px*2 + pz
            for (int k = tid; k < H; k += NT) {
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
            __syncthreads();

            for (int k0 = 0; k0 < H; k0 += KC) {
                // ---- step 0: geometry + a-table (f folded * trig) for the chunk
                if (tid < KC) chunk_geometry(P, c, tid, k0 + tid, rho2, rho);
                for (int idx = tid; idx < KC * (L + 1); idx += NT) {
                    const int k = idx % KC, m = idx / KC, kk = k0 + k;
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (kk < H) {
                        const int px = m & 1;
                        const float cmv = c.cm[m], smv = -c.sm[m];
                        // re: even in y, (-1)^m in x;  im: odd in y, -(-1)^m in x;  z parity = (l+m)&1
                        a.x = F[((px << 1) | 0) * H + kk] * cmv;
                        a.z = F[((px << 1) | 1) * H + kk] * cmv;
                        a.y = F[(4 | ((px ^ 1) << 1) | 0) * H + kk] * smv;
                        a.w = F[(4 | ((px ^ 1) << 1) | 1) * H + kk] * smv;
                    }
                    c.a4[k * (L + 1) + m] = a;
                }
                __syncthreads();
                // ---- step 1: recurrences -> operand panels in smem
                generate_panels(P, c, pass, warp, lane);
                __syncthreads();
                // ---- step 2: contraction over the chunk's voxels into the register tiles
#pragma unroll
                for (int k = 0; k < KC; ++k) {
#pragma unroll
                    for (int s = 0; s < TS; ++s) {
                        const float4 rv = *reinterpret_cast<const float4 *>(rptr[s] + k * RS);
                        const float4 pv = *reinterpret_cast<const float4 *>(pptr[s] + k * PS);
                        const float rr[4] = {rv.x, rv.y, rv.z, rv.w};
                        const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[s][i * 4 + j] = fmaf(rr[i], pp[j], acc[s][i * 4 + j]);
                    }
                }
            }
        }
        float *out = partial + (((size_t)b * P.npass + pass) * P.nsplit + split) * SLOT_WORDS;
#pragma unroll
        for (int s = 0; s < TS; ++s) {
            if (s * NT + tid < ntile) {
                float4 *o = reinterpret_cast<float4 *>(out + (size_t)(s * NT + tid) * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    o[q] = make_float4(acc[s][4 * q], acc[s][4 * q + 1], acc[s][4 * q + 2], acc[s][4 * q + 3]);
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

// lane L ends up with sum over lanes of v[L]  (31 shuffles instead of 32 * 5)
__device__ __forceinline__ float warp_transpose_reduce32(float v[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
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
__global__ void __launch_bounds__(NT, CTAS_PER_SM)
k_reconstruct(const DevPlan P, const float *__restrict__ otiles, int batch, int half_lo, int half_hi,
              float *__restrict__ pvol) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int L = P.order, N = P.dim, H = P.H;
    const SmemMap mp = smem_map(L, N, H);
    Ctx c = make_ctx(smem, mp);
    load_small_tables(P, c, smem, mp);
    float *red = reinterpret_cast<float *>(smem + mp.red);
    float *V = reinterpret_cast<float *>(smem + mp.V);
    // padding rows / columns of the panels are never written by the recurrences; they meet zero moment
    // entries in the tiles, so they must be finite: clear both panels once.
    for (int i = threadIdx.x; i < KC * (RS + PS); i += NT) c.Rs[i] = 0.0f;  // Rs and Ps are contiguous

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pass = blockIdx.x % P.npass, split = blockIdx.x / P.npass;
    const int par = pass & 1;  // parity of every l in this pass

    const float *rptr[TS], *pptr[TS];
#pragma unroll
    for (int s = 0; s < TS; ++s) {
        const int2 t = P.tiles[(pass * TS + s) * NT + tid];
        rptr[s] = c.Rs + t.x;
        pptr[s] = c.Ps + t.y;
    }
    // mirror class c = py*4 + px*2 + pz of tile column j (m0 even): j0 (m0,re) -> par, j1 (m0,im) -> 6|par,
    // j2 (m0+1,re) -> 2|(par^1), j3 (m0+1,im) -> 4|(par^1)

    const int nG = (half_hi - half_lo) * H;
    const size_t N3 = (size_t)N * N * N;
    for (int b = 0; b < batch; ++b) {
        float om[TS][16];
        const float *ob = otiles + ((size_t)b * P.npass + pass) * SLOT_WORDS;
#pragma unroll
        for (int s = 0; s < TS; ++s) {
            const float4 *o = reinterpret_cast<const float4 *>(ob + (size_t)(s * NT + tid) * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 x = __ldg(o + q);
                om[s][4 * q] = x.x; om[s][4 * q + 1] = x.y; om[s][4 * q + 2] = x.z; om[s][4 * q + 3] = x.w;
            }
        }
        float *pv_out = pvol + ((size_t)pass * batch + b) * N3;

        for (int g = split; g < nG; g += P.nsplit) {
            const int ip = half_lo + g / H, jp = g % H;
            const int i2 = N - 1 - ip, j2 = N - 1 - jp;
            const double y0 = P.side[ip], x0 = P.side[jp];
            const double rho2 = x0 * x0 + y0 * y0, rho = sqrt(rho2);
            __syncthreads();
            row_trig(P, c, x0, y0);
            __syncthreads();

            for (int k0 = 0; k0 < H; k0 += KC) {
                if (tid < KC) chunk_geometry(P, c, tid, k0 + tid, rho2, rho);
                for (int idx = tid; idx < KC * (L + 1); idx += NT) {
                    const int k = idx % KC, m = idx / KC;
                    const float cmv = c.cm[m], smv = c.sm[m];
                    c.a4[k * (L + 1) + m] = make_float4(cmv, smv, cmv, smv);
                }
                __syncthreads();
                generate_panels(P, c, pass, warp, lane);
                __syncthreads();
                float vacc[32];  // [k][j]
#pragma unroll
                for (int q = 0; q < 32; ++q) vacc[q] = 0.0f;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
#pragma unroll
                    for (int s = 0; s < TS; ++s) {
                        const float4 rv = *reinterpret_cast<const float4 *>(rptr[s] + k * RS);
                        const float4 pv = *reinterpret_cast<const float4 *>(pptr[s] + k * PS);
                        const float rr[4] = {rv.x, rv.y, rv.z, rv.w};
                        const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float t = om[s][j] * rr[0];
#pragma unroll
                            for (int i = 1; i < 4; ++i) t = fmaf(om[s][i * 4 + j], rr[i], t);
                            vacc[k * 4 + j] = fmaf(pp[j], t, vacc[k * 4 + j]);
                        }
                    }
                }
                const float tot = warp_transpose_reduce32(vacc, lane);
                red[warp * 32 + lane] = tot;
                __syncthreads();
                if (tid < 32) {
                    float s = 0.0f;
#pragma unroll
                    for (int w = 0; w < NWARP; ++w) s += red[w * 32 + tid];
                    const int k = tid >> 2, j = tid & 3;
                    if (k0 + k < H) V[j * H + k0 + k] = s;
                }
            }
            __syncthreads();
            // ---- unfold the 4 class sums of this pass into the 8 mirror voxels
            for (int k = tid; k < H; k += NT) {
                const int k2 = N - 1 - k;
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = 0.0f;
                const float V0 = V[k], V1 = V[H + k], V2 = V[2 * H + k], V3 = V[3 * H + k];
                if (par == 0) { v[0] = V0; v[6] = V1; v[3] = V2; v[5] = V3; }
                else          { v[1] = V0; v[7] = V1; v[2] = V2; v[4] = V3; }
                wht8(v);
#pragma unroll
                for (int img = 0; img < 8; ++img) {
                    const bool sy = img & 4, sx = img & 2, sz = img & 1;
                    if ((sy && i2 == ip) || (sx && j2 == jp) || (sz && k2 == k)) continue;
                    const int ii = sy ? i2 : ip, jj = sx ? j2 : jp, kk = sz ? k2 : k;
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

// FP32 FMA throughput probes for the roofline denominator
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
        unsigned long long p0, p1, p2, p3, px, py;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(px) : "f"(x), "f"(x));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(py) : "f"(y), "f"(y));
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(px), "l"(py));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(px), "l"(py));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(px), "l"(py));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(px), "l"(py));
            }
        }
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a4), "=f"(a5) : "l"(p2));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a6), "=f"(a7) : "l"(p3));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace z3d
