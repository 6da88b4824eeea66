// zernike3d_b200 - host planner and C ABI (see include/zernike3d_b200.h for the contract).
#include "../../include/zernike3d_b200.h"
#include "z3d_kernels.cuh"
#include "z3d_batched.cuh"
#include "z3d_stream.cuh"
#include "z3d_gen.cuh"
#include "z3d_synth.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace z3d;

namespace {

thread_local std::string g_err;
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(Z3D_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

inline int num_moments(int order) {
    int c = 0;
    for (int n = 0; n <= order; ++n)
        for (int l = n & 1; l <= n; l += 2) c += l + 1;
    return c;
}
inline int num_radial(int order) {
    int c = 0;
    for (int n = 0; n <= order; ++n) c += n / 2 + 1;
    return c;
}

// Every entry point runs on the plan's device and restores the caller's current device on return (a library call must not
// change torch.cuda.current_device() of the calling thread).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) err = cudaSetDevice(dev);
        else prev = -1;  // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define USE_DEVICE(dev)            \
    DeviceGuard _guard(dev);       \
    CUDA_TRY(_guard.err)

template <class T>
int upload(const std::vector<T> &h, T **d, std::vector<void *> &owned) {
    CUDA_TRY(cudaMalloc((void **)d, std::max<size_t>(1, h.size()) * sizeof(T)));
    owned.push_back(*d);
    if (!h.empty()) CUDA_TRY(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return Z3D_OK;
}

}  // namespace

struct z3d_plan {
    int order = 0, dim = 0, device = 0, H = 0, nmom = 0, nrad = 0, npass = 0, nsplit = 0, num_sms = 0;
    int smem_bytes = 0, useful_acc = 0, padded_acc = 0;
    DevPlan dev{};
    // batched analysis kernels (tiles of 32 or 64 volumes); absent when the order does not fit their fixed layout
    bool batched = false;
    int batched_min = 12;        // smallest (remaining) batch that is routed to the batched kernels
    int batched_wide = 33;       // smallest (remaining) batch that takes a 64-volume tile instead of a 32-volume one
    DevPlanB devb[2]{};          // [0]: 32 volumes per tile (8 tiles of 128 rows per row group), [1]: 64 (4 tiles)
    int2 *d_rowmap[2] = {nullptr, nullptr};
    int b_smem[2] = {0, 0};
    long long b_chunks = 0;      // chunks of the whole volume: row-group items x chunks per row
    BasisArgs basis{};           // inputs of k_build_basis; the tables themselves are built on first use
    bool basis_ready = false;
    float *d_Rt = nullptr, *d_Qt = nullptr, *d_trig = nullptr;
    // streamed-T variant of the batched path (z3d_stream.cuh): preferred whenever its T table fits in device memory
    bool stream = false, stream_ready = false;
    DevPlanS devs{};
    int s_smem = 0, s_ncols_all = 0;
    int4 *d_colsrc = nullptr;
    int2 *d_colmap = nullptr;
    float *d_Tt = nullptr;
    int tables_level = 0;        // 0: no basis tables (default), 1: R / Q tables, 2: R / Q + T tables (z3d_plan_prepare_tables)
    // table-free tensor-core variant (z3d_gen.cuh): the default for batches
    bool gen = false;
    int gen_min = 8;             // smallest (remaining) batch routed to it: a tile costs the same whatever it holds (6.4 per-volume launches at 128^3 / order 50)
    DevPlanG devg{};
    int g_smem = 0, g_ncols_all = 0, g_maxcols = 0;
    GenIssue g_issue{};          // per-set MMA issue table, passed in the kernel parameters
    int2 *d_gcolmap = nullptr;
    // ... its two-tile layout (128 volumes per generated T block: half the columns per set, half the splits), used while more than
    // one tile of volumes is left; absent when the order needs more sets than SMs (order > ~70)
    bool gen2 = false;           // the analysis kernel uses it (opt-in, Z3D_GEN2=1)
    bool g2_built = false;       // the layout exists (the tensor-core synthesis kernel runs on it)
    DevPlanG devg2{};
    int g2_smem = 0;
    GenIssue g2_issue{};
    int2 *d_g2colmap[2] = {nullptr, nullptr};   // (set, accumulator column) of every moment, per tile
    // tensor-core batched synthesis (z3d_synth.cuh) on the two-tile layout's column sets
    bool synth = false;
    int synth_min = 12;          // smallest batch of maps routed to it by the host layer (a tile costs the same whatever it holds)
    int synth_maxcols = 0, synth_off_t = 0, synth_slots = 0, synth_smem = 0;
    int *d_synth_colsrc = nullptr;   // [sets][maxcols] -> moment index or -1
    KindSets synth_kinds{};         // first partial-fold SLOT of every kind
    int *d_synth_pairinfo = nullptr; // [sets]: slot << 1 | exchange with the cluster partner (set ^ 1, same kind)
    int synth_nslots = 0;
    // reduce / prepare tables
    int *d_src = nullptr;        // [2*nmom]  pass*SLOT_WORDS + word
    double *d_scale = nullptr;   // [nmom]    coeff(n) * kappa_lm
    int *d_omap = nullptr;       // [npass*SLOT_WORDS] -> 2*mu+reim or -1
    double *d_ofac = nullptr;    // [nmom]    kappa_lm * w_m
    int *d_nl_base = nullptr, *d_nl_l = nullptr;  // invariants
    std::vector<void *> owned;
    // host-call staging
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    float *d_stage_vol[2] = {nullptr, nullptr};
    float *d_stage_mom[2] = {nullptr, nullptr};
    void *d_stage_ws[2] = {nullptr, nullptr};
    void *d_stage_synth_ws = nullptr;   // workspace of the tensor-core synthesis kernel for z3d_reconstruct_host (allocated on first use)
    size_t stage_synth_bytes = 0;
    int stage_batch = 0, stage_next = 0;
    bool staged_ok = false;
    // optional per-kernel timing (bench.py roofline): events around the dominant kernel of the last call
    bool profiling = false;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;       // synthesis kernel / fused path
    static constexpr int MAX_TIMED = 64;
    cudaEvent_t ev_t0[MAX_TIMED] = {}, ev_t1[MAX_TIMED] = {};  // one pair per analysis-kernel launch of the last call
    int n_timed = 0;
    bool last_was_decompose = false;
};

// ------------------------------------------------------------------------------------------------
extern "C" int z3d_version(void) { return 100; }
extern "C" const char *z3d_last_error(void) { return g_err.c_str(); }
extern "C" unsigned long long z3d_launch_count(void) { return g_launches.load(); }

extern "C" int z3d_num_moments(int order) { return order < 0 ? -1 : num_moments(order); }
extern "C" int z3d_num_radial(int order) { return order < 0 ? -1 : num_radial(order); }
extern "C" int z3d_num_harmonics(int order) { return order < 0 ? -1 : (order + 1) * (order + 2) / 2; }
extern "C" int z3d_nlm_table(int order, int *nlm) {
    if (order < 0 || !nlm) return fail(Z3D_ERR_INVALID, "z3d_nlm_table: bad arguments");
    int mu = 0;
    for (int n = 0; n <= order; ++n)
        for (int l = n & 1; l <= n; l += 2)
            for (int m = 0; m <= l; ++m, ++mu) {
                nlm[3 * mu] = n;
                nlm[3 * mu + 1] = l;
                nlm[3 * mu + 2] = m;
            }
    return Z3D_OK;
}

// ------------------------------------------------------------------------------------------------
// planner
// ------------------------------------------------------------------------------------------------
namespace {

struct HostPlan {
    int L, npass;
    std::vector<int> lpass, roffw, poffw, rbase, ownl, nown, ntiles, gen_tasks;
    std::vector<TileDesc> tiles;
    std::vector<float2> gd;     // [npass][GD_MAX] (gamma, -delta) of the parity-decoupled Legendre recurrence
    std::vector<int> goff;      // [npass][L+1]
    std::vector<float4> rk;     // [npass][RK_MAX]
    std::vector<double> kappa;  // [(L+1)*(L+1)] kappa[l*(L+1)+m]
    std::vector<double> beta;   // [(L+1)*(L+4)]  beta[m*(L+4)+l] of the monic Legendre recurrence, 0 outside m+2..L
    std::vector<int> src, omap;
    int useful = 0, padded = 0;
};

inline int padT(int x) { return (x + TILE - 1) / TILE * TILE; }

// pass(l): same parity inside a pass (the synthesis kernel relies on it), round-robin over l/2
inline int pass_of(int l, int npass) { return (l & 1) + 2 * ((l >> 1) % (npass / 2)); }

bool layout_fits(int L, int npass) {
    for (int p = 0; p < npass; ++p) {
        int rw = 0, pw = 0, nt = 0, nrk = 0, ngd = 8;
        for (int l = 0; l <= L; ++l) {
            if (pass_of(l, npass) != p) continue;
            const int nn = (L - l) / 2 + 1, m2 = 2 * (l + 1);
            rw += padT(nn);
            pw += padT(m2);
            nt += (padT(nn) / TILE) * (padT(m2) / TILE);
            nrk += nn;
        }
        for (int m = 0; m <= L; ++m) ngd += (L - m) / 2 + 2;
        if (rw > RS || pw > PS || nt > NT * TS || nrk > RK_MAX || ngd > GD_MAX) return false;
    }
    return true;
}

int build_host_plan(int L, HostPlan &hp) {
    hp.L = L;
    int npass = 2;
    while (npass <= 64 && !layout_fits(L, npass)) npass += 2;
    if (npass > 64)  // the Legendre coefficient table of a pass (GD_MAX entries of shared memory) holds orders up to 74
        return fail(Z3D_ERR_UNSUPPORTED, "order %d too large for this build (orders 0..74 are supported)", L);
    hp.npass = npass;
    const int S = L + 1;
    hp.lpass.assign(S, 0);
    hp.roffw.assign(S, 0);
    hp.poffw.assign(S, 0);
    hp.rbase.assign(S, 0);
    hp.ownl.assign((size_t)npass * S, -1);
    hp.nown.assign(npass, 0);
    hp.ntiles.assign(npass, 0);
    hp.tiles.assign((size_t)npass * NT * TS, TileDesc{0, 4, 0, 4});
    hp.gen_tasks.assign((size_t)npass * NWARP * MAXTASK, 0);

    // ---- coefficient tables (float64 on the host, like the reference's Python scalars)
    // Legendre: C1, C2 of zm:2017-2024, C3 of zm:2013; kappa_mm = P_mm / sin^m, kappa_lm = kappa_{l-1,m} C1/2
    hp.kappa.assign((size_t)S * S, 0.0);
    hp.beta.assign((size_t)S * (S + 3), 0.0);  // beta[m*(S+3) + l], zero for l <= m+1 and l > L
    std::vector<double> &beta = hp.beta;
    std::vector<double> C1((size_t)S * S, 0.0), C2((size_t)S * S, 0.0);
    for (int l = 1; l <= L; ++l)
        for (int m = 0; m < l; ++m) {
            const double c0 = std::sqrt((2.0 * l + 1.0) / ((double)(l + m) * (double)(l - m)));
            C1[(size_t)l * S + m] = c0 * std::sqrt(2.0 * l - 1.0);
            if (m < l - 1) C2[(size_t)l * S + m] = c0 * std::sqrt(((double)(l + m - 1) * (double)(l - m - 1)) / (2.0 * l - 3.0));
        }
    double kmm = 0.5 * std::sqrt(1.0 / M_PI);  // zm:2011
    for (int m = 0; m <= L; ++m) {
        if (m > 0) kmm *= -std::sqrt((2.0 * m + 1.0) / (2.0 * m));  // zm:2013-2014
        hp.kappa[(size_t)m * S + m] = kmm;
        for (int l = m + 1; l <= L; ++l) {
            hp.kappa[(size_t)l * S + m] = hp.kappa[(size_t)(l - 1) * S + m] * C1[(size_t)l * S + m] * 0.5;
            if (l >= m + 2)
                beta[(size_t)m * (S + 3) + l] = C2[(size_t)l * S + m] * 4.0 / (C1[(size_t)l * S + m] * C1[(size_t)(l - 1) * S + m]);
        }
    }
    // parity-decoupled form: Q_{l+2} = (4c^2 - gamma) Q_l - delta Q_{l-2}, gamma = beta_{l+2} + beta_{l+1},
    // delta = beta_{l+1} beta_l; chain (parity, m) starts at l0 = first l >= m of that parity.  One table per pass
    // (only the pass's parity), copied to shared memory by the kernels.
    hp.gd.assign((size_t)npass * GD_MAX, make_float2(0.f, 0.f));
    hp.goff.assign((size_t)npass * S, 0);
    for (int p = 0; p < npass; ++p) {
        const int par = p & 1;
        int pos = 8;  // readable pad in front: lanes whose chain has not started index a few entries before it
        for (int m = 0; m <= L; ++m) {
            hp.goff[(size_t)p * S + m] = pos;
            const double *bm = &beta[(size_t)m * (S + 3)];
            for (int l = m + ((m ^ par) & 1); l + 2 <= L; l += 2)
                hp.gd[(size_t)p * GD_MAX + pos++] = make_float2((float)(bm[l + 2] + bm[l + 1]), (float)(-(bm[l + 1] * bm[l])));
            pos++;  // one spare entry per chain
        }
    }
    // Radial: K1, K2, K3 of zm:1992-1998 for the pass's own l, l-major
    hp.rk.assign((size_t)npass * RK_MAX, make_float4(0.f, 0.f, 0.f, 0.f));
    {
        std::vector<int> pos(npass, 0);
        for (int l = 0; l <= L; ++l) {
            const int p = pass_of(l, npass);
            hp.rbase[l] = pos[p];
            for (int n = l; n <= L; n += 2) {
                float4 K = make_float4(0.f, 0.f, 1.f, 0.f);  // n == l: R = 1 * seed (R_ll = r^l, zm:1983-1984)
                if (n > l) {
                    const double k0 = (double)(n - l) * (n + l + 1) * (2 * n - 3);
                    const double k1 = (double)(2 * n - 1) * (2 * n + 1) * (2 * n - 3);
                    const double k2 = 0.5 * (double)(-2 * n + 1) * (double)((2 * l + 1) * (2 * l + 1)) - k1 / 2.0;
                    const double k3 = -1.0 * (double)(n - l - 2) * (n + l - 1) * (2 * n + 1);
                    K = make_float4((float)(k1 / k0), (float)(k2 / k0), (float)(k3 / k0), 0.f);
                }
                hp.rk[(size_t)p * RK_MAX + pos[p]++] = K;
            }
        }
    }

    // ---- per-pass panel layout and tile enumeration
    const int nmom = num_moments(L);
    hp.src.assign((size_t)2 * nmom, -1);
    hp.omap.assign((size_t)npass * SLOT_WORDS, -1);
    std::vector<int> mubase((size_t)S * S, -1);  // flat index of (n, l, 0)
    {
        int mu = 0;
        for (int n = 0; n <= L; ++n)
            for (int l = n & 1; l <= n; l += 2) {
                mubase[(size_t)n * S + l] = mu;
                mu += l + 1;
            }
    }
    for (int p = 0; p < npass; ++p) {
        int rw = 0, pw = 0, t = 0, no = 0;
        for (int l = 0; l <= L; ++l) {
            if (pass_of(l, npass) != p) continue;
            hp.lpass[l] = p;
            hp.ownl[(size_t)p * S + no++] = l;
            hp.roffw[l] = rw;
            hp.poffw[l] = pw;
            const int nn = (L - l) / 2 + 1, m2 = 2 * (l + 1);
            const int NB = padT(nn) / TILE, MB = padT(m2) / TILE;
            for (int nb = 0; nb < NB; ++nb)
                for (int mb = 0; mb < MB; ++mb, ++t) {
                    const int slot = t / NT, tid = t % NT;
                    const int roff = rw + TILE * nb, poff = pw + TILE * mb;
                    // Each lane loads its 8 R rows / 8 P columns as two LDS.128.  Eight lanes of a quarter warp that
                    // read consecutive 32-byte chunks would hit every 16-byte bank group twice, so lanes whose chunk
                    // index has bit 2 set load the upper half first; the permutation is undone in src / omap below.
                    const int fr = (roff / TILE >> 2) & 1, fp = (poff / TILE >> 2) & 1;
                    hp.tiles[((size_t)p * TS + slot) * NT + tid] =
                        TileDesc{roff + 4 * fr, roff + 4 * (1 - fr), poff + 4 * fp, poff + 4 * (1 - fp)};
                    auto rowof = [&](int ip, int e) { return (ip < 2 ? 4 * fr + 2 * ip : 4 * (1 - fr) + 2 * (ip - 2)) + e; };
                    auto colof = [&](int jp, int e) { return (jp < 2 ? 4 * fp + 2 * jp : 4 * (1 - fp) + 2 * (jp - 2)) + e; };
                    for (int ip = 0; ip < 4; ++ip)
                        for (int jp = 0; jp < 4; ++jp)
                            for (int x = 0; x < 2; ++x)
                                for (int h = 0; h < 2; ++h) {
                                    // D pair (x=0): (row e, col e);  X pair (x=1): (row e, col 1-e)
                                    const int ni = TILE * nb + rowof(ip, h), c = TILE * mb + colof(jp, x ? 1 - h : h);
                                    hp.padded++;
                                    if (ni >= nn || c >= m2) continue;
                                    const int n = l + 2 * ni, m = c >> 1, reim = c & 1;
                                    const int e = 2 * (mubase[(size_t)n * S + l] + m) + reim;
                                    const int word = (slot * NT + tid) * TW + ((ip * 4 + jp) * 2 + x) * 2 + h;
                                    hp.src[e] = p * SLOT_WORDS + word;
                                    hp.omap[(size_t)p * SLOT_WORDS + word] = e;
                                    hp.useful++;
                                }
                }
            rw += padT(nn);
            pw += padT(m2);
        }
        hp.nown[p] = no;
        hp.ntiles[p] = t;

        // ---- generation tasks: groups of GRP chains, longest-processing-time first over the warps
        struct Task { int code; int cost; };
        std::vector<Task> tasks;
        for (int m0 = 0; m0 <= L; m0 += GRP) tasks.push_back({(1 << 16) | m0, ((L - m0) / 2 + 1) * 20 + 80});
        for (int b = 0; b < no; b += GRP) {
            const int l0 = hp.ownl[(size_t)p * S + b];
            tasks.push_back({(2 << 16) | b, ((L - l0) / 2 + 1) * 16 + 60});
        }
        std::sort(tasks.begin(), tasks.end(), [](const Task &a, const Task &b) { return a.cost > b.cost; });
        int load[NWARP] = {0}, cnt[NWARP] = {0};
        for (const Task &tk : tasks) {
            int w = 0;
            for (int q = 1; q < NWARP; ++q)
                if (load[q] < load[w]) w = q;
            if (cnt[w] >= MAXTASK) return fail(Z3D_ERR_UNSUPPORTED, "order %d: too many generation tasks", L);
            hp.gen_tasks[((size_t)p * NWARP + w) * MAXTASK + cnt[w]++] = tk.code;
            load[w] += tk.cost;
        }
    }
    for (int e = 0; e < 2 * nmom; ++e)
        if (hp.src[e] < 0) return fail(Z3D_ERR_INVALID, "internal: output %d has no tile", e);
    return Z3D_OK;
}


// ---- planner of the batched (tensor-core) analysis path (z3d_batched.cuh) -------------------------------------------
struct HostPlanB {
    bool ok = false;
    int ngroups = 0, gpr = 0, nrounds = 0, nrad = 0, nq = 0;
    std::vector<int2> rowtab, rowmap;
    std::vector<int> tile_mpz, mpz, nmpz, ntile, qlo, qlen;
    // inputs of k_build_basis
    std::vector<int> rbase, goff, qoff;
    std::vector<float4> gd, rk;
};

// Returns ok = false (no error) when the order does not fit the kernel's fixed layout: callers then use the per-volume
// kernel for batches as well.
void build_batched_plan(int L, int num_sms, int tiles_per_group, const HostPlan &hp, HostPlanB &hb) {
    const int rows_per_group = tiles_per_group * BTM;
    const int S = L + 1;
    hb.nrad = num_radial(L);
    // radial table, l-major over ALL l (index in rk == index in the R table)
    hb.rbase.assign(S, 0);
    hb.rk.assign(hb.nrad + 8, make_float4(0.f, 0.f, 0.f, 0.f));
    {
        int pos = 0;
        for (int l = 0; l <= L; ++l) {
            hb.rbase[l] = pos;
            for (int n = l; n <= L; n += 2) {
                float4 K = make_float4(0.f, 0.f, 1.f, 0.f);  // n == l: R = 1 * seed (R_ll = r^l, zm:1983-1984)
                if (n > l) {  // K1, K2, K3 of zm:1992-1998
                    const double k0 = (double)(n - l) * (n + l + 1) * (2 * n - 3);
                    const double k1 = (double)(2 * n - 1) * (2 * n + 1) * (2 * n - 3);
                    const double k2 = 0.5 * (double)(-2 * n + 1) * (double)((2 * l + 1) * (2 * l + 1)) - k1 / 2.0;
                    const double k3 = -1.0 * (double)(n - l - 2) * (n + l - 1) * (2 * n + 1);
                    K = make_float4((float)(k1 / k0), (float)(k2 / k0), (float)(k3 / k0), 0.f);
                }
                hb.rk[pos++] = K;
            }
        }
    }
    // Legendre coefficients and Q table offsets of every m: pairs (E_i, O_i) = levels (m + 2i, m + 2i + 1)
    hb.goff.assign(S, 0);
    hb.qoff.assign(S, 0);
    {
        int qpos = 0;
        for (int m = 0; m <= L; ++m) {
            hb.qoff[m] = qpos;
            qpos += 2 * ((L - m) / 2 + 1);
            hb.goff[m] = (int)hb.gd.size();
            const double *bm = &hp.beta[(size_t)m * (S + 3)];
            auto B = [&](int l) { return (l >= 0 && l < S + 3) ? bm[l] : 0.0; };
            for (int i = 0; i < (L - m) / 2 + 2; ++i) {
                const int le = m + 2 * i, lo = le + 1;
                hb.gd.push_back(make_float4((float)(B(le + 2) + B(le + 1)), (float)(-(B(le + 1) * B(le))),
                                            (float)(B(lo + 2) + B(lo + 1)), (float)(-(B(lo + 1) * B(lo)))));
            }
        }
        hb.nq = qpos;
    }
    if (hb.nrad + 4 > 65535 || hb.nq > 65535) return;
    // rows of every (m, pz), in (m, pz) order, padded to tiles of BTM
    struct Row { int n, l, m; };
    std::vector<Row> rows;           // padded row list, n = -1 for padding
    std::vector<int> tile_code;      // m | pz << 16 per tile
    for (int m = 0; m <= L; ++m)
        for (int pz = 0; pz < 2; ++pz) {
            int cnt = 0;
            for (int l = m + pz; l <= L; l += 2)
                for (int n = l; n <= L; n += 2, ++cnt) rows.push_back({n, l, m});
            if (cnt == 0) continue;
            for (; cnt % BTM; ++cnt) rows.push_back({-1, 0, m});
            for (int t = (int)tile_code.size(); t < (int)rows.size() / BTM; ++t) tile_code.push_back(m | (pz << 16));
        }
    const int ntiles = (int)tile_code.size();
    const int ngroups = (ntiles + tiles_per_group - 1) / tiles_per_group;
    if (ngroups > num_sms) return;
    hb.ngroups = ngroups;
    {   // rounds (z3d_batched.cuh, WorkRange): time ~ sum over rounds of 1 / floor(SMs / groups in the round); take the largest
        // number of groups per round (best L2 reuse, fewest accumulator read-backs) within 6 % of the best
        auto cost = [&](int gpr) {
            double c = 0.0;
            for (int g0 = 0; g0 < ngroups; g0 += gpr) c += 1.0 / (num_sms / std::min(gpr, ngroups - g0));
            return c;
        };
        double best = 1e30;
        for (int gpr = 1; gpr <= ngroups; ++gpr) best = std::min(best, cost(gpr));
        for (int gpr = ngroups; gpr >= 1; --gpr)
            if (cost(gpr) <= 1.06 * best) {
                hb.gpr = gpr;
                break;
            }
        hb.nrounds = (ngroups + hb.gpr - 1) / hb.gpr;
    }
    hb.rowtab.assign((size_t)ngroups * rows_per_group, make_int2(hb.nrad, 0));  // padding rows: the zero R entry
    hb.tile_mpz.assign((size_t)ngroups * BTILES_MAX, 0);
    hb.mpz.assign((size_t)ngroups * BTILES_MAX, -1);
    hb.nmpz.assign(ngroups, 0);
    hb.ntile.assign(ngroups, 0);
    hb.qlo.assign(ngroups, 0);
    hb.qlen.assign(ngroups, 0);
    hb.rowmap.assign(num_moments(L), make_int2(-1, -1));
    std::vector<int> mubase((size_t)S * S, -1);
    {
        int mu = 0;
        for (int n = 0; n <= L; ++n)
            for (int l = n & 1; l <= n; l += 2) {
                mubase[(size_t)n * S + l] = mu;
                mu += l + 1;
            }
    }
    for (int g = 0; g < ngroups; ++g) {
        const int t0 = g * tiles_per_group, t1 = std::min(ntiles, t0 + tiles_per_group);
        const int mlo = tile_code[t0] & 0xffff, mhi = tile_code[t1 - 1] & 0xffff;
        hb.ntile[g] = t1 - t0;
        hb.qlo[g] = hb.qoff[mlo];
        hb.qlen[g] = hb.qoff[mhi] + 2 * ((L - mhi) / 2 + 1) - hb.qoff[mlo];
        if (hb.qlen[g] > BQ_MAX) return;
        int nm = 0, last = -1;
        for (int t = t0; t < t1; ++t) {
            if (tile_code[t] != last) {
                hb.mpz[(size_t)g * BTILES_MAX + nm++] = tile_code[t];
                last = tile_code[t];
            }
            hb.tile_mpz[(size_t)g * BTILES_MAX + (t - t0)] = nm - 1;
            for (int i = 0; i < BTM; ++i) {
                const Row &r = rows[(size_t)t * BTM + i];
                if (r.n < 0) continue;
                const int row = (t - t0) * BTM + i;
                hb.rowtab[(size_t)g * rows_per_group + row] =
                    make_int2(hb.rbase[r.l] + (r.n - r.l) / 2, hb.qoff[r.m] + (r.l - r.m) - hb.qlo[g]);
                hb.rowmap[mubase[(size_t)r.n * S + r.l] + r.m] = make_int2(g, row);
            }
        }
        hb.nmpz[g] = nm;
    }
    for (const int2 &r : hb.rowmap)
        if (r.x < 0) return;
    hb.ok = true;
}

// ---- planner of the streamed-T variant (z3d_stream.cuh) -------------------------------------------------------------
struct HostPlanS {
    bool ok = false;
    int nsets = 0, nsplits = 0, ncols_all = 0, t_bytes_max = 0;
    long long trec = 0;
    std::vector<StreamTabs> tabs;
    std::vector<int4> colsrc;
    std::vector<int2> colmap;
};

// Column sets: the moment rows of every (m, pz), padded to 16, are laid end to end in kind order ((m parity, pz), then
// m) and cut into sets of equal estimated shared-memory traffic per chunk (64 + 96 B per column for the T block and the
// tensor core's reads of it, ~28 KB per a panel), subject to <= 512 columns (tensor memory), <= S_PANELS panels, <= 2
// kinds.  sets x splits <= SMs; the number of splits is the one with the smallest estimated time.
void build_stream_plan(int L, int num_sms, const HostPlanB &hb, HostPlanS &hs, int force_splits) {
    const int S = L + 1;
    struct Row { int n, l, m; };
    struct Group { int kind, m, pz; std::vector<Row> rows; };
    std::vector<Group> groups;
    for (int mp = 0; mp < 2; ++mp)
        for (int pz = 0; pz < 2; ++pz)
            for (int m = mp; m <= L; m += 2) {
                Group g{(mp << 1) | pz, m, pz, {}};
                for (int l = m + pz; l <= L; l += 2)
                    for (int n = l; n <= L; n += 2) g.rows.push_back({n, l, m});
                if (!g.rows.empty()) groups.push_back(std::move(g));
            }
    struct Piece { int group, first, cols; };
    typedef std::vector<std::vector<Piece>> Sets;
    const double CA = 160.0, CB = 28000.0, CK = 16384.0;
    auto set_cost = [&](const std::vector<Piece> &st) {
        double c = 0.0;
        int k0 = -1, nk = 0;
        for (const Piece &p : st) {
            c += CB + CA * p.cols;
            if (groups[p.group].kind != k0) { k0 = groups[p.group].kind; ++nk; }
        }
        return c + CK * nk;
    };
    auto pack = [&](double cap) {
        Sets sets;
        std::vector<Piece> cur;
        double cost = 0.0;
        int cols = 0, segs = 0, k_first = -1, k_last = -1;
        auto close = [&]() {
            if (!cur.empty()) sets.push_back(cur);
            cur.clear(); cost = 0.0; cols = 0; segs = 0; k_first = k_last = -1;
        };
        for (int gi = 0; gi < (int)groups.size(); ++gi) {
            const int kind = groups[gi].kind;
            int rem = ((int)groups[gi].rows.size() + 15) / 16 * 16, first = 0;
            while (rem > 0) {
                const bool newkind = !cur.empty() && kind != k_last;
                const double kc = (cur.empty() || newkind) ? CK : 0.0;
                int room = std::min(512 - cols, (int)((cap - cost - CB - kc) / CA) / 16 * 16);
                if ((int)cur.size() >= S_PANELS || segs + 1 > S_SEGS || room < 16 || (newkind && (S_KINDS == 1 || k_first != k_last))) {
                    if (cur.empty()) room = 16; else { close(); continue; }
                }
                int take = std::min(rem, room);
                if (take > 256 && segs + 2 > S_SEGS) take = 256;
                if (cur.empty()) k_first = kind;
                k_last = kind;
                cur.push_back({gi, first, take});
                cost += CB + kc + CA * take;
                cols += take;
                segs += take > 256 ? 2 : 1;
                first += take;
                rem -= take;
            }
        }
        close();
        return sets;
    };
    Sets best;
    double best_t = 1e300;
    int best_splits = 0;
    for (int ns = 1; ns <= 12; ++ns) {
        if (force_splits > 0 && ns != force_splits) continue;
        const int target = num_sms / ns;
        if (target < 1) break;
        double lo = CB, hi = 4e6;
        if ((int)pack(hi).size() > target) continue;
        for (int it = 0; it < 40; ++it) {
            const double mid = 0.5 * (lo + hi);
            if ((int)pack(mid).size() > target) lo = mid; else hi = mid;
        }
        Sets st = pack(hi);
        double mx = 0.0;
        for (auto &x : st) mx = std::max(mx, set_cost(x));
        const double t = (mx + 20000.0) / ns;  // + per-chunk synchronisation overhead
        if (t < best_t) { best_t = t; best = st; best_splits = ns; }
    }
    if (best.empty()) return;
    hs.nsets = (int)best.size();
    hs.nsplits = best_splits;
    hs.tabs.assign(hs.nsets, StreamTabs{});
    hs.colmap.assign(num_moments(L), make_int2(-1, -1));
    std::vector<int> mubase((size_t)S * S, -1);
    {
        int mu = 0;
        for (int n = 0; n <= L; ++n)
            for (int l = n & 1; l <= n; l += 2) {
                mubase[(size_t)n * S + l] = mu;
                mu += l + 1;
            }
    }
    long long toff = 0;
    for (int si = 0; si < hs.nsets; ++si) {
        StreamTabs &T = hs.tabs[si];
        int cols = 0;
        for (const Piece &p : best[si]) cols += p.cols;
        T.ncols = cols;
        T.toff = toff;
        T.kind[0] = 0;
        int off = 0;
        for (const Piece &p : best[si]) {
            const Group &g = groups[p.group];
            int slot = -1;
            for (int q = 0; q < T.nkind; ++q)
                if (T.kind[q] == g.kind) slot = q;
            if (slot < 0) {
                if (T.nkind >= S_KINDS) return;
                slot = T.nkind;
                T.kind[T.nkind++] = g.kind;
            }
            if (T.npanel >= S_PANELS) return;
            const int j = T.npanel++;
            T.panel_code[j] = g.m | (g.pz << 16) | (slot << 24);
            for (int c0 = 0; c0 < p.cols; c0 += 256) {
                if (T.nseg >= S_SEGS) return;
                T.seg_cols[T.nseg] = std::min(256, p.cols - c0);
                T.seg_off[T.nseg] = off + c0;
                T.seg_panel[T.nseg] = j;
                ++T.nseg;
            }
            for (int c = 0; c < p.cols; ++c) {
                const int ri = p.first + c;
                int4 cs = make_int4(-1, 0, (int)(toff + (long long)(off + c) * 4), cols);
                if (ri < (int)g.rows.size()) {
                    const Row &r = g.rows[ri];
                    cs.x = hb.rbase[r.l] + (r.n - r.l) / 2;
                    cs.y = hb.qoff[r.m] + (r.l - r.m);
                    hs.colmap[mubase[(size_t)r.n * S + r.l] + r.m] = make_int2(si, off + c);
                }
                hs.colsrc.push_back(cs);
            }
            off += p.cols;
        }
        if (cols > 512 || toff + 16LL * cols > 0x7fffffffLL) return;
        hs.t_bytes_max = std::max(hs.t_bytes_max, cols * 64);
        toff += 16LL * cols;
    }
    hs.trec = toff;
    hs.ncols_all = (int)hs.colsrc.size();
    for (const int2 &r : hs.colmap)
        if (r.x < 0) return;
    hs.ok = true;
}

// ---- planner of the table-free variant (z3d_gen.cuh) ------------------------------------------------------------------
struct HostPlanG {
    bool ok = false;
    int nsets = 0, nsplits = 0, maxcols = 0, maxseedrows = 0, gd_entries = 0, rk_entries = 0, ncols_all = 0;
    double est_period = 0.0;
    std::vector<GenTabs> tabs;
    std::vector<int4> tasks;
    std::vector<float2> gd;
    std::vector<float4> rk;
    std::vector<int2> colmap;
};

// Column sets of the table-free kernel.  The moment rows of every (m, pz) are a list of CHAINS (l = m + pz, m + pz + 2, ..: the
// rows n = l, l + 2, .. of one chain are produced by one thread's recurrence), laid end to end in kind order and cut - at chain
// boundaries only - into sets of equal estimated tensor-pipe time per chunk: an MMA of N columns takes max(24, N / 2) clocks
// (scripts/experiments/tc_probe4.cu), three per segment.  Limits per set: accumulator columns + 2 a-panel slots of 16 columns
// per panel <= 512 (tensor memory), <= G_PANELS panels, <= G_SEGS segments, <= G_TASKS chain tasks, one kind.
// ntile = 2: TWO tiles of 64 volumes share every generated T block (two accumulator banks and two sets of a-panel slots in tensor
// memory), so a set holds half the columns and the chunk list is cut into half as many splits.
void build_gen_plan(int L, int num_sms, const HostPlan &hp, HostPlanG &hg, int force_splits, int ntile = 1) {
    const int S = L + 1;
    struct Chain { int l, len; };
    struct Group { int kind, m, pz; std::vector<Chain> chains; };
    std::vector<Group> groups;
    for (int mp = 0; mp < 2; ++mp)
        for (int pz = 0; pz < 2; ++pz)
            for (int m = mp; m <= L; m += 2) {
                Group g{(mp << 1) | pz, m, pz, {}};
                for (int l = m + pz; l <= L; l += 2) g.chains.push_back({l, (L - l) / 2 + 1});
                if (!g.chains.empty()) groups.push_back(std::move(g));
            }
    struct Piece { int group, first, count, rows; };   // chains [first, first + count) of the group, rows = sum of their lengths
    typedef std::vector<std::vector<Piece>> Sets;
    auto pad16 = [](int x) { return (x + 15) / 16 * 16; };
    auto piece_cost = [&](int rows) {
        // (a cost refitted to the measured duration of every CTA - ~1.1 clk per column and ~75 clk per panel on top of a fixed ~950
        // clk per chunk, -DZ3D_GEN_CTACLK - moves a few chains between neighbouring sets and changes nothing: 3.69 vs 3.63 ms)
        double c = 48.0;  // a panel: the a-panel team's work
        for (int c0 = 0, cols = pad16(rows); c0 < cols; c0 += 256) c += 3.0 * std::max(24.0, 0.5 * std::min(256, cols - c0));
        return c * ntile;
    };
    auto set_cost = [&](const std::vector<Piece> &st) {
        double c = 0.0;
        for (const Piece &p : st) c += piece_cost(p.rows);
        return c;
    };
    auto fits = [&](const std::vector<Piece> &st) {
        int cols = 0, segs = 0, chains = 0;
        for (const Piece &p : st) {
            cols += pad16(p.rows);
            segs += (pad16(p.rows) + 255) / 256;
            chains += p.count;
        }
        return ntile * (cols + g_aslots(ntile) * 16 * (int)st.size()) <= 512 && (int)st.size() <= G_PANELS && segs <= G_SEGS && 2 * chains <= G_TASKS;
    };
    auto pack = [&](double cap) {
        Sets sets;
        std::vector<Piece> cur;
        auto close = [&]() {
            if (!cur.empty()) sets.push_back(cur);
            cur.clear();
        };
        for (int gi = 0; gi < (int)groups.size(); ++gi) {
            const Group &g = groups[gi];
            if (!cur.empty() && groups[cur.back().group].kind != g.kind) close();
            for (int ci = 0; ci < (int)g.chains.size(); ++ci) {
                std::vector<Piece> trial = cur;
                if (!trial.empty() && trial.back().group == gi) {
                    trial.back().count += 1;
                    trial.back().rows += g.chains[ci].len;
                } else {
                    trial.push_back({gi, ci, 1, g.chains[ci].len});
                }
                if (!cur.empty() && (!fits(trial) || set_cost(trial) > cap)) {
                    close();
                    trial.assign(1, Piece{gi, ci, 1, g.chains[ci].len});
                }
                cur = trial;
            }
        }
        close();
        return sets;
    };
    Sets best;
    double best_t = 1e300;
    int best_splits = 0;
    const double PMIN = 300.0, POVH = 40.0;   // shortest sustainable chunk period, per-chunk overhead on top of the tensor time
    for (int ns = 1; ns <= 16; ++ns) {
        if (force_splits > 0 && ns != force_splits) continue;
        const int target = num_sms / ns;
        if (target < 1) break;
        double lo = 0.0, hi = 1e7;
        if ((int)pack(hi).size() > target) continue;
        for (int it = 0; it < 50; ++it) {
            const double mid = 0.5 * (lo + hi);
            if ((int)pack(mid).size() > target) lo = mid; else hi = mid;
        }
        Sets st = pack(hi);
        double mx = 0.0;
        for (auto &x : st) mx = std::max(mx, set_cost(x));
        const double t = (std::max(mx, PMIN) + POVH) / ns;
        if (t < best_t) { best_t = t; best = st; best_splits = ns; hg.est_period = std::max(mx, PMIN) + POVH; }
    }
    if (best.empty()) return;
    hg.nsets = (int)best.size();
    hg.nsplits = best_splits;
    // radial tables by parity of l: (K1, K2, K3, 0) of zm:1992-1998, l-major, entry 0 of a chain = (0, 0, 1): R_ll = r^l enters as seed
    std::vector<int> rkoff(S, 0);
    std::vector<float4> rkp[2];
    for (int l = 0; l <= L; ++l) {
        std::vector<float4> &t = rkp[l & 1];
        rkoff[l] = (int)t.size();
        for (int n = l; n <= L; n += 2) {
            float4 K = make_float4(0.f, 0.f, 1.f, 0.f);
            if (n > l) {
                const double k0 = (double)(n - l) * (n + l + 1) * (2 * n - 3);
                const double k1 = (double)(2 * n - 1) * (2 * n + 1) * (2 * n - 3);
                const double k2 = 0.5 * (double)(-2 * n + 1) * (double)((2 * l + 1) * (2 * l + 1)) - k1 / 2.0;
                const double k3 = -1.0 * (double)(n - l - 2) * (n + l - 1) * (2 * n + 1);
                K = make_float4((float)(k1 / k0), (float)(k2 / k0), (float)(k3 / k0), 0.f);
            }
            t.push_back(K);
        }
    }
    hg.rk_entries = (int)std::max(rkp[0].size(), rkp[1].size()) + 8;
    hg.rk.assign((size_t)2 * hg.rk_entries, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int par = 0; par < 2; ++par) std::copy(rkp[par].begin(), rkp[par].end(), hg.rk.begin() + (size_t)par * hg.rk_entries);
    std::vector<int> mubase((size_t)S * S, -1);
    {
        int mu = 0;
        for (int n = 0; n <= L; ++n)
            for (int l = n & 1; l <= n; l += 2) {
                mubase[(size_t)n * S + l] = mu;
                mu += l + 1;
            }
    }
    hg.tabs.assign(hg.nsets, GenTabs{});
    hg.tasks.assign((size_t)hg.nsets * G_TASKS, make_int4(0, 0, 0, 0));
    hg.colmap.assign(num_moments(L), make_int2(-1, -1));
    int maxgd = 0;
    for (const auto &st : best) {
        int rows = 0;
        for (const Piece &p : st) rows += p.first + p.count;
        maxgd = std::max(maxgd, rows);
    }
    hg.gd_entries = maxgd + 64;
    hg.gd.assign((size_t)hg.nsets * hg.gd_entries, make_float2(0.f, 0.f));
    for (int si = 0; si < hg.nsets; ++si) {
        GenTabs &T = hg.tabs[si];
        const auto &st = best[si];
        T.npanel = (int)st.size();
        T.kind = groups[st[0].group].kind;
        T.lpar = (groups[st[0].group].m + groups[st[0].group].pz) & 1;
        T.acol0 = 512 - ntile * g_aslots(ntile) * 16 * T.npanel;
        struct TaskC { int row0, len, rk, srow; };
        std::vector<TaskC> chains;
        int off = 0, soff = 0, gdo = 0;
        for (int j = 0; j < T.npanel; ++j) {
            const Piece &p = st[j];
            const Group &g = groups[p.group];
            const int cols = pad16(p.rows), l0 = g.m + g.pz;
            T.panel_m[j] = g.m;
            T.panel_pz[j] = g.pz;
            T.panel_nl[j] = p.first + p.count;   // the seed recurrence starts at l0 whichever chain the piece starts with
            T.panel_soff[j] = soff;
            T.panel_gd[j] = gdo;
            const double *bm = &hp.beta[(size_t)g.m * (S + 3)];
            auto B = [&](int l) { return (l >= 0 && l < S + 3) ? bm[l] : 0.0; };
            for (int i = 0; i + 1 < T.panel_nl[j]; ++i) {  // step l -> l + 2: (gamma_l, -delta_l)
                const int l = l0 + 2 * i;
                hg.gd[(size_t)si * hg.gd_entries + gdo + i] = make_float2((float)(B(l + 2) + B(l + 1)), (float)(-(B(l + 1) * B(l))));
            }
            gdo += T.panel_nl[j];
            for (int c0 = 0; c0 < cols; c0 += 256) {
                T.seg_cols[T.nseg] = std::min(256, cols - c0);
                T.seg_off[T.nseg] = off + c0;
                T.seg_panel[T.nseg] = j;
                ++T.nseg;
            }
            int r = 0;
            for (int ci = p.first; ci < p.first + p.count; ++ci) {
                const Chain &ch = g.chains[ci];
                chains.push_back({off + r, ch.len, rkoff[ch.l], soff + ci});
                for (int i = 0; i < ch.len; ++i) hg.colmap[mubase[(size_t)(ch.l + 2 * i) * S + ch.l] + g.m] = make_int2(si, off + r + i);
                r += ch.len;
            }
            soff += T.panel_nl[j];
            off += cols;
        }
        T.ncols = off;
        T.nseedrows = soff;
        hg.maxcols = std::max(hg.maxcols, off);
        hg.maxseedrows = std::max(hg.maxseedrows, soff);
        hg.ncols_all += off;
        // tasks: chains sorted by length (longest first), in blocks of 16 chains: 16 x k half 0, then 16 x k half 1 (a quarter
        // warp's 16-byte stores then belong to 8 different chains)
        std::stable_sort(chains.begin(), chains.end(), [](const TaskC &a, const TaskC &b) { return a.len > b.len; });
        // Within a warp's 16 chains, lanes 0-7 and 8-15 (and their k-half twins 16-23, 24-31) are the quarter-warps of the 16-byte
        // stores: a quarter-warp is conflict free when its 8 rows differ mod 8 (a row is 16 B, 8 rows span the 32 banks), and all
        // its lanes advance their rows in step.  Deal the block's chains into the two groups by row residue (ncu: 57 % of the T
        // stores' shared-memory wavefronts were conflict replays with the chains in plain length order).
        int nt = 0;
        for (size_t c0 = 0; c0 < chains.size(); c0 += 16) {
            const size_t n = std::min<size_t>(16, chains.size() - c0);
            std::vector<TaskC> grp[2], rest;
            bool used[2][8] = {{false}};
            for (size_t c = c0; c < c0 + n; ++c) {
                const int r = chains[c].row0 & 7;
                if (!used[0][r] && grp[0].size() < 8) { used[0][r] = true; grp[0].push_back(chains[c]); }
                else if (!used[1][r] && grp[1].size() < 8) { used[1][r] = true; grp[1].push_back(chains[c]); }
                else rest.push_back(chains[c]);
            }
            for (const TaskC &x : rest) (grp[0].size() < 8 && (grp[0].size() <= grp[1].size() || grp[1].size() >= 8) ? grp[0] : grp[1]).push_back(x);
            // a warp's 32 lanes: [group 0 | group 1] x k half 0, then the same x k half 1; a short last block keeps the quarter-warp
            // alignment with empty tasks (length 0: the lane idles)
            const TaskC none{0, 0, 0, 0};
            for (int kh = 0; kh < 2; ++kh)
                for (int g2 = 0; g2 < 2; ++g2)
                    for (size_t i = 0; i < 8; ++i) {
                        const TaskC &x = i < grp[g2].size() ? grp[g2][i] : none;
                        if (nt < G_TASKS) hg.tasks[(size_t)si * G_TASKS + nt] = make_int4(x.row0, x.len, x.rk, x.srow | (kh << 16));
                        ++nt;
                    }
        }
        T.ntask = nt;
        if (nt > G_TASKS || T.nseg > G_SEGS || ntile * (T.ncols + g_aslots(ntile) * 16 * T.npanel) > 512) return;
    }
    for (const int2 &r : hg.colmap)
        if (r.x < 0) return;
    hg.ok = true;
}

template <int NPH>
int set_kernel_attrs_t(int smem) {
    CUDA_TRY(cudaFuncSetAttribute(k_decompose<NPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_reconstruct<NPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_decompose<NPH>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_TRY(cudaFuncSetAttribute(k_reconstruct<NPH>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return Z3D_OK;
}

int set_kernel_attrs(int smem) {
    int rc = set_kernel_attrs_t<1>(smem);
    if (rc) return rc;
    CUDA_TRY(cudaFuncSetAttribute(k_decompose_batched<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_decompose_batched<32>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_TRY(cudaFuncSetAttribute(k_decompose_batched<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_decompose_batched<64>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_TRY(cudaFuncSetAttribute(k_decompose_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_decompose_stream, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_TRY(cudaFuncSetAttribute(k_decompose_gen<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_decompose_gen<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_TRY(cudaFuncSetAttribute(k_decompose_gen<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_decompose_gen<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_TRY(cudaFuncSetAttribute(k_synth_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_synth_gen, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return set_kernel_attrs_t<0>(smem);
}

}  // namespace

extern "C" int z3d_plan_create(int order, int dim, int device, z3d_plan **out) {
    if (!out) return fail(Z3D_ERR_INVALID, "z3d_plan_create: out is NULL");
    *out = nullptr;
    if (order < 0 || order > LMAX) return fail(Z3D_ERR_INVALID, "z3d_plan_create: order %d outside 0..%d", order, LMAX);
    if (dim < 2 || dim > 4096) return fail(Z3D_ERR_INVALID, "z3d_plan_create: bad dim %d (need 2..4096)", dim);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(Z3D_ERR_NODEVICE, "z3d_plan_create: no CUDA device (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(Z3D_ERR_INVALID, "z3d_plan_create: device %d of %d", device, ndev);
    USE_DEVICE(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(Z3D_ERR_NODEVICE, "z3d_plan_create: device %d is sm_%d%d; this build is sm_100a only", device,
                    prop.major, prop.minor);

    HostPlan hp;
    int rc = build_host_plan(order, hp);
    if (rc) return rc;

    z3d_plan *pl = new z3d_plan();
    pl->order = order;
    pl->dim = dim;
    pl->device = device;
    pl->H = (dim + 1) / 2;
    pl->nmom = num_moments(order);
    pl->nrad = num_radial(order);
    pl->npass = hp.npass;
    pl->num_sms = prop.multiProcessorCount;
    pl->nsplit = std::max(1, (prop.multiProcessorCount * CTAS_PER_SM) / hp.npass);
    pl->useful_acc = hp.useful;
    pl->padded_acc = hp.padded;
    const SmemMap mp = smem_map(dim, pl->H);
    pl->smem_bytes = mp.total;
    if ((size_t)mp.total > prop.sharedMemPerBlockOptin) {
        delete pl;
        return fail(Z3D_ERR_UNSUPPORTED, "plan needs %d B of shared memory per CTA (device allows %zu)", mp.total,
                    prop.sharedMemPerBlockOptin);
    }

    // side = np.linspace(-1/sqrt(3), 1/sqrt(3), dim)  (zm:2240): start + k*step, last point = stop
    std::vector<double> side(dim);
    {
        const double stop = 1.0 / std::sqrt(3.0), start = -stop, step = (stop - start) / (double)(dim - 1);
        for (int k = 0; k < dim; ++k) side[k] = (double)k * step + start;
        side[dim - 1] = stop;
    }
    // scale tables
    const int S = order + 1;
    std::vector<double> scale(pl->nmom), ofac(pl->nmom);
    std::vector<int> nl_base(pl->nrad), nl_l(pl->nrad);
    {
        int mu = 0, q = 0;
        const double d3 = (double)dim * dim * dim;
        for (int n = 0; n <= order; ++n) {
            const double coeff = (2.0 * (2.0 * n + 3.0)) / (3.0 * std::sqrt(3.0) * M_PI * d3);  // zm:2198
            for (int l = n & 1; l <= n; l += 2, ++q) {
                nl_base[q] = mu;
                nl_l[q] = l;
                for (int m = 0; m <= l; ++m, ++mu) {
                    const double k = hp.kappa[(size_t)l * S + m];
                    scale[mu] = coeff * k;
                    ofac[mu] = k * (m == 0 ? 1.0 : 2.0);
                }
            }
        }
    }

    DevPlan &d = pl->dev;
    d.order = order;
    d.dim = dim;
    d.H = pl->H;
    d.npass = pl->npass;
    d.nsplit = pl->nsplit;
    d.dim_vec4 = (dim % 4 == 0);
    auto &ow = pl->owned;
    double *d_side; float2 *d_gd; int *d_goff; float4 *d_rk; int *d_rbase, *d_roffw, *d_poffw, *d_ownl, *d_nown, *d_ntiles, *d_gen;
    TileDesc *d_tiles;
#define UP(h, dptr)                          \
    if ((rc = upload(h, &dptr, ow)) != 0) {  \
        z3d_plan_destroy(pl);                \
        return rc;                           \
    }
    UP(side, d_side) UP(hp.gd, d_gd) UP(hp.goff, d_goff) UP(hp.rk, d_rk) UP(hp.rbase, d_rbase) UP(hp.roffw, d_roffw)
    UP(hp.poffw, d_poffw) UP(hp.ownl, d_ownl) UP(hp.nown, d_nown) UP(hp.ntiles, d_ntiles)
    UP(hp.gen_tasks, d_gen) UP(hp.tiles, d_tiles)
    UP(hp.src, pl->d_src) UP(scale, pl->d_scale) UP(hp.omap, pl->d_omap) UP(ofac, pl->d_ofac)
    UP(nl_base, pl->d_nl_base) UP(nl_l, pl->d_nl_l)
    d.side = d_side; d.gd = d_gd; d.goff = d_goff; d.rk = d_rk; d.rbase = d_rbase; d.roffw = d_roffw; d.poffw = d_poffw;
    d.ownl = d_ownl; d.nown = d_nown; d.ntiles = d_ntiles; d.tiles = d_tiles; d.gen_tasks = d_gen;

    // dim >= 14 keeps every 8-voxel chunk and its z-mirror inside a row (k_fold_tile reads both)
    pl->batched = dim >= 14;
    HostPlanB hb0;
    for (int v = 0; v < 2 && pl->batched; ++v) {
        HostPlanB hb;
        build_batched_plan(order, prop.multiProcessorCount, v ? BCfg<64>::TILES : BCfg<32>::TILES, hp, hb);
        auto smem_of = [&](bool ad) { return v ? BCfg<64>::smem_bytes(hb.nrad, ad) : BCfg<32>::smem_bytes(hb.nrad, ad); };
        const bool adouble = (size_t)smem_of(true) <= prop.sharedMemPerBlockOptin;
        pl->b_smem[v] = smem_of(adouble);
        if (!hb.ok || (size_t)pl->b_smem[v] > prop.sharedMemPerBlockOptin) {
            pl->batched = false;
            break;
        }
        DevPlanB &b = pl->devb[v];
        b.order = order; b.dim = dim; b.H = pl->H; b.ngroups = hb.ngroups; b.gpr = hb.gpr; b.nrounds = hb.nrounds;
        b.nrad = hb.nrad; b.nq = hb.nq; b.adouble = adouble;
        int2 *d_rowtab; int *d_bm, *d_mpz, *d_nmpz, *d_ntile, *d_qlo, *d_qlen;
        UP(hb.rowtab, d_rowtab) UP(hb.tile_mpz, d_bm) UP(hb.mpz, d_mpz) UP(hb.nmpz, d_nmpz) UP(hb.ntile, d_ntile)
        UP(hb.qlo, d_qlo) UP(hb.qlen, d_qlen) UP(hb.rowmap, pl->d_rowmap[v])
        b.rowtab = d_rowtab; b.tile_mpz = d_bm; b.mpz = d_mpz; b.nmpz = d_nmpz; b.ntile = d_ntile;
        b.qlo = d_qlo; b.qlen = d_qlen;
        if (v == 0) {  // inputs of k_build_basis (the same for both tile widths)
            int *d_bgoff, *d_qoff, *d_brbase;
            float4 *d_bgd, *d_brk;
            UP(hb.goff, d_bgoff) UP(hb.qoff, d_qoff) UP(hb.rbase, d_brbase) UP(hb.gd, d_bgd) UP(hb.rk, d_brk)
            BasisArgs &a = pl->basis;
            a.order = order; a.dim = dim; a.H = pl->H; a.nrad = hb.nrad; a.nq = hb.nq; a.side = d_side;
            a.rk = d_brk; a.rbase = d_brbase; a.gd = d_bgd; a.goff = d_bgoff; a.qoff = d_qoff;
            pl->b_chunks = (long long)analysis_items(0, pl->H, pl->H) * ((pl->H + KC - 1) / KC);
            hb0 = hb;
        }
    }
    if (pl->batched) {  // streamed-T variant: tables only; the T table itself is built on first use (ensure_stream)
        HostPlanS hs;
        int force = 0;
        if (const char *e = getenv("Z3D_STREAM_SPLITS")) force = atoi(e);
        const char *off = getenv("Z3D_STREAM");
        if (!(off && atoi(off) == 0)) build_stream_plan(order, prop.multiProcessorCount, hb0, hs, force);
        if (hs.ok) {
            DevPlanS &d2 = pl->devs;
            d2.order = order; d2.dim = dim; d2.H = pl->H; d2.nsets = hs.nsets; d2.nsplits = hs.nsplits;
            d2.t_bytes_max = hs.t_bytes_max;
            d2.stage_bytes = hs.t_bytes_max + S_FK_BYTES + S_ABUF_BYTES;  // a slot: T block, folds of one kind, a panels
            d2.nstages = S_NTEAM;
            d2.trec = hs.trec;
            if (S_OFF_A + S_NTEAM * d2.stage_bytes <= (int)prop.sharedMemPerBlockOptin) {
                StreamTabs *d_tabs;
                UP(hs.tabs, d_tabs) UP(hs.colsrc, pl->d_colsrc) UP(hs.colmap, pl->d_colmap)
                d2.tabs = d_tabs;
                pl->s_smem = S_OFF_A + S_NTEAM * d2.stage_bytes;
                pl->s_ncols_all = hs.ncols_all;
                pl->stream = true;
            }
        }
    }
    if (dim >= 14) {  // table-free variant (dim >= 14 keeps every 8-voxel chunk and its z-mirror inside a row)
        int force = 0;
        if (const char *e = getenv("Z3D_GEN_SPLITS")) force = atoi(e);
        const char *off = getenv("Z3D_GEN");
        const char *off2 = getenv("Z3D_GEN2");
        for (int ntile = 1; ntile <= 2; ++ntile) {
            if (off && atoi(off) == 0) break;
            // two tiles per T block: correct and tested, but not faster yet (9.4 ms per 128 volumes against 2 x 3.7 ms at 128^3 /
            // order 50: with twice the a-panel work per chunk the a-panel teams and the single issuing warp become the bottleneck) -
            // opt-in with Z3D_GEN2=1
            if (ntile == 2 && !pl->gen) break;
            HostPlanG hg;
            build_gen_plan(order, prop.multiProcessorCount, hp, hg, force, ntile);
            if (!hg.ok) continue;
            DevPlanG &g = ntile == 1 ? pl->devg : pl->devg2;
            GenIssue &issue = ntile == 1 ? pl->g_issue : pl->g2_issue;
            g.order = order; g.dim = dim; g.H = pl->H; g.nsets = hg.nsets; g.nsplits = hg.nsplits;
            g.tslot_bytes = hg.maxcols * 64;
            g.gslot_bytes = (40 + hg.maxseedrows * 8) * 4;
            g.gd_bytes = hg.gd_entries * 8;
            g.rk_bytes = hg.rk_entries * 16;
            g.off_rk = (G_OFF_GD + g.gd_bytes + 15) / 16 * 16;
            g.off_g = g.off_rk + g.rk_bytes;
            g.off_f = (g.off_g + G_DG * g.gslot_bytes + 127) / 128 * 128;
            g.off_t = (g.off_f + G_FOLD_BYTES + 127) / 128 * 128;
            g.dbg = 0;
            if (const char *e = getenv("Z3D_GEN_DBG")) g.dbg = atoi(e);
            g.nslots = std::min(G_NS_MAX, ((int)prop.sharedMemPerBlockOptin - g.off_t) / g.tslot_bytes);
            if (const char *e = getenv("Z3D_GEN_SLOTS")) g.nslots = std::max(2, std::min(g.nslots, atoi(e)));
            if (g.nslots < G_NTT || hg.nsets > G_MAXSETS) continue;   // a T team's first chunk behind a read-back must find a free slot (see z3d_gen.cuh)
            GenTabs *d_tabs; int4 *d_tasks; float2 *d_ggd; float4 *d_grk;
            UP(hg.tabs, d_tabs) UP(hg.tasks, d_tasks) UP(hg.gd, d_ggd) UP(hg.rk, d_grk)
            g.tabs = d_tabs; g.tasks = d_tasks; g.gd = d_ggd; g.rk = d_grk; g.side = d_side;
            for (int i = 0; i < hg.nsets; ++i) {
                const GenTabs &T = hg.tabs[i];
                issue.hdr[i] = make_int4(T.ncols, T.nseg, T.npanel, T.acol0);
                for (int s2 = 0; s2 < G_SEGS; ++s2) {
                    const bool v = s2 < T.nseg;
                    const int n = v ? T.seg_cols[s2] : 16;
                    issue.seg[i][s2] = make_int4((int)((1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(SM_ROWS >> 4) << 24)),
                                                 v ? T.seg_panel[s2] * 16 : 0, v ? T.seg_off[s2] : 0, v ? 0 : -1);
                }
            }
            if (ntile == 1) {
                UP(hg.colmap, pl->d_gcolmap)
                pl->g_smem = g.off_t + g.nslots * g.tslot_bytes;
                pl->g_ncols_all = hg.ncols_all;
                pl->g_maxcols = hg.maxcols;
                pl->b_chunks = (long long)analysis_items(0, pl->H, pl->H) * ((pl->H + KC - 1) / KC);
                pl->gen = true;
            } else {
                std::vector<int2> cm1 = hg.colmap;   // tile 1's accumulators follow tile 0's
                for (int2 &c : cm1) c.y += hg.tabs[c.x].ncols;
                UP(hg.colmap, pl->d_g2colmap[0]) UP(cm1, pl->d_g2colmap[1])
                pl->g2_smem = g.off_t + g.nslots * g.tslot_bytes;
                pl->g2_built = true;
                pl->gen2 = off2 && atoi(off2) != 0;
                // tensor-core synthesis: which moment every column of every set holds, and the sets of every fold kind (the planner
                // lays the sets out kind by kind)
                pl->synth_maxcols = (hg.maxcols + 15) / 16 * 16;
                std::vector<int> colsrc((size_t)hg.nsets * pl->synth_maxcols, -1);
                for (int mu = 0; mu < pl->nmom; ++mu) colsrc[(size_t)hg.colmap[mu].x * pl->synth_maxcols + hg.colmap[mu].y] = mu;
                UP(colsrc, pl->d_synth_colsrc)
                bool ordered = true;
                for (int i = 1; i < hg.nsets; ++i) ordered = ordered && hg.tabs[i].kind >= hg.tabs[i - 1].kind;
                // CTAs run in clusters of two column sets (2p, 2p + 1): where both are of the same kind, the odd one sends its class
                // partials to the even one through distributed shared memory and only the sum is stored (one slot for the pair)
                std::vector<int> pairinfo(hg.nsets, 0), slot_kind;
                const bool pairs_on = !(getenv("Z3D_SYNTH_PAIRS") && atoi(getenv("Z3D_SYNTH_PAIRS")) == 0);
                for (int i = 0; i < hg.nsets; ++i) {
                    const bool second = (i & 1) && pairs_on && hg.tabs[i].kind == hg.tabs[i - 1].kind;
                    const bool first = !(i & 1) && pairs_on && i + 1 < hg.nsets && hg.tabs[i].kind == hg.tabs[i + 1].kind;
                    if (!second) slot_kind.push_back(hg.tabs[i].kind);
                    pairinfo[i] = (((int)slot_kind.size() - 1) << 1) | ((first || second) ? 1 : 0);
                }
                pl->synth_nslots = (int)slot_kind.size();
                for (int k = 0, i = 0; k <= 4; ++k) {
                    while (i < pl->synth_nslots && slot_kind[i] < k) ++i;
                    pl->synth_kinds.first[k] = i;
                }
                UP(pairinfo, pl->d_synth_pairinfo)
                // its T ring: 128 B per row (the one MN-major layout the tensor core reads for 32-bit operands), in place of the fold rings
                pl->synth_off_t = (g.off_g + sy_gring_bytes(g.gslot_bytes) + 1024 + SY_XBUF_BYTES + 1023) / 1024 * 1024;   // seed ring, 1 KB of barriers, exchange buffers, T ring
                pl->synth_slots = std::min(G_NS_MAX, ((int)prop.sharedMemPerBlockOptin - pl->synth_off_t - 1024) / (2 * g.tslot_bytes));
                pl->synth_smem = pl->synth_off_t + 1024 + pl->synth_slots * 2 * g.tslot_bytes;
                pl->synth = ordered && pl->synth_slots >= G_NTT && !(getenv("Z3D_SYNTH") && atoi(getenv("Z3D_SYNTH")) == 0);
            }
        }
    }
#undef UP

    // break-even against the per-volume kernel: a streamed-T tile costs the same whatever it holds (3.9 ms at 128^3 / order 50,
    // i.e. 6.2 volumes of the per-volume kernel; 4-5 at 64^3 and 96^3 / order 60); the R / Q table kernel breaks even at ~9
    if (pl->stream) pl->batched_min = 8;
    if (const char *e = getenv("Z3D_BATCHED_MIN")) pl->batched_min = std::max(1, atoi(e));  // tuning / tests
    if (const char *e = getenv("Z3D_BATCHED_WIDE")) pl->batched_wide = std::max(1, atoi(e));
    if (const char *e = getenv("Z3D_GEN_MIN")) pl->gen_min = std::max(1, atoi(e));
    if (const char *e = getenv("Z3D_SYNTH_MIN")) pl->synth_min = std::max(1, atoi(e));
    rc = set_kernel_attrs((int)prop.sharedMemPerBlockOptin);  // per-function attribute: plans of any size coexist
    if (rc) {
        z3d_plan_destroy(pl);
        return rc;
    }
    *out = pl;
    return Z3D_OK;
}

// Host-only: the column-set layout the streamed-T kernel would use (no device needed; tests/test_cabi.py checks its
// invariants for every supported order).
extern "C" int z3d_stream_layout(int order, int num_sms, int *out, int n_out) {
    if (order < 0 || order > LMAX || num_sms < 1 || !out) return fail(Z3D_ERR_INVALID, "z3d_stream_layout: bad arguments");
    HostPlan hp;
    int rc = build_host_plan(order, hp);
    if (rc) return rc;
    HostPlanB hb;
    build_batched_plan(order, num_sms, BCfg<64>::TILES, hp, hb);
    HostPlanS hs;
    if (hb.nrad > 0) build_stream_plan(order, num_sms, hb, hs, 0);
    int v[8] = {hs.ok ? 1 : 0, hs.nsets, hs.nsplits, hs.ncols_all, hs.t_bytes_max / 64, 0, 0, 0};
    if (hs.ok) {  // self-check: every set within the kernel's limits, every moment in exactly one column, columns in range
        int bad = 0, maxpanel = 0, maxseg = 0;
        std::vector<char> seen((size_t)hs.ncols_all, 0);
        std::vector<int> base(hs.nsets + 1, 0);
        for (int i = 0; i < hs.nsets; ++i) {
            const StreamTabs &T = hs.tabs[i];
            base[i + 1] = base[i] + T.ncols;
            int cols = 0;
            for (int g = 0; g < T.nseg; ++g) {
                if (T.seg_cols[g] < 16 || T.seg_cols[g] > 256 || T.seg_cols[g] % 16 || T.seg_off[g] != cols) ++bad;
                if (T.seg_panel[g] < 0 || T.seg_panel[g] >= T.npanel) ++bad;
                cols += T.seg_cols[g];
            }
            if (cols != T.ncols || T.ncols > 512 || T.npanel < 1 || T.npanel > S_PANELS || T.nseg > S_SEGS || T.nkind != 1) ++bad;
            if (T.toff != 16LL * base[i]) ++bad;
            maxpanel = std::max(maxpanel, T.npanel);
            maxseg = std::max(maxseg, T.nseg);
        }
        for (const int2 &c : hs.colmap) {
            if (c.x < 0 || c.x >= hs.nsets || c.y < 0 || c.y >= hs.tabs[c.x].ncols || seen[base[c.x] + c.y]++) ++bad;
        }
        if (hs.nsets * hs.nsplits > num_sms || (int)hs.colsrc.size() != hs.ncols_all || base[hs.nsets] != hs.ncols_all) ++bad;
        v[5] = bad;
        v[6] = maxpanel;
        v[7] = maxseg;
    }
    for (int i = 0; i < n_out && i < 8; ++i) out[i] = v[i];
    return Z3D_OK;
}

// Host-only: the column-set layout of the table-free kernel (no device needed; tests/test_cabi.py checks its invariants).
// out[]: 0 available, 1 sets, 2 splits, 3 columns (all sets), 4 widest set, 5 violated invariants, 6 most panels, 7 most
// segments, 8 most chain tasks, 9 most seed rows, 10 estimated clocks per chunk of the slowest set
extern "C" int z3d_gen_layout_tiles(int order, int num_sms, int ntile, int *out, int n_out);
extern "C" int z3d_gen_layout(int order, int num_sms, int *out, int n_out) { return z3d_gen_layout_tiles(order, num_sms, 1, out, n_out); }
// ... of the layout for `ntile` (1 or 2) tiles of 64 volumes per generated T block
extern "C" int z3d_gen_layout_tiles(int order, int num_sms, int ntile, int *out, int n_out) {
    if (order < 0 || order > LMAX || num_sms < 1 || !out || ntile < 1 || ntile > 2) return fail(Z3D_ERR_INVALID, "z3d_gen_layout: bad arguments");
    HostPlan hp;
    int rc = build_host_plan(order, hp);
    if (rc) return rc;
    HostPlanG hg;
    int force = 0;
    if (const char *e = getenv("Z3D_GEN_SPLITS")) force = atoi(e);
    build_gen_plan(order, num_sms, hp, hg, force, ntile);
    if (getenv("Z3D_GEN_DUMP"))   // developer aid: the sets on stderr
        for (int i = 0; i < hg.nsets && hg.ok; ++i) {
            const GenTabs &T = hg.tabs[i];
            fprintf(stderr, "set %d: kind %d cols %d segs %d panels %d tasks %d seedrows %d lpar %d acol0 %d |", i, T.kind, T.ncols, T.nseg, T.npanel,
                    T.ntask, T.nseedrows, T.lpar, T.acol0);
            for (int j = 0; j < T.npanel; ++j) fprintf(stderr, " (m %d pz %d nl %d soff %d gd %d)", T.panel_m[j], T.panel_pz[j], T.panel_nl[j], T.panel_soff[j], T.panel_gd[j]);
            fprintf(stderr, " | segs");
            for (int g = 0; g < T.nseg; ++g) fprintf(stderr, " %d@%d:p%d", T.seg_cols[g], T.seg_off[g], T.seg_panel[g]);
            fprintf(stderr, "\n");
        }
    int v[11] = {hg.ok ? 1 : 0, hg.nsets, hg.nsplits, hg.ncols_all, hg.maxcols, 0, 0, 0, 0, hg.maxseedrows, (int)hg.est_period};
    if (hg.ok) {
        int bad = 0;
        std::vector<std::vector<char>> seen(hg.nsets);
        for (int i = 0; i < hg.nsets; ++i) {
            const GenTabs &T = hg.tabs[i];
            seen[i].assign(T.ncols, 0);
            int cols = 0;
            for (int g = 0; g < T.nseg; ++g) {
                if (T.seg_cols[g] < 16 || T.seg_cols[g] > 256 || T.seg_cols[g] % 16 || T.seg_off[g] != cols) ++bad;
                if (T.seg_panel[g] < 0 || T.seg_panel[g] >= T.npanel) ++bad;
                cols += T.seg_cols[g];
            }
            if (cols != T.ncols || ntile * (T.ncols + g_aslots(ntile) * 16 * T.npanel) > 512 || T.npanel < 1 || T.npanel > G_PANELS || T.nseg > G_SEGS) ++bad;
            if (T.ntask > G_TASKS || T.acol0 != 512 - ntile * g_aslots(ntile) * 16 * T.npanel) ++bad;
            std::vector<char> covered(T.ncols, 0);   // every chain task writes its own rows, twice (two k halves)
            for (int t = 0; t < T.ntask; ++t) {
                const int4 tk = hg.tasks[(size_t)i * G_TASKS + t];
                if (tk.y == 0) continue;   // filler lane of a short block
                if (tk.y < 1 || tk.x < 0 || tk.x + tk.y > T.ncols || (tk.w & 0xffff) >= T.nseedrows || tk.z < 0 || tk.z + tk.y > hg.rk_entries) { ++bad; continue; }
                for (int r = 0; r < tk.y; ++r) covered[tk.x + r] += 1;
            }
            for (int c = 0; c < T.ncols; ++c) if (covered[c] != 0 && covered[c] != 2) ++bad;
            v[6] = std::max(v[6], T.npanel);
            v[7] = std::max(v[7], T.nseg);
            v[8] = std::max(v[8], T.ntask);
        }
        for (const int2 &c : hg.colmap)
            if (c.x < 0 || c.x >= hg.nsets || c.y < 0 || c.y >= hg.tabs[c.x].ncols || seen[c.x][c.y]++) ++bad;
        if (hg.nsets * hg.nsplits > num_sms) ++bad;
        v[5] = bad;
    }
    for (int i = 0; i < n_out && i < 11; ++i) out[i] = v[i];
    return Z3D_OK;
}

extern "C" int z3d_plan_destroy(z3d_plan *pl) {
    if (!pl) return Z3D_OK;
    DeviceGuard _guard(pl->device);
    if (pl->s_comp) cudaStreamSynchronize(pl->s_comp);   // submitted host-buffer batches (z3d_decompose_host_submit) may still be in flight
    if (pl->s_copy) cudaStreamSynchronize(pl->s_copy);
    for (void *p : pl->owned) cudaFree(p);
    if (pl->d_stage_synth_ws) cudaFree(pl->d_stage_synth_ws);
    if (pl->d_Rt) cudaFree(pl->d_Rt);
    if (pl->d_Qt) cudaFree(pl->d_Qt);
    if (pl->d_trig) cudaFree(pl->d_trig);
    if (pl->d_Tt) cudaFree(pl->d_Tt);
    for (int i = 0; i < 2; ++i) {
        if (pl->d_stage_vol[i]) cudaFree(pl->d_stage_vol[i]);
        if (pl->d_stage_mom[i]) cudaFree(pl->d_stage_mom[i]);
        if (pl->d_stage_ws[i]) cudaFree(pl->d_stage_ws[i]);
        if (pl->ev_in[i]) cudaEventDestroy(pl->ev_in[i]);
        if (pl->ev_done[i]) cudaEventDestroy(pl->ev_done[i]);
    }
    if (pl->ev_k0) cudaEventDestroy(pl->ev_k0);
    if (pl->ev_k1) cudaEventDestroy(pl->ev_k1);
    for (int i = 0; i < z3d_plan::MAX_TIMED; ++i) {
        if (pl->ev_t0[i]) cudaEventDestroy(pl->ev_t0[i]);
        if (pl->ev_t1[i]) cudaEventDestroy(pl->ev_t1[i]);
    }
    if (pl->s_copy) cudaStreamDestroy(pl->s_copy);
    if (pl->s_comp) cudaStreamDestroy(pl->s_comp);
    delete pl;
    return Z3D_OK;
}

extern "C" int z3d_plan_info(const z3d_plan *pl, int *info, int n) {
    if (!pl || !info) return fail(Z3D_ERR_INVALID, "z3d_plan_info: NULL argument");
    const long long t_mb = pl->stream ? (long long)((double)pl->b_chunks * (double)pl->devs.trec * 4.0 / 1048576.0) : 0;
    const int v[39] = {pl->order, pl->dim, pl->nmom, pl->npass, pl->npass * pl->nsplit, NT, TS, pl->smem_bytes,
                       KC, pl->num_sms, pl->useful_acc, pl->padded_acc,
                       pl->batched ? pl->batched_min : 0, pl->devb[0].ngroups, pl->devb[1].ngroups, pl->batched ? pl->batched_wide : 0,
                       pl->devb[0].nrounds, pl->devb[1].nrounds,
                       // streamed-T variant: 18 state (0 unavailable, 1 planned - the T table is built on the first batched call and
                       // may still be refused for lack of device memory, 2 table resident), 19 column sets, 20 splits of the chunk
                       // list, 21 moment rows padded to 16 (all sets), 22 T table MiB, 23 chunks of 8 folded voxels
                       (pl->batched && pl->stream) ? (pl->stream_ready ? 2 : 1) : 0, pl->devs.nsets, pl->devs.nsplits, pl->s_ncols_all,
                       (int)std::min<long long>(t_mb, 0x7fffffff), (int)std::min<long long>(pl->b_chunks, 0x7fffffff),
                       // table-free tensor-core variant: 24 available, 25 column sets, 26 splits, 27 moment rows padded to 16 per
                       // chain piece, 28 T-ring slots, 29 smallest batch routed to it, 30 basis-table level in use, 31 widest set
                       pl->gen ? 1 : 0, pl->devg.nsets, pl->devg.nsplits, pl->g_ncols_all, pl->devg.nslots, pl->gen ? pl->gen_min : 0,
                       pl->tables_level, pl->g_maxcols,
                       // its two-tile layout (128 volumes per generated T block): 32 available, 33 column sets, 34 splits, 35 T-ring slots
                       pl->gen2 ? 1 : 0, pl->devg2.nsets, pl->devg2.nsplits, pl->devg2.nslots,
                       // tensor-core batched synthesis: 36 available, 37 smallest batch of maps the host layer routes to it, 38 partial-fold slots
                       pl->synth ? 1 : 0, pl->synth ? pl->synth_min : 0, pl->synth ? pl->synth_nslots : 0};
    for (int i = 0; i < n && i < 39; ++i) info[i] = v[i];
    return Z3D_OK;
}

// The per-CTA partial sums of the analysis kernel are 24 MB per volume at order 50.  A batch is processed in groups of
// a few volumes (kernel + reduction per group, same scratch) so that the partials never leave the 126 MB L2: the
// launch's DRAM traffic is then the volumes themselves, and the scratch stays small for any batch.
static int decompose_group(const z3d_plan *pl) {
    const size_t per_vol = (size_t)pl->npass * pl->nsplit * SLOT_WORDS * sizeof(float);
    return (int)std::max<size_t>(1, ((size_t)64 << 20) / per_vol);
}
static size_t ws_decompose(const z3d_plan *pl, int batch) {
    size_t w = (size_t)std::min(batch, decompose_group(pl)) * pl->npass * pl->nsplit * SLOT_WORDS * sizeof(float);
    if (pl->tables_level > 0 && pl->batched && batch >= pl->batched_min)  // per-(CTA, segment) partial sums + the folds of one tile of volumes
        w = std::max(w, (size_t)pl->num_sms * std::max(pl->devb[0].nrounds, pl->devb[1].nrounds) * BSLOT_WORDS * sizeof(float) +
                            (size_t)pl->b_chunks * ((pl->stream || batch >= pl->batched_wide) ? BCfg<64>::F_BYTES : BCfg<32>::F_BYTES));
    if (pl->gen && batch >= pl->gen_min)  // table-free tensor-core kernel: per-CTA partial sums + the folds of one tile of 64 volumes
        w = std::max(w, (size_t)pl->num_sms * BSLOT_WORDS * sizeof(float) + (size_t)pl->b_chunks * 4 * S_FK_BYTES * ((pl->gen2 && batch > SNV) ? 2 : 1));
    return w;
}
static size_t ws_reconstruct(const z3d_plan *pl, int batch) {
    const size_t N3 = (size_t)pl->dim * pl->dim * pl->dim;
    return ((size_t)batch * pl->npass * SLOT_WORDS + (size_t)pl->npass * batch * N3) * sizeof(float);
}
extern "C" size_t z3d_workspace_bytes(const z3d_plan *pl, int batch) {
    if (!pl || batch < 1) return 0;
    return std::max(ws_decompose(pl, batch), ws_reconstruct(pl, batch));
}

static int check_range(const z3d_plan *pl, int batch, int lo, int hi, const char *who) {
    if (!pl) return fail(Z3D_ERR_INVALID, "%s: plan is NULL", who);
    if (batch < 1) return fail(Z3D_ERR_INVALID, "%s: batch %d < 1", who, batch);
    if (lo < 0 || hi > pl->H || lo > hi) return fail(Z3D_ERR_INVALID, "%s: pair range [%d, %d) outside [0, %d]", who, lo, hi, pl->H);
    return Z3D_OK;
}

static int decompose_impl(const z3d_plan *pl, const float *d_vol, int batch, int half_lo, int half_hi, int part, int nparts,
                          float *d_moments, void *d_ws, size_t ws_bytes, void *stream);

extern "C" int z3d_decompose(const z3d_plan *pl, const float *d_vol, int batch, int half_lo, int half_hi,
                             float *d_moments, void *d_ws, size_t ws_bytes, void *stream) {
    return decompose_impl(pl, d_vol, batch, half_lo, half_hi, 0, 1, d_moments, d_ws, ws_bytes, stream);
}

extern "C" int z3d_decompose_part(const z3d_plan *pl, const float *d_vol, int batch, int part, int nparts,
                                  float *d_moments, void *d_ws, size_t ws_bytes, void *stream) {
    if (!pl) return fail(Z3D_ERR_INVALID, "z3d_decompose_part: plan is NULL");
    if (nparts < 1 || part < 0 || part >= nparts) return fail(Z3D_ERR_INVALID, "z3d_decompose_part: part %d of %d", part, nparts);
    return decompose_impl(pl, d_vol, batch, 0, pl->H, part, nparts, d_moments, d_ws, ws_bytes, stream);
}

// Basis tables of the batched path (what is left of CalculatePolynomials, zm:2237-2323): built on the first batched
// call, on that call's stream.  When device memory is short the plan silently keeps to the per-volume kernel.
static int ensure_basis(z3d_plan *pl, cudaStream_t st) {
    if (pl->basis_ready || !pl->batched) return Z3D_OK;
    const size_t items = (size_t)analysis_items(0, pl->H, pl->H);
    const size_t rb = (size_t)pl->b_chunks * pl->basis.nrad * KC * sizeof(float), qb = (size_t)pl->b_chunks * pl->basis.nq * KC * sizeof(float);
    const size_t tb = items * 4 * (LMAX + 1) * sizeof(float);
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    if (rb + qb + tb > free_b / 2 || cudaMalloc((void **)&pl->d_Rt, rb) != cudaSuccess ||
        cudaMalloc((void **)&pl->d_Qt, qb) != cudaSuccess || cudaMalloc((void **)&pl->d_trig, tb) != cudaSuccess) {
        cudaGetLastError();
        if (pl->d_Rt) cudaFree(pl->d_Rt);
        if (pl->d_Qt) cudaFree(pl->d_Qt);
        pl->d_Rt = pl->d_Qt = pl->d_trig = nullptr;
        return Z3D_OK;
    }
    pl->basis.Rt = pl->d_Rt; pl->basis.Qt = pl->d_Qt; pl->basis.trig = pl->d_trig;
    for (int v = 0; v < 2; ++v) {
        pl->devb[v].Rt = pl->d_Rt; pl->devb[v].Qt = pl->d_Qt; pl->devb[v].trig = pl->d_trig;
    }
    CUDA_TRY(cudaMemsetAsync(pl->d_trig, 0, tb, st));
    k_build_basis<<<(unsigned)pl->b_chunks, 1024, 0, st>>>(pl->basis);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    pl->basis_ready = true;
    return Z3D_OK;
}

// T table of the streamed variant (13.7 GB at 128^3 / order 50): built from the R / Q tables on the first batched call when it
// fits in 45 % of the free device memory; otherwise the plan keeps to k_decompose_batched (R / Q tables, products in-kernel).
static int ensure_stream(z3d_plan *pl, cudaStream_t st, size_t max_bytes) {
    if (pl->stream_ready || !pl->stream || !pl->batched || !pl->basis_ready) return Z3D_OK;
    const size_t bytes = (size_t)pl->b_chunks * (size_t)pl->devs.trec * sizeof(float);
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    size_t limit = (size_t)(0.45 * (double)free_b);
    if (max_bytes) limit = std::min(limit, max_bytes);
    if (bytes > limit || cudaMalloc((void **)&pl->d_Tt, bytes) != cudaSuccess) {
        cudaGetLastError();
        pl->d_Tt = nullptr;
        return Z3D_OK;
    }
    pl->devs.Tt = pl->d_Tt;
    pl->devs.trig = pl->d_trig;
    k_build_T<<<(unsigned)pl->b_chunks, 1024, 0, st>>>(pl->d_Rt, pl->d_Qt, pl->basis.nrad, pl->basis.nq, pl->d_colsrc,
                                                        pl->s_ncols_all, pl->devs.trec, pl->d_Tt);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    pl->stream_ready = true;
    return Z3D_OK;
}

// Opt-in basis tables (NOT needed by the default path, which generates everything in registers): level 1 builds the folded
// R_nl / Q_lm factor tables (k_decompose_batched), level 2 also the product table T = R Q that k_decompose_stream reads with the
// TMA engine (13.7 GB at 128^3 / order 50).  Synchronous: the tables are complete when the call returns, so later calls on any
// stream may use them.  max_bytes = 0: up to 45 % of the free device memory.  Returns the level actually reached in *level_out
// (a table that does not fit is not an error).  Level 0 frees the tables and returns the plan to the table-free kernels.
extern "C" int z3d_plan_prepare_tables(z3d_plan *pl, int level, size_t max_bytes, int *level_out) {
    if (!pl || level < 0 || level > 2) return fail(Z3D_ERR_INVALID, "z3d_plan_prepare_tables: bad arguments");
    USE_DEVICE(pl->device);
    CUDA_TRY(cudaDeviceSynchronize());
    if (level == 0) {
        if (pl->d_Tt) cudaFree(pl->d_Tt);
        if (pl->d_Rt) cudaFree(pl->d_Rt);
        if (pl->d_Qt) cudaFree(pl->d_Qt);
        if (pl->d_trig) cudaFree(pl->d_trig);
        pl->d_Tt = pl->d_Rt = pl->d_Qt = pl->d_trig = nullptr;
        pl->basis_ready = pl->stream_ready = false;
        pl->tables_level = 0;
        if (level_out) *level_out = 0;
        return Z3D_OK;
    }
    int reached = 0;
    if (pl->batched) {
        int rc = ensure_basis(pl, nullptr);
        if (rc) return rc;
        if (pl->basis_ready) reached = 1;
        if (reached == 1 && level == 2 && pl->stream) {
            rc = ensure_stream(pl, nullptr, max_bytes);
            if (rc) return rc;
            if (pl->stream_ready) reached = 2;
        }
        CUDA_TRY(cudaDeviceSynchronize());
    }
    pl->tables_level = reached;
    if (level_out) *level_out = reached;
    return Z3D_OK;
}

static int decompose_impl(const z3d_plan *pl, const float *d_vol, int batch, int half_lo, int half_hi, int part, int nparts,
                          float *d_moments, void *d_ws, size_t ws_bytes, void *stream) {
    int rc = check_range(pl, batch, half_lo, half_hi, "z3d_decompose");
    if (rc) return rc;
    if (!d_vol || !d_moments || !d_ws) return fail(Z3D_ERR_INVALID, "z3d_decompose: NULL buffer");
    if (ws_bytes < ws_decompose(pl, batch))
        return fail(Z3D_ERR_WORKSPACE, "z3d_decompose: workspace %zu < %zu bytes", ws_bytes, ws_decompose(pl, batch));
    USE_DEVICE(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    float *partial = (float *)d_ws;
    const int grp = decompose_group(pl);
    const size_t N3 = (size_t)pl->dim * pl->dim * pl->dim;
    z3d_plan *mpl = const_cast<z3d_plan *>(pl);  // timing bookkeeping only
    mpl->n_timed = 0;
    mpl->last_was_decompose = true;
    int done = 0;
    // batches: tiles of up to BNV volumes through the batched kernel (whole-volume calls only); what is left below
    // the threshold goes through the per-volume kernel
    const bool whole = half_lo == 0 && half_hi == pl->H;
    // default for batches: the table-free tensor-core kernel (tiles of up to 64 volumes; a tile costs the same whatever it holds)
    // while more than one tile is left: two tiles per launch, which share every generated T block
    while (pl->gen && pl->tables_level == 0 && whole && batch - done >= pl->gen_min) {
        const bool two = pl->gen2 && batch - done > SNV;
        const int nb = std::min(two ? 2 * SNV : SNV, batch - done);
        const DevPlanG &dp = two ? pl->devg2 : pl->devg;
        const float *vol = d_vol + (size_t)done * N3;
        float *folds = partial + (size_t)pl->num_sms * BSLOT_WORDS;
        const size_t tile_words = (size_t)pl->b_chunks * S_FK_BYTES;   // 4 kinds x S_FK_BYTES per chunk, in floats
        float *mom = d_moments + (size_t)done * 2 * pl->nmom;
        k_fold_tile_gen<<<(unsigned)pl->b_chunks, 2 * SNV * KC, 0, st>>>(vol, std::min(nb, SNV), pl->dim, pl->H, folds);
        if (two) k_fold_tile_gen<<<(unsigned)pl->b_chunks, 2 * SNV * KC, 0, st>>>(vol + (size_t)SNV * N3, nb - SNV, pl->dim, pl->H, folds + tile_words);
        CUDA_TRY(cudaGetLastError());
        const bool timed = pl->profiling && mpl->n_timed < z3d_plan::MAX_TIMED;
        if (timed) CUDA_TRY(cudaEventRecord(mpl->ev_t0[mpl->n_timed], st));
        if (two) k_decompose_gen<2><<<dim3(dp.nsplits, dp.nsets), GNT, pl->g2_smem, st>>>(dp, pl->g2_issue, folds, tile_words, part, nparts, partial);
        else k_decompose_gen<1><<<dim3(dp.nsplits, dp.nsets), GNT, pl->g_smem, st>>>(dp, pl->g_issue, folds, 0, part, nparts, partial);
        CUDA_TRY(cudaGetLastError());
        if (timed) CUDA_TRY(cudaEventRecord(mpl->ev_t1[mpl->n_timed++], st));
        const unsigned rblocks = (unsigned)((pl->nmom * SM_ROWS + 255) / 256);
        if (two) {
            k_reduce_stream<<<rblocks, 256, 0, st>>>(partial, pl->d_g2colmap[0], pl->d_scale, pl->nmom, dp.nsplits, SNV, mom);
            k_reduce_stream<<<rblocks, 256, 0, st>>>(partial, pl->d_g2colmap[1], pl->d_scale, pl->nmom, dp.nsplits, nb - SNV, mom + (size_t)SNV * 2 * pl->nmom);
        } else {
            k_reduce_stream<<<rblocks, 256, 0, st>>>(partial, pl->d_gcolmap, pl->d_scale, pl->nmom, dp.nsplits, nb, mom);
        }
        CUDA_TRY(cudaGetLastError());
        g_launches += two ? 5 : 3;
        done += nb;
    }
    // opt-in (z3d_plan_prepare_tables): the table-fed kernels
    while (pl->tables_level == 2 && pl->batched && pl->stream_ready && whole && batch - done >= pl->batched_min) {
        // streamed-T variant: tiles of up to 64 volumes (a tile costs the same whatever it holds)
        const int nb = std::min(SNV, batch - done);
        const DevPlanS &dp = pl->devs;
        const float *vol = d_vol + (size_t)done * N3;
        float *folds = partial + (size_t)pl->num_sms * std::max(pl->devb[0].nrounds, pl->devb[1].nrounds) * BSLOT_WORDS;
        float *mom = d_moments + (size_t)done * 2 * pl->nmom;
        k_fold_tile_kind<<<(unsigned)pl->b_chunks, 2 * SNV * KC, 0, st>>>(vol, nb, pl->dim, pl->H, folds);
        CUDA_TRY(cudaGetLastError());
        const bool timed = pl->profiling && mpl->n_timed < z3d_plan::MAX_TIMED;
        if (timed) CUDA_TRY(cudaEventRecord(mpl->ev_t0[mpl->n_timed], st));
        k_decompose_stream<<<dp.nsets * dp.nsplits, SNT, pl->s_smem, st>>>(dp, folds, part, nparts, partial);
        CUDA_TRY(cudaGetLastError());
        if (timed) CUDA_TRY(cudaEventRecord(mpl->ev_t1[mpl->n_timed++], st));
        k_reduce_stream<<<(pl->nmom * SM_ROWS + 255) / 256, 256, 0, st>>>(partial, pl->d_colmap, pl->d_scale, pl->nmom, dp.nsplits, nb, mom);
        CUDA_TRY(cudaGetLastError());
        g_launches += 3;
        done += nb;
    }
    while (pl->tables_level >= 1 && pl->batched && pl->basis_ready && whole && batch - done >= pl->batched_min) {
        // a tile of up to 64 volumes when enough are left, else of up to 32 (a tile costs the same whatever it holds)
        const int wide = (batch - done >= pl->batched_wide) ? 1 : 0;
        const int nb = std::min(wide ? 64 : 32, batch - done);
        const DevPlanB &dp = pl->devb[wide];
        const float *vol = d_vol + (size_t)done * N3;
        float *folds = partial + (size_t)pl->num_sms * std::max(pl->devb[0].nrounds, pl->devb[1].nrounds) * BSLOT_WORDS;
        float *mom = d_moments + (size_t)done * 2 * pl->nmom;
        if (wide) k_fold_tile<64><<<(unsigned)pl->b_chunks, 2 * 64 * KC, 0, st>>>(vol, nb, pl->dim, pl->H, folds);
        else k_fold_tile<32><<<(unsigned)pl->b_chunks, 2 * 32 * KC, 0, st>>>(vol, nb, pl->dim, pl->H, folds);
        CUDA_TRY(cudaGetLastError());
        // one launch per round of dp.gpr row groups: groups x floor(SMs / groups) CTAs sweep the chunk list in lock step
        const bool timed = pl->profiling && mpl->n_timed < z3d_plan::MAX_TIMED;
        if (timed) CUDA_TRY(cudaEventRecord(mpl->ev_t0[mpl->n_timed], st));
        for (int r = 0; r < dp.nrounds; ++r) {
            const int g0 = r * dp.gpr, gr = std::min(dp.gpr, dp.ngroups - g0), cpg = pl->num_sms / gr;
            float *slots = partial + (size_t)r * pl->num_sms * BSLOT_WORDS;
            if (wide) k_decompose_batched<64><<<gr * cpg, BNT, pl->b_smem[1], st>>>(dp, folds, part, nparts, g0, cpg, slots);
            else k_decompose_batched<32><<<gr * cpg, BNT, pl->b_smem[0], st>>>(dp, folds, part, nparts, g0, cpg, slots);
            CUDA_TRY(cudaGetLastError());
        }
        if (timed) CUDA_TRY(cudaEventRecord(mpl->ev_t1[mpl->n_timed++], st));
        if (wide)
            k_reduce_batched<64><<<(pl->nmom * 128 + 255) / 256, 256, 0, st>>>(partial, pl->d_rowmap[1], pl->d_scale, pl->nmom, dp.ngroups,
                                                                             dp.gpr, pl->num_sms, nb, mom);
        else
            k_reduce_batched<32><<<(pl->nmom * 64 + 255) / 256, 256, 0, st>>>(partial, pl->d_rowmap[0], pl->d_scale, pl->nmom, dp.ngroups,
                                                                            dp.gpr, pl->num_sms, nb, mom);
        CUDA_TRY(cudaGetLastError());
        g_launches += 2 + dp.nrounds;
        done += nb;
    }
    for (int b0 = done; b0 < batch; b0 += grp) {
        const int nb = std::min(grp, batch - b0);
        const float *vol = d_vol + (size_t)b0 * N3;
        const bool timed = pl->profiling && mpl->n_timed < z3d_plan::MAX_TIMED;
        if (timed) CUDA_TRY(cudaEventRecord(mpl->ev_t0[mpl->n_timed], st));
        if (pl->npass == 2)
            k_decompose<1><<<pl->npass * pl->nsplit, NT, pl->smem_bytes, st>>>(pl->dev, vol, nb, half_lo, half_hi, part, nparts, partial);
        else
            k_decompose<0><<<pl->npass * pl->nsplit, NT, pl->smem_bytes, st>>>(pl->dev, vol, nb, half_lo, half_hi, part, nparts, partial);
        CUDA_TRY(cudaGetLastError());
        if (timed) CUDA_TRY(cudaEventRecord(mpl->ev_t1[mpl->n_timed++], st));
        dim3 g((2 * pl->nmom + 255) / 256, nb);
        k_reduce_moments<<<g, 256, 0, st>>>(partial, pl->d_src, pl->d_scale, pl->nmom, pl->npass, pl->nsplit, nb,
                                            d_moments + (size_t)b0 * 2 * pl->nmom);
        CUDA_TRY(cudaGetLastError());
        g_launches += 2;
    }
    return Z3D_OK;
}

extern "C" int z3d_reconstruct(const z3d_plan *pl, const float *d_moments, int batch, int half_lo, int half_hi,
                               float *d_vol, void *d_ws, size_t ws_bytes, void *stream) {
    int rc = check_range(pl, batch, half_lo, half_hi, "z3d_reconstruct");
    if (rc) return rc;
    if (!d_vol || !d_moments || !d_ws) return fail(Z3D_ERR_INVALID, "z3d_reconstruct: NULL buffer");
    if (ws_bytes < ws_reconstruct(pl, batch))
        return fail(Z3D_ERR_WORKSPACE, "z3d_reconstruct: workspace %zu < %zu bytes", ws_bytes, ws_reconstruct(pl, batch));
    if (half_lo == half_hi) return Z3D_OK;
    USE_DEVICE(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    float *otiles = (float *)d_ws;
    float *pvol = otiles + (size_t)batch * pl->npass * SLOT_WORDS;
    dim3 g((pl->npass * SLOT_WORDS + 255) / 256, batch);
    k_prepare_tiles<<<g, 256, 0, st>>>(d_moments, pl->d_omap, pl->d_ofac, pl->nmom, pl->npass, batch, otiles);
    CUDA_TRY(cudaGetLastError());
    const_cast<z3d_plan *>(pl)->last_was_decompose = false;
    if (pl->profiling) CUDA_TRY(cudaEventRecord(pl->ev_k0, st));
    if (pl->npass == 2)
        k_reconstruct<1><<<pl->npass * pl->nsplit, NT, pl->smem_bytes, st>>>(pl->dev, otiles, batch, half_lo, half_hi, pvol);
    else
        k_reconstruct<0><<<pl->npass * pl->nsplit, NT, pl->smem_bytes, st>>>(pl->dev, otiles, batch, half_lo, half_hi, pvol);
    CUDA_TRY(cudaGetLastError());
    if (pl->profiling) CUDA_TRY(cudaEventRecord(pl->ev_k1, st));
    const size_t total = (size_t)2 * (half_hi - half_lo) * pl->dim * pl->dim * batch;
    k_sum_passes<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pvol, pl->npass, batch, pl->dim, half_lo, half_hi, d_vol);
    CUDA_TRY(cudaGetLastError());
    g_launches += 3;
    return Z3D_OK;
}

// ---- tensor-core batched synthesis (z3d_synth.cuh): tiles of 64 maps, the chunk list in parts so that the partial folds of a part
// (sets x chunks x 8 KB) stay below ~3.5 GB (every part is a launch that reloads O' into tensor memory: 3 parts at 128^3 / order 50)
static int synth_parts(const z3d_plan *pl) {
    const double bytes = (double)pl->synth_nslots * (double)pl->b_chunks * 2.0 * SM_ROWS * KC * sizeof(float);
    return std::max(1, (int)std::ceil(bytes / 3.5e9));
}
static size_t synth_part_floats(const z3d_plan *pl) {
    const long long per = (pl->b_chunks + synth_parts(pl) - 1) / synth_parts(pl) + 1;
    return (size_t)pl->synth_nslots * (size_t)per * 2 * SM_ROWS * KC;
}
static size_t synth_op_floats(const z3d_plan *pl) { return (size_t)pl->devg2.nsets * SM_ROWS * 2 * pl->synth_maxcols; }

extern "C" size_t z3d_reconstruct_batched_workspace_bytes(const z3d_plan *pl, int batch) {
    if (!pl || batch < 1 || !pl->synth) return 0;
    return (synth_op_floats(pl) + synth_part_floats(pl)) * sizeof(float);
}

extern "C" int z3d_reconstruct_batched(const z3d_plan *pl, const float *d_moments, int batch, float *d_vol, void *d_ws, size_t ws_bytes,
                                       void *stream) {
    if (!pl) return fail(Z3D_ERR_INVALID, "z3d_reconstruct_batched: plan is NULL");
    if (batch < 1 || !d_moments || !d_vol || !d_ws) return fail(Z3D_ERR_INVALID, "z3d_reconstruct_batched: bad arguments");
    if (!pl->synth) return fail(Z3D_ERR_UNSUPPORTED, "z3d_reconstruct_batched: no tensor-core synthesis layout for order %d / box %d", pl->order, pl->dim);
    if (ws_bytes < z3d_reconstruct_batched_workspace_bytes(pl, batch))
        return fail(Z3D_ERR_WORKSPACE, "z3d_reconstruct_batched: workspace %zu < %zu bytes", ws_bytes, z3d_reconstruct_batched_workspace_bytes(pl, batch));
    USE_DEVICE(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    float *Op = (float *)d_ws, *Fs = Op + synth_op_floats(pl);
    const DevPlanG &dp = pl->devg2;
    const size_t N3 = (size_t)pl->dim * pl->dim * pl->dim;
    const int nparts = synth_parts(pl);
    const_cast<z3d_plan *>(pl)->last_was_decompose = false;
    for (int b0 = 0; b0 < batch; b0 += SNV) {
        const int nb = std::min(SNV, batch - b0);
        k_prepare_synth<<<dim3(dp.nsets, SM_ROWS), 128, 0, st>>>(d_moments + (size_t)b0 * pl->nmom * 2, pl->d_synth_colsrc, pl->d_ofac,
                                                                 pl->synth_maxcols, pl->nmom, nb, Op);
        CUDA_TRY(cudaGetLastError());
        ++g_launches;
        if (pl->profiling && b0 == 0) CUDA_TRY(cudaEventRecord(pl->ev_k0, st));
        for (int part = 0; part < nparts; ++part) {
            const long long gbeg = pl->b_chunks * part / nparts, glen = pl->b_chunks * (part + 1) / nparts - gbeg;
            if (glen <= 0) continue;
            {   // clusters of two CTAs along the set dimension (an odd number of sets gets one padding CTA)
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(dp.nsplits, (dp.nsets + 1) / 2 * 2);
                cfg.blockDim = dim3(GNT);
                cfg.dynamicSmemBytes = pl->synth_smem;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = 1;
                attr[0].val.clusterDim.y = 2;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                CUDA_TRY(cudaLaunchKernelEx(&cfg, k_synth_gen, dp, (const float *)Op, pl->synth_maxcols, pl->synth_off_t, pl->synth_slots, part, nparts,
                                            (const int *)pl->d_synth_pairinfo, Fs));
            }
            CUDA_TRY(cudaGetLastError());
            k_unfold_tile<<<(unsigned)glen, 2 * SNV * KC, 0, st>>>(Fs, gbeg, glen, pl->synth_kinds, nb, pl->dim, pl->H, d_vol + (size_t)b0 * N3);
            CUDA_TRY(cudaGetLastError());
            g_launches += 2;
        }
        if (pl->profiling && b0 == 0) CUDA_TRY(cudaEventRecord(pl->ev_k1, st));
    }
    return Z3D_OK;
}

extern "C" int z3d_profile_enable(z3d_plan *pl, int on) {
    if (!pl) return fail(Z3D_ERR_INVALID, "z3d_profile_enable: plan is NULL");
    USE_DEVICE(pl->device);
    if (on && !pl->ev_k0) {
        CUDA_TRY(cudaEventCreate(&pl->ev_k0));
        CUDA_TRY(cudaEventCreate(&pl->ev_k1));
        for (int i = 0; i < z3d_plan::MAX_TIMED; ++i) {
            CUDA_TRY(cudaEventCreate(&pl->ev_t0[i]));
            CUDA_TRY(cudaEventCreate(&pl->ev_t1[i]));
        }
    }
    pl->profiling = on != 0;
    return Z3D_OK;
}

extern "C" int z3d_profile_last_ms(z3d_plan *pl, float *ms) {
    if (!pl || !ms || !pl->ev_k0) return fail(Z3D_ERR_INVALID, "z3d_profile_last_ms: profiling was not enabled");
    if (pl->last_was_decompose) {  // sum over the analysis-kernel launches of the last call (one per volume group)
        float tot = 0.f;
        for (int i = 0; i < pl->n_timed; ++i) {
            float t = 0.f;
            CUDA_TRY(cudaEventSynchronize(pl->ev_t1[i]));
            CUDA_TRY(cudaEventElapsedTime(&t, pl->ev_t0[i], pl->ev_t1[i]));
            tot += t;
        }
        *ms = tot;
        return Z3D_OK;
    }
    CUDA_TRY(cudaEventSynchronize(pl->ev_k1));
    CUDA_TRY(cudaEventElapsedTime(ms, pl->ev_k0, pl->ev_k1));
    return Z3D_OK;
}

extern "C" int z3d_invariants(const z3d_plan *pl, const float *d_moments, int batch, float *d_inv, void *stream) {
    if (!pl || !d_moments || !d_inv || batch < 1) return fail(Z3D_ERR_INVALID, "z3d_invariants: bad arguments");
    USE_DEVICE(pl->device);
    dim3 g((pl->nrad + 127) / 128, batch);
    k_invariants<<<g, 128, 0, (cudaStream_t)stream>>>(d_moments, pl->d_nl_base, pl->d_nl_l, pl->nrad, pl->nmom, batch, d_inv);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return Z3D_OK;
}

extern "C" int z3d_zcc(const z3d_plan *pl, const float *d_moments_a, const float *d_moments_b, int batch, float *d_out, void *stream) {
    if (!pl || !d_moments_a || !d_moments_b || !d_out || batch < 1) return fail(Z3D_ERR_INVALID, "z3d_zcc: bad arguments");
    if (batch > 65535) return fail(Z3D_ERR_INVALID, "z3d_zcc: at most 65535 pairs per call");
    USE_DEVICE(pl->device);
    dim3 g(pl->order + 1, batch);
    k_zcc<<<g, 32, 0, (cudaStream_t)stream>>>(d_moments_a, d_moments_b, pl->order, pl->nmom, batch, d_out);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return Z3D_OK;
}

extern "C" int z3d_crop_bin(const float *d_in, int batch, int n_in, int cz, int cy, int cx, int crop, int n_out, float *d_out,
                            void *stream) {
    if (!d_in || !d_out || batch < 1 || batch > 65535 || n_in < 1 || n_out < 1 || crop < 0)
        return fail(Z3D_ERR_INVALID, "z3d_crop_bin: bad arguments");
    int nc = n_in, oz = 0, oy = 0, ox = 0;
    if (crop > 0) {  // v[c - half : c + half] on every axis (zm:2796-2800)
        const int half = (crop - crop % 2) / 2;
        nc = 2 * half;
        oz = cz - half; oy = cy - half; ox = cx - half;
        if (nc < 1 || oz < 0 || oy < 0 || ox < 0 || oz + nc > n_in || oy + nc > n_in || ox + nc > n_in)
            return fail(Z3D_ERR_INVALID, "z3d_crop_bin: crop box %d around (%d, %d, %d) leaves the %d^3 volume", crop, cz, cy, cx, n_in);
    }
    if (n_out > nc) return fail(Z3D_ERR_INVALID, "z3d_crop_bin: output box %d larger than the cropped box %d", n_out, nc);
    const size_t per = (size_t)n_out * n_out * n_out;
    dim3 g((unsigned)((per + 255) / 256), batch);
    k_crop_bin<<<g, 256, 0, (cudaStream_t)stream>>>(d_in, n_in, oz, oy, ox, nc, n_out, d_out);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return Z3D_OK;
}

extern "C" int z3d_scale_moments(const z3d_plan *pl, const float *d_dims, int ncomp, const float *d_latents, int npoints,
                                 float *d_out, void *stream) {
    if (!pl || !d_dims || !d_out || ncomp < 0 || npoints < 1 || (ncomp > 0 && !d_latents))
        return fail(Z3D_ERR_INVALID, "z3d_scale_moments: bad arguments");
    if (npoints > 65535) return fail(Z3D_ERR_INVALID, "z3d_scale_moments: at most 65535 latent coordinates per call");
    USE_DEVICE(pl->device);
    const int nelem = 2 * pl->nmom;
    dim3 g((nelem + 255) / 256, npoints);
    k_scale_moments<<<g, 256, 0, (cudaStream_t)stream>>>(d_dims, ncomp, d_latents, npoints, nelem, d_out);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return Z3D_OK;
}

// ------------------------------------------------------------------------------------------------
// host-buffer calls: double-buffered groups of volumes, H2D on one stream, compute + D2H on another
// ------------------------------------------------------------------------------------------------
static void free_staging(z3d_plan *pl) {
    for (int i = 0; i < 2; ++i) {
        if (pl->d_stage_vol[i]) cudaFree(pl->d_stage_vol[i]);
        if (pl->d_stage_mom[i]) cudaFree(pl->d_stage_mom[i]);
        if (pl->d_stage_ws[i]) cudaFree(pl->d_stage_ws[i]);
        if (pl->ev_in[i]) cudaEventDestroy(pl->ev_in[i]);
        if (pl->ev_done[i]) cudaEventDestroy(pl->ev_done[i]);
        pl->d_stage_vol[i] = pl->d_stage_mom[i] = nullptr;
        pl->d_stage_ws[i] = nullptr;
        pl->ev_in[i] = pl->ev_done[i] = nullptr;
    }
    if (pl->d_stage_synth_ws) cudaFree(pl->d_stage_synth_ws);
    pl->d_stage_synth_ws = nullptr;
    pl->stage_synth_bytes = 0;
    if (pl->s_copy) cudaStreamDestroy(pl->s_copy);
    if (pl->s_comp) cudaStreamDestroy(pl->s_comp);
    pl->s_copy = pl->s_comp = nullptr;
    pl->staged_ok = false;
}

// Staging of the host-buffer calls; `staged_ok` is set only when every allocation succeeded (a partial failure frees everything,
// so a later call starts from scratch instead of finding half-initialised buffers).  The caller holds the device guard.
static int ensure_staging_impl(z3d_plan *pl) {
    const size_t N3 = (size_t)pl->dim * pl->dim * pl->dim;
    // group size: one tile of the batched kernels (64 volumes) when that path exists and a sixteenth of the free device memory
    // holds the two stages; otherwise ~32 MiB of volume data per stage, at least 1 volume
    int gb = (int)std::min<size_t>(32, std::max<size_t>(1, ((size_t)32 << 20) / (N3 * sizeof(float))));
    if (pl->gen || pl->tables_level > 0) {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = 2 * ((size_t)64 * N3 * sizeof(float) + z3d_workspace_bytes(pl, 64));
        if (need <= free_b / 16) gb = 64;
    }
    pl->stage_batch = gb;
    CUDA_TRY(cudaStreamCreateWithFlags(&pl->s_copy, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&pl->s_comp, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&pl->ev_in[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&pl->ev_done[i], cudaEventDisableTiming));
        CUDA_TRY(cudaMalloc((void **)&pl->d_stage_vol[i], (size_t)gb * N3 * sizeof(float)));
        CUDA_TRY(cudaMalloc((void **)&pl->d_stage_mom[i], (size_t)gb * pl->nmom * 2 * sizeof(float)));
        CUDA_TRY(cudaMalloc(&pl->d_stage_ws[i], z3d_workspace_bytes(pl, gb)));
    }
    pl->staged_ok = true;
    return Z3D_OK;
}
static int ensure_staging(z3d_plan *pl) {
    if (pl->staged_ok) return Z3D_OK;
    const int rc = ensure_staging_impl(pl);
    if (rc) {
        const std::string keep = g_err;
        free_staging(pl);
        cudaGetLastError();
        g_err = keep;
    }
    return rc;
}

// an error in the middle of a host-buffer call must not leave copies to / from the caller's buffers in flight
static int host_call_fail(z3d_plan *pl, int rc) {
    const std::string keep = g_err;
    if (pl->s_comp) cudaStreamSynchronize(pl->s_comp);
    if (pl->s_copy) cudaStreamSynchronize(pl->s_copy);
    cudaGetLastError();
    g_err = keep;
    return rc;
}
#define HOST_TRY(expr)                                                                                                    \
    do {                                                                                                                  \
        cudaError_t _e = (expr);                                                                                          \
        if (_e != cudaSuccess)                                                                                            \
            return host_call_fail(pl, fail(Z3D_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__)); \
    } while (0)

// enqueue the staged groups of one host-buffer analysis call; the stage counter lives in the plan so that consecutive submissions keep
// alternating the two staging buffers
static int decompose_host_enqueue(z3d_plan *pl, const float *h_vol, int batch, float *h_mom) {
    int rc = ensure_staging(pl);
    if (rc) return rc;
    const size_t N3 = (size_t)pl->dim * pl->dim * pl->dim;
    const int gb = pl->stage_batch;
    for (int b0 = 0; b0 < batch; b0 += gb, pl->stage_next ^= 1) {
        const int stage = pl->stage_next;
        const int nb = std::min(gb, batch - b0);
        // the stage's previous compute + D2H must be finished before its input buffer is overwritten
        HOST_TRY(cudaStreamWaitEvent(pl->s_copy, pl->ev_done[stage], 0));
        HOST_TRY(cudaMemcpyAsync(pl->d_stage_vol[stage], h_vol + (size_t)b0 * N3, (size_t)nb * N3 * sizeof(float),
                                 cudaMemcpyHostToDevice, pl->s_copy));
        HOST_TRY(cudaEventRecord(pl->ev_in[stage], pl->s_copy));
        HOST_TRY(cudaStreamWaitEvent(pl->s_comp, pl->ev_in[stage], 0));
        rc = z3d_decompose(pl, pl->d_stage_vol[stage], nb, 0, pl->H, pl->d_stage_mom[stage], pl->d_stage_ws[stage],
                           z3d_workspace_bytes(pl, gb), pl->s_comp);
        if (rc) return host_call_fail(pl, rc);
        HOST_TRY(cudaMemcpyAsync(h_mom + (size_t)b0 * pl->nmom * 2, pl->d_stage_mom[stage],
                                 (size_t)nb * pl->nmom * 2 * sizeof(float), cudaMemcpyDeviceToHost, pl->s_comp));
        HOST_TRY(cudaEventRecord(pl->ev_done[stage], pl->s_comp));
    }
    return Z3D_OK;
}

extern "C" int z3d_host_wait(z3d_plan *pl) {
    if (!pl) return fail(Z3D_ERR_INVALID, "z3d_host_wait: plan is NULL");
    if (!pl->staged_ok) return Z3D_OK;
    USE_DEVICE(pl->device);
    HOST_TRY(cudaStreamSynchronize(pl->s_comp));
    HOST_TRY(cudaStreamSynchronize(pl->s_copy));
    return Z3D_OK;
}

extern "C" int z3d_decompose_host(z3d_plan *pl, const float *h_vol, int batch, float *h_mom) {
    if (!pl || !h_vol || !h_mom || batch < 1) return fail(Z3D_ERR_INVALID, "z3d_decompose_host: bad arguments");
    USE_DEVICE(pl->device);
    const int rc = decompose_host_enqueue(pl, h_vol, batch, h_mom);
    if (rc) return rc;
    return z3d_host_wait(pl);
}

extern "C" int z3d_decompose_host_submit(z3d_plan *pl, const float *h_vol, int batch, float *h_mom) {
    if (!pl || !h_vol || !h_mom || batch < 1) return fail(Z3D_ERR_INVALID, "z3d_decompose_host_submit: bad arguments");
    USE_DEVICE(pl->device);
    return decompose_host_enqueue(pl, h_vol, batch, h_mom);
}

extern "C" int z3d_reconstruct_host(z3d_plan *pl, const float *h_mom, int batch, float *h_vol) {
    if (!pl || !h_vol || !h_mom || batch < 1) return fail(Z3D_ERR_INVALID, "z3d_reconstruct_host: bad arguments");
    USE_DEVICE(pl->device);
    int rc = ensure_staging(pl);
    if (rc) return rc;
    const size_t N3 = (size_t)pl->dim * pl->dim * pl->dim;
    const int gb = pl->stage_batch;
    int stage = 0;
    for (int b0 = 0; b0 < batch; b0 += gb, stage ^= 1) {
        const int nb = std::min(gb, batch - b0);
        HOST_TRY(cudaStreamWaitEvent(pl->s_copy, pl->ev_done[stage], 0));
        HOST_TRY(cudaMemcpyAsync(pl->d_stage_mom[stage], h_mom + (size_t)b0 * pl->nmom * 2,
                                 (size_t)nb * pl->nmom * 2 * sizeof(float), cudaMemcpyHostToDevice, pl->s_copy));
        HOST_TRY(cudaEventRecord(pl->ev_in[stage], pl->s_copy));
        HOST_TRY(cudaStreamWaitEvent(pl->s_comp, pl->ev_in[stage], 0));
        if (pl->synth && nb >= pl->synth_min) {   // a group of maps worth a tile of the tensor-core synthesis kernel
            const size_t need = z3d_reconstruct_batched_workspace_bytes(pl, nb);
            if (pl->stage_synth_bytes < need) {
                HOST_TRY(cudaStreamSynchronize(pl->s_comp));
                if (pl->d_stage_synth_ws) cudaFree(pl->d_stage_synth_ws);
                pl->d_stage_synth_ws = nullptr;
                pl->stage_synth_bytes = 0;
                HOST_TRY(cudaMalloc(&pl->d_stage_synth_ws, need));
                pl->stage_synth_bytes = need;
            }
            rc = z3d_reconstruct_batched(pl, pl->d_stage_mom[stage], nb, pl->d_stage_vol[stage], pl->d_stage_synth_ws, pl->stage_synth_bytes, pl->s_comp);
        } else {
            rc = z3d_reconstruct(pl, pl->d_stage_mom[stage], nb, 0, pl->H, pl->d_stage_vol[stage], pl->d_stage_ws[stage],
                                 z3d_workspace_bytes(pl, gb), pl->s_comp);
        }
        if (rc) return host_call_fail(pl, rc);
        HOST_TRY(cudaMemcpyAsync(h_vol + (size_t)b0 * N3, pl->d_stage_vol[stage], (size_t)nb * N3 * sizeof(float),
                                 cudaMemcpyDeviceToHost, pl->s_comp));
        HOST_TRY(cudaEventRecord(pl->ev_done[stage], pl->s_comp));
    }
    HOST_TRY(cudaStreamSynchronize(pl->s_comp));
    HOST_TRY(cudaStreamSynchronize(pl->s_copy));
    return Z3D_OK;
}

extern "C" int z3d_fma_peak(int device, int variant, int iters, double *tflops) {
    if (!tflops || iters < 1) return fail(Z3D_ERR_INVALID, "z3d_fma_peak: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) {
        cudaGetLastError();
        return fail(Z3D_ERR_NODEVICE, "z3d_fma_peak: no such CUDA device");
    }
    USE_DEVICE(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 4, threads = 512;
    float *d_out;
    CUDA_TRY(cudaMalloc((void **)&d_out, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    k_fma_peak<<<blocks, threads>>>(iters, d_out, variant);  // warm-up
    CUDA_TRY(cudaEventRecord(e0));
    k_fma_peak<<<blocks, threads>>>(iters, d_out, variant);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    g_launches += 2;
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = (double)blocks * threads * (double)iters * 8.0 * 8.0 * 2.0;
    *tflops = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return Z3D_OK;
}

// Measured dense TF32 tensor-pipe peak (tcgen05.mma kind::tf32, M 128, N 256, K 8 back to back on every SM): the denominator of
// bench.py's roofline for the batched analysis kernel.
extern "C" int z3d_tf32_peak(int device, int iters, double *tflops) {
    if (!tflops || iters < 1) return fail(Z3D_ERR_INVALID, "z3d_tf32_peak: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) {
        cudaGetLastError();
        return fail(Z3D_ERR_NODEVICE, "z3d_tf32_peak: no such CUDA device");
    }
    USE_DEVICE(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    k_tf32_peak<<<prop.multiProcessorCount, 128>>>(iters);  // warm-up
    CUDA_TRY(cudaEventRecord(e0));
    k_tf32_peak<<<prop.multiProcessorCount, 128>>>(iters);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaGetLastError());
    g_launches += 2;
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *tflops = (double)prop.multiProcessorCount * (double)iters * 8.0 * (2.0 * 128 * 256 * 8) / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return Z3D_OK;
}

#if defined(Z3D_GEN_DEBUG) && !defined(Z3D_GEN_ABL)
// developer aid: mapped host memory the kernel's bounded waits report into (survives a trapped context)
static int *g_trap_host = nullptr;
extern "C" int z3d_debug_trap_arm(void) {
    if (!g_trap_host) {
        if (cudaHostAlloc((void **)&g_trap_host, 4096, cudaHostAllocMapped) != cudaSuccess) return -1;
        memset(g_trap_host, 0, 4096);
        int *d = nullptr;
        if (cudaHostGetDevicePointer((void **)&d, g_trap_host, 0) != cudaSuccess) return -2;
        if (cudaMemcpyToSymbol(z3d::g_gen_trap, &d, sizeof d) != cudaSuccess) return -3;
    }
    return 0;
}
extern "C" int z3d_debug_trap_read(int *out, int n) {
    if (!g_trap_host) return -1;
    for (int i = 0; i < n && i < 1024; ++i) out[i] = g_trap_host[i];
    return 0;
}
#endif
#ifdef Z3D_GEN_TIMING
extern "C" int z3d_debug_gen(long long *out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, z3d::g_gdbg, sizeof(long long) * 32);
    if (reset) {
        long long z[32] = {0};
        cudaMemcpyToSymbol(z3d::g_gdbg, z, sizeof z);
    }
    return 0;
}
#endif

#ifdef Z3D_GEN_CTACLK
extern "C" int z3d_debug_gen_cta(long long *out, int n) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, z3d::g_cta_clk, sizeof(long long) * std::min(n, 5 * 1024));
    return 0;
}
#endif

#ifdef Z3D_DEBUG_TIMING
extern "C" int z3d_debug_dump(long long *out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, z3d::g_dbg, sizeof(long long) * 8);
    if (reset) {
        long long z[8] = {0};
        cudaMemcpyToSymbol(z3d::g_dbg, z, sizeof z);
    }
    return 0;
}
#endif

// ------------------------------------------------------------------------------------------------
// peer-memory communicator for the fused reduce + all-reduce (one process per GPU, CUDA IPC)
// ------------------------------------------------------------------------------------------------
struct z3d_comm {
    int device = 0, rank = -1, world = 0;
    size_t max_elems = 0, bytes = 0;
    float *local = nullptr;
    PeerBufs peers{};
    bool opened[MAX_RANKS] = {};
    int *d_err = nullptr;
    unsigned epoch = 0;
};

extern "C" int z3d_comm_create(const z3d_plan *pl, int batch_max, z3d_comm **out, unsigned char *handle64) {
    if (!pl || !out || !handle64 || batch_max < 1) return fail(Z3D_ERR_INVALID, "z3d_comm_create: bad arguments");
    USE_DEVICE(pl->device);
    z3d_comm *c = new z3d_comm();
    c->device = pl->device;
    c->max_elems = (size_t)batch_max * 2 * pl->nmom;
    c->bytes = comm_bytes(c->max_elems);
    CUDA_TRY(cudaMalloc((void **)&c->local, c->bytes));
    CUDA_TRY(cudaMemset(c->local, 0, c->bytes));
    CUDA_TRY(cudaMalloc((void **)&c->d_err, sizeof(int)));
    CUDA_TRY(cudaMemset(c->d_err, 0, sizeof(int)));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->local));
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    CUDA_TRY(cudaDeviceSynchronize());
    *out = c;
    return Z3D_OK;
}

extern "C" int z3d_comm_connect(z3d_comm *c, int rank, int world, const unsigned char *handles /* [world][64] */) {
    if (!c || !handles || world < 1 || world > MAX_RANKS || rank < 0 || rank >= world)
        return fail(Z3D_ERR_INVALID, "z3d_comm_connect: bad arguments (world <= %d)", MAX_RANKS);
    USE_DEVICE(c->device);
    c->rank = rank;
    c->world = world;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            c->peers.data[r] = c->local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * r, 64);
        void *p = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peers.data[r] = (float *)p;
        c->opened[r] = true;
    }
    return Z3D_OK;
}

extern "C" int z3d_comm_destroy(z3d_comm *c) {
    if (!c) return Z3D_OK;
    DeviceGuard _guard(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < MAX_RANKS; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peers.data[r]);
    if (c->local) cudaFree(c->local);
    if (c->d_err) cudaFree(c->d_err);
    delete c;
    return Z3D_OK;
}

extern "C" int z3d_decompose_allreduce(const z3d_plan *pl, z3d_comm *c, const float *d_vol, int batch, float *d_moments,
                                       void *d_ws, size_t ws_bytes, void *stream) {
    if (!pl || !c || c->rank < 0) return fail(Z3D_ERR_INVALID, "z3d_decompose_allreduce: communicator not connected");
    if (!d_vol || !d_moments || !d_ws || batch < 1) return fail(Z3D_ERR_INVALID, "z3d_decompose_allreduce: bad arguments");
    // the per-CTA partials of the analysis kernel cover decompose_group(pl) volumes (workspace contract of z3d_workspace_bytes):
    // a larger batch runs group by group, one exchange (epoch) per group
    const int grp = decompose_group(pl);
    if ((size_t)std::min(batch, grp) * 2 * pl->nmom > c->max_elems)
        return fail(Z3D_ERR_INVALID, "z3d_decompose_allreduce: communicator created for fewer volumes per exchange (batch_max)");
    if (ws_bytes < (size_t)std::min(batch, grp) * pl->npass * pl->nsplit * SLOT_WORDS * sizeof(float))
        return fail(Z3D_ERR_WORKSPACE, "z3d_decompose_allreduce: workspace too small");
    USE_DEVICE(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    float *partial = (float *)d_ws;
    const size_t N3 = (size_t)pl->dim * pl->dim * pl->dim;
    const_cast<z3d_plan *>(pl)->last_was_decompose = false;
    for (int b0 = 0; b0 < batch; b0 += grp) {
        const int nb = std::min(grp, batch - b0);
        dim3 g((2 * pl->nmom + 255) / 256, nb);
        // a block of the exchange kernel waits for its counterpart blocks on the peers: keep the whole grid co-resident so that
        // progress never depends on the order in which the ranks' block schedulers hand out blocks; one flag per block
        if ((int)(g.x * g.y) > std::min(pl->num_sms * 4, P2P_MAX_BLOCKS))
            return fail(Z3D_ERR_UNSUPPORTED, "z3d_decompose_allreduce: order %d too large for the one-shot exchange (use "
                                             "z3d_decompose_part + a library all-reduce)", pl->order);
        const float *vol = d_vol + (size_t)b0 * N3;
        if (pl->profiling && b0 == 0) CUDA_TRY(cudaEventRecord(pl->ev_k0, st));
        if (pl->npass == 2)
            k_decompose<1><<<pl->npass * pl->nsplit, NT, pl->smem_bytes, st>>>(pl->dev, vol, nb, 0, pl->H, c->rank, c->world, partial);
        else
            k_decompose<0><<<pl->npass * pl->nsplit, NT, pl->smem_bytes, st>>>(pl->dev, vol, nb, 0, pl->H, c->rank, c->world, partial);
        CUDA_TRY(cudaGetLastError());
        if (pl->profiling && b0 == 0) CUDA_TRY(cudaEventRecord(pl->ev_k1, st));
        c->epoch += 1;
        k_reduce_allreduce_p2p<<<g, 256, 0, st>>>(partial, pl->d_src, pl->d_scale, pl->nmom, pl->npass, pl->nsplit, nb, c->peers,
                                                  c->rank, c->world, c->max_elems, c->epoch, c->d_err,
                                                  d_moments + (size_t)b0 * 2 * pl->nmom);
        CUDA_TRY(cudaGetLastError());
        g_launches += 2;
    }
    return Z3D_OK;
}

extern "C" int z3d_comm_error(z3d_comm *c) {  // 1 after an exchange timed out (results were set to NaN)
    if (!c) return -1;
    int e = 0;
    DeviceGuard _guard(c->device);
    cudaMemcpy(&e, c->d_err, sizeof(int), cudaMemcpyDeviceToHost);
    return e;
}
