/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's mode-'3' spharm decomposition /
 * reconstruction path (charbj/zernike3d, zernike_master.py).  It is the checker for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it.  Nothing under zernike3d_b200/ links, imports or calls it.
 *
 * Parity status: the reference holds no tests, fixtures or golden vectors, so this
 * restatement is pinned against outputs of the UNMODIFIED reference run in the authoring container
 * (oracle/ref_shim.py + oracle/make_golden.py -> tests/golden/ref_*.npz; checked by
 * tests/test_oracle_vs_reference.py).
 *
 * Two arithmetic modes:
 *   ZO_FAITHFUL (0): the reference's dtype flow, statement by statement - float64 evaluation of each
 *       recurrence step from float32-rounded predecessors, float32 / complex64 stores
 *       (zernike_master.py:2290, 2298, 1985-2002, 2208).  The voxel sum of the analysis step is kept in
 *       float64 (the reference sums complex64 inside BLAS in an unspecified order, zm:2210); the
 *       synthesis step accumulates in float32 in the reference's own (n, l, m) order (zm:2611-2613,
 *       2701), which is fully determined because that step is elementwise.
 *   ZO_EXACT (1): the same formulas in float64 throughout (accuracy yardstick, and the oracle for
 *       box sizes whose materialised basis the reference cannot hold, e.g. 256^3 = 223 GB).
 *
 * Everything is evaluated voxel by voxel (the reference's array statements are elementwise, so a
 * streaming evaluation performs the identical scalar operations without the N_basis * N^3 memory).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---- tiny static-partition thread pool (libgomp is not in the image) ---- */
static int g_threads = 0;
int zo_max_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
void zo_set_threads(int n) { g_threads = n > 0 ? n : 0; }
typedef void (*zo_range_fn)(void *ctx, int tid, long lo, long hi);
typedef struct {
    zo_range_fn fn;
    void *ctx;
    int tid;
    long lo, hi;
} zo_task;
static void *zo_trampoline(void *a) {
    zo_task *t = (zo_task *)a;
    t->fn(t->ctx, t->tid, t->lo, t->hi);
    return NULL;
}
/* contiguous static split of [0, total) over nthreads workers */
static void zo_parallel(int nthreads, long total, zo_range_fn fn, void *ctx) {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    zo_task *tk = (zo_task *)malloc(sizeof(zo_task) * nthreads);
    for (int t = 0; t < nthreads; ++t) {
        tk[t].fn = fn;
        tk[t].ctx = ctx;
        tk[t].tid = t;
        tk[t].lo = total * t / nthreads;
        tk[t].hi = total * (t + 1) / nthreads;
        if (t > 0) pthread_create(&th[t], NULL, zo_trampoline, &tk[t]);
    }
    zo_trampoline(&tk[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(tk);
}

#define ZO_FAITHFUL 0
#define ZO_EXACT 1

/* ---- index helpers: order of sorted (n, l, m), m >= 0  (indexing.py:191-228, zm:378-382) ---- */
int zo_num_moments(int order) {
    int c = 0;
    for (int n = 0; n <= order; ++n)
        for (int l = n & 1; l <= n; l += 2) c += l + 1;
    return c;
}
int zo_num_radial(int order) {
    int c = 0;
    for (int n = 0; n <= order; ++n) c += n / 2 + 1;
    return c;
}
static inline int lm_index(int l, int m) { return l * (l + 1) / 2 + m; }

/* side = np.linspace(-1/sqrt(3), 1/sqrt(3), dim)   (zm:2240) */
static void make_side(int dim, double *side) {
    const double stop = 1.0 / sqrt(3.0), start = -stop;
    const double step = (stop - start) / (double)(dim - 1);
    for (int k = 0; k < dim; ++k) side[k] = (double)k * step + start;
    if (dim > 1) side[dim - 1] = stop;
}

typedef struct {
    int order, nlm, mode;
    double *c1, *c2, *c3; /* per (l, m): zm:2013, 2017-2018, 2022-2024 */
} coef_t;

static void coef_init(coef_t *c, int order, int mode) {
    c->order = order;
    c->mode = mode;
    c->nlm = (order + 1) * (order + 2) / 2;
    c->c1 = (double *)calloc(c->nlm, sizeof(double));
    c->c2 = (double *)calloc(c->nlm, sizeof(double));
    c->c3 = (double *)calloc(c->nlm, sizeof(double));
    for (int l = 1; l <= order; ++l)
        for (int m = 0; m <= l; ++m) {
            int i = lm_index(l, m);
            if (l == m) {
                c->c3[i] = sqrt((2.0 * l + 1.0) / (2.0 * l));
            } else {
                double c0 = sqrt((2.0 * l + 1.0) / (double)((l + m) * (l - m)));
                c->c1[i] = c0 * sqrt(2.0 * l - 1.0);
                if (m < l - 1) c->c2[i] = c0 * sqrt((double)((l + m - 1) * (l - m - 1)) / (2.0 * l - 3.0));
            }
        }
}
static void coef_free(coef_t *c) {
    free(c->c1);
    free(c->c2);
    free(c->c3);
}

/* Normalised associated Legendre P_lm(theta) incl. Condon-Shortley phase; Y_lm_recurrence zm:2006-2026,
 * iterated in sorted (l, m) order with the float32 store of zm:2290.  p[] has (order+1)(order+2)/2
 * entries.  In faithful mode each entry holds the float32-rounded value. */
static void legendre_all(const coef_t *c, double theta, double *p) {
    const int L = c->order;
    const double st = sin(theta), ct = cos(theta);
    const int faithful = (c->mode == ZO_FAITHFUL);
    double v = 1.0 * sqrt(1.0 / M_PI) * 0.5; /* zm:2010-2011 (ones_like(theta): never NaN) */
    p[0] = faithful ? (double)(float)v : v;
    for (int l = 1; l <= L; ++l)
        for (int m = 0; m <= l; ++m) {
            int i = lm_index(l, m);
            if (l == m)
                v = -c->c3[i] * st * p[lm_index(l - 1, m - 1)];
            else if (m == l - 1)
                v = c->c1[i] * ct * p[lm_index(l - 1, m)];
            else
                v = c->c1[i] * ct * p[lm_index(l - 1, m)] - c->c2[i] * p[lm_index(l - 2, m)];
            p[i] = faithful ? (double)(float)v : v;
        }
}

/* R_nl(r); R_nl_recurrence zm:1980-2002 iterated in sorted (n, l) order with the float32 store.
 * Storage: rad[n * (L+1) + l] (only l <= n, n-l even are written). */
static void radial_all(int L, int mode, double r, double *rad) {
    const int faithful = (mode == ZO_FAITHFUL);
    const int S = L + 1;
    for (int n = 0; n <= L; ++n)
        for (int l = n & 1; l <= n; l += 2) {
            double v;
            if (l == n) {
                v = pow(r, (double)n);
            } else if (l == n - 2) {
                v = (n + 0.5) * pow(r, (double)n) - (n - 0.5) * pow(r, (double)(n - 2));
            } else {
                double k0 = (double)(n - l) * (double)(n + l + 1) * (double)(2 * n - 3);
                double k1 = (double)(2 * n - 1) * (double)(2 * n + 1) * (double)(2 * n - 3);
                double k2 = 0.5 * (double)(-2 * n + 1) * (double)((2 * l + 1) * (2 * l + 1)) - (k1 / 2.0);
                double k3 = -1.0 * (double)(n - l - 2) * (double)(n + l - 1) * (double)(2 * n + 1);
                double K1 = k1 / k0, K2 = k2 / k0, K3 = k3 / k0;
                v = (K1 * (r * r) + K2) * rad[(n - 2) * S + l] + K3 * rad[(n - 4) * S + l];
            }
            if (faithful) {
                float f = (float)v;
                if (isnan(f)) f = 0.0f; /* nan_to_num zm:2318 */
                v = (double)f;
            }
            rad[n * S + l] = v;
        }
}

typedef struct {
    double r, theta, phi;
} polar_t;

/* cartesian_2_polar zm:1886-1890 with the 'xy' meshgrid of zm:2241: [i,j,k] -> (y_i, x_j, z_k) */
static polar_t polar_of(const double *side, int i, int j, int k) {
    double x = side[j], y = side[i], z = side[k];
    polar_t p;
    p.r = sqrt(x * x + y * y + z * z);
    p.theta = acos(z / p.r);
    p.phi = atan2(y, x);
    return p;
}

/* Y_lm = nan_to_num(P_lm * exp(1j m phi)).astype(complex64)  zm:2297-2298 */
static inline void y_of(int mode, double p, double cm, double sm, double *yr, double *yi) {
    double a = p * cm, b = p * sm;
    if (mode == ZO_FAITHFUL) {
        float fa = (float)a, fb = (float)b;
        if (isnan(fa)) fa = 0.0f;
        if (isnan(fb)) fb = 0.0f;
        a = fa;
        b = fb;
    } else {
        if (isnan(a)) a = 0.0;
        if (isnan(b)) b = 0.0;
    }
    *yr = a;
    *yi = b;
}

/* basis at one voxel, for unit tests: rad_out[num_radial] in sorted (n,l) order,
 * y_out[2 * (L+1)(L+2)/2] (re, im) in (l, m) order. */
int zo_basis_at(int dim, int order, int mode, int i, int j, int k, double *rad_out, double *y_out) {
    double *side = (double *)malloc(sizeof(double) * dim);
    make_side(dim, side);
    coef_t c;
    coef_init(&c, order, mode);
    const int S = order + 1;
    double *p = (double *)malloc(sizeof(double) * c.nlm);
    double *rad = (double *)calloc((size_t)S * S, sizeof(double));
    polar_t q = polar_of(side, i, j, k);
    legendre_all(&c, q.theta, p);
    radial_all(order, mode, q.r, rad);
    int t = 0;
    for (int n = 0; n <= order; ++n)
        for (int l = n & 1; l <= n; l += 2) rad_out[t++] = rad[n * S + l];
    for (int l = 0; l <= order; ++l)
        for (int m = 0; m <= l; ++m) {
            double yr, yi;
            y_of(mode, p[lm_index(l, m)], cos((double)m * q.phi), sin((double)m * q.phi), &yr, &yi);
            y_out[2 * lm_index(l, m)] = yr;
            y_out[2 * lm_index(l, m) + 1] = yi;
        }
    free(side);
    free(p);
    free(rad);
    coef_free(&c);
    return 0;
}

/* Analysis: chi_nlm = coeff(n) * sum_v vol[v] * R_nl[v] * conj(Y_lm[v]),  m >= 0 only
 * (_calculate_Z_GPU zm:2194-2217; coeff zm:2198).  Rows i in [i_lo, i_hi) of array axis 0 are summed
 * (i_lo = 0, i_hi = dim for the whole volume; sub-ranges give the unreduced slab partials used by the
 * multi-GPU tests).  out32: complex64 as (re, im) float pairs in sorted (n, l, m) order; out64
 * (optional) the same before the final float32 rounding. */
typedef struct {
    const float *vol;
    int dim, order, mode, i_lo, nmom;
    const double *side;
    const coef_t *c;
    double *accs;
} dec_ctx;

static void dec_worker(void *vctx, int tid, long lo, long hi) {
    dec_ctx *x = (dec_ctx *)vctx;
    const int dim = x->dim, order = x->order, mode = x->mode, S = order + 1;
    const coef_t *c = x->c;
    double *acc = x->accs + (size_t)tid * 2 * x->nmom;
    double *p = (double *)malloc(sizeof(double) * c->nlm);
    double *rad = (double *)calloc((size_t)S * S, sizeof(double));
    double *cm = (double *)malloc(sizeof(double) * S);
    double *sm = (double *)malloc(sizeof(double) * S);
    double *yr = (double *)malloc(sizeof(double) * c->nlm);
    double *yi = (double *)malloc(sizeof(double) * c->nlm);
    for (long row = lo; row < hi; ++row) {
        const int i = x->i_lo + (int)(row / dim), j = (int)(row % dim);
        for (int k = 0; k < dim; ++k) {
            const double f = (double)x->vol[((size_t)i * dim + j) * dim + k];
            polar_t q = polar_of(x->side, i, j, k);
            legendre_all(c, q.theta, p);
            radial_all(order, mode, q.r, rad);
            for (int m = 0; m <= order; ++m) {
                cm[m] = cos((double)m * q.phi);
                sm[m] = sin((double)m * q.phi);
            }
            for (int l = 0; l <= order; ++l)
                for (int m = 0; m <= l; ++m) {
                    int t = lm_index(l, m);
                    y_of(mode, p[t], cm[m], sm[m], &yr[t], &yi[t]);
                }
            int mu = 0;
            for (int n = 0; n <= order; ++n)
                for (int l = n & 1; l <= n; l += 2) {
                    const double R = rad[n * S + l];
                    const int base = lm_index(l, 0);
                    if (mode == ZO_FAITHFUL) {
                        /* Z = R * conj(Y) in complex64 (zm:2208), volume * Z in complex64 */
                        const float Rf = (float)R, ff = (float)f;
                        for (int m = 0; m <= l; ++m, ++mu) {
                            float zr = Rf * (float)yr[base + m];
                            float zi = -(Rf * (float)yi[base + m]);
                            acc[2 * mu] += (double)(ff * zr);
                            acc[2 * mu + 1] += (double)(ff * zi);
                        }
                    } else {
                        const double fR = f * R;
                        for (int m = 0; m <= l; ++m, ++mu) {
                            acc[2 * mu] += fR * yr[base + m];
                            acc[2 * mu + 1] -= fR * yi[base + m];
                        }
                    }
                }
        }
    }
    free(p);
    free(rad);
    free(cm);
    free(sm);
    free(yr);
    free(yi);
}

int zo_decompose(const float *vol, int dim, int order, int mode, int i_lo, int i_hi, float *out32,
                 double *out64) {
    if (dim < 2 || order < 0 || i_lo < 0 || i_hi > dim || i_lo > i_hi) return 1;
    const int nmom = zo_num_moments(order);
    double *side = (double *)malloc(sizeof(double) * dim);
    make_side(dim, side);
    coef_t c;
    coef_init(&c, order, mode);
    const int nthreads = zo_max_threads();
    double *accs = (double *)calloc((size_t)nthreads * 2 * nmom, sizeof(double));
    dec_ctx ctx = {vol, dim, order, mode, i_lo, nmom, side, &c, accs};
    zo_parallel(nthreads, (long)(i_hi - i_lo) * dim, dec_worker, &ctx);
    int mu = 0;
    for (int n = 0; n <= order; ++n) {
        const double coeff = (2.0 * (2.0 * n + 3.0)) / (3.0 * sqrt(3.0) * M_PI * ((double)dim * dim * dim));
        for (int l = n & 1; l <= n; l += 2)
            for (int m = 0; m <= l; ++m, ++mu) {
                double re = 0.0, im = 0.0;
                for (int t = 0; t < nthreads; ++t) {
                    re += accs[(size_t)t * 2 * nmom + 2 * mu];
                    im += accs[(size_t)t * 2 * nmom + 2 * mu + 1];
                }
                re *= coeff;
                im *= coeff;
                if (out64) {
                    out64[2 * mu] = re;
                    out64[2 * mu + 1] = im;
                }
                if (out32) {
                    out32[2 * mu] = (float)re;
                    out32[2 * mu + 1] = (float)im;
                }
            }
    }
    free(accs);
    free(side);
    coef_free(&c);
    return 0;
}

/* Synthesis: out = Re sum_{n,l,m>=0} [ mom R Y + (m != 0) conj(mom) R conj(Y) ]
 * (_calculate_cum_GPU zm:2596-2614, Reconstruct zm:2686-2715).  The spurious "+1j" initial values
 * (zm:2597, 2687) only touch the imaginary part, which cp.real() discards (zm:2715).
 * Rows i in [i_lo, i_hi) of axis 0 are written to out[(i - i_lo), :, :]. */
typedef struct {
    const float *mom32;
    int dim, order, mode, i_lo;
    const double *side;
    const coef_t *c;
    float *out;
} rec_ctx;

static void rec_worker(void *vctx, int tid, long lo, long hi) {
    (void)tid;
    rec_ctx *x = (rec_ctx *)vctx;
    const int dim = x->dim, order = x->order, mode = x->mode, S = order + 1;
    const coef_t *c = x->c;
    const float *mom32 = x->mom32;
    double *p = (double *)malloc(sizeof(double) * c->nlm);
    double *rad = (double *)calloc((size_t)S * S, sizeof(double));
    double *cm = (double *)malloc(sizeof(double) * S);
    double *sm = (double *)malloc(sizeof(double) * S);
    double *yr = (double *)malloc(sizeof(double) * c->nlm);
    double *yi = (double *)malloc(sizeof(double) * c->nlm);
    for (long row = lo; row < hi; ++row) {
        const int i = x->i_lo + (int)(row / dim), j = (int)(row % dim);
        for (int k = 0; k < dim; ++k) {
            polar_t q = polar_of(x->side, i, j, k);
            legendre_all(c, q.theta, p);
            radial_all(order, mode, q.r, rad);
            for (int m = 0; m <= order; ++m) {
                cm[m] = cos((double)m * q.phi);
                sm[m] = sin((double)m * q.phi);
            }
            for (int l = 0; l <= order; ++l)
                for (int m = 0; m <= l; ++m) {
                    int t = lm_index(l, m);
                    y_of(mode, p[t], cm[m], sm[m], &yr[t], &yi[t]);
                }
            int mu = 0;
            double result;
            if (mode == ZO_FAITHFUL) {
                volatile float cumulative = 0.0f;
                for (int n = 0; n <= order; ++n) {
                    volatile float grid = 0.0f;
                    for (int l = n & 1; l <= n; l += 2) {
                        const float Rf = (float)rad[n * S + l];
                        const int base = lm_index(l, 0);
                        for (int m = 0; m <= l; ++m, ++mu) {
                            /* cp.multiply(R, Y) in complex64, then mom * (.) in complex64 */
                            volatile float a = Rf * (float)yr[base + m];
                            volatile float b = Rf * (float)yi[base + m];
                            const float mr = mom32[2 * mu], mi = mom32[2 * mu + 1];
                            volatile float t1 = mr * a, t2 = mi * b;
                            float re = t1 - t2;
                            if (isnan(re)) re = 0.0f;
                            grid = grid + re;
                            if (m != 0) grid = grid + re; /* conj(mom) R conj(Y): same real part */
                        }
                    }
                    cumulative = cumulative + grid;
                }
                result = (double)cumulative;
            } else {
                double cumulative = 0.0;
                for (int n = 0; n <= order; ++n)
                    for (int l = n & 1; l <= n; l += 2) {
                        const double R = rad[n * S + l];
                        const int base = lm_index(l, 0);
                        for (int m = 0; m <= l; ++m, ++mu) {
                            double re = (double)mom32[2 * mu] * (R * yr[base + m]) -
                                        (double)mom32[2 * mu + 1] * (R * yi[base + m]);
                            if (isnan(re)) re = 0.0;
                            cumulative += (m != 0) ? 2.0 * re : re;
                        }
                    }
                result = cumulative;
            }
            x->out[((size_t)(i - x->i_lo) * dim + j) * dim + k] = (float)result;
        }
    }
    free(p);
    free(rad);
    free(cm);
    free(sm);
    free(yr);
    free(yi);
}

int zo_reconstruct(const float *mom32, int dim, int order, int mode, int i_lo, int i_hi, float *out) {
    if (dim < 2 || order < 0 || i_lo < 0 || i_hi > dim || i_lo > i_hi) return 1;
    double *side = (double *)malloc(sizeof(double) * dim);
    make_side(dim, side);
    coef_t c;
    coef_init(&c, order, mode);
    rec_ctx ctx = {mom32, dim, order, mode, i_lo, side, &c, out};
    zo_parallel(zo_max_threads(), (long)(i_hi - i_lo) * dim, rec_worker, &ctx);
    free(side);
    coef_free(&c);
    return 0;
}
