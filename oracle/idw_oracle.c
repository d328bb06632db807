/*
 * oracle/idw_oracle.c -- CPU restatement of Dewberry/v2r's IDW gridding path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (v2r_b200/, include/v2r_idw.h) never links, imports or calls anything in here.
 *
 * Parity status: PINNED to +-0.5 by the reference's own golden rasters
 *   features/idw/idw_test_files/idw_correct_step1-1.asc  (p = 1.7, 16x13, serial + 3x2 chunks)
 *   features/idw/idw_test_files/idw_correct_step2-2.asc  (orphan vector: step-2 grid, step-1 keys)
 * (tests/test_oracle_golden.py).  Beyond the Int16 rounding of those files the reference's
 * tests pin nothing; the 1e-12 bar is pinned by this restatement cross-checked against an
 * 80-bit long-double evaluation (oracle_full_solve_ld) and glibc pow().
 * The reference itself (Go 1.18 + GDAL) cannot be built in this image: no go, no libgdal.
 *
 * Third-party arithmetic restated here: the Go standard library `math` package
 * (go.mod:3 "go 1.18", Dockerfile:3 golang:1.18.1) -- Pow, Modf, Frexp, Ldexp, Log, Exp,
 * Sqrt.  go_pow() follows the published algorithm of src/math/pow.go; go_log()/go_exp()
 * follow the portable src/math/log.go and src/math/exp.go (FreeBSD e_log.c / e_exp.c
 * derivations).  On amd64 Go 1.18 substitutes an assembly Exp that differs from the
 * portable one by <= 1 ulp; that is four orders of magnitude inside the 1e-12 bar.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -lpthread  (Go/amd64 never fuses a*b+c).
 *
 * Every function cites the reference file:line (relative to the v2r repository) it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------
 * Go `math` restatement
 * ---------------------------------------------------------------------------------------- */

/* Go's float64 -> int conversion on amd64 (CVTTSD2SQ): truncation toward zero; NaN and
 * out-of-range values give the "integer indefinite" 0x8000000000000000.
 * Used by tools/utils.go:70-72 (PointToPair) and :133-137 (GetDimensions). */
int64_t go_int(double v)
{
    if (!(v > -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)v; /* C truncates toward zero as well */
}

/* src/math/log.go (portable): log(x) = k*ln2 + log(1+f), f reduced to [sqrt(2)/2-1, sqrt(2)-1],
 * s = f/(2+f), R(z) 7-term minimax in z = s*s. */
double go_log(double x)
{
    static const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10,
                        L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01,
                        L3 = 2.857142874366239149e-01, L4 = 2.222219843214978396e-01,
                        L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
                        L7 = 1.479819860511658591e-01;
    if (isnan(x) || (isinf(x) && x > 0)) return x;
    if (x < 0) return NAN;
    if (x == 0) return -INFINITY;
    int ki;
    double f1 = frexp(x, &ki);
    if (f1 < 0.70710678118654752440 /* Sqrt2/2 */) { f1 *= 2; ki--; }
    double f = f1 - 1;
    double k = (double)ki;
    double s = f / (2 + f);
    double s2 = s * s;
    double s4 = s2 * s2;
    double t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)));
    double t2 = s4 * (L2 + s4 * (L4 + s4 * L6));
    double R = t1 + t2;
    double hfsq = 0.5 * f * f;
    return k * Ln2Hi - ((hfsq - (s * (hfsq + R) + k * Ln2Lo)) - f);
}

/* src/math/exp.go (portable): x = k*ln2 + r, exp(r) by the 5-term Remez rational, scaled by 2^k. */
double go_exp(double x)
{
    static const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10,
                        Log2e = 1.44269504088896338700e+00, Overflow = 7.09782712893383973096e+02,
                        Underflow = -7.45133219101941108420e+02, NearZero = 1.0 / (1 << 28),
                        P1 = 1.66666666666666657415e-01, P2 = -2.77777777770155933842e-03,
                        P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                        P5 = 4.13813679705723846039e-08;
    if (isnan(x) || (isinf(x) && x > 0)) return x;
    if (isinf(x) && x < 0) return 0;
    if (x > Overflow) return INFINITY;
    if (x < Underflow) return 0;
    if (-NearZero < x && x < NearZero) return 1 + x;
    int k = 0;
    if (x < 0) k = (int)(Log2e * x - 0.5);
    else if (x > 0) k = (int)(Log2e * x + 0.5);
    double hi = x - (double)k * Ln2Hi;
    double lo = (double)k * Ln2Lo;
    /* expmulti(hi, lo, k) */
    double r = hi - lo;
    double t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1 - ((lo - (r * c) / (2 - c)) - hi);
    return ldexp(y, k);
}

static int go_is_odd_int(double x)
{
    if (fabs(x) >= 9007199254740992.0 /* 1<<53 */) return 0;
    double xi, xf = modf(x, &xi);
    return xf == 0 && (((int64_t)xi) & 1) == 1;
}

/* src/math/pow.go (Go 1.18): special cases, then x**y = x**yf * x**yi with yf in (-1/2, 1/2]
 * evaluated as Exp(yf*Log(x)) and the integer part by repeated squaring on Frexp mantissas;
 * negative y inverts the mantissa product before the final Ldexp. */
double go_pow(double x, double y)
{
    if (y == 0 || x == 1) return 1;
    if (y == 1) return x;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0) {
        if (y < 0) {
            if (go_is_odd_int(y)) return copysign(INFINITY, x);
            return INFINITY;
        }
        if (y > 0) {
            if (go_is_odd_int(y)) return x;
            return 0;
        }
    }
    if (isinf(y)) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }
    if (isinf(x)) {
        if (x < 0) return go_pow(1 / x, -y);
        if (y < 0) return 0;
        if (y > 0) return INFINITY;
    }
    if (y == 0.5) return sqrt(x);
    if (y == -0.5) return 1 / sqrt(x);

    double yi, yf = modf(fabs(y), &yi);
    if (yf != 0 && x < 0) return NAN;
    if (yi >= 9.223372036854775808e18 /* 1<<63 */) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }

    double a1 = 1.0;
    int ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf--; yi++; }
        a1 = go_exp(yf * go_log(x));
    }
    int xe;
    double x1 = frexp(x, &xe);
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (xe < -(1 << 12) || (1 << 12) < xe) {
            ae += xe;
            break;
        }
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    if (y < 0) { a1 = 1 / a1; ae = -ae; }
    return ldexp(a1, ae);
}

/* ------------------------------------------------------------------------------------------
 * tools/utils.go restatement
 * ---------------------------------------------------------------------------------------- */

/* tools/utils.go:133-137 GetDimensions: cols = int((1 + xMax - xMin)/xStep), rows likewise. */
void oracle_get_dimensions(double x_min, double x_max, double x_step, double y_min, double y_max,
                           double y_step, int64_t *rows, int64_t *cols)
{
    *cols = go_int((1 + x_max - x_min) / x_step);
    *rows = go_int((1 + y_max - y_min) / y_step);
}

/* tools/utils.go:70-72 PointToPair: R from Y, C from X, both truncating toward zero. */
void oracle_point_to_pair(double x, double y, double x_min, double x_step, double y_min,
                          double y_step, int64_t *r, int64_t *c)
{
    *r = go_int((y - y_min) / y_step);
    *c = go_int((x - x_min) / x_step);
}

/* tools/utils.go:156-159 euclidDist: Pow(Pow(dx,2) + Pow(dy,2), .5). */
double oracle_euclid_dist(double x1, double y1, double x2, double y2)
{
    double total = go_pow(x1 - x2, 2) + go_pow(y1 - y2, 2);
    return go_pow(total, .5);
}

/* tools/utils.go:166-168 DistExp(p0, p, exp) = Pow(euclidDist(p, p0), -exp); p is the data
 * point (first operand of the subtraction), p0 the grid node. */
double oracle_dist_exp(double cell_x, double cell_y, double px, double py, double exp)
{
    return go_pow(oracle_euclid_dist(px, py, cell_x, cell_y), -exp);
}

/* --- (r,c) -> slot hash table standing in for Go's map[OrderedPair]Point ------------------ */
typedef struct {
    int64_t *kr, *kc;
    int64_t *slot; /* index into the point arrays, -1 = empty */
    uint64_t mask;
} keymap;

static uint64_t key_hash(int64_t r, int64_t c)
{
    uint64_t h = (uint64_t)r * 0x9E3779B97F4A7C15ull ^ ((uint64_t)c + 0x7F4A7C15u) * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    return h;
}

static int keymap_init(keymap *m, int64_t n)
{
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 2 + 1) cap <<= 1;
    m->mask = cap - 1;
    m->kr = (int64_t *)malloc(cap * sizeof(int64_t));
    m->kc = (int64_t *)malloc(cap * sizeof(int64_t));
    m->slot = (int64_t *)malloc(cap * sizeof(int64_t));
    if (!m->kr || !m->kc || !m->slot) return -1;
    for (uint64_t i = 0; i < cap; i++) m->slot[i] = -1;
    return 0;
}

static void keymap_free(keymap *m)
{
    free(m->kr);
    free(m->kc);
    free(m->slot);
}

/* returns pointer to the slot cell for (r,c): existing entry or the empty cell to fill */
static int64_t *keymap_find(keymap *m, int64_t r, int64_t c)
{
    uint64_t i = key_hash(r, c) & m->mask;
    while (m->slot[i] >= 0 && !(m->kr[i] == r && m->kc[i] == c)) i = (i + 1) & m->mask;
    if (m->slot[i] < 0) { m->kr[i] = r; m->kc[i] = c; }
    return &m->slot[i];
}

/* tools/utils.go:97-119 MakeCoordSpace.  The reference receives points from one goroutine per
 * point, i.e. in scheduler order; the oracle fixes that order to INPUT order, which is one valid
 * execution.  A colliding arrival replaces X,Y with its own and Weight with (new + old)/2
 * (pairwise running average, :111-113).  Output order = first-insertion order of each key (a Go
 * map has no order; any order is valid).  Returns the number of distinct keys. */
int64_t oracle_make_coord_space(int64_t n, const double *x, const double *y, const double *w,
                                double x_min, double x_step, double y_min, double y_step,
                                double *ox, double *oy, double *ow, int64_t *okr, int64_t *okc)
{
    keymap m;
    if (keymap_init(&m, n) != 0) return -1;
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t r, c;
        oracle_point_to_pair(x[i], y[i], x_min, x_step, y_min, y_step, &r, &c);
        int64_t *s = keymap_find(&m, r, c);
        double wi = w[i];
        if (*s >= 0) {
            wi = (wi + ow[*s]) / 2;
        } else {
            *s = cnt++;
            okr[*s] = r;
            okc[*s] = c;
        }
        ox[*s] = x[i];
        oy[*s] = y[i];
        ow[*s] = wi;
    }
    keymap_free(&m);
    return cnt;
}

/* ------------------------------------------------------------------------------------------
 * features/idw restatement
 * ---------------------------------------------------------------------------------------- */

/* features/idw/idw_utils.go:8-27 calculateIDW for one cell.  RCToPoint (tools/utils.go:78-82):
 * px = xMin + float64(c)*xStep (separate multiply and add).  Exact-hit test is the key lookup
 * (:13-16), not a distance test.  The reference iterates the Go map in random order; the oracle
 * sums in array order. */
/* mode 0: array order (the oracle's fixed choice).  mode 1: REVERSE array order -- another valid execution of the
 * reference, whose `range` over the Go map visits the points in a random order every time.  mode 2: the reference's
 * own float64 terms (same DistExp, same w*h product) accumulated in 80-bit long double: the order-independent value
 * that every execution of the reference scatters around by its float64 summation error. */
static double cell_idw_mode(const keymap *m_const, int64_t n, const double *x, const double *y,
                            const double *w, double x_min, double x_step, double y_min, double y_step,
                            double exp, int64_t r, int64_t c, int mode);

static double cell_idw(const keymap *m_const, int64_t n, const double *x, const double *y,
                       const double *w, double x_min, double x_step, double y_min, double y_step,
                       double exp, int64_t r, int64_t c)
{
    return cell_idw_mode(m_const, n, x, y, w, x_min, x_step, y_min, y_step, exp, r, c, 0);
}

static double cell_idw_mode(const keymap *m_const, int64_t n, const double *x, const double *y,
                            const double *w, double x_min, double x_step, double y_min, double y_step,
                            double exp, int64_t r, int64_t c, int mode)
{
    keymap *m = (keymap *)m_const;
    double px = x_min + (double)c * x_step;
    double py = y_min + (double)r * y_step;
    /* read-only probe (never inserts: table has spare capacity and we do not write slot) */
    uint64_t i = key_hash(r, c) & m->mask;
    while (m->slot[i] >= 0) {
        if (m->kr[i] == r && m->kc[i] == c) return w[m->slot[i]];
        i = (i + 1) & m->mask;
    }
    if (mode == 2) {
        long double num = 0.0L, den = 0.0L;
        for (int64_t k = 0; k < n; k++) {
            double h = oracle_dist_exp(px, py, x[k], y[k], exp);
            num += (long double)(w[k] * h);
            den += (long double)h;
        }
        return (double)(num / den);
    }
    double num = 0.0, den = 0.0;
    for (int64_t q = 0; q < n; q++) {
        int64_t k = mode == 1 ? n - 1 - q : q;
        double h = oracle_dist_exp(px, py, x[k], y[k], exp);
        num += w[k] * h;
        den += h;
    }
    return num / den;
}

static int build_keymap(keymap *m, int64_t n, const int64_t *kr, const int64_t *kc)
{
    if (keymap_init(m, n) != 0) return -1;
    for (int64_t i = 0; i < n; i++) {
        int64_t *s = keymap_find(m, kr[i], kc[i]);
        *s = i; /* duplicate keys: last one wins, like repeated Go map assignment */
    }
    return 0;
}

/* features/idw/idw.go:34-40 FullSolve hot loop restricted to rows [row0, row0+nrows): serial,
 * row-major, out[(r-row0)*cols + c].  nrows == rows, row0 == 0 is the whole FullSolve grid. */
int oracle_solve_rows(int64_t n, const double *x, const double *y, const double *w,
                      const int64_t *kr, const int64_t *kc, double x_min, double x_step,
                      double y_min, double y_step, int64_t cols, int64_t row0, int64_t nrows,
                      double exp, double *out)
{
    keymap m;
    if (build_keymap(&m, n, kr, kc) != 0) return -1;
    for (int64_t r = row0; r < row0 + nrows; r++)
        for (int64_t c = 0; c < cols; c++)
            out[(r - row0) * cols + c] = cell_idw(&m, n, x, y, w, x_min, x_step, y_min, y_step, exp, r, c);
    keymap_free(&m);
    return 0;
}

/* The same per-cell evaluation on an arbitrary list of cells (rr[i], cc[i]): lets the tests
 * spot-check full-size rasters (BASELINE.json configs c2..c5) without running the whole grid on
 * the CPU.  nthreads > 1 splits the list statically. */
typedef struct {
    const keymap *m; int64_t n; const double *x, *y, *w; double x_min, x_step, y_min, y_step, exp;
    const int64_t *rr, *cc;
    int mode; int64_t lo, hi; double *out;
} cells_job;

static void *cells_worker(void *arg)
{
    cells_job *j = (cells_job *)arg;
    for (int64_t i = j->lo; i < j->hi; i++)
        j->out[i] = cell_idw_mode(j->m, j->n, j->x, j->y, j->w, j->x_min, j->x_step, j->y_min, j->y_step, j->exp,
                                  j->rr[i], j->cc[i], j->mode);
    return NULL;
}

int oracle_solve_cells_mode(int64_t n, const double *x, const double *y, const double *w,
                            const int64_t *kr, const int64_t *kc, double x_min, double x_step,
                            double y_min, double y_step, int64_t ncells, const int64_t *rr,
                            const int64_t *cc, double exp, int nthreads, int mode, double *out);

int oracle_solve_cells(int64_t n, const double *x, const double *y, const double *w,
                       const int64_t *kr, const int64_t *kc, double x_min, double x_step,
                       double y_min, double y_step, int64_t ncells, const int64_t *rr,
                       const int64_t *cc, double exp, int nthreads, double *out)
{
    return oracle_solve_cells_mode(n, x, y, w, kr, kc, x_min, x_step, y_min, y_step, ncells, rr, cc, exp, nthreads, 0, out);
}

/* mode: 0 = array order, 1 = reverse order, 2 = the reference's float64 terms summed in long double (cell_idw_mode) */
int oracle_solve_cells_mode(int64_t n, const double *x, const double *y, const double *w,
                            const int64_t *kr, const int64_t *kc, double x_min, double x_step,
                            double y_min, double y_step, int64_t ncells, const int64_t *rr,
                            const int64_t *cc, double exp, int nthreads, int mode, double *out)
{
    keymap m;
    if (build_keymap(&m, n, kr, kc) != 0) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    cells_job jobs[256];
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads; t++) {
        cells_job j = {&m, n, x, y, w, x_min, x_step, y_min, y_step, exp, rr, cc, mode,
                       ncells * t / nthreads, ncells * (t + 1) / nthreads, out};
        jobs[t] = j;
    }
    for (int t = 1; t < nthreads; t++)
        if (pthread_create(&th[started], NULL, cells_worker, &jobs[t]) == 0) started++;
        else cells_worker(&jobs[t]);
    cells_worker(&jobs[0]);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    keymap_free(&m);
    return 0;
}

int oracle_full_solve(int64_t n, const double *x, const double *y, const double *w,
                      const int64_t *kr, const int64_t *kc, double x_min, double x_step,
                      double y_min, double y_step, int64_t rows, int64_t cols, double exp, double *out)
{
    return oracle_solve_rows(n, x, y, w, kr, kc, x_min, x_step, y_min, y_step, cols, 0, rows, exp, out);
}

typedef struct {
    const keymap *m;
    int64_t n;
    const double *x, *y, *w;
    double x_min, x_step, y_min, y_step, exp;
    int64_t cols, row0, nrows, chunk_r, chunk_c, jobs_c, njobs;
    atomic_llong next; /* the jobs channel: workers pull the next chunk index */
    double *out;
} chunk_pool;

/* features/idw/idw_chunk.go:27-41 makeGridIDW: one worker draining the jobs channel. */
static void *chunk_worker(void *arg)
{
    chunk_pool *p = (chunk_pool *)arg;
    for (;;) {
        int64_t j = atomic_fetch_add(&p->next, 1);
        if (j >= p->njobs) break;
        int64_t r0 = p->row0 + (j / p->jobs_c) * p->chunk_r, c0 = (j % p->jobs_c) * p->chunk_c;
        int64_t r1 = r0 + p->chunk_r < p->row0 + p->nrows ? r0 + p->chunk_r : p->row0 + p->nrows;
        int64_t c1 = c0 + p->chunk_c < p->cols ? c0 + p->chunk_c : p->cols;
        for (int64_t r = r0; r < r1; r++)
            for (int64_t c = c0; c < c1; c++)
                p->out[(r - p->row0) * p->cols + c] = cell_idw(p->m, p->n, p->x, p->y, p->w, p->x_min,
                                                              p->x_step, p->y_min, p->y_step, p->exp, r, c);
    }
    return NULL;
}

int oracle_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* features/idw/idw_chunk.go:43-105 ChunkSolve: chunk jobs enumerated row-major (:69-75) with
 * Min() clamps for ragged edges, oversize chunk -> quarter per axis (:48-53), a pool of workers
 * pulls jobs from a channel (:59-68).  The goroutine pool becomes a pthread pool over the same
 * job list; nthreads <= 0 means all online cores (Go schedules on GOMAXPROCS = all cores).
 * Rows restricted to [row0,row0+nrows) so a band can be timed (bench cpu_baseline). */
int oracle_chunk_solve_rows(int64_t n, const double *x, const double *y, const double *w,
                            const int64_t *kr, const int64_t *kc, double x_min, double x_step,
                            double y_min, double y_step, int64_t rows, int64_t cols, int64_t row0,
                            int64_t nrows, int64_t chunk_r, int64_t chunk_c, double exp,
                            int nthreads, double *out)
{
    if (chunk_r > rows || chunk_c > cols) { chunk_r = rows / 4; chunk_c = cols / 4; }
    if (chunk_r <= 0 || chunk_c <= 0) return -2; /* the reference would loop forever / panic */
    keymap m;
    if (build_keymap(&m, n, kr, kc) != 0) return -1;
    chunk_pool p;
    p.m = &m; p.n = n; p.x = x; p.y = y; p.w = w;
    p.x_min = x_min; p.x_step = x_step; p.y_min = y_min; p.y_step = y_step; p.exp = exp;
    p.cols = cols; p.row0 = row0; p.nrows = nrows; p.chunk_r = chunk_r; p.chunk_c = chunk_c;
    p.jobs_c = (cols + chunk_c - 1) / chunk_c;
    p.njobs = ((nrows + chunk_r - 1) / chunk_r) * p.jobs_c;
    atomic_init(&p.next, 0);
    p.out = out;
    if (nthreads <= 0) nthreads = oracle_num_threads();
    if ((int64_t)nthreads > p.njobs) nthreads = (int)p.njobs;
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int t = 0; t < nthreads - 1; t++)
        if (pthread_create(&th[started], NULL, chunk_worker, &p) == 0) started++;
    chunk_worker(&p);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    keymap_free(&m);
    return 0;
}

int oracle_chunk_solve(int64_t n, const double *x, const double *y, const double *w,
                       const int64_t *kr, const int64_t *kc, double x_min, double x_step,
                       double y_min, double y_step, int64_t rows, int64_t cols, int64_t chunk_r,
                       int64_t chunk_c, double exp, int nthreads, double *out)
{
    return oracle_chunk_solve_rows(n, x, y, w, kr, kc, x_min, x_step, y_min, y_step, rows, cols, 0,
                                   rows, chunk_r, chunk_c, exp, nthreads, out);
}

/* cmd/idw.go:153 exponent list: exp := es; exp < ee; exp += ei, accumulated in float64.
 * Returns the count; writes at most max_out values. */
int64_t oracle_exp_range(double es, double ee, double ei, double *out, int64_t max_out)
{
    int64_t n = 0;
    for (double e = es; e < ee; e += ei) {
        if (n < max_out) out[n] = e;
        n++;
        if (n > (1 << 20)) break;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * "Truth" evaluation in 80-bit long double (64-bit mantissa): same semantics, pairwise-free
 * plain summation, used only to show that oracle and GPU both sit inside 1e-12 of the exact
 * value on small grids.  Also returns the scale  sum|w_i| h_i / sum h_i  the tolerance is
 * defined on.
 * ---------------------------------------------------------------------------------------- */
int oracle_full_solve_ld(int64_t n, const double *x, const double *y, const double *w,
                         const int64_t *kr, const int64_t *kc, double x_min, double x_step,
                         double y_min, double y_step, int64_t rows, int64_t cols, double exp,
                         double *out, double *scale)
{
    keymap m;
    if (build_keymap(&m, n, kr, kc) != 0) return -1;
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++) {
            double px = x_min + (double)c * x_step, py = y_min + (double)r * y_step;
            int hit = 0;
            uint64_t i = key_hash(r, c) & m.mask;
            while (m.slot[i] >= 0) {
                if (m.kr[i] == r && m.kc[i] == c) {
                    out[r * cols + c] = w[m.slot[i]];
                    if (scale) scale[r * cols + c] = fabs(w[m.slot[i]]);
                    hit = 1;
                    break;
                }
                i = (i + 1) & m.mask;
            }
            if (hit) continue;
            long double num = 0, den = 0, anum = 0;
            for (int64_t k = 0; k < n; k++) {
                long double dx = (long double)x[k] - px, dy = (long double)y[k] - py;
                long double d2 = dx * dx + dy * dy;
                long double h = (exp == 0) ? 1.0L : powl(d2, -(long double)exp / 2);
                num += w[k] * h;
                anum += fabsl((long double)w[k]) * h;
                den += h;
            }
            out[r * cols + c] = (double)(num / den);
            if (scale) scale[r * cols + c] = (double)(anum / den);
        }
    keymap_free(&m);
    return 0;
}
