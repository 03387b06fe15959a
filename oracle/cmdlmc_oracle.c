/*
 * cmdlmc_oracle.c -- CPU restatement of the cMD/LMC per-frame hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (cmdlmc_b200/) is CUDA-only and fails loudly without its extension.
 *
 * Every function restates one function of the reference (paths relative to /root/reference)
 * in plain C99, written from the algorithm's description (no code copied), compiled WITHOUT
 * -ffast-math and with -ffp-contract=off so that it is a fixed, reproducible definition.
 * Parity status: PINNED for A1-A13 -- checked in tests/test_oracle_golden.py against the
 * reference's own compiled Cython (oracle/_ref) and its Python layers imported in place, and
 * against the committed golden vectors in tests/golden/ generated from them.
 * PARITY UNPINNED for the legacy pieces that are not in the reference tree as code
 * (activation-energy / exponential rates A9', legacy LMC sweep A14, GSL-style MT19937 stream):
 * restated from the in-tree specification text only (mdlmc/IO/config_parser.py:182-189,322-349).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* kind: 0 = orthorhombic (AtomBoxCubic), 1 = general cell (AtomBoxMonoclinic).
 * box layout (double[21]): [0..2] periodic_boundaries_extended (ortho),
 *                          [3..11] h   row-major, columns are the cell vectors (PBCHelper.pyx:255-258)
 *                          [12..20] h_inv row-major (PBCHelper.pyx:259) */
typedef struct {
    int kind;
    double pbc[3];
    double h[9];
    double hinv[9];
    /* AtomBoxWater distance conversions (PBCHelper.pyx:278-351): conv 0 none, 1 linear, 2 ramp */
    int conv;
    double ca, cb, cd0, cleft, cright;
} orc_box;

ORC_API int orc_box_size(void) { return (int)sizeof(orc_box); }

ORC_API void orc_box_init(orc_box *bx, int kind, const double *pbc3, const double *h9,
                          const double *hinv9, int conv, const double *conv5)
{
    memset(bx, 0, sizeof(*bx));
    bx->kind = kind;
    if (pbc3) memcpy(bx->pbc, pbc3, 3 * sizeof(double));
    if (h9) memcpy(bx->h, h9, 9 * sizeof(double));
    if (hinv9) memcpy(bx->hinv, hinv9, 9 * sizeof(double));
    bx->conv = conv;
    if (conv5) { bx->ca = conv5[0]; bx->cb = conv5[1]; bx->cd0 = conv5[2];
                 bx->cleft = conv5[3]; bx->cright = conv5[4]; }
}

/* ---- A1: mdlmc/cython_exts/atoms/numpyatom.pyx:33-42 (diff_ptr) ------------------------- */
static void diff_ortho(const double *a1, const double *a2, const double *pbc, double *d)
{
    for (int i = 0; i < 3; i++) {
        d[i] = a2[i] - a1[i];
        while (d[i] < -pbc[i] / 2) d[i] += pbc[i];
        while (d[i] > pbc[i] / 2) d[i] -= pbc[i];
    }
}

/* ---- A2: numpyatom.pyx:173-179 (length_ptr) ---------------------------------------------- */
static double length_ortho(const double *a1, const double *a2, const double *pbc)
{
    double d[3];
    diff_ortho(a1, a2, pbc, d);
    return sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

/* ---- math_helper.pyx:50-60 (matrix_mult_ptr): in-place row-major 3x3 times vector,
 *      accumulation order ((0 + m0 v0) + m1 v1) + m2 v2 ------------------------------------ */
static void matvec3(const double *m, double *v)
{
    double r[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[i] += m[3 * i + j] * v[j];
    v[0] = r[0]; v[1] = r[1]; v[2] = r[2];
}

/* ---- A3: numpyatom.pyx:61-74 (diff_ptr_nonortho); round == C99 round (cython_gsl.round) --- */
static void diff_general(const double *a1, const double *a2, const double *h, const double *hinv,
                         double *d)
{
    for (int i = 0; i < 3; i++) d[i] = a2[i] - a1[i];
    matvec3(hinv, d);
    for (int i = 0; i < 3; i++) d[i] -= round(d[i]);
    matvec3(h, d);
}

/* ---- A4: numpyatom.pyx:101-123 (length_nonortho_bruteforce_ptr): 27 images, min^2 starts 1e6 */
static double length_general(const double *a1, const double *a2, const double *h,
                             const double *hinv)
{
    double d[3], mind = 1e6;
    diff_general(a1, a2, h, hinv, d);
    for (int i = -1; i < 2; i++)
        for (int j = -1; j < 2; j++)
            for (int k = -1; k < 2; k++) {
                double v[3];
                for (int dim = 0; dim < 3; dim++)
                    v[dim] = d[dim] + i * h[3 * dim] + j * h[3 * dim + 1] + k * h[3 * dim + 2];
                double dist = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
                if (dist < mind) mind = dist;
            }
    return sqrt(mind);
}

/* math_helper.pyx:16-23 (dot_product_ptr) */
static double dot3(const double *a, const double *b)
{
    double r = 0;
    for (int i = 0; i < 3; i++) r += a[i] * b[i];
    return r;
}

/* ---- A5: numpyatom.pyx:244-264 (angle_ptr): angle between (a2-a1) and (a4-a3), ortho wrap -- */
static double angle_ortho(const double *a1, const double *a2, const double *a3, const double *a4,
                          const double *pbc)
{
    double v1[3], v2[3];
    for (int i = 0; i < 3; i++) {
        v1[i] = a2[i] - a1[i];
        v2[i] = a4[i] - a3[i];
        while (v1[i] > pbc[i] / 2) v1[i] -= pbc[i];
        while (v1[i] < -pbc[i] / 2) v1[i] += pbc[i];
        while (v2[i] > pbc[i] / 2) v2[i] -= pbc[i];
        while (v2[i] < -pbc[i] / 2) v2[i] += pbc[i];
    }
    return acos(dot3(v1, v2) / sqrt(dot3(v1, v1)) / sqrt(dot3(v2, v2)));
}

/* ---- A5: numpyatom.pyx:280-291 (angle_ptr_nonortho): fractional wrap only (no 27 images) --- */
static double angle_general(const double *a1, const double *a2, const double *a3,
                            const double *a4, const double *h, const double *hinv)
{
    double v1[3], v2[3];
    diff_general(a1, a2, h, hinv, v1);
    diff_general(a3, a4, h, hinv, v2);
    return acos(dot3(v1, v2) / sqrt(dot3(v1, v1)) / sqrt(dot3(v2, v2)));
}

/* ---- PBCHelper.pyx:318-324, 342-351 (convert_distance) ------------------------------------ */
static double convert_distance(const orc_box *bx, double d)
{
    if (bx->conv == 1) {
        if (bx->cleft < d && d < bx->cright) return bx->ca * d + bx->cb;
        return d;
    }
    if (bx->conv == 2) {
        if (bx->cleft < d && d < bx->cright) {
            if (d < bx->cd0) return bx->cb;
            return bx->ca * (d - bx->cd0) + bx->cb;
        }
        return d;
    }
    return d;
}

/* ---- A6 dispatch: PBCHelper.pyx:228-239 (Cubic), :262-275 (Monoclinic), :282-303 (Water) --- */
static double box_length(const orc_box *bx, const double *a, const double *b)
{
    double d = bx->kind == 0 ? length_ortho(a, b, bx->pbc) : length_general(a, b, bx->h, bx->hinv);
    return convert_distance(bx, d);
}

static void box_distance(const orc_box *bx, const double *a, const double *b, double *out)
{
    if (bx->kind == 0) diff_ortho(a, b, bx->pbc, out);
    else diff_general(a, b, bx->h, bx->hinv, out);
}

/* AtomBox.length, PBCHelper.pyx:74-85 */
ORC_API void orc_length(const orc_box *bx, const double *a, const double *b, long n, double *out)
{
    for (long i = 0; i < n; i++) out[i] = box_length(bx, a + 3 * i, b + 3 * i);
}

/* AtomBox.distance, PBCHelper.pyx:56-70 (vector, no 27-image search, no water conversion) */
ORC_API void orc_distance(const orc_box *bx, const double *a, const double *b, long n, double *out)
{
    for (long i = 0; i < n; i++) box_distance(bx, a + 3 * i, b + 3 * i, out + 3 * i);
}

/* AtomBox.length_all_to_all, PBCHelper.pyx:88-95 */
ORC_API void orc_length_all_to_all(const orc_box *bx, const double *a, long n, const double *b,
                                   long m, double *out)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++)
        for (long j = 0; j < m; j++) out[i * m + j] = box_length(bx, a + 3 * i, b + 3 * j);
}

/* AtomBox.angle, PBCHelper.pyx:133-134, 237-239, 273-275: vertex is the second atom */
ORC_API void orc_angle(const orc_box *bx, const double *a1, const double *a2, const double *a3,
                       long n, double *out)
{
    for (long i = 0; i < n; i++) {
        const double *p1 = a1 + 3 * i, *p2 = a2 + 3 * i, *p3 = a3 + 3 * i;
        out[i] = bx->kind == 0 ? angle_ortho(p2, p1, p2, p3, bx->pbc)
                               : angle_general(p2, p1, p2, p3, bx->h, bx->hinv);
    }
}

/* AtomBox.next_neighbor, PBCHelper.pyx:153-167 (box_multiplier == 1: the multiplied loop of the
 * reference reads past the frame and is undefined).  Strict '<', first minimum wins. */
ORC_API void orc_next_neighbor(const orc_box *bx, const double *pos, const double *frame, long n,
                               int *idx, double *dist)
{
    double best = 1e30;
    int bi = -1;
    for (long j = 0; j < n; j++) {
        double l = box_length(bx, pos, frame + 3 * j);
        if (l < best) { best = l; bi = (int)j; }
    }
    *idx = bi;
    *dist = best;
}

/* ---- A7: mdlmc/topo/topology.py:55-72 (get_topology_bruteforce) ---------------------------
 * pairs j<i evaluated as length(frame[i], frame[j]); kept if dist <= rc (rc = cutoff + buffer
 * added by the caller exactly as Python does); symmetric; a distance of exactly 0.0 is dropped
 * (sparse zero of the LIL matrix); output in LIL->COO order = row-major, columns ascending.
 * Returns the number of directed pairs P, or -(needed) if cap is too small. */
ORC_API long orc_topology_bruteforce(const orc_box *bx, const double *frame, long n, double rc,
                                     int *row, int *col, double *dist, long cap)
{
    double *m = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
    if (!m) return -1;
#pragma omp parallel for schedule(dynamic, 8)
    for (long i = 0; i < n; i++) {
        m[i * n + i] = 0.0;
        for (long j = 0; j < i; j++) {
            double d = box_length(bx, frame + 3 * i, frame + 3 * j);
            double v = (d <= rc) ? d : 0.0;
            m[i * n + j] = v;
            m[j * n + i] = v;
        }
    }
    long p = 0;
    for (long i = 0; i < n; i++)
        for (long j = 0; j < n; j++)
            if (m[i * n + j] != 0.0) {
                if (p < cap) { row[p] = (int)i; col[p] = (int)j; dist[p] = m[i * n + j]; }
                p++;
            }
    free(m);
    return p <= cap ? p : -p;
}

/* ---- A8 pieces: mdlmc/topo/topology.py:80-114 --------------------------------------------- */
/* line 110: dist = atombox.length(frame[row], frame[col]) */
ORC_API void orc_pairs_refresh(const orc_box *bx, const double *frame, const int *row,
                               const int *col, long p, double *dist)
{
    for (long k = 0; k < p; k++)
        dist[k] = box_length(bx, frame + 3 * (long)row[k], frame + 3 * (long)col[k]);
}

/* lines 96-107: running displacement and the rebuild decision for ONE new frame.
 * displacement[] is updated in place (+= dr, dr = length(last, cur) or 0 for the first frame);
 * returns 1 when the two largest entries sum to more than `buffer` (strict >), in which case
 * the caller rebuilds and the displacement restarts from 0.  n >= 2. */
ORC_API int orc_verlet_step(const orc_box *bx, const double *last, const double *cur, long n,
                            double buffer, double *displacement)
{
    double m1 = -INFINITY, m2 = -INFINITY; /* m1 >= m2: two largest */
    for (long i = 0; i < n; i++) {
        double dr = last ? box_length(bx, last + 3 * i, cur + 3 * i) : 0.0;
        displacement[i] += dr;
        double v = displacement[i];
        if (v > m1) { m2 = m1; m1 = v; }
        else if (v > m2) m2 = v;
    }
    /* np.sort(displacement)[-2:] -> (second largest, largest); sum is commutative */
    if (m2 + m1 > buffer) {
        for (long i = 0; i < n; i++) displacement[i] = 0.0;
        return 1;
    }
    return 0;
}

/* ---- A9: mdlmc/LMC/jumprate_generators.py:33-34 (Fermi), :42-43 (FermiAngle) -------------- */
/* ---- A9' (PARITY UNPINNED; spec text mdlmc/IO/config_parser.py:322-349):
 *      kind 0 Fermi            w = a / (1 + exp((x - b) / c))             par = a,b,c
 *      kind 1 FermiAngle       0 where theta < theta0 else Fermi          par = a,b,c,theta0
 *      kind 2 ActivationEnergy E = a (x-d0) / sqrt(b + 1/(x-d0)^2), w = A exp(-E/(kB T)),
 *                              w = A for x <= d0 (our choice)             par = A,a,b,d0,T
 *      kind 3 Exponential      w = a exp(b x)                             par = a,b          */
#define ORC_KB_EV 8.617333262e-5
static double rate_eval(int kind, const double *par, double x, double theta)
{
    switch (kind) {
    case 0: return par[0] / (1 + exp((x - par[1]) / par[2]));
    case 1: return theta < par[3] ? 0.0 : par[0] / (1 + exp((x - par[1]) / par[2]));
    case 2: {
        double u = x - par[3];
        if (!(u > 0)) return par[0];
        double e = par[1] * u / sqrt(par[2] + 1.0 / (u * u));
        return par[0] * exp(-e / (ORC_KB_EV * par[4]));
    }
    case 3: return par[0] * exp(par[1] * x);
    }
    return 0.0;
}

ORC_API void orc_rates(int kind, const double *par, const double *x, const double *theta, long n,
                       double *out)
{
    for (long i = 0; i < n; i++) out[i] = rate_eval(kind, par, x[i], theta ? theta[i] : 0.0);
}

/* ---- NumPy float64 add.reduce (np.sum) summation order, used at MDMC.py:85 ----------------
 * pairwise summation with 8 accumulators in blocks of <=128 (NumPy's documented algorithm). */
static double pairwise_sum(const double *a, long n)
{
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}

ORC_API double orc_np_sum(const double *a, long n) { return pairwise_sum(a, n); }

/* ---- A10-A12: mdlmc/LMC/MDMC.py:77-171, 229-248 -------------------------------------------
 * One KMC replica in exact-replay mode.  The per-frame topology + rates are given as CSR-like
 * concatenated arrays: frame f owns pairs [fptr[f], fptr[f+1]).  `u` is derived from the uniform
 * stream the reference would draw from the global legacy RandomState after its initial shuffle,
 * strictly alternating:  u[2e] = -np.log(1 - np.random.random()) (MDMC.py:148; evaluated by NumPy
 * on the host so that no libm/NumPy log difference can enter -- the reference feeds rounding
 * noise back into kmc_time with gain S_0/S_t per event, quirk Q1),  u[2e+1] -> the u of
 * np.random.uniform(0,S) == S*u (MDMC.py:110).  All quirks of the reference are kept:
 *   Q1 current_rate is the total of frame 0 forever (MDMC.py:146);
 *   Q2 a same-frame event re-masks the arrays of the last consumed frame (MDMC.py:98,105-108);
 *   Q5 np.sum (pairwise) for totals, sequential cumsum for selection; Q6 Python // and %;
 *   Q7 searchsorted side='left'.
 * Stops cleanly when the trajectory is exhausted (the reference raises RuntimeError there) or
 * after max_events.  Outputs per event: ev_frame (sweep), ev_dframe, ev_time, ev_start,
 * ev_dest, ev_proton.  frame_event[f] = index of the event whose cached-frame flush yields
 * frame f (MDMC.py:94-96), i.e. frame f is stamped with time ev_time[frame_event[f]] and seen
 * with the lattice before that event's move; -1 if never yielded.  Returns number of events. */
static double py_floordiv(double a, double b)
{
    /* CPython float_floor_div */
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0 && ((b < 0) != (mod < 0))) div -= 1.0;
    if (div != 0) {
        double fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
        return fl;
    }
    return copysign(0.0, a / b);
}

static double py_mod(double a, double b)
{
    double mod = fmod(a, b);
    if (mod != 0) { if ((b < 0) != (mod < 0)) mod += b; }
    else mod = copysign(0.0, b);
    return mod;
}

/* ---- A11 alone: MDMC.py:121-171 (fastforward_to_next_jump) on a given stream of per-frame
 * total rates (cycled when `cycle` != 0, like itertools.cycle in tests/LMC/test_MDMC.py).
 * u[e] = -np.log(1 - r_e) for the e-th np.random.random() draw r_e (NumPy-evaluated).  rows[e] = (sweep, delta_frame, kmc_time). */
ORC_API long orc_fastforward(const double *rates, long nrates, int cycle, double dt,
                             const double *u, long nevents, double *rows)
{
    long pos = 0, nev = 0, sweep = 0;
    double kmc_time = 0.0;
#define NEXT_RATE(dst) do { if (pos >= nrates) { if (!cycle) return nev; pos = 0; } dst = rates[pos++]; } while (0)
    double current_rate;
    NEXT_RATE(current_rate);
    while (nev < nevents) {
        double time_selector = u[nev]; /* = -np.log(1 - np.random.random()) evaluated by NumPy */
        double t_trial = time_selector / current_rate;
        long delta_frame;
        if (py_floordiv(kmc_time + t_trial, dt) == py_floordiv(kmc_time, dt)) {
            kmc_time += t_trial;
            delta_frame = 0;
        } else {
            double delta_t = dt - py_mod(kmc_time, dt);
            delta_frame = 1;
            double current_probsum = current_rate * delta_t, next_rate, next_probsum;
            NEXT_RATE(next_rate);
            next_probsum = current_probsum + next_rate * dt;
            while (next_probsum < time_selector) {
                delta_frame += 1;
                current_probsum = next_probsum;
                NEXT_RATE(next_rate);
                next_probsum = current_probsum + next_rate * dt;
            }
            double rest = time_selector - current_probsum;
            delta_t += (delta_frame - 1) * dt + rest / next_rate;
            kmc_time += delta_t;
        }
        sweep += delta_frame;
        rows[3 * nev] = (double)sweep; rows[3 * nev + 1] = (double)delta_frame;
        rows[3 * nev + 2] = kmc_time;
        nev++;
    }
#undef NEXT_RATE
    return nev;
}

typedef struct {
    const long *fptr; const int *start; const int *dest; const double *omega;
    long nframes; long next_frame;     /* next frame the generator would consume */
    int *lattice;
    /* last consumed frame (remember_last_element, MDMC.py:83-84): its allowed mask */
    long last_frame; unsigned char *mask; double *scratch; long maxp;
} kmc_state;

/* jumprate_generator + np.sum, MDMC.py:229-238,85.  returns 0 at end of trajectory */
static int kmc_next_rate(kmc_state *s, double *total)
{
    if (s->next_frame >= s->nframes) return 0;
    long f = s->next_frame++;
    long p0 = s->fptr[f], p1 = s->fptr[f + 1], m = 0;
    for (long p = p0; p < p1; p++) {
        int ok = s->lattice[s->start[p]] > 0 && !(s->lattice[s->dest[p]] > 0);
        s->mask[p - p0] = (unsigned char)ok;
        if (ok) s->scratch[m++] = s->omega[p];
    }
    s->last_frame = f;
    *total = pairwise_sum(s->scratch, m);
    return 1;
}

ORC_API long orc_kmc_replay(const long *fptr, const int *start, const int *dest,
                            const double *omega, long nframes, int *lattice, long nsites,
                            double dt, const double *u, long max_events,
                            long *ev_frame, long *ev_dframe, double *ev_time, int *ev_start,
                            int *ev_dest, int *ev_proton, long *frame_event,
                            int *lattice_trace /* [max_events, nsites] after each event or NULL */)
{
    kmc_state s;
    long maxp = 0;
    for (long f = 0; f < nframes; f++)
        if (fptr[f + 1] - fptr[f] > maxp) maxp = fptr[f + 1] - fptr[f];
    s.fptr = fptr; s.start = start; s.dest = dest; s.omega = omega;
    s.nframes = nframes; s.next_frame = 0; s.lattice = lattice; s.last_frame = -1;
    s.maxp = maxp;
    s.mask = (unsigned char *)malloc((size_t)maxp + 1);
    s.scratch = (double *)malloc(sizeof(double) * ((size_t)maxp + 1));
    for (long f = 0; f < nframes; f++) frame_event[f] = -1;

    long nev = 0, sweep = 0, flushed = 0;
    double kmc_time = 0.0, current_rate;
    if (!kmc_next_rate(&s, &current_rate)) goto done;
    while (nev < max_events) {
        double time_selector = u[2 * nev]; /* = -np.log(1 - np.random.random()), MDMC.py:148 */
        double t_trial = time_selector / current_rate;
        long delta_frame;
        if (py_floordiv(kmc_time + t_trial, dt) == py_floordiv(kmc_time, dt)) {
            kmc_time += t_trial;
            delta_frame = 0;
        } else {
            double delta_t = dt - py_mod(kmc_time, dt);
            delta_frame = 1;
            double current_probsum = current_rate * delta_t, next_rate, next_probsum;
            if (!kmc_next_rate(&s, &next_rate)) goto done;
            next_probsum = current_probsum + next_rate * dt;
            while (next_probsum < time_selector) {
                delta_frame += 1;
                current_probsum = next_probsum;
                if (!kmc_next_rate(&s, &next_rate)) goto done;
                next_probsum = current_probsum + next_rate * dt;
            }
            double rest = time_selector - current_probsum;
            delta_t += (delta_frame - 1) * dt + rest / next_rate;
            kmc_time += delta_t;
        }
        sweep += delta_frame;
        /* flush of the frame cache (MDMC.py:94-96): frames consumed since the last event */
        for (; flushed < s.next_frame; flushed++) frame_event[flushed] = nev;
        /* move_proton (MDMC.py:101-119) on the arrays of the last consumed frame */
        long f = s.last_frame, p0 = fptr[f], p1 = fptr[f + 1];
        double cum = 0.0;
        long m = 0;
        for (long p = p0; p < p1; p++) {
            if (!s.mask[p - p0]) continue;
            if (lattice[start[p]] > 0 && !(lattice[dest[p]] > 0)) {
                cum += omega[p];            /* np.cumsum: sequential */
                s.scratch[m++] = cum;
            }
        }
        if (m == 0) break; /* reference: IndexError on cumsum[-1] */
        double draw = 0.0 + (s.scratch[m - 1] - 0.0) * u[2 * nev + 1];
        long lo = 0, hi = m; /* searchsorted side='left' */
        while (lo < hi) {
            long mid = lo + (hi - lo) / 2;
            if (s.scratch[mid] < draw) lo = mid + 1; else hi = mid;
        }
        long sel = -1, cnt = 0;
        for (long p = p0; p < p1; p++) {
            if (!s.mask[p - p0]) continue;
            if (lattice[start[p]] > 0 && !(lattice[dest[p]] > 0)) {
                if (cnt == lo) { sel = p; break; }
                cnt++;
            }
        }
        if (sel < 0) break; /* draw beyond the last entry: reference would IndexError */
        int si = start[sel], di = dest[sel], proton = lattice[si];
        lattice[di] = proton;
        lattice[si] = 0;
        ev_frame[nev] = sweep; ev_dframe[nev] = delta_frame; ev_time[nev] = kmc_time;
        ev_start[nev] = si; ev_dest[nev] = di; ev_proton[nev] = proton;
        if (lattice_trace) memcpy(lattice_trace + nev * nsites, lattice, sizeof(int) * (size_t)nsites);
        nev++;
    }
done:
    free(s.mask);
    free(s.scratch);
    return nev;
}

/* ---- A13: mdlmc/LMC/output.py:17-49 (MeanSquareDisplacement) ------------------------------ */
/* determine_proton_positions, output.py:25-30 */
ORC_API void orc_proton_positions(const double *pos, const int *lattice, long nsites, double *out)
{
    for (long s = 0; s < nsites; s++)
        if (lattice[s] > 0) {
            long l = lattice[s] - 1;
            out[3 * l] = pos[3 * s]; out[3 * l + 1] = pos[3 * s + 1]; out[3 * l + 2] = pos[3 * s + 2];
        }
}

/* update_displacement, output.py:35-43: displacement += atombox.distance(snapshot, new) */
ORC_API void orc_msd_update(const orc_box *bx, double *snapshot, double *displacement,
                            const double *pos, const int *lattice, long nsites, long nprot)
{
    double *np_ = (double *)calloc((size_t)nprot * 3, sizeof(double));
    orc_proton_positions(pos, lattice, nsites, np_);
    for (long k = 0; k < nprot; k++) {
        double d[3];
        box_distance(bx, snapshot + 3 * k, np_ + 3 * k, d);
        for (int c = 0; c < 3; c++) displacement[3 * k + c] += d[c];
    }
    memcpy(snapshot, np_, sizeof(double) * (size_t)nprot * 3);
    free(np_);
}

/* CovalentAutocorrelation.calculate, output.py:13-14 */
ORC_API long orc_autocorr(const int *lattice, const int *lattice0, long nsites)
{
    long c = 0;
    for (long s = 0; s < nsites; s++) c += (lattice[s] == lattice0[s]) && (lattice[s] != 0);
    return c;
}

/* ---- MT19937 (Matsumoto & Nishimura 2002 init_genrand) -- the generator behind both NumPy's
 * legacy RandomState and GSL's default gsl_rng_mt19937.  PARITY UNPINNED for the GSL flavour:
 * the reference tree has no GSL RNG call site; gsl_rng_uniform = u32 / 2^32 and
 * gsl_rng_uniform_int(n) = rejection on u32 / (0xffffffff / n) are GSL's documented forms. */
typedef struct { uint32_t mt[624]; int idx; } orc_mt;

ORC_API int orc_mt_size(void) { return (int)sizeof(orc_mt); }

ORC_API void orc_mt_seed(orc_mt *g, uint32_t s)
{
    g->mt[0] = s;
    for (int i = 1; i < 624; i++)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

ORC_API uint32_t orc_mt_u32(orc_mt *g)
{
    if (g->idx >= 624) {
        for (int k = 0; k < 624; k++) {
            uint32_t y = (g->mt[k] & 0x80000000u) | (g->mt[(k + 1) % 624] & 0x7fffffffu);
            g->mt[k] = g->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}

/* NumPy legacy random_sample(): 53-bit double from two draws */
ORC_API double orc_mt_double53(orc_mt *g)
{
    uint32_t a = orc_mt_u32(g) >> 5, b = orc_mt_u32(g) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

ORC_API double orc_gsl_uniform(orc_mt *g) { return orc_mt_u32(g) / 4294967296.0; }

ORC_API uint32_t orc_gsl_uniform_int(orc_mt *g, uint32_t n)
{
    uint32_t scale = 0xffffffffu / n, k;
    do { k = orc_mt_u32(g) / scale; } while (k >= n);
    return k;
}

/* ---- A14 (PARITY UNPINNED): legacy LMC sweep, restated from the specification text
 * mdlmc/IO/config_parser.py:182-189 ("A sweep is the number of single proton jump attempts,
 * after which (on average) each oxygen bond has been selected once") and the .bak tests
 * (tests/cython_exts/LMC/test_LMCRoutine.py.bak:33-42, tests/LMC/test_MDMC.py.bak:61-84).
 * One sweep on one frame = P attempts; attempt a draws a pair index k = pick[a] uniformly in
 * [0,P) and a uniform acc[a] in [0,1); the hop start[k]->dest[k] happens iff start is occupied,
 * dest is empty and acc[a] < prob[k]  (prob = omega(d) * dt, 0 beyond cutoff_radius).
 * Both streams are consumed for every attempt.  jumpmatrix (nsites x nsites, may be NULL) counts
 * hops per (start,dest) (sweep_with_jumpmatrix).  Returns the number of hops in this sweep. */
ORC_API long orc_lmc_sweep(const int *start, const int *dest, const double *prob, long p,
                           int *lattice, const int *pick, const double *acc, long nsites,
                           long *jumpmatrix)
{
    long jumps = 0;
    for (long a = 0; a < p; a++) {
        long k = pick[a];
        int si = start[k], di = dest[k];
        if (lattice[si] != 0 && lattice[di] == 0 && acc[a] < prob[k]) {
            lattice[di] = lattice[si];
            lattice[si] = 0;
            jumps++;
            if (jumpmatrix) jumpmatrix[(long)si * nsites + di]++;
        }
    }
    return jumps;
}

/* ---- throughput helper for bench.py's C "port" baseline: per frame all-pairs topology +
 * rates over a block of frames, OpenMP over frames.  Returns total directed pairs found. */
ORC_API long orc_bench_frames(const orc_box *bx, const double *frames, long nframes, long n,
                              double rc, int rate_kind, const double *par, double *rate_sum)
{
    long total = 0;
    double rs = 0.0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total, rs)
    for (long f = 0; f < nframes; f++) {
        const double *fr = frames + 3 * n * f;
        for (long i = 0; i < n; i++)
            for (long j = 0; j < i; j++) {
                double d = box_length(bx, fr + 3 * i, fr + 3 * j);
                if (d <= rc && d != 0.0) {
                    total += 2;
                    rs += 2 * rate_eval(rate_kind, par, d, 0.0);
                }
            }
    }
    *rate_sum = rs;
    return total;
}
