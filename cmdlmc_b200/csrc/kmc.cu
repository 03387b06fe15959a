// kmc.cu -- time-dependent-rate KMC over the per-frame topology (rows A10-A13 of SURVEY.md 8):
//   jumprate_generator / filter_allowed_transitions   MDMC.py:229-248
//   KMCLattice.fastforward_to_next_jump                MDMC.py:121-171
//   KMCLattice.move_proton / continuous_output         MDMC.py:101-119, 77-99
//   MeanSquareDisplacement / CovalentAutocorrelation   output.py:6-49, MDMC.py:179-208
//
// One WARP per replica, several replicas per CTA.  All replicas walk the frames of a topology
// block in the same order; the warps of a CTA are kept on the same frame by one barrier per
// frame so that the frame's (start, dest, omega) arrays are served from L1 for every replica
// after the first.  Per-replica state (lattice, occupancy bits, the allowed-mask of the last
// consumed frame) lives in shared memory; scalars are kept redundantly in every lane.
//
// Exactness: the state machine reproduces the reference's control flow including its quirks
// (Q1 stale current_rate, Q2 same-frame events re-mask the last consumed
// frame, Q3/Q4 cached frames stamped with the event time and seen with the pre-jump lattice,
// Q6 Python // and %, Q7 searchsorted side='left').  Two arithmetic flavours:
//   replay mode  every sum in NumPy's own order (np.sum pairwise tree, sequential np.cumsum), no
//                FMA contraction, time selectors taken from the host stream: bit-identical to the
//                CPU oracle on the same rates (the reference amplifies rounding noise through Q1,
//                so anything less drifts apart after a few hundred events);
//   Philox mode  warp-parallel sums in a fixed order (deterministic in the seed), for throughput.
// Every decision closer than 1e-9 relative to its boundary is counted by the tie audit
// (cmd_kmc_tie_count).
#include <math.h>
#include <stdlib.h>

#include "pbc.cuh"
#include "philox.cuh"
#include "tma.cuh"

#define KMC_PHASE_START 0
#define KMC_PHASE_SCAN 2
#define KMC_PHASE_HALT 3

struct KmcState {  // one per replica, global memory, persistent across cmd_kmc_advance calls
    double kmc_time, current_rate, time_selector, current_probsum, delta_t;
    long long sweep, delta_frame, n_events, site_updates, draws, frames_seen, n_rows, pending_row;
    long long cursor;   // position in the current replay stream
    long long log_pos;  // events logged since the last cmd_kmc_set_event_log
    int phase, reason;  // reason: 1 replay stream exhausted, 2 no allowed transition
    double u_sel;       // Philox mode: selection uniform drawn together with the time selector
    long long ev_resolved;  // logged events whose jump distance has been filled in (k_resolve_ev_dist)
};

// HydroniumTopology colvars inside the KMC (topology.py:213-232,260-353): the distance of a
// transition is rescaled depending on how long its proton has been sitting on the start site
struct HydParams {
    int on;
    RateParams rate;     // jump rate evaluated on the rescaled distance
    int tkind;           // 0 none, 1 ReLUTransformation, 2 InterpolatedTransformation
    double tpar[5];      // ReLU: a, b, d0, left_bound, right_bound
    const double *tx, *ty;   // interpolation table (device)
    int nt;
    double relax;        // DistanceInterpolator.relaxation_time; <= 0: no interpolator
    double frame_dt;     // trajectory time step: frame.time = frame_number * frame_dt
};

struct cmd_kmc {
    BoxParams bx;
    int n_sites, n_replicas, n_protons_max;
    double dt;
    int rng_mode;
    uint64_t seed;
    int replica_first, replica_step;   // global id of local replica r = first + r * step (Philox)
    int *d_lattice;      // [R][n_sites]
    int *d_lattice0;     // [R][n_sites] autocorrelation reference (output.py:10-11)
    KmcState *d_state;   // [R]
    double *d_u;         // replay stream [R][n_u]
    int64_t n_u;
    // event log
    int64_t ev_cap;
    long long *d_ev_frame;
    double *d_ev_time;
    int *d_ev_start, *d_ev_dest, *d_ev_proton;
    double *d_ev_dist;   // O-O distance of the jump pair at the event (jumpstat histogram)
    // observables
    int reset_freq, print_freq;
    int64_t row_cap;
    double *d_rows;      // [R][row_cap][6]
    double *d_snapshot;  // [R][n_sites][3]  indexed by proton label - 1
    double *d_disp;      // [R][n_sites][3]
    unsigned long long *d_ties;
    int64_t frames_total;
    int64_t obs_offset;  // observable frame numbers run ahead of the walked frames by this much
    HydParams hyd;
    double *d_tlast;     // [R][n_sites] time of the last jump per proton label - 1 (-1: never)
    // occupancy histogram: frames a site was seen occupied, kept as (closed intervals, open since)
    unsigned int *d_occ_count;   // [R][n_sites]
    int *d_occ_since;            // [R][n_sites] consumed-frame count at which the site was filled
    double *d_tx, *d_ty;
    // exact-replay scratch (per replica): the compacted allowed list of the last consumed frame
    void *d_exact;
    size_t exact_bytes;
    int solo_enabled;    // CMDLMC_B200_KMC_SOLO=0 keeps few-replica runs on the warp-per-replica kernel
    double sel_margin;   // CMDLMC_B200_KMC_SELECT_MARGIN scales the solo kernel's selection margin
};

struct KmcArgs {
    int n_sites, n_replicas, rng_mode, replicas_per_cta, mask_words, occ_words;
    int replica_first, replica_step;
    double dt, inv_dt;
    uint64_t seed;
    int64_t stride, nframes, frames_base, obs_base, n_u, ev_cap, row_cap;
    int reset_freq, print_freq;
    const int *start, *dest, *counts;
    const double *omega, *positions, *u, *dist;
    int *lattice, *lattice0;
    KmcState *state;
    long long *ev_frame;
    double *ev_time;
    int *ev_start, *ev_dest, *ev_proton;
    double *ev_dist;
    double *rows, *snapshot, *disp;
    unsigned long long *ties;
    HydParams hyd;
    double *tlast;
    unsigned int *occ_count;
    int *occ_since;
    // streaming (Philox) kernel: row index of the block's lists, smem ring geometry
    int fast, ro_pitch, nst_max;
    const int *rowoff;
    // exact-replay scratch, one slice per replica (see kmc_consume_exact)
    double sel_margin;   // solo kernel: relative safety margin of the parallel selection
    int exact, x_smem;   // x_smem: the scratch of the replica lives in shared memory
    int64_t x_cap, x_leaves;
    double *x_comp, *x_cum, *x_lsum;
    int *x_cidx, *x_loff, *x_ln;
};

// CPython / NumPy float floor-division and modulo (MDMC.py:152,156)
__device__ __forceinline__ double py_floordiv(double a, double b)
{
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

__device__ __forceinline__ double py_mod(double a, double b)
{
    double mod = fmod(a, b);
    if (mod != 0) { if ((b < 0) != (mod < 0)) mod += b; }
    else mod = copysign(0.0, b);
    return mod;
}

// floor(a / b) and a - floor(a / b) * b for finite a >= 0, b > 0 without fmod's division loop and
// without a division: the quotient estimate a * (1 / b) is at most one off, and one exact FMA
// remainder per candidate settles it.  Same values as py_floordiv / py_mod (both are the exact floor
// of the real quotient and the exact remainder).
__device__ __forceinline__ double floor_div_pos(double a, double b, double inv_b, double *rem)
{
    double q = floor(a * inv_b);
    double r = fma(-q, b, a);
    if (r < 0.0) { q -= 1.0; r = fma(-q, b, a); }
    else if (r >= b) { q += 1.0; r = fma(-q, b, a); }
    *rem = r;
    return q;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct WarpCtx {
    int lane;
    int *lat;            // [n_sites] labels
    unsigned *occ;       // occupancy bits
    unsigned *mask0;     // allowed-at-consumption bits of the last consumed frame
    int64_t base;        // element offset of the last consumed frame
    int p;               // its pair count
    // exact-replay mode: the allowed transitions of the last consumed frame, compacted in list
    // order (what jumprate_generator yields, MDMC.py:229-238), in global scratch
    double *comp, *cum, *lsum;
    int *cidx, *loff, *ln;
    int m;
    // streaming kernel: psum[stage][lane] = allowed rate of the pairs k = stage*1024 + i*32 + lane,
    // the frame's row index (shared memory) and the running per-lane total
    double *psum;
    const int *ro;
    double lane_total;
    double scan_inc;     // inclusive scan of the lane totals of the frame at element offset scan_base
    int64_t scan_base;
    int nst;
    double *tlast;       // hydronium: [n_sites] last jump time per proton label - 1 (shared memory)
    double t_frame;      // hydronium: frame.time of the frame being consumed
    // solo kernel (one CTA per replica): the state machine runs on the first warp, the other warps
    // join it for the per-frame and per-event array work; `leader` is the one thread with side
    // effects (lane 0 of the replica's warp in every kernel)
    bool solo, leader;
    int tid, nthr;
    unsigned *kbits;     // [mask_words] allowed bits of the last consumed frame in list order
    int *wpre;           // [mask_words] exclusive popc prefix of kbits
    int *si;             // [64] ints: results / warp totals
    double *sd;          // [64] doubles: results / warp totals
    unsigned *cpair;     // [x_cap] (start << 16) | dest of the compacted transitions
    unsigned *praw;      // [x_cap] the same for every pair of the frame, list order (staging)
    double *lsum2;       // second value buffer of the tree combine
    int *loff2, *ln2;    // second node buffer of the tree walk
    unsigned *tflags;    // [SOLO_LEVELS][tlw] "node was split" bits per tree level
    int *tcnt;           // [SOLO_LEVELS + 1] nodes per level, [SOLO_LEVELS + 1] = number of levels
    int tlw;
    // replay stream look-ahead: (uc0, uc1) = stream[uc_pos, uc_pos + 1], (un0, un1) the next pair
    double uc0, uc1, un0, un1;
    long long uc_pos;
    // Philox look-ahead: time selector and selection uniform of event nx_event, computed while the
    // selection of the previous event waits for its loads (kmc_move_fast)
    double nx_ts, nx_usel, nx_trial, cur_rate;
    long long nx_event;
    int scan_par;        // solo scans alternate between two sets of warp-total slots
    int p_next;          // pair count of the next frame, requested a frame ahead (-1: none)
#ifdef SOLO_PROFILE
    long long prof[16], prof_t;
#endif
};

#define SOLO_LEVELS 20

// -DSOLO_PROFILE: cycle counters of the first thread per phase, added to ties[2 + i] (tools only)
#ifdef SOLO_PROFILE
#define SOLO_T(i) do { const long long t_ = clock64(); if (c.tid == 0) c.prof[i] += t_ - c.prof_t; c.prof_t = t_; } while (0)
#else
#define SOLO_T(i) do { } while (0)
#endif

__host__ __device__ inline size_t solo_scratch_bytes(int64_t cap, int64_t leaves)
{
    const size_t lw = (size_t)(leaves + 31) / 32;
    size_t b = (size_t)cap * 24 + (size_t)leaves * 32 + (size_t)SOLO_LEVELS * lw * 4 + (SOLO_LEVELS + 4) * 4;
    return (b + 15) / 16 * 16;
}

// replay stream values of the event at `cursor`; the following event's pair is requested at the
// same time, so that its latency hides behind this event's work
__device__ __forceinline__ void replay_fetch(const KmcArgs &a, WarpCtx &c, int r, long long cursor)
{
    if (c.uc_pos == cursor) return;
    const double *u = a.u + (int64_t)r * a.n_u;
    if (c.uc_pos + 2 == cursor) { c.uc0 = c.un0; c.uc1 = c.un1; }
    else {
        c.uc0 = cursor < a.n_u ? u[cursor] : 0.0;
        c.uc1 = cursor + 1 < a.n_u ? u[cursor + 1] : 0.0;
    }
    c.uc_pos = cursor;
    c.un0 = cursor + 2 < a.n_u ? u[cursor + 2] : 0.0;
    c.un1 = cursor + 3 < a.n_u ? u[cursor + 3] : 0.0;
}


__device__ __forceinline__ bool occupied(const WarpCtx &c, int s) { return (c.occ[s >> 5] >> (s & 31)) & 1u; }


// ---- HydroniumTopology: rescaled distance of one transition --------------------------------------
// ReLUTransformation.__call__ (topology.py:286-290), InterpolatedTransformation.__call__ (:328-333,
// scipy interp1d kind="linear": slope * (x - x_lo) + y_lo), DistanceInterpolator.__call__ (:349-353)
// and transform_distances (:213-232), in NumPy's operation order without FMA contraction.
__device__ __forceinline__ double hyd_transform(const HydParams &h, double d)
{
    if (h.tkind == 1) {
        if (d <= h.tpar[3] || h.tpar[4] <= d) return d;
        return d < h.tpar[2] ? h.tpar[1] : __dadd_rn(__dmul_rn(h.tpar[0], __dadd_rn(d, -h.tpar[2])), h.tpar[1]);
    }
    if (h.tkind == 2) {
        const double x_min = h.tx[0], x_max = h.tx[h.nt - 1];
        if (!(x_min <= d && d <= x_max)) return d < x_min ? h.ty[0] : d;
        // np.searchsorted(x, d) (side='left'), clipped to [1, n - 1]
        int lo = 0, hi = h.nt;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (h.tx[mid] < d) lo = mid + 1; else hi = mid; }
        int idx = lo < 1 ? 1 : (lo > h.nt - 1 ? h.nt - 1 : lo);
        const double x_lo = h.tx[idx - 1], x_hi = h.tx[idx], y_lo = h.ty[idx - 1], y_hi = h.ty[idx];
        const double slope = __ddiv_rn(__dadd_rn(y_hi, -y_lo), __dadd_rn(x_hi, -x_lo));
        const double y = __dadd_rn(__dmul_rn(slope, __dadd_rn(d, -x_lo)), y_lo);
        return y < x_min ? h.ty[0] : y;
    }
    return d;
}

__device__ __forceinline__ double hyd_rate(const HydParams &h, const WarpCtx &c, int start, double d)
{
    const int proton = c.lat[start];
    const double tl = c.tlast[proton - 1];
    const double relaxed = hyd_transform(h, d);
    double dd = relaxed;
    if (h.relax > 0.0) {
        const double res = tl >= 0.0 ? __dadd_rn(c.t_frame, -tl) : INFINITY;
        const double ratio = fmin(__ddiv_rn(res, h.relax), 1.0);
        dd = __dadd_rn(__dmul_rn(__dadd_rn(1.0, -ratio), d), __dmul_rn(ratio, relaxed));
    }
    return rate_eval(h.rate, dd, 0.0);
}

// jumprate_generator + np.sum (MDMC.py:229-238, :85): total rate of the allowed transitions of
// frame f for this replica; records the allowed mask (remember_last_element, MDMC.py:83-84)
__device__ double kmc_consume(const KmcArgs &a, WarpCtx &c, int64_t f)
{
    const int p = a.counts[f];
    const int64_t base = f * a.stride;
    c.base = base;
    c.p = p;
    double s = 0.0;
    for (int k0 = 0; k0 < p; k0 += 32) {
        int k = k0 + c.lane;
        bool ok = false;
        if (k < p) {
            int st = __ldg(a.start + base + k), de = __ldg(a.dest + base + k);
            ok = occupied(c, st) && !occupied(c, de);
            if (ok) s += __ldg(a.omega + base + k);
        }
        unsigned bits = __ballot_sync(0xffffffffu, ok);
        if (c.lane == 0) c.mask0[k0 >> 5] = bits;
    }
    __syncwarp();
    return warp_sum(s);
}

// leaves of NumPy's pairwise recursion over m elements, left to right: (offset, length <= 128)
__device__ int np_sum_leaves(int m, int *loff, int *ln)
{
    int nleaf = 0;
    int so[40], sn[40], sp = 1;
    so[0] = 0; sn[0] = m;
    while (sp) {
        --sp;
        int off = so[sp], n = sn[sp];
        if (n <= 128) { loff[nleaf] = off; ln[nleaf] = n; nleaf++; }
        else {
            int n2 = n / 2;
            n2 -= n2 % 8;
            so[sp] = off + n2; sn[sp] = n - n2; sp++;   // right half below, left half on top
            so[sp] = off; sn[sp] = n2; sp++;
        }
    }
    return nleaf;
}

// the leaf sums combined along the same recursion tree: pairwise(left) + pairwise(right)
__device__ double np_sum_combine(int m, const double *lsum)
{
    double ret = 0.0;
    double val[40];
    int sn[40], sp = 1, li = 0;
    signed char ph[40];
    bool have = false;
    sn[0] = m; ph[0] = 0;
    while (sp > 0) {
        const int t = sp - 1;
        if (have) {
            if (ph[t] == 1) {  // left value arrived: descend into the right half
                val[t] = ret; ph[t] = 2; have = false;
                int n = sn[t], n2 = n / 2;
                n2 -= n2 % 8;
                sn[sp] = n - n2; ph[sp] = 0; sp++;
            } else {
                ret = __dadd_rn(val[t], ret);
                sp--;
            }
        } else {
            int n = sn[t];
            if (n <= 128) { ret = lsum[li++]; have = true; sp--; }
            else {
                ph[t] = 1;
                int n2 = n / 2;
                n2 -= n2 % 8;
                sn[sp] = n2; ph[sp] = 0; sp++;
            }
        }
    }
    return ret;
}

// ---- exact-replay arithmetic ------------------------------------------------------------------
// The reference's time stepping feeds rounding noise back into kmc_time with gain S_0/S_t per
// event (quirk Q1: the partial-frame interval always uses the total rate of frame 0), so a replay
// only stays bit-identical if every sum is formed in NumPy's own order:
//   np.sum   (MDMC.py:85)  -> pairwise summation, 8 accumulators per block of <= 128 elements,
//                             halves rounded down to a multiple of 8 (NumPy's add.reduce)
//   np.cumsum (MDMC.py:109) -> strictly sequential
// np_sum_warp evaluates exactly that tree over the compacted allowed rates c.comp[0..m).
__device__ double np_sum_warp(WarpCtx &c, int m)
{
    const double *a = c.comp;
    if (m < 8) {  // every lane forms the same sequential sum
        double res = 0.;
        for (int i = 0; i < m; i++) res = __dadd_rn(res, a[i]);
        return res;
    }
    // 1. leaves of the recursion, left to right (lane 0)
    int nleaf = 0;
    if (c.lane == 0) nleaf = np_sum_leaves(m, c.loff, c.ln);
    nleaf = __shfl_sync(0xffffffffu, nleaf, 0);
    __syncwarp();
    // 2. leaf sums: 8 lanes = the 8 accumulators of one leaf, 4 leaves per pass
    const int g = c.lane >> 3, j = c.lane & 7;
    for (int l0 = 0; l0 < nleaf; l0 += 4) {
        const int l = l0 + g;
        const bool act = l < nleaf;
        int off = 0, n = 0;
        if (act) { off = c.loff[l]; n = c.ln[l]; }
        double r = 0.0;
        if (act) {
            r = a[off + j];
            for (int i = 8; i < n - (n % 8); i += 8) r = __dadd_rn(r, a[off + i + j]);
        }
        // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); IEEE addition commutes, so every lane of the
        // group ends with the same bits
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
        if (act && j == 0) {
            for (int i = n - (n % 8); i < n; i++) r = __dadd_rn(r, a[off + i]);
            c.lsum[l] = r;
        }
    }
    __syncwarp();
    // 3. combine the leaves along the recursion tree (lane 0)
    double ret = 0.0;
    if (c.lane == 0) ret = np_sum_combine(m, c.lsum);
    return __shfl_sync(0xffffffffu, ret, 0);
}

// exact-replay form of kmc_consume: compacts the allowed transitions (list order) into the
// replica's scratch and sums them like np.sum
__device__ double kmc_consume_exact(const KmcArgs &a, WarpCtx &c, int64_t f)
{
    const int p = a.counts[f];
    const int64_t base = f * a.stride;
    c.base = base;
    c.p = p;
    int m = 0;
    // 128 pairs per trip: all loads of the four groups are issued before anything depends on them
    // (a verification run has one warp per SM and nothing else to hide the latency behind)
    for (int k0 = 0; k0 < p; k0 += 128) {
        int st[4], de[4];
        double om[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = k0 + 32 * u + c.lane;
            st[u] = k < p ? __ldg(a.start + base + k) : 0;
            de[u] = k < p ? __ldg(a.dest + base + k) : 0;
            om[u] = k < p ? __ldg((a.hyd.on ? a.dist : a.omega) + base + k) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = k0 + 32 * u + c.lane;
            ok[u] = k < p && occupied(c, st[u]) && !occupied(c, de[u]);
            // hydronium: the rate depends on this replica's lattice and jump times
            if (a.hyd.on && ok[u]) om[u] = hyd_rate(a.hyd, c, st[u], om[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned bits = __ballot_sync(0xffffffffu, ok[u]);
            if (ok[u]) {
                const int e = m + __popc(bits & ((1u << c.lane) - 1u));
                c.comp[e] = om[u];
                c.cidx[e] = k0 + 32 * u + c.lane;
            }
            m += __popc(bits);
        }
    }
    __syncwarp();
    c.m = m;
    return np_sum_warp(c, m);
}

// exact-replay form of kmc_move (MDMC.py:101-119): the arrays of the last consumed frame were
// filtered at consumption (c.comp / c.cidx); they are filtered again with the current lattice
// (it differs after a same-frame event, Q2), np.cumsum runs sequentially, draw = S*u,
// searchsorted(side='left').
__device__ bool kmc_move_exact(const KmcArgs &a, WarpCtx &c, double u, int *o_start, int *o_dest,
                               int *o_proton, int *o_index, unsigned long long *ties)
{
    const int m = c.m;
    const int64_t base = c.base;
    bool any = false;
    // 1. re-mask in parallel: cum[e] <- the rate if transition e is still allowed, else +0.0 (adding
    //    +0.0 leaves a running sum unchanged, like the entry being absent from np.cumsum's input)
    for (int e0 = 0; e0 < m; e0 += 128) {
        int st[4], de[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + 32 * u + c.lane;
            const int k = e < m ? c.cidx[e] : 0;
            st[u] = e < m ? __ldg(a.start + base + k) : 0;
            de[u] = e < m ? __ldg(a.dest + base + k) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + 32 * u + c.lane;
            const bool ok = e < m && occupied(c, st[u]) && !occupied(c, de[u]);
            if (e < m) c.cum[e] = ok ? c.comp[e] : 0.0;
            const unsigned okbits = __ballot_sync(0xffffffffu, ok);
            any |= okbits != 0u;
            if (c.lane == 0 && e0 + 32 * u < m) c.mask0[(e0 >> 5) + u] = okbits;
        }
    }
    __syncwarp();
    // 2. np.cumsum: strictly sequential additions, one lane, eight values per trip so that only
    //    the additions themselves are on the dependency chain
    double cum = 0.0;
    if (c.lane == 0) {
        int e = 0;
        for (; e + 8 <= m; e += 8) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = c.cum[e + q];
#pragma unroll
            for (int q = 0; q < 8; q++) { cum = __dadd_rn(cum, v[q]); v[q] = cum; }
#pragma unroll
            for (int q = 0; q < 8; q++) c.cum[e + q] = v[q];
        }
        for (; e < m; e++) { cum = __dadd_rn(cum, c.cum[e]); c.cum[e] = cum; }
    }
    cum = __shfl_sync(0xffffffffu, cum, 0);
    __syncwarp();
    if (!any) return false;  // empty cumsum: IndexError upstream
    const double total = cum;
    const double draw = __dadd_rn(0.0, __dmul_rn(__dadd_rn(total, -0.0), u));  // uniform(0, S)
    int found = -1;
    for (int e0 = 0; e0 < m && found < 0; e0 += 32) {
        int e = e0 + c.lane;
        bool ok = e < m && ((c.mask0[e0 >> 5] >> c.lane) & 1u);
        double cv = ok ? c.cum[e] : 0.0;
        unsigned hit = __ballot_sync(0xffffffffu, ok && cv >= draw);  // side='left'
        if (hit) {
            int l = __ffs(hit) - 1;
            found = e0 + l;
            double cl = __shfl_sync(0xffffffffu, cv, l);
            double prev = cl - c.comp[found];
            if (c.lane == 0 && (fabs(cl - draw) < 1e-9 * total || fabs(draw - prev) < 1e-9 * total))
                atomicAdd(ties, 1ull);
        }
    }
    if (found < 0) return false;  // cannot happen for u < 1 (draw <= cumsum[-1])
    const int k = c.cidx[found];
    int st = __ldg(a.start + base + k), de = __ldg(a.dest + base + k);
    int proton = c.lat[st];
    __syncwarp();
    if (c.lane == 0) {
        c.lat[de] = proton;
        c.lat[st] = 0;
        c.occ[de >> 5] |= 1u << (de & 31);
        c.occ[st >> 5] &= ~(1u << (st & 31));
    }
    __syncwarp();
    *o_start = st; *o_dest = de; *o_proton = proton; *o_index = k;
    return true;
}

// move_proton (MDMC.py:101-119) on the last consumed frame: re-mask, cumsum, draw = S*u,
// searchsorted(left), move the label.  Returns false when nothing is allowed (the reference
// raises IndexError there).
__device__ bool kmc_move(const KmcArgs &a, WarpCtx &c, double u, int *o_start, int *o_dest,
                         int *o_proton, int *o_index, unsigned long long *ties)
{
    const int p = c.p;
    const int64_t base = c.base;
    double s = 0.0;
    int n_allowed = 0;
    for (int k0 = 0; k0 < p; k0 += 32) {
        int k = k0 + c.lane;
        if (k < p && ((c.mask0[k0 >> 5] >> c.lane) & 1u)) {
            int st = __ldg(a.start + base + k), de = __ldg(a.dest + base + k);
            if (occupied(c, st) && !occupied(c, de)) { s += __ldg(a.omega + base + k); n_allowed++; }
        }
    }
    const double total = warp_sum(s);
    if (!__any_sync(0xffffffffu, n_allowed > 0)) return false;  // empty cumsum: IndexError upstream
    const double draw = 0.0 + (total - 0.0) * u;  // np.random.uniform(0, cumsum[-1])
    double running = 0.0;
    int found = -1, last_ok = -1;
    for (int k0 = 0; k0 < p && found < 0; k0 += 32) {
        int k = k0 + c.lane;
        double om = 0.0;
        bool ok = false;
        if (k < p && ((c.mask0[k0 >> 5] >> c.lane) & 1u)) {
            int st = __ldg(a.start + base + k), de = __ldg(a.dest + base + k);
            ok = occupied(c, st) && !occupied(c, de);
            if (ok) om = __ldg(a.omega + base + k);
        }
        double inc = om;  // inclusive scan across the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(0xffffffffu, inc, o);
            if (c.lane >= o) inc += t;
        }
        double cum = running + inc;
        unsigned okbits = __ballot_sync(0xffffffffu, ok);
        unsigned hit = __ballot_sync(0xffffffffu, ok && cum >= draw);  // side='left'
        if (hit) {
            int l = __ffs(hit) - 1;
            found = k0 + l;
            double cl = __shfl_sync(0xffffffffu, cum, l);
            double prev = cl - __shfl_sync(0xffffffffu, om, l);
            if (c.lane == 0 && (fabs(cl - draw) < 1e-9 * total || fabs(draw - prev) < 1e-9 * total))
                atomicAdd(ties, 1ull);
        }
        if (okbits) last_ok = k0 + 31 - __clz(okbits);
        running += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (found < 0) {
        // draw beyond the parallel running total by rounding only: the sequential cumsum's last
        // entry is the reference's own draw bound -> take the last allowed transition
        if (last_ok < 0) return false;
        found = last_ok;
        if (c.lane == 0) atomicAdd(ties, 1ull);
    }
    int st = __ldg(a.start + base + found), de = __ldg(a.dest + base + found);
    int proton = c.lat[st];
    __syncwarp();
    if (c.lane == 0) {
        c.lat[de] = proton;
        c.lat[st] = 0;
        c.occ[de >> 5] |= 1u << (de & 31);
        c.occ[st >> 5] &= ~(1u << (st & 31));
    }
    __syncwarp();
    *o_start = st; *o_dest = de; *o_proton = proton; *o_index = found;
    return true;
}


// ---- streaming (Philox) flavour ------------------------------------------------------------------
// The frame's (start, dest, omega) arrays come through a TMA-fed shared-memory ring shared by all
// replicas of the CTA; a lane accumulates the allowed rates of ITS pairs (k mod 32 == lane) per
// 1024-pair stage without any shuffle.  Selection walks the transitions in (lane, stage, i) order
// instead of list order -- the probability of picking transition k is omega_k / S either way.
__device__ __forceinline__ void kmc_consume_stage(WarpCtx &c, const int *__restrict__ s_start,
                                                  const int *__restrict__ s_dest,
                                                  const double *__restrict__ s_omega, int k0, int cnt,
                                                  int stage)
{
    double acc = 0.0;
    const int full = cnt & ~31;
    int i0 = 0;
    // 128 pairs per trip, software-pipelined by hand: all loads of the four groups first, then the
    // occupancy look-ups, then the votes; no branch, one 16-byte store of the four mask words
    for (; i0 + 128 <= full; i0 += 128) {
        int st[4], de[4];
        double om[4];
        unsigned os[4], od[4], bits[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = i0 + 32 * u + c.lane;
            st[u] = s_start[k]; de[u] = s_dest[k]; om[u] = s_omega[k];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) { os[u] = c.occ[st[u] >> 5]; od[u] = c.occ[de[u] >> 5]; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool ok = ((os[u] >> (st[u] & 31)) & ~(od[u] >> (de[u] & 31)) & 1u) != 0u;
            acc += ok ? om[u] : 0.0;
            bits[u] = __ballot_sync(0xffffffffu, ok);
        }
        if (c.lane == 0)
            *(uint4 *)(c.mask0 + ((k0 + i0) >> 5)) = make_uint4(bits[0], bits[1], bits[2], bits[3]);
    }
    for (; i0 < full; i0 += 32) {
        const int k = i0 + c.lane;
        const int st = s_start[k], de = s_dest[k];
        const double om = s_omega[k];
        const bool ok = ((c.occ[st >> 5] >> (st & 31)) & ~(c.occ[de >> 5] >> (de & 31)) & 1u) != 0u;
        acc += ok ? om : 0.0;
        const unsigned bits = __ballot_sync(0xffffffffu, ok);
        if (c.lane == 0) c.mask0[(k0 + i0) >> 5] = bits;
    }
    if (full < cnt) {   // ragged tail of the frame (entries behind cnt are not initialised)
        const int k = full + c.lane;
        bool ok = false;
        if (k < cnt) {
            const int st = s_start[k], de = s_dest[k];
            ok = occupied(c, st) && !occupied(c, de);
            if (ok) acc += s_omega[k];
        }
        const unsigned bits = __ballot_sync(0xffffffffu, ok);
        if (c.lane == 0) c.mask0[(k0 + full) >> 5] = bits;
    }
    c.psum[stage * 32 + c.lane] = acc;
    c.lane_total += acc;
}

// Selection for the streaming kernel.  The partial sums describe the transitions allowed when the
// frame was consumed.  A same-frame follow-up event must choose among those that are STILL allowed
// (Q2: the reference re-filters the filtered arrays, so transitions only ever leave the set).
// Instead of maintaining the sums, a transition is drawn from the consumed set with probability
// omega_k / S and redrawn (fresh Philox numbers) when it is no longer allowed -- which leaves
// exactly the distribution omega_k / S' over the still-allowed ones.  After 24 rejections the
// exact masked scan over the global arrays (kmc_move) takes over.
__device__ bool kmc_move_fast(const KmcArgs &a, WarpCtx &c, double u, int r, long long event,
                              int *o_start, int *o_dest, int *o_proton, int *o_index,
                              unsigned long long *ties)
{
    // lane totals -> inclusive scan over lanes; the last value is the total the draw refers to.
    // The totals belong to the frame, not to the event: the scan is kept for the frame's further
    // events (same values, same bits).
    const double lt = c.lane_total;   // == sum over the frame's stages of psum[s][lane], same order
    if (c.scan_base != c.base) {
        double v = lt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, v, o);
            if (c.lane >= o) v += t;
        }
        c.scan_inc = v;
        c.scan_base = c.base;
    }
    const double inc = c.scan_inc;
    const double total = __shfl_sync(0xffffffffu, inc, 31);
    const unsigned have = __ballot_sync(0xffffffffu, lt > 0.0);
    if (!have || !(total > 0.0)) return false;   // nothing was allowed: IndexError upstream
    for (int attempt = 0; attempt < 24; attempt++) {
        if (attempt > 0) {
            uint32_t ctr[4] = {(uint32_t)event, (uint32_t)((uint64_t)event >> 32),
                               (uint32_t)(a.replica_first + r * a.replica_step), (uint32_t)attempt};
            philox4x32_10(ctr, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
            u = u53(ctr[0], ctr[1]);
        }
        const double draw = total * u;
        const unsigned hit = __ballot_sync(0xffffffffu, lt > 0.0 && inc >= draw);
        const int L = hit ? __ffs(hit) - 1 : 31 - __clz(have);
        double run = __shfl_sync(0xffffffffu, inc - lt, L);   // rate in front of lane L's pairs
        // lane L's stages, in order (every lane walks the same values)
        int sfound = -1;
        double run_before = run;
        for (int s = 0; s < c.nst; s++) {
            const double v = c.psum[s * 32 + L];
            if (v > 0.0) {
                sfound = s;
                run_before = run;
                if (run + v >= draw) break;
                run += v;
            }
        }
        if (sfound < 0) return false;
        run = run_before;
        // its <= 32 pairs k = sfound*1024 + i*32 + L, one per lane
        const int k = sfound * 1024 + c.lane * 32 + L;
        const bool ok = k < c.p && ((c.mask0[k >> 5] >> L) & 1u);
        const double om = ok ? __ldg(a.omega + c.base + k) : 0.0;
        // the pair itself comes along with its rate: one trip to L2 per attempt instead of two
        const int st_l = ok ? __ldg(a.start + c.base + k) : 0, de_l = ok ? __ldg(a.dest + c.base + k) : 0;
        if (attempt == 0) {
            // while the loads are under way: the NEXT event's Philox draw and its logarithm
            uint32_t ctr[4] = {(uint32_t)(event + 1), (uint32_t)((uint64_t)(event + 1) >> 32),
                               (uint32_t)(a.replica_first + r * a.replica_step), 0u};
            philox4x32_10(ctr, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
            c.nx_ts = -log(1 - u53(ctr[0], ctr[1]));
            c.nx_usel = u53(ctr[2], ctr[3]);
            c.nx_trial = c.nx_ts / c.cur_rate;
            c.nx_event = event + 1;
        }
        double inc2 = om;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, inc2, o);
            if (c.lane >= o) inc2 += t;
        }
        const unsigned okbits = __ballot_sync(0xffffffffu, ok);
        if (!okbits) return false;
        const unsigned hit2 = __ballot_sync(0xffffffffu, ok && run + inc2 >= draw);
        int isel;
        if (hit2) isel = __ffs(hit2) - 1;
        else {   // partial sums and the scan round differently: the last allowed pair of the stage
            isel = 31 - __clz(okbits);
            if (c.lane == 0) atomicAdd(ties, 1ull);
        }
        const int found = sfound * 1024 + isel * 32 + L;
        const int st = __shfl_sync(0xffffffffu, st_l, isel), de = __shfl_sync(0xffffffffu, de_l, isel);
        if (!(occupied(c, st) && !occupied(c, de))) continue;   // left the set since consumption
        const int proton = c.lat[st];
        __syncwarp();
        if (c.lane == 0) {
            c.lat[de] = proton;
            c.lat[st] = 0;
            c.occ[de >> 5] |= 1u << (de & 31);
            c.occ[st >> 5] &= ~(1u << (st & 31));
        }
        __syncwarp();
        *o_start = st; *o_dest = de; *o_proton = proton; *o_index = found;
        return true;
    }
    return kmc_move(a, c, u, o_start, o_dest, o_proton, o_index, ties);
}


// ---- solo mode: ONE CTA PER REPLICA (exact arithmetic, few replicas: `mdmc`, verification runs) ---
// A single replica is sequential in time, so the only parallelism is inside a frame.  The first
// warp runs the scalar state machine (kmc_after_consume / kmc_run_until_frame_needed, as in the
// warp-per-replica kernels); all warps share the array work of a frame (solo_consume) and of an
// event (solo_move), the helper warps waiting in solo_helpers for the first warp's requests.
//   consume  allowed bits in list order -> popc prefix -> compaction (the arrays jumprate_generator
//            yields, MDMC.py:229-238) -> np.sum's pairwise tree with one 8-lane group per leaf
//   move     np.cumsum is strictly sequential in the reference, but move_proton (MDMC.py:101-119)
//            only uses the INDEX searchsorted returns.  The index is decided on a parallel prefix
//            sum whenever the draw is further than the worst-case rounding distance between any
//            two summation orders from the interval's ends (margin = sel_margin * m * S, with
//            sel_margin = 2^-50 against a bound of 2 m 2^-53 S per order); otherwise -- once in
//            ~1e8 events -- the leader repeats the reference's sequential sum (counted in
//            cmd_kmc_selection_fallbacks).  The decision is the reference's either way.

// exclusive CTA-wide scans (one barrier; the warp-total slots alternate between two sets); every
// thread obtains the same total
__device__ __forceinline__ int solo_scan_int(WarpCtx &c, int v, int *total)
{
    const int wid = c.tid >> 5, nw = c.nthr >> 5;
    int *slot = c.si + 16 + 16 * (c.scan_par & 1);
    c.scan_par++;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (c.lane >= o) inc += t;
    }
    if (c.lane == 31) slot[wid] = inc;
    __syncthreads();
    const int x = c.lane < nw ? slot[c.lane] : 0;
    int winc = x;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (c.lane >= o) winc += t;
    }
    *total = __shfl_sync(0xffffffffu, winc, 15);
    return __shfl_sync(0xffffffffu, winc - x, wid) + inc - v;
}

__device__ __forceinline__ double solo_scan_double(WarpCtx &c, double v, double *total)
{
    const int wid = c.tid >> 5, nw = c.nthr >> 5;
    double *slot = c.sd + 16 + 16 * (c.scan_par & 1);
    c.scan_par++;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (c.lane >= o) inc += t;
    }
    double excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (c.lane == 0) excl = 0.0;
    if (c.lane == 31) slot[wid] = inc;
    __syncthreads();
    const double x = c.lane < nw ? slot[c.lane] : 0.0;
    double winc = x;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, winc, o);
        if (c.lane >= o) winc += t;
    }
    *total = __shfl_sync(0xffffffffu, winc, 15);
    return __shfl_sync(0xffffffffu, winc - x, wid) + excl;
}

// np.sum's recursion tree over m elements, built breadth first by ONE WARP: a node (offset, n)
// with n > 128 splits into (offset, n2) and (offset + n2, n - n2), n2 = n / 2 rounded down to a
// multiple of 8 (NumPy's pairwise_sum).  Level l keeps its "was split" bits in tflags[l] and its
// node count in tcnt[l]; the leaves end up, left to right, in (loff, ln).  Returns their number.
__device__ int np_tree_build_warp(WarpCtx &c, int m)
{
    const unsigned lt = (1u << c.lane) - 1u;
    int *off_a = c.loff, *n_a = c.ln, *off_b = c.loff2, *n_b = c.ln2;
    if (c.lane == 0) { off_a[0] = 0; n_a[0] = m; }
    __syncwarp();
    int count = 1, level = 0;
    for (;; level++) {
        int carry = 0;
        for (int i0 = 0; i0 < count; i0 += 32) {
            const int i = i0 + c.lane;
            const bool valid = i < count;
            const int off = valid ? off_a[i] : 0, n = valid ? n_a[i] : 0;
            const bool split = n > 128 && level < SOLO_LEVELS;
            const unsigned bits = __ballot_sync(0xffffffffu, split);
            if (c.lane == 0) c.tflags[level * c.tlw + (i0 >> 5)] = bits;
            const int pos = i + carry + __popc(bits & lt);
            if (valid) {
                if (split) {
                    const int n2 = (n >> 1) & ~7;
                    off_b[pos] = off; n_b[pos] = n2;
                    off_b[pos + 1] = off + n2; n_b[pos + 1] = n - n2;
                } else {
                    off_b[pos] = off; n_b[pos] = n;
                }
            }
            carry += __popc(bits);
        }
        if (c.lane == 0) c.tcnt[level] = count;
        __syncwarp();
        if (carry == 0) break;   // (off_a, n_a) holds the leaves
        count += carry;
        int *t = off_a; off_a = off_b; off_b = t;
        t = n_a; n_a = n_b; n_b = t;
    }
    if (off_a != c.loff) {
        for (int i = c.lane; i < count; i += 32) { c.loff[i] = off_a[i]; c.ln[i] = n_a[i]; }
    }
    if (c.lane == 0) c.tcnt[SOLO_LEVELS + 1] = level;   // tcnt[level] == count == number of leaves
    __syncwarp();
    return count;
}

// the leaf sums (c.lsum) combined along that tree, deepest level first: pairwise(left) + pairwise(right)
__device__ double np_tree_combine_warp(WarpCtx &c)
{
    const unsigned lt = (1u << c.lane) - 1u;
    const int levels = c.tcnt[SOLO_LEVELS + 1];
    double *va = c.lsum, *vb = c.lsum2;
    for (int level = levels - 1; level >= 0; level--) {
        const int count = c.tcnt[level];
        int carry = 0;
        for (int i0 = 0; i0 < count; i0 += 32) {
            const int i = i0 + c.lane;
            const unsigned bits = c.tflags[level * c.tlw + (i0 >> 5)];
            const int pos = i + carry + __popc(bits & lt);
            if (i < count) vb[i] = (bits >> c.lane) & 1u ? __dadd_rn(va[pos], va[pos + 1]) : va[pos];
            carry += __popc(bits);
        }
        __syncwarp();
        double *t = va; va = vb; vb = t;
    }
    return va[0];
}

__device__ double solo_consume(const KmcArgs &a, WarpCtx &c, int64_t f)
{
    SOLO_T(0);   // frame bookkeeping / observables
    const int p = c.p_next >= 0 ? c.p_next : a.counts[f];
    c.p_next = f + 1 < a.nframes ? a.counts[f + 1] : -1;   // in flight until the next frame needs it
    const int64_t base = f * a.stride;
    const int T = c.nthr, tid = c.tid;
    c.base = base;
    c.p = p;
    const double *val = (a.hyd.on ? a.dist : a.omega) + base;
    if (f + 1 < a.nframes) {
        // the next frame's arrays on their way into L2 while this frame is worked on (its pair count
        // is about this one's; a line too many or too few does not matter)
        const int *ns = a.start + base + a.stride, *nd = a.dest + base + a.stride;
        const double *nv = val + a.stride;
        for (int o = tid * 32; o < p; o += T * 32) {
            asm volatile("prefetch.global.L2 [%0];" :: "l"(ns + o));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(nd + o));
        }
        for (int o = tid * 16; o < p; o += T * 16) asm volatile("prefetch.global.L2 [%0];" :: "l"(nv + o));
    }
    // 1. allowed bits in list order, eight independent pairs per thread and trip; rate and pair code
    //    wait in shared memory (c.cum is idle between moves) for their compacted position
    for (int k0 = 0; k0 < p; k0 += 8 * T) {
        int st[8], de[8];
        double om[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int k = k0 + u * T + tid;
            st[u] = k < p ? __ldg(a.start + base + k) : 0;
            de[u] = k < p ? __ldg(a.dest + base + k) : 0;
            om[u] = k < p ? __ldg(val + k) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int k = k0 + u * T + tid;
            if (k0 + u * T < p) {   // uniform over the CTA
                const bool ok = k < p && occupied(c, st[u]) && !occupied(c, de[u]);
                const unsigned bits = __ballot_sync(0xffffffffu, ok);
                if (c.lane == 0 && k < p) c.kbits[k >> 5] = bits;
                if (ok) {
                    c.cum[k] = a.hyd.on ? hyd_rate(a.hyd, c, st[u], om[u]) : om[u];
                    c.praw[k] = ((unsigned)st[u] << 16) | (unsigned)de[u];
                }
            }
        }
    }
    __syncthreads();
    SOLO_T(1);   // pass A
    // 2. compacted position of every allowed transition
    const int nwords = (p + 31) >> 5;
    int m = 0;
    for (int w0 = 0; w0 < nwords; w0 += T) {
        const int wi = w0 + tid;
        const int v = wi < nwords ? __popc(c.kbits[wi]) : 0;
        int tot;
        const int ex = solo_scan_int(c, v, &tot);
        if (wi < nwords) c.wpre[wi] = m + ex;
        m += tot;
    }
    c.m = m;
    __syncthreads();
    SOLO_T(2);   // scan
    // 3. the last warp lays out np.sum's tree while the others compact (rate, pair) in list order:
    //    the arrays jumprate_generator yields (MDMC.py:229-238)
    if (tid >= T - 32) {
        if (m >= 8) {
            const int nl = np_tree_build_warp(c, m);
            if (c.lane == 0) c.si[1] = nl;
        }
    } else {
        const unsigned lt = (1u << c.lane) - 1u;
        for (int k = tid; k < p; k += T - 32) {
            const unsigned bits = c.kbits[k >> 5];
            if ((bits >> c.lane) & 1u) {
                const int e = c.wpre[k >> 5] + __popc(bits & lt);
                c.comp[e] = c.cum[k];
                c.cpair[e] = c.praw[k];
            }
        }
    }
    __syncthreads();
    SOLO_T(3);   // tree + compaction
    if (m < 8) {  // np.sum of a short array: plain loop
        double res = 0.;
        for (int i = 0; i < m; i++) res = __dadd_rn(res, c.comp[i]);
        return res;
    }
    // 4. one 8-lane group per leaf: the 8 accumulators of NumPy's unrolled loop
    const int nleaf = c.si[1];
    const int g = tid >> 3, j = tid & 7, ng = T >> 3;
    const double *x = c.comp;
    for (int l0 = 0; l0 < nleaf; l0 += ng) {
        const int l = l0 + g;
        const bool act = l < nleaf;
        int off = 0, n = 0;
        if (act) { off = c.loff[l]; n = c.ln[l]; }
        double r = 0.0;
        if (act) {
            r = x[off + j];
            for (int i = 8; i < n - (n % 8); i += 8) r = __dadd_rn(r, x[off + i + j]);
        }
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
        if (act && j == 0) {
            for (int i = n - (n % 8); i < n; i++) r = __dadd_rn(r, x[off + i]);
            c.lsum[l] = r;
        }
    }
    __syncthreads();
    SOLO_T(4);   // leaf sums
    if (tid < 32) {
        const double tot = np_tree_combine_warp(c);
        if (tid == 0) c.sd[0] = tot;
    }
    __syncthreads();
    SOLO_T(5);   // combine
    return c.sd[0];
}

// list index k of the e-th allowed transition of the last consumed frame (kbits / wpre)
__device__ int solo_list_index(const WarpCtx &c, int e)
{
    int lo = 0, hi = (c.p + 31) >> 5;   // last word with wpre <= e
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (c.wpre[mid] <= e) lo = mid; else hi = mid;
    }
    unsigned bits = c.kbits[lo];
    for (int q = e - c.wpre[lo]; q > 0; q--) bits &= bits - 1;
    return lo * 32 + __ffs(bits) - 1;
}

#define SOLO_PER 8   // compacted transitions per thread held in registers during a move

__device__ bool solo_move(const KmcArgs &a, WarpCtx &c, double u, int64_t slot, int *o_start,
                          int *o_dest, int *o_proton, int *o_index, unsigned long long *ties)
{
    const int m = c.m, T = c.nthr, tid = c.tid;
    SOLO_T(6);   // scalar state machine up to the move
    if (tid == 0) { c.si[2] = 0; c.si[3] = -1; c.si[4] = 0; }
    // 1. filter_allowed_transitions with the current lattice (it differs after a same-frame event):
    //    thread-contiguous runs of the compacted list, the re-masked rates stay in registers
    const int per = (m + T - 1) / T;
    const int lo = min(tid * per, m), hi = min(lo + per, m);
    const bool in_regs = per <= SOLO_PER;
    double x[SOLO_PER];
    double loc = 0.0;
    if (in_regs) {
#pragma unroll
        for (int i = 0; i < SOLO_PER; i++) {
            const int e = lo + i;
            x[i] = 0.0;
            if (e < hi) {
                const unsigned pr = c.cpair[e];
                if (occupied(c, pr >> 16) && !occupied(c, pr & 0xffff)) x[i] = c.comp[e];
            }
            loc += x[i];
        }
    } else {
        for (int e = lo; e < hi; e++) {
            const unsigned pr = c.cpair[e];
            loc += occupied(c, pr >> 16) && !occupied(c, pr & 0xffff) ? c.comp[e] : 0.0;
        }
    }
    SOLO_T(12);  // move: re-mask
    // 2. parallel prefix
    double total;
    const double pre = solo_scan_double(c, loc, &total);
    SOLO_T(13);  // move: prefix scan
    const double draw = total * u;
    const double eps = a.sel_margin * (double)m * total;
    // 3. searchsorted(side='left'): the first entry whose running sum reaches the draw
    double run = pre;
    if (in_regs) {
#pragma unroll
        for (int i = 0; i < SOLO_PER; i++) {
            const double prev = run;
            run += x[i];
            if (prev < draw && draw <= run) {
                atomicAdd(&c.si[2], 1);
                c.si[3] = lo + i;
                c.si[4] = (draw - prev > eps && run - draw > eps) ? 1 : 0;
                if (run - draw < 1e-9 * total || draw - prev < 1e-9 * total) atomicAdd(ties, 1ull);
            }
        }
    } else {
        for (int e = lo; e < hi; e++) {
            const double prev = run;
            const unsigned pr = c.cpair[e];
            run += occupied(c, pr >> 16) && !occupied(c, pr & 0xffff) ? c.comp[e] : 0.0;
            if (prev < draw && draw <= run) {
                atomicAdd(&c.si[2], 1);
                c.si[3] = e;
                c.si[4] = (draw - prev > eps && run - draw > eps) ? 1 : 0;
                if (run - draw < 1e-9 * total || draw - prev < 1e-9 * total) atomicAdd(ties, 1ull);
            }
        }
    }
    __syncthreads();
    SOLO_T(7);   // move: re-mask + scan + search
    int found = c.si[3];
    if (c.si[2] != 1 || !c.si[4]) {
        // too close to an interval end for any summation order but the reference's own (or nothing is
        // allowed, or the total is zero): the leader repeats move_proton literally -- filter, strictly
        // sequential np.cumsum, draw = uniform(0, S), searchsorted(side='left')
        __syncthreads();
        if (tid == 0) {
            double cum = 0.0;
            int any = 0;
            for (int e = 0; e < m; e++) {
                const unsigned pr = c.cpair[e];
                if (occupied(c, pr >> 16) && !occupied(c, pr & 0xffff)) { cum = __dadd_rn(cum, c.comp[e]); any = 1; }
            }
            int fnd = -1;
            if (any) {
                const double d = __dadd_rn(0.0, __dmul_rn(__dadd_rn(cum, -0.0), u));   // uniform(0, S)
                double run2 = 0.0;
                for (int e = 0; e < m && fnd < 0; e++) {
                    const unsigned pr = c.cpair[e];
                    if (occupied(c, pr >> 16) && !occupied(c, pr & 0xffff)) {
                        run2 = __dadd_rn(run2, c.comp[e]);
                        if (run2 >= d) fnd = e;
                    }
                }
                atomicAdd(ties + 1, 1ull);
            }
            c.si[3] = fnd;
        }
        __syncthreads();
        found = c.si[3];
    }
    if (found < 0) return false;  // nothing allowed: IndexError upstream
    const unsigned pr = c.cpair[found];
    const int st = pr >> 16, de = pr & 0xffff;
    // Nobody reads the lattice between the barrier above and the next barrier every thread takes
    // (the first warp's next request / end of frame), so the leader moves the label right away and
    // no barrier closes the move.  The proton label is the leader's alone.
    int proton = 0;
    if (tid == 0) {
        proton = c.lat[st];
        c.lat[de] = proton;
        c.lat[st] = 0;
        c.occ[de >> 5] |= 1u << (de & 31);
        c.occ[st >> 5] &= ~(1u << (st & 31));
    }
    if (tid == 32 && slot >= 0)   // a helper logs the pair's flat index (see kmc_event)
        ((long long *)a.ev_dist)[slot] = c.base + solo_list_index(c, found);
    SOLO_T(8);   // move: tail
    *o_start = st; *o_dest = de; *o_proton = proton; *o_index = 0;
    return true;
}

// first warp: wake the helper warps for one event, then take part in it
__device__ bool solo_move_request(const KmcArgs &a, WarpCtx &c, double u, int64_t slot, int *o_start,
                                  int *o_dest, int *o_proton, int *o_index, unsigned long long *ties)
{
    if (c.lane == 0) { c.si[8] = 1; c.sd[1] = u; ((long long *)c.sd)[2] = slot; }
    __syncthreads();
    return solo_move(a, c, u, slot, o_start, o_dest, o_proton, o_index, ties);
}

// helper warps: serve the first warp's events until it is done with the frame; true = replica halted
__device__ bool solo_helpers(const KmcArgs &a, WarpCtx &c)
{
    for (;;) {
        __syncthreads();
        const int cmd = c.si[8];
        if (cmd != 1) return cmd == 2;
        int d0, d1, d2, d3;
        solo_move(a, c, c.sd[1], ((long long *)c.sd)[2], &d0, &d1, &d2, &d3, a.ties);
    }
}

// observables on a consumed frame (MDMC.py:198-208 + output.py); the lattice a frame is seen
// with is the lattice at consumption == the pre-jump lattice at the flush (Q4)
__device__ void kmc_observe(const KmcArgs &a, const BoxParams &bx, WarpCtx &c, int r, int64_t f,
                            KmcState &st)
{
    const int64_t gf = a.obs_base + f;
    const double *pos = a.positions + f * (int64_t)a.n_sites * 3;
    double *snap = a.snapshot + (int64_t)r * a.n_sites * 3;
    double *disp = a.disp + (int64_t)r * a.n_sites * 3;
    int *lat0 = a.lattice0 + (int64_t)r * a.n_sites;
    if (gf == 0) {  // MDMC.py:193-196: first frame initialises both observables
        for (int s = c.lane; s < a.n_sites; s += 32) {
            int l = c.lat[s];
            lat0[s] = l;
            if (l > 0) {
                for (int k = 0; k < 3; k++) { snap[3 * (l - 1) + k] = pos[3 * s + k]; disp[3 * (l - 1) + k] = 0.0; }
            }
        }
        __syncwarp();
        return;
    }
    const bool reset = a.reset_freq > 0 && (gf % a.reset_freq) == 0;
    double m[3] = {0, 0, 0};
    int nprot = 0, same = 0;
    for (int s = c.lane; s < a.n_sites; s += 32) {
        int l = c.lat[s];
        if (reset) lat0[s] = l;
        if (l > 0) {
            nprot++;
            double pa[3], pb[3], d[3];
            for (int k = 0; k < 3; k++) { pa[k] = snap[3 * (l - 1) + k]; pb[k] = pos[3 * s + k]; }
            distance_exact(bx, pa, pb, d);  // output.py:41: atombox.distance(snapshot, new)
            for (int k = 0; k < 3; k++) {
                double v = __dadd_rn(reset ? 0.0 : disp[3 * (l - 1) + k], d[k]);
                disp[3 * (l - 1) + k] = v;
                snap[3 * (l - 1) + k] = pb[k];
                m[k] += v * v;
            }
            same += (lat0[s] == l);
        }
    }
    __syncwarp();
    if (a.print_freq > 0 && (gf % a.print_freq) == 0) {
        for (int k = 0; k < 3; k++) m[k] = warp_sum(m[k]);
        for (int o = 16; o > 0; o >>= 1) {
            nprot += __shfl_xor_sync(0xffffffffu, nprot, o);
            same += __shfl_xor_sync(0xffffffffu, same, o);
        }
        if (c.lane == 0 && st.n_rows < a.row_cap) {
            double *row = a.rows + ((int64_t)r * a.row_cap + st.n_rows) * 6;
            row[0] = (double)gf;
            row[1] = nan("");
            row[2] = m[0] / nprot; row[3] = m[1] / nprot; row[4] = m[2] / nprot;
            row[5] = (double)same;
        }
        if (st.n_rows < a.row_cap) st.n_rows++;
    }
}

// kmc_observe for the solo kernel: the sites are shared by the whole CTA (one warp walking 400 sites
// through global memory every frame would cost as much as the frame's KMC work).  The per-site
// arithmetic is kmc_observe's; the MSD sums of a printed row are formed in a different (fixed) order.
__device__ void solo_observe(const KmcArgs &a, const BoxParams &bx, WarpCtx &c, int r, int64_t f,
                             KmcState &st)
{
    const int64_t gf = a.obs_base + f;
    const double *pos = a.positions + f * (int64_t)a.n_sites * 3;
    double *snap = a.snapshot + (int64_t)r * a.n_sites * 3;
    double *disp = a.disp + (int64_t)r * a.n_sites * 3;
    int *lat0 = a.lattice0 + (int64_t)r * a.n_sites;
    if (f + 1 < a.nframes)   // next frame's positions towards L2
        for (int o = c.tid * 16; o < 3 * a.n_sites; o += c.nthr * 16)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(pos + (int64_t)a.n_sites * 3 + o));
    if (gf == 0) {  // MDMC.py:193-196: first frame initialises both observables
        for (int s = c.tid; s < a.n_sites; s += c.nthr) {
            const int l = c.lat[s];
            lat0[s] = l;
            if (l > 0)
                for (int k = 0; k < 3; k++) { snap[3 * (l - 1) + k] = pos[3 * s + k]; disp[3 * (l - 1) + k] = 0.0; }
        }
        return;
    }
    const bool reset = a.reset_freq > 0 && (gf % a.reset_freq) == 0;
    const bool printing = a.print_freq > 0 && (gf % a.print_freq) == 0;
    double m[3] = {0, 0, 0};
    int nprot = 0, same = 0;
    for (int s = c.tid; s < a.n_sites; s += c.nthr) {
        const int l = c.lat[s];
        if (reset) lat0[s] = l;
        if (l > 0) {
            nprot++;
            double pa[3], pb[3], d[3];
            for (int k = 0; k < 3; k++) { pa[k] = snap[3 * (l - 1) + k]; pb[k] = pos[3 * s + k]; }
            distance_exact(bx, pa, pb, d);  // output.py:41: atombox.distance(snapshot, new)
            for (int k = 0; k < 3; k++) {
                const double v = __dadd_rn(reset ? 0.0 : disp[3 * (l - 1) + k], d[k]);
                disp[3 * (l - 1) + k] = v;
                snap[3 * (l - 1) + k] = pb[k];
                m[k] += v * v;
            }
            same += (lat0[s] == l);
        }
    }
    if (!printing) return;
    // block sums: lanes, then the warps in order (two of the scan's slot sets are free here)
    for (int k = 0; k < 3; k++) m[k] = warp_sum(m[k]);
    for (int o = 16; o > 0; o >>= 1) {
        nprot += __shfl_xor_sync(0xffffffffu, nprot, o);
        same += __shfl_xor_sync(0xffffffffu, same, o);
    }
    const int wid = c.tid >> 5, nw = c.nthr >> 5;
    __syncthreads();
    if (c.lane == 0) {
        c.sd[16 + wid] = m[0]; c.sd[32 + wid] = m[1]; c.sd[48 + wid] = m[2];
        c.si[16 + wid] = nprot; c.si[32 + wid] = same;
    }
    __syncthreads();
    if (c.tid < 32) {
        double t[3] = {0, 0, 0};
        int np = 0, sm = 0;
        for (int w = 0; w < nw; w++) {
            t[0] += c.sd[16 + w]; t[1] += c.sd[32 + w]; t[2] += c.sd[48 + w];
            np += c.si[16 + w]; sm += c.si[32 + w];
        }
        if (c.leader && st.n_rows < a.row_cap) {
            double *row = a.rows + ((int64_t)r * a.row_cap + st.n_rows) * 6;
            row[0] = (double)gf;
            row[1] = nan("");
            row[2] = t[0] / np; row[3] = t[1] / np; row[4] = t[2] / np;
            row[5] = (double)sm;
        }
        if (st.n_rows < a.row_cap) st.n_rows++;
    }
    __syncthreads();
}

// event bookkeeping shared by both branches of fastforward_to_next_jump: stamp the cached
// frames' rows with the event time (Q3), select + move, log.  Returns false if the replica halts.
__device__ bool kmc_event(const KmcArgs &a, WarpCtx &c, int r, KmcState &st)
{
    if (c.leader)
        for (long long q = st.pending_row; q < st.n_rows; q++)
            a.rows[((int64_t)r * a.row_cap + q) * 6 + 1] = st.kmc_time;
    st.pending_row = st.n_rows;
    double u;
    if (a.rng_mode == CMD_RNG_REPLAY) {
        if (st.cursor + 1 >= a.n_u) { st.reason = 1; return false; }
        replay_fetch(a, c, r, st.cursor);
        u = c.uc1;
    } else {
        u = st.u_sel;   // drawn with the time selector of this event (same Philox counter)
    }
    int es, ed, ep, ek;
    c.cur_rate = st.current_rate;   // for the look-ahead of the next event's trial time
    // the jump distance is logged as the flat index of the pair; k_resolve_ev_dist turns it into the
    // distance after the kernel (a dependent global load here would stall the replica)
    const int64_t slot = st.log_pos < a.ev_cap ? (int64_t)r * a.ev_cap + st.log_pos : -1;
    const bool moved = c.solo ? solo_move_request(a, c, u, slot, &es, &ed, &ep, &ek, a.ties)
                       : a.exact ? kmc_move_exact(a, c, u, &es, &ed, &ep, &ek, a.ties)
                       : a.fast ? kmc_move_fast(a, c, u, r, st.n_events, &es, &ed, &ep, &ek, a.ties)
                                : kmc_move(a, c, u, &es, &ed, &ep, &ek, a.ties);
    if (!moved) { st.reason = 2; return false; }
    if (a.hyd.on && c.leader) c.tlast[ep - 1] = st.kmc_time;   // update_time_of_last_jump (MDMC.py:99)
    if (a.occ_count && c.leader) {
        // occupancy histogram: the frames consumed so far saw `es` occupied since occ_since[es]; the
        // following ones see `ed` occupied
        const int64_t o = (int64_t)r * a.n_sites;
        a.occ_count[o + es] += (unsigned int)(st.frames_seen - a.occ_since[o + es]);
        a.occ_since[o + ed] = (int)st.frames_seen;
    }
    __syncwarp();
    if (c.leader && st.log_pos < a.ev_cap) {
        int64_t q = (int64_t)r * a.ev_cap + st.log_pos;
        a.ev_frame[q] = st.sweep;
        a.ev_time[q] = st.kmc_time;
        a.ev_start[q] = es; a.ev_dest[q] = ed; a.ev_proton[q] = ep;
        if (!c.solo) ((long long *)a.ev_dist)[q] = c.base + ek;   // solo: a helper warp wrote it
    }
    st.n_events++;
    if (a.ev_cap > 0) st.log_pos++;   // counts on past the capacity: cmd_kmc_events_dropped reports the excess
    st.draws += 2;
    st.cursor += 2;
    return true;
}

// runs the reference's `while True` loop (MDMC.py:147-171) until it needs another frame
__device__ void kmc_run_until_frame_needed(const KmcArgs &a, WarpCtx &c, int r, KmcState &st)
{
    for (;;) {
        bool have_trial = false;
        if (a.rng_mode == CMD_RNG_REPLAY) {
            if (st.cursor + 1 >= a.n_u) { st.phase = KMC_PHASE_HALT; st.reason = 1; return; }
            // the stream carries -np.log(1 - np.random.random()) as the host evaluated it
            replay_fetch(a, c, r, st.cursor);
            st.time_selector = c.uc0;                                 // MDMC.py:148
        } else {
            if (c.nx_event == st.n_events) {   // drawn ahead by kmc_move_fast
                st.time_selector = c.nx_ts;
                st.u_sel = c.nx_usel;
                have_trial = true;
            } else {
                uint32_t ctr[4] = {(uint32_t)st.n_events, (uint32_t)((uint64_t)st.n_events >> 32),
                                   (uint32_t)(a.replica_first + r * a.replica_step), 0u};
                philox4x32_10(ctr, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                st.time_selector = -log(1 - u53(ctr[0], ctr[1]));         // MDMC.py:148
                st.u_sel = u53(ctr[2], ctr[3]);                           // MDMC.py:110, same counter
            }
        }
        // Q1: the rate of frame 0, forever (so the quotient can be formed ahead as well)
        double t_trial = have_trial ? c.nx_trial : st.time_selector / st.current_rate;
        double x = st.kmc_time + t_trial;
        // kmc_time >= 0, dt > 0: Python's // and % (MDMC.py:152,156) are the exact floor and the exact
        // remainder; floor_div_pos returns the same values without fmod's division loop
        double rem_x, rem_t;
        const bool same_frame = floor_div_pos(x, a.dt, a.inv_dt, &rem_x) ==
                                floor_div_pos(st.kmc_time, a.dt, a.inv_dt, &rem_t);
        if (!a.fast && c.leader && (rem_x < 1e-9 * a.dt || a.dt - rem_x < 1e-9 * a.dt))
            atomicAdd(a.ties, 1ull);   // tie audit of the floor-division decision
        if (same_frame) {
            st.kmc_time = x;
            st.delta_frame = 0;
            if (!kmc_event(a, c, r, st)) { st.phase = KMC_PHASE_HALT; return; }
        } else {
            st.delta_t = a.dt - rem_t;
            st.delta_frame = 1;
            st.current_probsum = st.current_rate * st.delta_t;
            st.phase = KMC_PHASE_SCAN;
            return;
        }
    }
}

// the reference's per-frame step once the frame's total allowed rate is known (MDMC.py:146-171)
__device__ void kmc_after_consume(const KmcArgs &a, WarpCtx &c, int r, KmcState &st, double rate)
{
    if (st.phase == KMC_PHASE_START) {  // MDMC.py:146
        st.current_rate = rate;
        kmc_run_until_frame_needed(a, c, r, st);
    } else {  // KMC_PHASE_SCAN, MDMC.py:158-165
        // explicit roundings: no FMA contraction anywhere in the decision arithmetic
        double next_probsum = __dadd_rn(st.current_probsum, __dmul_rn(rate, a.dt));
        if (c.leader && fabs(next_probsum - st.time_selector) < 1e-9 * st.time_selector)
            atomicAdd(a.ties, 1ull);
        if (next_probsum < st.time_selector) {
            st.delta_frame += 1;
            st.current_probsum = next_probsum;
        } else {
            double rest = st.time_selector - st.current_probsum;
            st.delta_t = __dadd_rn(st.delta_t,
                           __dadd_rn(__dmul_rn((double)(st.delta_frame - 1), a.dt),
                                     __ddiv_rn(rest, rate)));
            st.kmc_time += st.delta_t;
            st.sweep += st.delta_frame;
            if (!kmc_event(a, c, r, st)) st.phase = KMC_PHASE_HALT;
            else kmc_run_until_frame_needed(a, c, r, st);
        }
    }
}

__global__ void __launch_bounds__(512, 1) k_kmc_advance(const __grid_constant__ BoxParams bx,
                                                        const __grid_constant__ KmcArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * a.replicas_per_cta + w;
    const bool active = r < a.n_replicas;
    const size_t base_bytes = ((size_t)a.n_sites * 4 + (size_t)a.occ_words * 4 + (size_t)a.mask_words * 4 + 7) / 8 * 8;
    const size_t per_warp = base_bytes + (a.hyd.on ? (size_t)a.n_sites * 8 : 0);
    WarpCtx c;
    c.lane = lane;
    c.solo = false; c.leader = lane == 0; c.tid = lane; c.nthr = 32;
    c.kbits = nullptr; c.wpre = nullptr; c.si = nullptr; c.sd = nullptr;
    c.uc_pos = -8; c.uc0 = c.uc1 = c.un0 = c.un1 = 0.0; c.scan_par = 0; c.nx_event = -1; c.nx_ts = c.nx_usel = c.nx_trial = 0.0; c.cur_rate = 1.0; c.p_next = -1;
    c.lat = (int *)(smem_raw + per_warp * w);
    c.occ = (unsigned *)(c.lat + a.n_sites);
    c.mask0 = c.occ + a.occ_words;
    c.tlast = (double *)(smem_raw + per_warp * w + base_bytes);
    c.t_frame = 0.0;
    c.base = 0;
    c.p = 0;
    c.m = 0;
    c.comp = c.cum = c.lsum = nullptr;
    c.cidx = c.loff = c.ln = nullptr;
    if (a.exact && active) {
        if (a.x_smem) {   // one replica per CTA: comp | cum | lsum | cidx | loff | ln behind the state
            unsigned char *x = smem_raw + ((per_warp * a.replicas_per_cta + 15) / 16) * 16;
            c.comp = (double *)x;
            c.cum = c.comp + a.x_cap;
            c.lsum = c.cum + a.x_cap;
            c.cidx = (int *)(c.lsum + a.x_leaves);
            c.loff = c.cidx + a.x_cap;
            c.ln = c.loff + a.x_leaves;
        } else {
            c.comp = a.x_comp + (int64_t)r * a.x_cap;
            c.cum = a.x_cum + (int64_t)r * a.x_cap;
            c.cidx = a.x_cidx + (int64_t)r * a.x_cap;
            c.lsum = a.x_lsum + (int64_t)r * a.x_leaves;
            c.loff = a.x_loff + (int64_t)r * a.x_leaves;
            c.ln = a.x_ln + (int64_t)r * a.x_leaves;
        }
    }
    KmcState st;
    memset(&st, 0, sizeof(st));
    st.phase = KMC_PHASE_HALT;
    if (active) {
        st = a.state[r];
        for (int s = lane; s < a.n_sites; s += 32) c.lat[s] = a.lattice[(int64_t)r * a.n_sites + s];
        if (a.hyd.on)
            for (int s = lane; s < a.n_sites; s += 32) c.tlast[s] = a.tlast[(int64_t)r * a.n_sites + s];
        __syncwarp();
        for (int q = lane; q < a.occ_words; q += 32) {
            unsigned bits = 0;
            for (int b = 0; b < 32; b++) {
                int s = q * 32 + b;
                if (s < a.n_sites && c.lat[s] > 0) bits |= 1u << b;
            }
            c.occ[q] = bits;
        }
        __syncwarp();
    }
    for (int64_t f = 0; f < a.nframes; f++) {
        if (active && st.phase != KMC_PHASE_HALT) {
            if (a.positions) kmc_observe(a, bx, c, r, f, st);
            c.t_frame = __dmul_rn((double)(a.frames_base + f), a.hyd.frame_dt);   // frame.time
            double rate = a.exact ? kmc_consume_exact(a, c, f) : kmc_consume(a, c, f);
            st.site_updates += c.p;
            st.frames_seen++;
            kmc_after_consume(a, c, r, st, rate);
        }
        __syncthreads();  // keep the CTA's replicas on the same frame (L1 reuse of the frame data)
    }
    if (active) {
        __syncwarp();
        for (int s = lane; s < a.n_sites; s += 32) a.lattice[(int64_t)r * a.n_sites + s] = c.lat[s];
        if (a.hyd.on)
            for (int s = lane; s < a.n_sites; s += 32) a.tlast[(int64_t)r * a.n_sites + s] = c.tlast[s];
        if (lane == 0) a.state[r] = st;
    }
}


#ifndef SOLO_THREADS
#define SOLO_THREADS 384   // measured on C2: 512 threads 25.0, 384 22.1, 256 22.9 us per frame (C1: 6.3 / 5.7 / 5.1)
#endif

// one CTA per replica, exact arithmetic (see "solo mode" above)
__global__ void __launch_bounds__(SOLO_THREADS, 1) k_kmc_solo(const __grid_constant__ BoxParams bx,
                                                              const __grid_constant__ KmcArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int r = blockIdx.x;
    // lat | occ | mask0 | (tlast) | kbits | wpre | si | sd | exact scratch
    const size_t base_bytes = ((size_t)a.n_sites * 4 + (size_t)a.occ_words * 4 + (size_t)a.mask_words * 4 + 7) / 8 * 8;
    const size_t state_bytes = base_bytes + (a.hyd.on ? (size_t)a.n_sites * 8 : 0);
    WarpCtx c;
    c.lane = lane;
    c.solo = true; c.leader = tid == 0; c.tid = tid; c.nthr = blockDim.x;
    c.lat = (int *)smem_raw;
    c.occ = (unsigned *)(c.lat + a.n_sites);
    c.mask0 = c.occ + a.occ_words;
    c.tlast = (double *)(smem_raw + base_bytes);
    c.kbits = (unsigned *)(smem_raw + state_bytes);
    c.wpre = (int *)(c.kbits + a.mask_words);
    c.si = c.wpre + a.mask_words;   // 2 * mask_words ints behind an 8-byte boundary: still aligned
    c.sd = (double *)(c.si + 64);
    c.t_frame = 0.0;
    c.base = 0; c.p = 0; c.m = 0;
    c.psum = nullptr; c.ro = nullptr; c.lane_total = 0.0; c.nst = 0;
    {   // comp | cum (raw rates) | lsum | lsum2 | cpair | praw | loff | ln | loff2 | ln2 | tflags | tcnt
        unsigned char *x = a.x_smem ? (unsigned char *)(c.sd + 64)
                                    : (unsigned char *)a.x_comp + (size_t)r * solo_scratch_bytes(a.x_cap, a.x_leaves);
        c.comp = (double *)x;
        c.cum = c.comp + a.x_cap;
        c.lsum = c.cum + a.x_cap;
        c.lsum2 = c.lsum + a.x_leaves;
        c.cpair = (unsigned *)(c.lsum2 + a.x_leaves);
        c.praw = c.cpair + a.x_cap;
        c.loff = (int *)(c.praw + a.x_cap);
        c.ln = c.loff + a.x_leaves;
        c.loff2 = c.ln + a.x_leaves;
        c.ln2 = c.loff2 + a.x_leaves;
        c.tlw = (int)((a.x_leaves + 31) / 32);
        c.tflags = (unsigned *)(c.ln2 + a.x_leaves);
        c.tcnt = (int *)(c.tflags + (size_t)SOLO_LEVELS * c.tlw);
        c.cidx = nullptr;
    }
    c.uc_pos = -8; c.uc0 = c.uc1 = c.un0 = c.un1 = 0.0; c.scan_par = 0; c.nx_event = -1; c.nx_ts = c.nx_usel = c.nx_trial = 0.0; c.cur_rate = 1.0; c.p_next = -1;
#ifdef SOLO_PROFILE
    for (int i = 0; i < 16; i++) c.prof[i] = 0;
    c.prof_t = clock64();
    const long long prof_c0 = c.prof_t;
    long long prof_n0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_n0));
#endif
    KmcState st = a.state[r];
    for (int s = tid; s < a.n_sites; s += blockDim.x) c.lat[s] = a.lattice[(int64_t)r * a.n_sites + s];
    if (a.hyd.on)
        for (int s = tid; s < a.n_sites; s += blockDim.x) c.tlast[s] = a.tlast[(int64_t)r * a.n_sites + s];
    __syncthreads();
    for (int q = tid; q < a.occ_words; q += blockDim.x) {
        unsigned bits = 0;
        for (int b = 0; b < 32; b++) {
            const int s = q * 32 + b;
            if (s < a.n_sites && c.lat[s] > 0) bits |= 1u << b;
        }
        c.occ[q] = bits;
    }
    __syncthreads();
    const bool first_warp = tid < 32;
    bool halt = st.phase == KMC_PHASE_HALT;   // the same for every thread: read from a.state[r]
    for (int64_t f = 0; f < a.nframes && !halt; f++) {
        if (a.positions) solo_observe(a, bx, c, r, f, st);
        c.t_frame = __dmul_rn((double)(a.frames_base + f), a.hyd.frame_dt);   // frame.time
        const double rate = solo_consume(a, c, f);
        if (first_warp) {
            st.site_updates += c.p;
            st.frames_seen++;
            kmc_after_consume(a, c, r, st, rate);   // events call the helpers in (solo_move_request)
            SOLO_T(9);   // scalar state machine after the last move of the frame
            halt = st.phase == KMC_PHASE_HALT;
            if (lane == 0) c.si[8] = halt ? 2 : 0;
            __syncthreads();
        } else {
            halt = solo_helpers(a, c);
        }
    }
    __syncthreads();
    for (int s = tid; s < a.n_sites; s += blockDim.x) a.lattice[(int64_t)r * a.n_sites + s] = c.lat[s];
    if (a.hyd.on)
        for (int s = tid; s < a.n_sites; s += blockDim.x) a.tlast[(int64_t)r * a.n_sites + s] = c.tlast[s];
    if (tid == 0) a.state[r] = st;
#ifdef SOLO_PROFILE
    if (tid == 0 && r == 0) {
        long long prof_n1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_n1));
        c.prof[10] = clock64() - prof_c0;
        c.prof[11] = prof_n1 - prof_n0;
        for (int i = 0; i < 16; i++) atomicAdd(a.ties + 2 + i, (unsigned long long)c.prof[i]);
    }
#endif
}

#define KMC_STAGE 1024   // pairs per ring stage: 8 KB omega + 4 KB start + 4 KB dest
#define KMC_NSTG 8       // ring depth: a replica busy with events may lag seven stages (~ a frame)

// Streaming KMC kernel (Philox mode).  Warp-specialised: one PRODUCER warp feeds a four-stage
// shared-memory ring with TMA bulk copies of the frames' (start, dest, omega) arrays; every other
// warp is one replica consuming the ring (full / empty mbarriers, no CTA-wide barrier in the
// loop).  Per replica-frame the work is one branch-free pass over the pairs out of shared memory;
// events cost a few scans thanks to the per-(stage, lane) partial sums.
__global__ void __maxnreg__(112) k_kmc_stream(const __grid_constant__ BoxParams bx,
                                                       const __grid_constant__ KmcArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int nrep = a.replicas_per_cta;          // consumer warps; warp `nrep` is the producer
    // ring stage b: omega at b*8K, start at NSTG*8K + b*4K, dest at NSTG*12K + b*4K (computed from the
    // shared base every time so that the loads stay LDS, not generic)
    double *const ring_omega = (double *)smem_raw;
    int *const ring_start = (int *)(smem_raw + KMC_NSTG * KMC_STAGE * 8);
    int *const ring_dest = (int *)(smem_raw + KMC_NSTG * KMC_STAGE * 12);
    unsigned char *q = smem_raw + KMC_NSTG * KMC_STAGE * 16;
    uint64_t *full = (uint64_t *)q, *empty = full + KMC_NSTG;
    q += 2 * KMC_NSTG * 8;
    if (tid == 0) {
        for (int b = 0; b < KMC_NSTG; b++) { mbar_init(&full[b], 1); mbar_init(&empty[b], nrep); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (w == nrep) {
        // ---------------- producer ---------------------------------------------------------------
        if (lane == 0) {
            unsigned it = 0;
            for (int64_t f = 0; f < a.nframes; f++) {
                const int p = a.counts[f];
                const int nst = p > 0 ? (p + KMC_STAGE - 1) / KMC_STAGE : 1;
                for (int cs = 0; cs < nst; cs++, it++) {
                    const int b = it % KMC_NSTG;
                    if (it >= KMC_NSTG) mbar_wait_sleep(&empty[b], ((it / KMC_NSTG) - 1) & 1);
                    const int64_t off = f * a.stride + (int64_t)cs * KMC_STAGE;
                    int64_t e = p > 0 ? a.stride - (int64_t)cs * KMC_STAGE : 0;   // inside the slot
                    if (e > KMC_STAGE) e = KMC_STAGE;
                    if (e > 0) {
                        mbar_expect_tx(&full[b], (uint32_t)e * 16u);
                        tma_load_1d(ring_start + b * KMC_STAGE, a.start + off, (uint32_t)e * 4u, &full[b]);
                        tma_load_1d(ring_dest + b * KMC_STAGE, a.dest + off, (uint32_t)e * 4u, &full[b]);
                        tma_load_1d(ring_omega + b * KMC_STAGE, a.omega + off, (uint32_t)e * 8u, &full[b]);
                    } else {
                        mbar_arrive(&full[b]);   // a frame without pairs: an empty stage
                    }
                }
            }
        }
        return;
    }

    // ---------------- consumers: one replica per warp -------------------------------------------
    const int r = blockIdx.x * nrep + w;
    const bool active = r < a.n_replicas;
    // per warp: psum | mask0 (16-byte aligned: written four words at a time) | lattice | occupancy
    const size_t mask_bytes = ((size_t)a.mask_words * 4 + 15) / 16 * 16;
    const size_t per_warp = ((size_t)a.nst_max * 32 * 8 + mask_bytes + (size_t)a.n_sites * 4 +
                             (size_t)a.occ_words * 4 + 15) / 16 * 16;
    WarpCtx c;
    c.lane = lane;
    c.solo = false; c.leader = lane == 0; c.tid = lane; c.nthr = 32;
    c.kbits = nullptr; c.wpre = nullptr; c.si = nullptr; c.sd = nullptr;
    c.uc_pos = -8; c.uc0 = c.uc1 = c.un0 = c.un1 = 0.0; c.scan_par = 0; c.nx_event = -1; c.nx_ts = c.nx_usel = c.nx_trial = 0.0; c.cur_rate = 1.0; c.p_next = -1;
    c.psum = (double *)(q + per_warp * w);
    c.mask0 = (unsigned *)(c.psum + (size_t)a.nst_max * 32);
    c.lat = (int *)((unsigned char *)c.mask0 + mask_bytes);
    c.occ = (unsigned *)(c.lat + a.n_sites);
    c.base = 0; c.p = 0; c.m = 0; c.nst = 0; c.lane_total = 0.0; c.ro = nullptr;
    c.scan_base = -1; c.scan_inc = 0.0;
    c.comp = c.cum = c.lsum = nullptr;
    c.cidx = c.loff = c.ln = nullptr;
    KmcState st;
    memset(&st, 0, sizeof(st));
    st.phase = KMC_PHASE_HALT;
    if (active) {
        st = a.state[r];
        for (int s = lane; s < a.n_sites; s += 32) c.lat[s] = a.lattice[(int64_t)r * a.n_sites + s];
        __syncwarp();
        for (int qq = lane; qq < a.occ_words; qq += 32) {
            unsigned bits = 0;
            for (int b = 0; b < 32; b++) {
                int s = qq * 32 + b;
                if (s < a.n_sites && c.lat[s] > 0) bits |= 1u << b;
            }
            c.occ[qq] = bits;
        }
        __syncwarp();
    }
    unsigned it = 0;
    for (int64_t f = 0; f < a.nframes; f++) {
        const int p = a.counts[f];
        const int nst = p > 0 ? (p + KMC_STAGE - 1) / KMC_STAGE : 1;
        c.base = f * a.stride;
        c.p = p;
        c.nst = nst;
        c.lane_total = 0.0;
        const bool run = active && st.phase != KMC_PHASE_HALT;
        if (run && a.positions) kmc_observe(a, bx, c, r, f, st);
        for (int cs = 0; cs < nst; cs++, it++) {
            const int b = it % KMC_NSTG;
            mbar_wait(&full[b], (it / KMC_NSTG) & 1);
            if (run) {
                const int cnt = min(KMC_STAGE, p - cs * KMC_STAGE);
                kmc_consume_stage(c, ring_start + b * KMC_STAGE, ring_dest + b * KMC_STAGE,
                                  ring_omega + b * KMC_STAGE, cs * KMC_STAGE, cnt > 0 ? cnt : 0, cs);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[b]);   // this replica is done with the stage
        }
        if (run) {
            const double rate = warp_sum(c.lane_total);
            st.site_updates += p;
            st.frames_seen++;
            kmc_after_consume(a, c, r, st, rate);
        }
    }
    if (active) {
        __syncwarp();
        for (int s = lane; s < a.n_sites; s += 32) a.lattice[(int64_t)r * a.n_sites + s] = c.lat[s];
        if (lane == 0) a.state[r] = st;
    }
}

// ------------------------------------------------------------------ host side ------------------
extern "C" void cmd_kmc_destroy(cmd_kmc *k)
{
    if (!k) return;
    cudaStreamSynchronize(cmd_global().stream);
    cudaFree(k->d_lattice); cudaFree(k->d_lattice0); cudaFree(k->d_state); cudaFree(k->d_u);
    cudaFree(k->d_ev_frame); cudaFree(k->d_ev_time); cudaFree(k->d_ev_start); cudaFree(k->d_ev_dest);
    cudaFree(k->d_ev_proton); cudaFree(k->d_ev_dist); cudaFree(k->d_rows); cudaFree(k->d_snapshot); cudaFree(k->d_disp);
    cudaFree(k->d_ties); cudaFree(k->d_exact); cudaFree(k->d_tlast); cudaFree(k->d_tx); cudaFree(k->d_ty);
    cudaFree(k->d_occ_count); cudaFree(k->d_occ_since);
    free(k);
}

#define KALLOC(ptr, bytes)                                                     \
    if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) {                 \
        cudaGetLastError();                                                    \
        return cmd_set_error(CMD_ENOMEM, "cudaMalloc failed for %s (%zu bytes)", #ptr, (size_t)(bytes)); \
    }

extern "C" int cmd_kmc_create(const cmd_box *box, int n_sites, int n_replicas, const int *lattices,
                              double dt, int rng_mode, uint64_t seed, cmd_kmc **out)
{
    CMD_REQUIRE_INIT();
    if (!box || !out || !lattices || n_sites < 1 || n_replicas < 1 || !(dt > 0))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    if (rng_mode != CMD_RNG_REPLAY && rng_mode != CMD_RNG_PHILOX)
        return cmd_set_error(CMD_EINVAL, "bad rng mode");
    cmd_kmc *k = (cmd_kmc *)calloc(1, sizeof(cmd_kmc));
    if (!k) return cmd_set_error(CMD_ENOMEM, "out of host memory");
    k->bx = box->p;
    k->n_sites = n_sites;
    k->n_replicas = n_replicas;
    k->dt = dt;
    k->rng_mode = rng_mode;
    k->seed = seed;
    k->replica_first = 0;
    k->replica_step = 1;
    {   // test / A-B knobs of the solo kernel
        const char *e = getenv("CMDLMC_B200_KMC_SOLO");
        k->solo_enabled = !(e && e[0] == '0');
        e = getenv("CMDLMC_B200_KMC_SELECT_MARGIN");
        const double scale = e ? atof(e) : 1.0;
        k->sel_margin = ldexp(1.0, -50) * (scale > 0.0 ? scale : 1.0);
    }
    cudaStream_t st = cmd_global().stream;
    size_t nl = (size_t)n_replicas * n_sites;
    int rc = CMD_OK;
    do {
        if (cudaMalloc((void **)&k->d_lattice, nl * 4) != cudaSuccess ||
            cudaMalloc((void **)&k->d_lattice0, nl * 4) != cudaSuccess ||
            cudaMalloc((void **)&k->d_state, (size_t)n_replicas * sizeof(KmcState)) != cudaSuccess ||
            cudaMalloc((void **)&k->d_ties, 256) != cudaSuccess) {
            cudaGetLastError();
            rc = cmd_set_error(CMD_ENOMEM, "cudaMalloc failed for the KMC state");
            break;
        }
        cudaMemcpyAsync(k->d_lattice, lattices, nl * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(k->d_lattice0, lattices, nl * 4, cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(k->d_state, 0, (size_t)n_replicas * sizeof(KmcState), st);
        cudaMemsetAsync(k->d_ties, 0, 256, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) {
            rc = cmd_set_error(CMD_ECUDA, "KMC state upload failed: %s",
                               cudaGetErrorString(cudaGetLastError()));
            break;
        }
    } while (0);
    if (rc) { cmd_kmc_destroy(k); return rc; }
    *out = k;
    return CMD_OK;
}

__global__ void k_kmc_reset_cursor(KmcState *st, int n, int what)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) {
        if (what == 0) st[r].cursor = 0;
        else { st[r].log_pos = 0; st[r].ev_resolved = 0; }
    }
}

__global__ void k_max_count(const int *__restrict__ counts, int64_t n, int *__restrict__ out)
{
    int m = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = max(m, counts[i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// jump distances of the events logged since the last call: flat pair index -> distance (kmc_event)
__global__ void __launch_bounds__(128)
k_resolve_ev_dist(KmcState *__restrict__ state, double *__restrict__ ev_dist, int64_t ev_cap,
                  const double *__restrict__ dist)
{
    const int r = blockIdx.x;
    const long long from = state[r].ev_resolved;
    long long to = state[r].log_pos;
    if (to > ev_cap) to = ev_cap;
    for (long long q = from + threadIdx.x; q < to; q += blockDim.x) {
        const long long idx = ((const long long *)ev_dist)[(int64_t)r * ev_cap + q];
        ev_dist[(int64_t)r * ev_cap + q] = dist[idx];
    }
    __syncthreads();
    if (threadIdx.x == 0 && to > from) state[r].ev_resolved = to;
}

__global__ void k_fill_double(double *p, int64_t n, double v)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

extern "C" int cmd_kmc_set_hydronium(cmd_kmc *k, int rate_kind, const double rate_par[CMD_RATE_NPAR],
                                     int transform_kind, const double tpar[5], const double *h_table_x,
                                     const double *h_table_y, int n_table, double relaxation_time,
                                     double frame_time_step)
{
    CMD_REQUIRE_INIT();
    if (!k || !rate_par || rate_kind < 0 || rate_kind > CMD_RATE_EXP || rate_kind == CMD_RATE_FERMI_ANGLE ||
        transform_kind < 0 || transform_kind > 2)
        return cmd_set_error(CMD_EINVAL, "bad argument");
    if (transform_kind == 1 && !tpar) return cmd_set_error(CMD_EINVAL, "ReLU parameters missing");
    if (transform_kind == 2 && (!h_table_x || !h_table_y || n_table < 2))
        return cmd_set_error(CMD_EINVAL, "interpolation table needs at least two points");
    if (k->frames_total > 0) return cmd_set_error(CMD_ESTATE, "set the hydronium mode before the first advance");
    cudaStream_t st = cmd_global().stream;
    HydParams &h = k->hyd;
    memset(&h, 0, sizeof(h));
    h.on = 1;
    h.rate.kind = rate_kind;
    memcpy(h.rate.par, rate_par, sizeof(h.rate.par));
    h.tkind = transform_kind;
    if (tpar) memcpy(h.tpar, tpar, sizeof(h.tpar));
    h.relax = relaxation_time;
    h.frame_dt = frame_time_step;
    if (transform_kind == 2) {
        cudaFree(k->d_tx); cudaFree(k->d_ty);
        k->d_tx = k->d_ty = nullptr;
        KALLOC(k->d_tx, (size_t)n_table * 8);
        KALLOC(k->d_ty, (size_t)n_table * 8);
        CMD_CUDA(cudaMemcpyAsync(k->d_tx, h_table_x, (size_t)n_table * 8, cudaMemcpyHostToDevice, st));
        CMD_CUDA(cudaMemcpyAsync(k->d_ty, h_table_y, (size_t)n_table * 8, cudaMemcpyHostToDevice, st));
        h.tx = k->d_tx; h.ty = k->d_ty; h.nt = n_table;
    }
    if (!k->d_tlast) KALLOC(k->d_tlast, (size_t)k->n_replicas * k->n_sites * 8);
    // _time_of_last_jump_vec = -1 for every proton (topology.py:209)
    k_fill_double<<<64, 256, 0, st>>>(k->d_tlast, (int64_t)k->n_replicas * k->n_sites, -1.0);
    CMD_LAUNCHED();
    CMD_CUDA(cudaStreamSynchronize(st));
    return CMD_OK;
}

extern "C" int cmd_kmc_get_last_jump_times(const cmd_kmc *k, double *h_tlast)
{
    CMD_REQUIRE_INIT();
    if (!k || !h_tlast || !k->d_tlast) return cmd_set_error(CMD_ESTATE, "hydronium mode is not enabled");
    cudaStream_t st = cmd_global().stream;
    CMD_CUDA(cudaMemcpyAsync(h_tlast, k->d_tlast, (size_t)k->n_replicas * k->n_sites * 8, cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    return CMD_OK;
}

extern "C" int cmd_kmc_enable_occupancy(cmd_kmc *k)
{
    CMD_REQUIRE_INIT();
    if (!k) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (k->frames_total > 0) return cmd_set_error(CMD_ESTATE, "enable the occupancy histogram before the first advance");
    if (!k->d_occ_count) {
        const size_t n = (size_t)k->n_replicas * k->n_sites;
        KALLOC(k->d_occ_count, n * 4);
        KALLOC(k->d_occ_since, n * 4);
        CMD_CUDA(cudaMemsetAsync(k->d_occ_count, 0, n * 4, cmd_global().stream));
        CMD_CUDA(cudaMemsetAsync(k->d_occ_since, 0, n * 4, cmd_global().stream));
    }
    return CMD_OK;
}

// counts[s] = number of (replica, consumed frame) pairs that saw site s occupied
extern "C" int cmd_kmc_get_occupancy(const cmd_kmc *k, int64_t *h_counts, int64_t *h_replica_frames)
{
    CMD_REQUIRE_INIT();
    if (!k || !h_counts) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (!k->d_occ_count) return cmd_set_error(CMD_ESTATE, "the occupancy histogram is not enabled");
    cudaStream_t st = cmd_global().stream;
    const size_t R = (size_t)k->n_replicas, n = (size_t)k->n_sites;
    unsigned int *cnt = (unsigned int *)malloc(R * n * 4);
    int *since = (int *)malloc(R * n * 4), *lat = (int *)malloc(R * n * 4);
    KmcState *hs = (KmcState *)malloc(R * sizeof(KmcState));
    if (!cnt || !since || !lat || !hs) {
        free(cnt); free(since); free(lat); free(hs);
        return cmd_set_error(CMD_ENOMEM, "out of host memory");
    }
    cudaError_t e = cudaMemcpyAsync(cnt, k->d_occ_count, R * n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(since, k->d_occ_since, R * n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(lat, k->d_lattice, R * n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hs, k->d_state, R * sizeof(KmcState), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) {
        int64_t frames = 0;
        for (size_t s = 0; s < n; s++) h_counts[s] = 0;
        for (size_t r = 0; r < R; r++) {
            frames += hs[r].frames_seen;
            for (size_t s = 0; s < n; s++) {
                int64_t c = cnt[r * n + s];
                if (lat[r * n + s] > 0) c += hs[r].frames_seen - since[r * n + s];   // still open
                h_counts[s] += c;
            }
        }
        if (h_replica_frames) *h_replica_frames = frames;
    }
    free(cnt); free(since); free(lat); free(hs);
    CMD_CUDA(e);
    return CMD_OK;
}

extern "C" int cmd_kmc_set_replica_ids(cmd_kmc *k, int first, int step)
{
    if (!k || first < 0 || step < 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    k->replica_first = first;
    k->replica_step = step;
    return CMD_OK;
}

extern "C" int cmd_kmc_get_status(const cmd_kmc *k, int *phase, int *reason, int64_t *cursor)
{
    CMD_REQUIRE_INIT();
    if (!k) return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    KmcState *hs = (KmcState *)malloc((size_t)k->n_replicas * sizeof(KmcState));
    if (!hs) return cmd_set_error(CMD_ENOMEM, "out of host memory");
    cudaError_t e = cudaMemcpyAsync(hs, k->d_state, (size_t)k->n_replicas * sizeof(KmcState),
                                    cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess)
        for (int r = 0; r < k->n_replicas; r++) {
            if (phase) phase[r] = hs[r].phase;
            if (reason) reason[r] = hs[r].reason;
            if (cursor) cursor[r] = hs[r].cursor;
        }
    free(hs);
    CMD_CUDA(e);
    return CMD_OK;
}

extern "C" int cmd_kmc_set_replay_stream(cmd_kmc *k, const double *u, int64_t n)
{
    CMD_REQUIRE_INIT();
    if (!k || !u || n < 2) return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    CMD_CUDA(cudaStreamSynchronize(st));
    cudaFree(k->d_u);
    k->d_u = nullptr;
    KALLOC(k->d_u, (size_t)k->n_replicas * n * 8);
    CMD_CUDA(cudaMemcpyAsync(k->d_u, u, (size_t)k->n_replicas * n * 8, cudaMemcpyHostToDevice, st));
    k_kmc_reset_cursor<<<(k->n_replicas + 255) / 256, 256, 0, st>>>(k->d_state, k->n_replicas, 0);
    CMD_LAUNCHED();
    CMD_CUDA(cudaStreamSynchronize(st));
    k->n_u = n;
    return CMD_OK;
}

extern "C" int cmd_kmc_set_event_log(cmd_kmc *k, int64_t cap)
{
    CMD_REQUIRE_INIT();
    if (!k || cap < 0) return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    if (cap > k->ev_cap) {
        CMD_CUDA(cudaStreamSynchronize(st));
        cudaFree(k->d_ev_frame); cudaFree(k->d_ev_time); cudaFree(k->d_ev_start);
        cudaFree(k->d_ev_dest); cudaFree(k->d_ev_proton); cudaFree(k->d_ev_dist);
        k->d_ev_frame = nullptr; k->d_ev_time = nullptr; k->d_ev_dist = nullptr;
        k->d_ev_start = k->d_ev_dest = k->d_ev_proton = nullptr;
        k->ev_cap = 0;
        size_t n = (size_t)k->n_replicas * cap;
        KALLOC(k->d_ev_frame, n * 8);
        KALLOC(k->d_ev_time, n * 8);
        KALLOC(k->d_ev_start, n * 4);
        KALLOC(k->d_ev_dest, n * 4);
        KALLOC(k->d_ev_proton, n * 4);
        KALLOC(k->d_ev_dist, n * 8);
        k->ev_cap = cap;
    }
    if (cap == 0) k->ev_cap = 0;
    k_kmc_reset_cursor<<<(k->n_replicas + 255) / 256, 256, 0, st>>>(k->d_state, k->n_replicas, 1);
    CMD_LAUNCHED();
    return CMD_OK;
}

extern "C" int cmd_kmc_set_observables(cmd_kmc *k, int reset_frequency, int print_frequency)
{
    CMD_REQUIRE_INIT();
    if (!k || reset_frequency < 0 || print_frequency < 0) return cmd_set_error(CMD_EINVAL, "bad argument");
    k->reset_freq = reset_frequency;
    k->print_freq = print_frequency;
    if (print_frequency > 0 && !k->d_snapshot) {
        size_t n = (size_t)k->n_replicas * k->n_sites * 3 * 8;
        KALLOC(k->d_snapshot, n);
        KALLOC(k->d_disp, n);
        CMD_CUDA(cudaMemsetAsync(k->d_snapshot, 0, n, cmd_global().stream));
        CMD_CUDA(cudaMemsetAsync(k->d_disp, 0, n, cmd_global().stream));
    }
    return CMD_OK;
}

// frame number 0 of the observables from positions handed in by the host (see the header)
__global__ void k_obs_seed(const int *__restrict__ lattice, const double *__restrict__ pos, int n_sites,
                           int *__restrict__ lattice0, double *__restrict__ snapshot,
                           double *__restrict__ disp)
{
    const int r = blockIdx.x;
    for (int s = threadIdx.x; s < n_sites; s += blockDim.x) {
        const int l = lattice[(int64_t)r * n_sites + s];
        lattice0[(int64_t)r * n_sites + s] = l;
        if (l > 0)
            for (int k = 0; k < 3; k++) {
                snapshot[((int64_t)r * n_sites + (l - 1)) * 3 + k] = pos[3 * s + k];
                disp[((int64_t)r * n_sites + (l - 1)) * 3 + k] = 0.0;
            }
    }
}

extern "C" int cmd_kmc_seed_observables(cmd_kmc *k, const double *h_positions)
{
    CMD_REQUIRE_INIT();
    if (!k || !h_positions) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (k->print_freq <= 0 || !k->d_snapshot)
        return cmd_set_error(CMD_ESTATE, "call cmd_kmc_set_observables first");
    if (k->frames_total != 0) return cmd_set_error(CMD_ESTATE, "frames have been walked already");
    cudaStream_t st = cmd_global().stream;
    void *buf;
    int rc = cmd_scratch(4, (size_t)k->n_sites * 24, &buf);
    if (rc) return rc;
    CMD_CUDA(cudaMemcpyAsync(buf, h_positions, (size_t)k->n_sites * 24, cudaMemcpyHostToDevice, st));
    k_obs_seed<<<k->n_replicas, 256, 0, st>>>(k->d_lattice, (const double *)buf, k->n_sites, k->d_lattice0,
                                              k->d_snapshot, k->d_disp);
    CMD_LAUNCHED();
    CMD_CUDA(cudaStreamSynchronize(st));   // h_positions is borrowed for the call only
    k->obs_offset = 1;
    return CMD_OK;
}

static int kmc_resolve_events(cmd_kmc *k, const double *d_dist)
{
    if (k->ev_cap > 0) {
        k_resolve_ev_dist<<<k->n_replicas, 128, 0, cmd_global().stream>>>(k->d_state, k->d_ev_dist, k->ev_cap, d_dist);
        CMD_LAUNCHED();
    }
    return CMD_OK;
}

extern "C" int cmd_kmc_advance(cmd_kmc *k, const cmd_topo *t, const double *d_positions)
{
    CMD_REQUIRE_INIT();
    if (!k || !t) return cmd_set_error(CMD_EINVAL, "bad argument");
    const int *d_start, *d_dest, *d_counts;
    const double *d_dist, *d_omega;
    int rc = cmd_topo_device_arrays(t, &d_start, &d_dest, &d_dist, &d_omega, &d_counts);
    if (rc) return rc;
    if (k->rng_mode == CMD_RNG_REPLAY && !k->d_u)
        return cmd_set_error(CMD_ESTATE, "replay mode needs cmd_kmc_set_replay_stream first");
    const bool obs = k->print_freq > 0;
    if (obs && !d_positions)
        return cmd_set_error(CMD_EINVAL, "observables need the donor positions of the block");
    CmdGlobal &g = cmd_global();
    cudaStream_t st = g.stream;
    int64_t nframes = cmd_topo_nframes(t), stride = cmd_topo_stride(t);
    if (k->hyd.on) {
        // HydroniumTopology: the transitions of a frame are the k nearest listed neighbours of every
        // site (cmd_topo_nearest); their rates are evaluated per replica inside the kernel
        if ((rc = cmd_topo_near_arrays(t, &d_start, &d_dest, &d_dist, &d_counts, &stride))) return rc;
        d_omega = d_dist;
    }
    // observable rows: grow to hold this block's prints
    if (obs) {
        int64_t need = (k->frames_total + k->obs_offset + nframes) / k->print_freq + 2;
        if (need > k->row_cap) {
            int64_t cap = need + need / 2 + 16;
            double *nr = nullptr;
            KALLOC(nr, (size_t)k->n_replicas * cap * 6 * 8);
            if (k->d_rows) {
                CMD_CUDA(cudaMemcpy2DAsync(nr, cap * 48, k->d_rows, k->row_cap * 48, k->row_cap * 48,
                                           k->n_replicas, cudaMemcpyDeviceToDevice, st));
                CMD_CUDA(cudaStreamSynchronize(st));
                cudaFree(k->d_rows);
            }
            k->d_rows = nr;
            k->row_cap = cap;
        }
    }
    KmcArgs a;
    memset(&a, 0, sizeof(a));
    a.n_sites = k->n_sites; a.n_replicas = k->n_replicas; a.rng_mode = k->rng_mode;
    a.replica_first = k->replica_first; a.replica_step = k->replica_step;
    a.dt = k->dt; a.inv_dt = 1.0 / k->dt; a.seed = k->seed; a.stride = stride; a.nframes = nframes;
    a.frames_base = k->frames_total; a.obs_base = k->frames_total + k->obs_offset; a.n_u = k->n_u; a.ev_cap = k->ev_cap; a.row_cap = k->row_cap;
    a.reset_freq = k->reset_freq; a.print_freq = k->print_freq;
    a.start = d_start; a.dest = d_dest; a.counts = d_counts; a.omega = d_omega; a.dist = d_dist;
    a.positions = obs ? d_positions : nullptr;
    a.u = k->d_u; a.lattice = k->d_lattice; a.lattice0 = k->d_lattice0; a.state = k->d_state;
    a.ev_frame = k->d_ev_frame; a.ev_time = k->d_ev_time; a.ev_start = k->d_ev_start;
    a.ev_dest = k->d_ev_dest; a.ev_proton = k->d_ev_proton; a.ev_dist = k->d_ev_dist;
    a.rows = k->d_rows; a.snapshot = k->d_snapshot; a.disp = k->d_disp; a.ties = k->d_ties;
    a.hyd = k->hyd;
    a.tlast = k->d_tlast;
    a.occ_count = k->d_occ_count;
    a.occ_since = k->d_occ_since;
    // reference-order arithmetic in replay mode -- and for hydronium runs in either RNG mode
    a.exact = (k->rng_mode == CMD_RNG_REPLAY || k->hyd.on) ? 1 : 0;
    if (a.exact) {
        // per replica: comp f64[cap], cum f64[cap], lsum f64[leaves], cidx i32[cap],
        // loff i32[leaves], ln i32[leaves]
        a.x_cap = stride;
        a.x_leaves = stride / 64 + 4;
        size_t R = (size_t)k->n_replicas;
        size_t need = R * ((size_t)a.x_cap * 20 + (size_t)a.x_leaves * 16);
        {   // the solo kernel rounds its capacity up to 64 entries
            const int64_t cap64 = (stride + 63) / 64 * 64;
            const size_t solo = R * solo_scratch_bytes(cap64, cap64 / 64 + 4);
            if (need < solo) need = solo;
        }
        if (need > k->exact_bytes) {
            CMD_CUDA(cudaStreamSynchronize(st));
            cudaFree(k->d_exact);
            k->d_exact = nullptr;
            k->exact_bytes = 0;
            KALLOC(k->d_exact, need);
            k->exact_bytes = need;
        }
        a.x_comp = (double *)k->d_exact;
        a.x_cum = a.x_comp + R * a.x_cap;
        a.x_lsum = a.x_cum + R * a.x_cap;
        a.x_cidx = (int *)(a.x_lsum + R * a.x_leaves);
        a.x_loff = a.x_cidx + R * a.x_cap;
        a.x_ln = a.x_loff + R * a.x_leaves;
    }
    a.occ_words = (k->n_sites + 31) / 32;
    a.mask_words = (int)((stride + 31) / 32) + 1;
    int rpc = (k->n_replicas + g.sm_count - 1) / g.sm_count;
    if (rpc < 1) rpc = 1;
    if (rpc > 16) rpc = 16;
    if (k->rng_mode == CMD_RNG_PHILOX && !k->hyd.on && cmd_topo_n_atoms(t) == k->n_sites) {
        // streaming kernel: TMA-fed shared-memory ring + per-(stage, lane) partial sums
        const int *d_rowoff;
        if ((rc = cmd_topo_row_offsets(t, &d_rowoff))) return rc;
        a.fast = 1;
        a.rowoff = d_rowoff;
        a.ro_pitch = cmd_ro_pitch(k->n_sites);
        a.nst_max = (int)((stride + KMC_STAGE - 1) / KMC_STAGE);
        if (a.nst_max < 1) a.nst_max = 1;
        const size_t ring = (size_t)KMC_NSTG * KMC_STAGE * 16 + 2 * KMC_NSTG * 8;
        const size_t mask_bytes = ((size_t)a.mask_words * 4 + 15) / 16 * 16;
        const size_t per_warp = ((size_t)a.nst_max * 32 * 8 + mask_bytes + (size_t)a.n_sites * 4 +
                                 (size_t)a.occ_words * 4 + 15) / 16 * 16;
        while (rpc > 1 && ring + per_warp * rpc > 200 * 1024) rpc--;
        if (ring + per_warp * rpc <= 226 * 1024) {
            a.replicas_per_cta = rpc;
            const size_t smem = ring + per_warp * rpc;
            CMD_CUDA(cudaFuncSetAttribute(k_kmc_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int blocks = (k->n_replicas + rpc - 1) / rpc;
            k_kmc_stream<<<blocks, (rpc + 1) * 32, smem, st>>>(k->bx, a);   // + the producer warp
            CMD_LAUNCHED();
            k->frames_total += nframes;
            return kmc_resolve_events(k, d_dist);
        }
        a.fast = 0;   // state too large for the ring: the plain kernel below
    }
    size_t per_warp = ((size_t)a.n_sites * 4 + (size_t)a.occ_words * 4 + (size_t)a.mask_words * 4 + 7) / 8 * 8 +
                      (k->hyd.on ? (size_t)a.n_sites * 8 : 0);
    while (rpc > 1 && per_warp * rpc > 200 * 1024) rpc--;
    if (per_warp * rpc > 226 * 1024)
        return cmd_set_error(CMD_ECAPACITY, "KMC per-replica state (%zu bytes) exceeds shared memory", per_warp);
    a.replicas_per_cta = rpc;
    size_t smem = per_warp * rpc;
    if (a.exact && rpc == 1 && k->solo_enabled && k->n_sites < 65536) {
        // few replicas (`mdmc`, verification runs): one CTA per replica.  Its scratch is sized by the
        // block's largest pair count, not by the per-frame capacity.
        int pmax = 0;
        CMD_CUDA(cudaMemsetAsync(k->d_ties + 31, 0, 8, st));
        k_max_count<<<64, 256, 0, st>>>(d_counts, nframes, (int *)(k->d_ties + 31));
        CMD_LAUNCHED();
        CMD_CUDA(cudaMemcpyAsync(&pmax, k->d_ties + 31, sizeof(int), cudaMemcpyDeviceToHost, st));
        CMD_CUDA(cudaStreamSynchronize(st));
        if (pmax < 0 || pmax > stride) pmax = (int)stride;
        a.x_cap = (pmax + 63) / 64 * 64;
        a.x_leaves = a.x_cap / 64 + 4;
        const size_t aux = (size_t)a.mask_words * 8 + 64 * 4 + 64 * 8;
        const size_t xbytes = solo_scratch_bytes(a.x_cap, a.x_leaves);
        const size_t xoff = ((per_warp + 7) / 8) * 8 + aux;
        size_t need = xoff;
        if (xoff + xbytes <= 226 * 1024) {
            a.x_smem = 1;
            need = xoff + xbytes;
        }
        if (need <= 226 * 1024) {
            a.sel_margin = k->sel_margin;
            CMD_CUDA(cudaFuncSetAttribute(k_kmc_solo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
            // short lists: fewer warps, cheaper barriers
            k_kmc_solo<<<k->n_replicas, a.x_cap <= 4096 ? 256 : SOLO_THREADS, need, st>>>(k->bx, a);
            CMD_LAUNCHED();
            k->frames_total += nframes;
            return kmc_resolve_events(k, d_dist);
        }
        a.x_smem = 0;
        a.x_cap = stride;   // the warp-per-replica kernel below slices the scratch by the capacity
        a.x_leaves = stride / 64 + 4;
    }
    if (a.exact && rpc == 1) {   // scratch in shared memory
        const size_t xbytes = (size_t)a.x_cap * 20 + (size_t)a.x_leaves * 16;
        const size_t xoff = ((per_warp + 15) / 16) * 16;
        if (xoff + xbytes <= 226 * 1024) {
            a.x_smem = 1;
            smem = xoff + xbytes;
        }
    }
    CMD_CUDA(cudaFuncSetAttribute(k_kmc_advance, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (k->n_replicas + rpc - 1) / rpc;
    k_kmc_advance<<<blocks, rpc * 32, smem, st>>>(k->bx, a);
    CMD_LAUNCHED();
    k->frames_total += nframes;
    return kmc_resolve_events(k, d_dist);
}

extern "C" int cmd_kmc_get_state(const cmd_kmc *k, int *lattices, double *time, int64_t *frame,
                                 int64_t *n_events, int64_t *site_updates)
{
    CMD_REQUIRE_INIT();
    if (!k) return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    KmcState *hs = (KmcState *)malloc((size_t)k->n_replicas * sizeof(KmcState));
    if (!hs) return cmd_set_error(CMD_ENOMEM, "out of host memory");
    cudaError_t e = cudaMemcpyAsync(hs, k->d_state, (size_t)k->n_replicas * sizeof(KmcState),
                                    cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && lattices)
        e = cudaMemcpyAsync(lattices, k->d_lattice, (size_t)k->n_replicas * k->n_sites * 4,
                            cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess)
        for (int r = 0; r < k->n_replicas; r++) {
            if (time) time[r] = hs[r].kmc_time;
            if (frame) frame[r] = hs[r].sweep;
            if (n_events) n_events[r] = hs[r].n_events;
            if (site_updates) site_updates[r] = hs[r].site_updates;
        }
    free(hs);
    CMD_CUDA(e);
    return CMD_OK;
}

extern "C" int cmd_kmc_get_events(const cmd_kmc *k, int replica, int64_t capacity, int64_t *n,
                                  int64_t *frame, double *time, int *start, int *dest, int *proton)
{
    CMD_REQUIRE_INIT();
    if (!k || replica < 0 || replica >= k->n_replicas || !n) return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    KmcState hs;
    CMD_CUDA(cudaMemcpyAsync(&hs, k->d_state + replica, sizeof(hs), cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    int64_t have = hs.log_pos < k->ev_cap ? hs.log_pos : k->ev_cap;
    if (have > capacity) have = capacity;
    *n = have;
    int64_t off = (int64_t)replica * k->ev_cap;
    if (have > 0) {
        if (frame) CMD_CUDA(cudaMemcpyAsync(frame, k->d_ev_frame + off, have * 8, cudaMemcpyDeviceToHost, st));
        if (time) CMD_CUDA(cudaMemcpyAsync(time, k->d_ev_time + off, have * 8, cudaMemcpyDeviceToHost, st));
        if (start) CMD_CUDA(cudaMemcpyAsync(start, k->d_ev_start + off, have * 4, cudaMemcpyDeviceToHost, st));
        if (dest) CMD_CUDA(cudaMemcpyAsync(dest, k->d_ev_dest + off, have * 4, cudaMemcpyDeviceToHost, st));
        if (proton) CMD_CUDA(cudaMemcpyAsync(proton, k->d_ev_proton + off, have * 4, cudaMemcpyDeviceToHost, st));
        CMD_CUDA(cudaStreamSynchronize(st));
    }
    return CMD_OK;
}

extern "C" int cmd_kmc_get_event_distances(const cmd_kmc *k, int replica, int64_t capacity,
                                           int64_t *n, double *dist)
{
    CMD_REQUIRE_INIT();
    if (!k || replica < 0 || replica >= k->n_replicas || !n) return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    KmcState hs;
    CMD_CUDA(cudaMemcpyAsync(&hs, k->d_state + replica, sizeof(hs), cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    int64_t have = hs.log_pos < k->ev_cap ? hs.log_pos : k->ev_cap;
    if (have > capacity) have = capacity;
    *n = have;
    if (have > 0 && dist) {
        CMD_CUDA(cudaMemcpyAsync(dist, k->d_ev_dist + (int64_t)replica * k->ev_cap, have * 8,
                                 cudaMemcpyDeviceToHost, st));
        CMD_CUDA(cudaStreamSynchronize(st));
    }
    return CMD_OK;
}

// jump-pair distances of the logged events of every replica -> histogram (jumpstat numerator)
__global__ void __launch_bounds__(256)
k_jump_hist(const KmcState *__restrict__ state, const double *__restrict__ ev_dist, int64_t ev_cap,
            int n_replicas, double lo, double inv_width, int nbins, unsigned long long *__restrict__ hist)
{
    const int r = blockIdx.x;
    if (r >= n_replicas) return;
    long long have = state[r].log_pos < ev_cap ? state[r].log_pos : ev_cap;
    for (long long e = threadIdx.x; e < have; e += blockDim.x) {
        const double d = ev_dist[(int64_t)r * ev_cap + e];
        const double b = floor((d - lo) * inv_width);
        if (b >= 0 && b < nbins) atomicAdd(hist + (int)b, 1ull);
    }
}

extern "C" int cmd_kmc_jump_histogram_dev(const cmd_kmc *k, double lo, double hi, int nbins,
                                          unsigned long long *d_hist)
{
    CMD_REQUIRE_INIT();
    if (!k || !d_hist || nbins < 1 || !(hi > lo)) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (k->ev_cap <= 0) return cmd_set_error(CMD_ESTATE, "the event log is disabled");
    k_jump_hist<<<k->n_replicas, 256, 0, cmd_global().stream>>>(k->d_state, k->d_ev_dist, k->ev_cap,
                                                                 k->n_replicas, lo, nbins / (hi - lo),
                                                                 nbins, d_hist);
    CMD_LAUNCHED();
    return CMD_OK;
}

extern "C" int cmd_kmc_jump_histogram(const cmd_kmc *k, double lo, double hi, int nbins,
                                      int64_t *h_hist)
{
    CMD_REQUIRE_INIT();
    if (!h_hist || nbins < 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    void *d;
    int rc = cmd_scratch(3, (size_t)nbins * 8, &d);
    if (rc) return rc;
    cudaStream_t st = cmd_global().stream;
    CMD_CUDA(cudaMemsetAsync(d, 0, (size_t)nbins * 8, st));
    if ((rc = cmd_kmc_jump_histogram_dev(k, lo, hi, nbins, (unsigned long long *)d))) return rc;
    int64_t *tmp = (int64_t *)malloc((size_t)nbins * 8);
    if (!tmp) return cmd_set_error(CMD_ENOMEM, "out of host memory");
    cudaError_t e = cudaMemcpyAsync(tmp, d, (size_t)nbins * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) for (int b = 0; b < nbins; b++) h_hist[b] += tmp[b];
    free(tmp);
    CMD_CUDA(e);
    return CMD_OK;
}

extern "C" int cmd_kmc_get_observables(const cmd_kmc *k, int replica, int64_t capacity, int64_t *n,
                                       double *rows)
{
    CMD_REQUIRE_INIT();
    if (!k || replica < 0 || replica >= k->n_replicas || !n) return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    KmcState hs;
    CMD_CUDA(cudaMemcpyAsync(&hs, k->d_state + replica, sizeof(hs), cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    // only rows whose frames were flushed by an event carry a time stamp (MDMC.py:94-96)
    int64_t have = hs.pending_row;
    if (!rows) { *n = have; return CMD_OK; }   // query: how many rows there are
    if (have > capacity) have = capacity;
    *n = have;
    if (have > 0 && rows) {
        CMD_CUDA(cudaMemcpyAsync(rows, k->d_rows + (int64_t)replica * k->row_cap * 6, have * 48,
                                 cudaMemcpyDeviceToHost, st));
        CMD_CUDA(cudaStreamSynchronize(st));
    }
    return CMD_OK;
}

extern "C" int cmd_kmc_get_observable_rows(const cmd_kmc *k, int replica, int64_t first, int64_t count,
                                           double *rows)
{
    CMD_REQUIRE_INIT();
    if (!k || replica < 0 || replica >= k->n_replicas || first < 0 || count < 0 || (count && !rows))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    KmcState hs;
    CMD_CUDA(cudaMemcpyAsync(&hs, k->d_state + replica, sizeof(hs), cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    if (first + count > hs.pending_row)
        return cmd_set_error(CMD_EINVAL, "rows [%lld, %lld) asked for, %lld time-stamped rows exist",
                             (long long)first, (long long)(first + count), (long long)hs.pending_row);
    if (count) {
        CMD_CUDA(cudaMemcpyAsync(rows, k->d_rows + ((int64_t)replica * k->row_cap + first) * 6, count * 48,
                                 cudaMemcpyDeviceToHost, st));
        CMD_CUDA(cudaStreamSynchronize(st));
    }
    return CMD_OK;
}

extern "C" int64_t cmd_kmc_selection_fallbacks(const cmd_kmc *k)
{
    if (!k || !cmd_global().inited) return -1;
    unsigned long long v = 0;
    cudaStream_t st = cmd_global().stream;
    if (cudaMemcpyAsync(&v, k->d_ties + 1, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    cudaStreamSynchronize(st);
    return (int64_t)v;
}

// tools only: raw counter i of the tie / fallback / profile block
extern "C" int64_t cmd_kmc_debug_counter(const cmd_kmc *k, int i)
{
    if (!k || !cmd_global().inited || i < 0 || i >= 32) return -1;
    unsigned long long v = 0;
    cudaStream_t st = cmd_global().stream;
    if (cudaMemcpyAsync(&v, k->d_ties + i, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    cudaStreamSynchronize(st);
    return (int64_t)v;
}

// events that did not fit the log since the last cmd_kmc_set_event_log, summed over the replicas
__global__ void k_events_dropped(const KmcState *__restrict__ state, int n_replicas, int64_t ev_cap,
                                 unsigned long long *__restrict__ out)
{
    unsigned long long v = 0;
    for (int r = threadIdx.x; r < n_replicas; r += blockDim.x)
        if (state[r].log_pos > ev_cap) v += (unsigned long long)(state[r].log_pos - ev_cap);
    if (v) atomicAdd(out, v);
}

extern "C" int64_t cmd_kmc_events_dropped(const cmd_kmc *k)
{
    if (!k || !cmd_global().inited) return -1;
    if (k->ev_cap <= 0) return 0;
    cudaStream_t st = cmd_global().stream;
    void *buf;
    if (cmd_scratch(4, 8, &buf)) return -1;
    if (cudaMemsetAsync(buf, 0, 8, st) != cudaSuccess) return -1;
    k_events_dropped<<<1, 256, 0, st>>>(k->d_state, k->n_replicas, k->ev_cap, (unsigned long long *)buf);
    unsigned long long v = 0;
    if (cudaMemcpyAsync(&v, buf, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    cudaStreamSynchronize(st);
    return (int64_t)v;
}

extern "C" int64_t cmd_kmc_tie_count(const cmd_kmc *k)
{
    if (!k || !cmd_global().inited) return -1;
    unsigned long long v = 0;
    cudaStream_t st = cmd_global().stream;
    if (cudaMemcpyAsync(&v, k->d_ties, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    cudaStreamSynchronize(st);
    return (int64_t)v;
}
