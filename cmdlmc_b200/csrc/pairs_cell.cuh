// pairs_cell.cuh -- cell-list neighbour search for boxes too large for the one-CTA-per-frame
// dense kernel (row A7 of SURVEY.md section 8 at the sizes of configs C3 / C5).
//
// Same contract as pairs_dense.cuh: the list of get_topology_bruteforce (topology.py:55-72) --
// both directions, row-major, columns ascending, every distance and every `dist <= cutoff+buffer`
// decision in the reference's FP64 arithmetic -- but only pairs of atoms in adjacent cells are
// looked at.  Cells live in FRACTIONAL space (any cell shape): cell index along axis c is the top
// bits of the 2^-32 fixed-point fractional coordinate times nc[c], and nc[c] is chosen so that a
// cell is at least cutoff+buffer thick (perpendicular height / nc[c] >= radius), hence two atoms
// within the radius sit in the same or in adjacent cells.  An axis with fewer than 3 cells is
// collapsed to one cell (all its atoms are candidates; the exact stage finds the image).
//
//   k_cell_build   one CTA per frame: fixed-point coordinates, counting sort into cells
//   k_cell_pairs   one CTA per (x, y) column of cells: the column and its half shell staged in
//                  shared memory, every unordered pair once; a thread per home atom runs the
//                  filter (pairs_dense.cuh filter_pair), survivors go to a CTA-wide list that is
//                  evaluated exactly with every thread busy, hits appended to both rows of a
//                  fixed-capacity scratch
//   k_cell_scan    per frame: row counts -> row offsets (the LIL->COO order is row-major)
//   k_cell_emit    warp per 32 rows: rank of each entry inside its row (columns ascending),
//                  write of (start, dest, dist, omega); k_cell_rsum: ordered rate sums
#pragma once
#include "pairs_dense.cuh"

struct CellGrid {
    int nc[3];     // cells per fractional axis
    int span[3];   // 1: the axis is collapsed (only offset 0), 3: offsets -1, 0, +1
    int ncell;
};

__device__ __forceinline__ int cell_axis(int q, int nc)
{
    // (the 64-bit product form of this is mis-folded by nvcc 12.9 when the result is multiplied
    // by another grid dimension; __umulhi is not)
    return (int)__umulhi((unsigned)q, (unsigned)nc);
}

// grid.x = frames of this batch; frame = ids ? ids[first + blockIdx.x] : first + blockIdx.x
template <int KIND>
__global__ void __launch_bounds__(1024, 1)
k_cell_build(const __grid_constant__ BoxParams bx, const __grid_constant__ CellGrid cg,
             const double *__restrict__ frames, const int *__restrict__ ids,
             const int *__restrict__ n_ids, int first, int n, int4 *__restrict__ fxu,
             int *__restrict__ slot, int4 *__restrict__ sorted, int *__restrict__ cell_start)
{
    extern __shared__ int cnt[];   // [ncell + 1] counts -> exclusive starts, then 34 scan words
    int *scan = cnt + cg.ncell + 1;
    if (n_ids && first + (int)blockIdx.x >= *n_ids) return;
    const int64_t f = ids ? ids[first + blockIdx.x] : first + blockIdx.x;
    const int b = blockIdx.x, tid = threadIdx.x;
    const double *fr = frames + f * (int64_t)n * 3;
    int4 *my_fxu = fxu + (int64_t)b * n;
    int *my_slot = slot + (int64_t)b * n;
    for (int c = tid; c <= cg.ncell; c += blockDim.x) cnt[c] = 0;
    __syncthreads();
    for (int a = tid; a < n; a += blockDim.x) {
        const double x = __ldg(fr + 3 * a), y = __ldg(fr + 3 * a + 1), z = __ldg(fr + 3 * a + 2);
        int q[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double v = KIND == 0 ? (c == 0 ? x : c == 1 ? y : z) * bx.hinv[4 * c]
                                 : fma(bx.hinv[3 * c + 2], z, fma(bx.hinv[3 * c + 1], y, bx.hinv[3 * c] * x));
            q[c] = (int)(unsigned)(unsigned long long)__double2ll_rn(v * 4294967296.0);
        }
        const int cell = (cell_axis(q[0], cg.nc[0]) * cg.nc[1] + cell_axis(q[1], cg.nc[1])) * cg.nc[2] +
                         cell_axis(q[2], cg.nc[2]);
        my_fxu[a] = make_int4(q[0], q[1], q[2], cell);
        my_slot[a] = atomicAdd(&cnt[cell], 1);
    }
    __syncthreads();
    // exclusive scan of the counts, 1024 cells per round
    int carry = 0;
    for (int c0 = 0; c0 < cg.ncell; c0 += blockDim.x) {
        const int c = c0 + tid;
        const int v = c < cg.ncell ? cnt[c] : 0;
        int tot;
        const int ex = block_exclusive_scan(v, scan, &scan[33]);
        tot = scan[33];
        if (c < cg.ncell) cnt[c] = carry + ex;
        carry += tot;
        __syncthreads();
    }
    if (tid == 0) cnt[cg.ncell] = carry;
    __syncthreads();
    int *cs = cell_start + (int64_t)b * (cg.ncell + 1);
    for (int c = tid; c <= cg.ncell; c += blockDim.x) cs[c] = cnt[c];
    int4 *my_sorted = sorted + (int64_t)b * n;
    for (int a = tid; a < n; a += blockDim.x) {
        const int4 p = my_fxu[a];
        my_sorted[cnt[p.w] + my_slot[a]] = make_int4(p.x, p.y, p.z, a);
    }
}

#define CELL_TPB 128          // threads per CTA: one home atom each
#define CELL_STAGE_CAP 768    // atoms of the <= 5 staged columns
#define CELL_PLIST_CAP 512    // filtered pairs per warp and exact round

// exact evaluation of one filtered pair (reference arithmetic); a hit is appended to BOTH rows
// (one atomicAdd per row on its counter; rows are unordered here, k_cell_emit sorts them)
template <int KIND, bool IMAGES>
__device__ __forceinline__ void cell_exact_pair(const BoxParams &bx, const double *__restrict__ fr,
                                                int i, int j, double rc, double t2, int rowcap,
                                                int64_t row0, int *__restrict__ rcnt,
                                                int *__restrict__ tmp_j, double *__restrict__ tmp_d,
                                                int *__restrict__ cap_need, unsigned long long &my_ties)
{
    const double pa[3] = {__ldg(fr + 3 * i), __ldg(fr + 3 * i + 1), __ldg(fr + 3 * i + 2)};
    const double pb[3] = {__ldg(fr + 3 * j), __ldg(fr + 3 * j + 1), __ldg(fr + 3 * j + 2)};
    double d[3], d2;
    if (KIND == 0) {
        diff_ortho_exact(bx, pa, pb, d);
        d2 = norm2_exact(d);
    } else {
        diff_general_norm_exact(bx, pa, pb, d);
        d2 = IMAGES ? min_image_norm2_kept(bx, d) : fmin(1e6, norm2_exact(d));
    }
    const double dist = convert_distance(bx, sqrt(d2));
    const bool hit = (bx.conv == CMD_CONV_NONE ? d2 <= t2 : dist <= rc) && dist != 0.0;
    if (fabs(dist - rc) <= 1e-11 * rc) my_ties++;
    if (hit) {
        const int pi = atomicAdd(rcnt + i, 1), pj = atomicAdd(rcnt + j, 1);
        if (pi < rowcap) {
            const int64_t at = (row0 + i) * rowcap + pi;
            tmp_j[at] = j; tmp_d[at] = dist;
        }
        if (pj < rowcap) {
            const int64_t at = (row0 + j) * rowcap + pj;
            tmp_j[at] = i; tmp_d[at] = dist;
        }
        if (pi >= rowcap || pj >= rowcap) atomicMax(cap_need, max(pi, pj) + 1);
    }
}

// two filtered pairs at once (the second only if `two`): loads first, then both FP64 chains, then
// the appends
template <int KIND, bool IMAGES>
__device__ __forceinline__ void cell_exact_two(const BoxParams &bx, const double *__restrict__ fr,
                                               int2 p0, int2 p1, bool two, double rc, double t2,
                                               int rowcap, int64_t row0, int *__restrict__ rcnt,
                                               int *__restrict__ tmp_j, double *__restrict__ tmp_d,
                                               int *__restrict__ cap_need, unsigned long long &my_ties)
{
    const int2 pr[2] = {p0, p1};
    double pa[2][3], pb[2][3], d2[2], dist[2];
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            pa[q][c] = __ldg(fr + 3 * pr[q].x + c);
            pb[q][c] = __ldg(fr + 3 * pr[q].y + c);
        }
#pragma unroll
    for (int q = 0; q < 2; q++) {
        double d[3];
        if (KIND == 0) {
            diff_ortho_exact(bx, pa[q], pb[q], d);
            d2[q] = norm2_exact(d);
        } else {
            diff_general_norm_exact(bx, pa[q], pb[q], d);
            d2[q] = IMAGES ? min_image_norm2_kept(bx, d) : fmin(1e6, norm2_exact(d));
        }
    }
#pragma unroll
    for (int q = 0; q < 2; q++) dist[q] = convert_distance(bx, sqrt(d2[q]));
    bool hit[2];
    int pi[2] = {0, 0}, pj[2] = {0, 0};
#pragma unroll
    for (int q = 0; q < 2; q++) {
        hit[q] = (q == 0 || two) && (bx.conv == CMD_CONV_NONE ? d2[q] <= t2 : dist[q] <= rc) && dist[q] != 0.0;
        if ((q == 0 || two) && fabs(dist[q] - rc) <= 1e-11 * rc) my_ties++;
        if (hit[q]) { pi[q] = atomicAdd(rcnt + pr[q].x, 1); pj[q] = atomicAdd(rcnt + pr[q].y, 1); }
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
        if (!hit[q]) continue;
        if (pi[q] < rowcap) {
            const int64_t at = (row0 + pr[q].x) * rowcap + pi[q];
            tmp_j[at] = pr[q].y; tmp_d[at] = dist[q];
        }
        if (pj[q] < rowcap) {
            const int64_t at = (row0 + pr[q].y) * rowcap + pj[q];
            tmp_j[at] = pr[q].x; tmp_d[at] = dist[q];
        }
        if (pi[q] >= rowcap || pj[q] >= rowcap) atomicMax(cap_need, max(pi[q], pj[q]) + 1);
    }
}

// the pair list is full: evaluate on the spot (cold; kept out of line, it has a dozen call sites)
template <int KIND, bool IMAGES>
__device__ __noinline__ void cell_exact_pair_cold(const BoxParams &bx, const double *__restrict__ fr,
                                                  int i, int j, double rc, double t2, int rowcap,
                                                  int64_t row0, int *__restrict__ rcnt,
                                                  int *__restrict__ tmp_j, double *__restrict__ tmp_d,
                                                  int *__restrict__ cap_need, unsigned long long *ties)
{
    unsigned long long my = 0;
    cell_exact_pair<KIND, IMAGES>(bx, fr, i, j, rc, t2, rowcap, row0, rcnt, tmp_j, tmp_d, cap_need, my);
    if (my) atomicAdd(ties, my);
}

// grid = (columns * segments, frames of the batch), block = 32 .. CELL_TPB threads (about one per
// atom of a column).  One CTA per (x, y) column
// of cells (or per z segment of `zseg` cells of it when the batch is too small to fill the GPU
// with whole columns).  The sorted order runs z fastest, so a column is ONE contiguous run of the
// sorted atoms.  The CTA stages its own column and the four columns of the "half shell"
// ((0,+1), (+1,-1), (+1,0), (+1,+1); offsets (dx, dy, dz) > 0 lexicographically) in shared
// memory -- every unordered pair of adjacent cells is then looked at exactly once.  One thread
// per home atom walks the z window of each staged column through the FP32 / fixed-point filter
// and appends the survivors to a CTA-wide pair list; the list is evaluated in the reference's FP64
// arithmetic with every thread busy.
template <int KIND, bool IMAGES>
__global__ void __launch_bounds__(CELL_TPB)
k_cell_pairs(const __grid_constant__ BoxParams bx, const __grid_constant__ FilterParams fp,
             const __grid_constant__ CellGrid cg, const double *__restrict__ frames,
             const int *__restrict__ ids, const int *__restrict__ n_ids, int first, int n,
             double rc, double t2, int rowcap, int zseg, const int4 *__restrict__ sorted,
             const int *__restrict__ cell_start, int *__restrict__ rowcount,
             int *__restrict__ tmp_j, double *__restrict__ tmp_d, int *__restrict__ cap_need,
             unsigned long long *__restrict__ ties)
{
    __shared__ int4 stage[CELL_STAGE_CAP];
    extern __shared__ int2 plist_dyn[];    // [warps][CELL_PLIST_CAP]
    __shared__ int cs_s[5][68];          // cell starts of the staged columns (nc[2] <= 64)
    __shared__ int col_of[5], s_off[5], s_np[CELL_TPB / 32], s_next, s_staged;
    if (n_ids && first + (int)blockIdx.y >= *n_ids) return;
    const int64_t f = ids ? ids[first + blockIdx.y] : first + blockIdx.y;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int nz = cg.nc[2];
    const int nseg = (nz + zseg - 1) / zseg;
    const int column = blockIdx.x / nseg, seg = blockIdx.x - column * nseg;
    const int cy = column % cg.nc[1], cx = column / cg.nc[1];
    const int z0 = seg * zseg, z1 = min(z0 + zseg, nz);
    const double *fr = frames + f * (int64_t)n * 3;
    const int4 *srt = sorted + (int64_t)b * n;
    const int *cs = cell_start + (int64_t)b * (cg.ncell + 1);
    int *rcnt = rowcount + (int64_t)b * n;
    const int64_t row0 = (int64_t)b * n;

    // the staged columns: 0 = own, then the half shell (-1: absent)
    if (tid == 0) {
        const int yw = cg.span[1] == 3 ? 1 : 0;
        const int yp = cy + 1 >= cg.nc[1] ? 0 : cy + 1, ym = cy - 1 < 0 ? cg.nc[1] - 1 : cy - 1;
        const int xp = cx + 1 >= cg.nc[0] ? 0 : cx + 1;
        col_of[0] = (cx * cg.nc[1] + cy) * nz;
        col_of[1] = yw ? (cx * cg.nc[1] + yp) * nz : -1;
        col_of[2] = cg.span[0] == 3 && yw ? (xp * cg.nc[1] + ym) * nz : -1;
        col_of[3] = cg.span[0] == 3 ? (xp * cg.nc[1] + cy) * nz : -1;
        col_of[4] = cg.span[0] == 3 && yw ? (xp * cg.nc[1] + yp) * nz : -1;
    }
    if (tid < CELL_TPB / 32) s_np[tid] = 0;
    __syncthreads();
    for (int k = tid; k < 5 * (nz + 1); k += blockDim.x) {
        const int c = k / (nz + 1), z = k - c * (nz + 1);
        cs_s[c][z] = col_of[c] >= 0 ? __ldg(cs + col_of[c] + z) : 0;
    }
    __syncthreads();
    if (tid == 0) {   // whole columns are staged; a segment CTA or an overfull one reads global memory
        int tot = 0;
        for (int c = 0; c < 5; c++) { s_off[c] = tot; tot += col_of[c] >= 0 ? cs_s[c][nz] - cs_s[c][0] : 0; }
        s_staged = (nseg == 1 && tot <= CELL_STAGE_CAP) ? 1 : 0;
        s_next = cs_s[0][z0];
    }
    __syncthreads();
    const bool staged = s_staged != 0;
    if (staged) {
        for (int c = 0; c < 5; c++) {
            if (col_of[c] < 0) continue;
            const int g0 = cs_s[c][0], cnt = cs_s[c][nz] - g0;
            for (int q = tid; q < cnt; q += blockDim.x) stage[s_off[c] + q] = __ldg(srt + g0 + q);
        }
    }
    __syncthreads();   // the last CTA-wide barrier: from here on the warps run on their own

    const int zw = cg.span[2] == 3 ? 1 : 0;
    const bool zall = 2 * zw + 1 >= nz;           // the z window is the whole column
    const int home_hi = cs_s[0][z1];
    int2 *plist = plist_dyn + wp * CELL_PLIST_CAP;
    int *my_np = &s_np[wp];
    unsigned long long my_ties = 0;
    for (;;) {
        int hb = 0;
        if (lane == 0) hb = atomicAdd(&s_next, 32);   // the warp's next 32 home atoms
        hb = __shfl_sync(0xffffffffu, hb, 0);
        if (hb >= home_hi) break;
        const int k = hb + lane;
        if (k < home_hi) {
            const int4 me = staged ? stage[k - cs_s[0][0]] : __ldg(srt + k);
            int z = z0;
            while (cs_s[0][z + 1] <= k) z++;        // the home atom's cell
            // <= 2 index ranges per column (the z window wraps at the column ends); ONE loop site
#pragma unroll 1
            for (int r = 0; r < 10; r++) {
                const int c = r >> 1;
                const bool second = r & 1;
                if (col_of[c] < 0) continue;
                const int *cc = cs_s[c];
                int lo = 0, hi = 0;
                if (c == 0) {
                    // own column: the rest of the home cell, then cell z + 1 (wrapped)
                    if (!zw) { if (!second) { lo = k + 1; hi = cc[z + 1]; } }
                    else if (zall) {   // three cells: z + 1 is distinct, z + 2 = z - 1 is that cell's job
                        const int zn = z + 1 < nz ? z + 1 : 0;
                        lo = second ? cc[zn] : k + 1;
                        hi = second ? cc[zn + 1] : cc[z + 1];
                    } else if (z + 1 < nz) { if (!second) { lo = k + 1; hi = cc[z + 2]; } }
                    else { lo = second ? cc[0] : k + 1; hi = second ? cc[1] : cc[nz]; }
                } else {
                    // half-shell column: cells z - zw .. z + zw (wrapped)
                    if (zall) { if (!second) { lo = cc[0]; hi = cc[nz]; } }
                    else if (z - 1 < 0) { lo = second ? cc[0] : cc[nz - 1]; hi = second ? cc[z + 2] : cc[nz]; }
                    else if (z + 1 >= nz) { lo = second ? cc[0] : cc[z - 1]; hi = second ? cc[1] : cc[nz]; }
                    else if (!second) { lo = cc[z - 1]; hi = cc[z + 2]; }
                }
                const int4 *bp = staged ? stage + (s_off[c] - cc[0]) : srt;
                for (int q = lo; q < hi; q++) {
                    const int4 cand = bp[q];
                    if (filter_pair<KIND, IMAGES>(fp, me, cand)) {
                        const int e = atomicAdd(my_np, 1);
                        if (e < CELL_PLIST_CAP) plist[e] = make_int2(me.w, cand.w);
                        else cell_exact_pair_cold<KIND, IMAGES>(bx, fr, me.w, cand.w, rc, t2, rowcap, row0,
                                                               rcnt, tmp_j, tmp_d, cap_need, ties);
                    }
                }
            }
        }
        __syncwarp();
        const int np = min(*my_np, CELL_PLIST_CAP);
        // two pairs per lane and trip: the coordinate loads and the row-counter atomics of one
        // overlap the arithmetic of the other
        for (int e = lane; e < np; e += 64) {
            const bool two = e + 32 < np;
            const int2 pr0 = plist[e], pr1 = plist[two ? e + 32 : e];
            cell_exact_two<KIND, IMAGES>(bx, fr, pr0, pr1, two, rc, t2, rowcap, row0, rcnt, tmp_j, tmp_d,
                                         cap_need, my_ties);
        }
        __syncwarp();
        if (lane == 0) *my_np = 0;
        __syncwarp();
    }
    if (my_ties) atomicAdd(ties, my_ties);
}

// per frame: rowcount -> exclusive row offsets (n + 1 entries), frame total -> out_counts
__global__ void __launch_bounds__(1024, 1)
k_cell_scan(const int *__restrict__ ids, const int *__restrict__ n_ids, int first, int n,
            int64_t stride, const int *__restrict__ rowcount, int *__restrict__ rowoff,
            int *__restrict__ out_counts, uint8_t *__restrict__ out_rebuilt,
            double *__restrict__ out_rate_sum, int *__restrict__ out_rowoff,
            int *__restrict__ err)
{
    __shared__ int scan[40];
    if (n_ids && first + (int)blockIdx.x >= *n_ids) return;
    const int64_t f = ids ? ids[first + blockIdx.x] : first + blockIdx.x;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int *rcnt = rowcount + (int64_t)b * n;
    int *ro = rowoff + (int64_t)b * (n + 1);
    int *ro2 = out_rowoff ? out_rowoff + f * (int64_t)cmd_ro_pitch(n) : nullptr;
    int carry = 0;
    for (int c0 = 0; c0 < n; c0 += blockDim.x) {
        const int c = c0 + tid;
        const int v = c < n ? rcnt[c] : 0;
        const int ex = block_exclusive_scan(v, scan, &scan[33]);
        const int tot = scan[33];
        if (c < n) {
            ro[c] = carry + ex;
            if (ro2) ro2[c] = carry + ex;
        }
        carry += tot;
        __syncthreads();
    }
    if (tid == 0) {
        ro[n] = carry;
        if (ro2) ro2[n] = carry;
        out_counts[f] = carry > stride ? -carry : carry;
        if (out_rebuilt) out_rebuilt[f] = 1;
        if (out_rate_sum) out_rate_sum[f] = 0.0;   // k_cell_emit accumulates into it
        if (carry > stride) atomicMax(err, carry);
    }
}

// grid = (ceil(n / 256), frames of the batch), block = 256: a warp writes 32 consecutive rows, one
// output entry per lane and trip.  The rows arrive unordered; the entry's place inside its row is
// the number of smaller columns in that row (columns are distinct), counted by the lane itself --
// the neighbouring lanes read the same few rows, so the loop runs out of L1.  The writes of a warp
// cover one contiguous range of the four output arrays.
__global__ void __launch_bounds__(256)
k_cell_emit(const __grid_constant__ RateParams rp, const int *__restrict__ ids,
            const int *__restrict__ n_ids, int first, int n, int64_t stride, int rowcap,
            const int *__restrict__ rowoff, const int *__restrict__ tmp_j,
            const double *__restrict__ tmp_d,
            int *__restrict__ out_start,
            int *__restrict__ out_dest, double *__restrict__ out_dist,
            double *__restrict__ out_omega, double *__restrict__ part)
{
    if (n_ids && first + (int)blockIdx.y >= *n_ids) return;
    const int64_t f = ids ? ids[first + blockIdx.y] : first + blockIdx.y;
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int wglob = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int r0 = wglob * 32;
    double *my_part = part ? part + (int64_t)b * (gridDim.x * 8) + wglob : nullptr;
    if (r0 >= n) { if (my_part && lane == 0) *my_part = 0.0; return; }
    const int *ro = rowoff + (int64_t)b * (n + 1);
    if (ro[n] > stride) { if (my_part && lane == 0) *my_part = 0.0; return; }   // overflow: k_cell_scan
    const int myrow = min(r0 + lane, n);
    const int myoff = ro[myrow];                       // offsets of rows r0 .. r0+31 (clamped)
    const int g0 = __shfl_sync(0xffffffffu, myoff, 0);
    const int g1 = ro[min(r0 + 32, n)];
    const int64_t base = f * stride;
    double rsum = 0.0;
    for (int gb = g0; gb < g1; gb += 32) {   // warp-uniform trip count: the shuffles need all lanes
        const int g = gb + lane;
        // the row holding entry g: last row of the 32 with offset <= g
        int lo = 0;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const int probe = __shfl_sync(0xffffffffu, myoff, min(lo + s, 31));
            if (lo + s <= 31 && probe <= g) lo += s;
        }
        const int off_lo = __shfl_sync(0xffffffffu, myoff, lo);
        const int off_nx = __shfl_sync(0xffffffffu, myoff, min(lo + 1, 31));
        if (g < g1) {
            const int r = r0 + lo;
            const int cnt = (lo < 31 ? off_nx : g1) - off_lo;
            const int64_t rbase = ((int64_t)b * n + r) * rowcap;
            const int j = tmp_j[rbase + (g - off_lo)];
            const double dist = tmp_d[rbase + (g - off_lo)];
            int rank = 0;
#pragma unroll 4
            for (int q = 0; q < cnt; q++) rank += __ldg(tmp_j + rbase + q) < j;
            const double om = rate_eval(rp, dist, 0.0);
            rsum += om;
            const int64_t at = base + off_lo + rank;
            out_start[at] = r; out_dest[at] = j;
            out_dist[at] = dist; out_omega[at] = om;
        }
    }
    if (my_part) {
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_down_sync(0xffffffffu, rsum, o);
        if (lane == 0) *my_part = rsum;
    }
}

// per frame: the warps' partial rate sums in a fixed order (deterministic, unlike atomics)
__global__ void __launch_bounds__(256)
k_cell_rsum(const int *__restrict__ ids, const int *__restrict__ n_ids, int first, int nparts,
            const double *__restrict__ part, double *__restrict__ out_rate_sum)
{
    __shared__ double red[256];
    if (n_ids && first + (int)blockIdx.x >= *n_ids) return;
    const int64_t f = ids ? ids[first + blockIdx.x] : first + blockIdx.x;
    const double *p = part + (int64_t)blockIdx.x * nparts;
    double acc = 0.0;
    for (int k = threadIdx.x; k < nparts; k += 256) acc += p[k];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out_rate_sum[f] = red[0];
}
