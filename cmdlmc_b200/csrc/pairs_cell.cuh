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
//   k_cell_pairs   one CTA per (x, y) column of cells, a warp per home cell: lanes hold the
//                  candidates of the half shell (every unordered pair of adjacent cells once),
//                  the home atoms pass by in a warp-uniform loop through the FP32 / fixed-point
//                  filter (pairs_dense.cuh filter_pair); survivors are compacted by
//                  __ballot_sync / __popc into a per-warp list, evaluated 64 at a time in the
//                  reference's FP64 arithmetic, hits appended to both rows of a fixed-capacity
//                  scratch
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

#define CELL_WARPS 4          // warps per CTA: each walks home cells of the CTA's column on its own
#define CELL_TPB (32 * CELL_WARPS)
#ifndef CELL_UNROLL
#define CELL_UNROLL 1
#endif
#ifndef CELL_MINB
#define CELL_MINB 8            // resident CTAs per SM the register budget is cut for
#endif
#define CELL_PLIST 96         // filtered pairs a warp holds: drained 64 at a time, < 32 appended per step

// numpyatom.pyx:33-42 for a difference that is at most 1.5 box lengths long (everything the
// filter lets through): the reference's loops run at most once, in one direction.  `cold` is
// raised for anything else (the caller then takes wrap_ortho_exact).
__device__ __forceinline__ double wrap_ortho_once(double d, double L, double hL, bool &cold)
{
    cold |= !(fabs(d) <= 1.5 * L);
    if (d < -hL) d = __dadd_rn(d, L);
    else if (d > hL) d = __dadd_rn(d, -L);
    return d;
}

// Exact evaluation (reference arithmetic) of the first `cnt` (<= 64) pairs of a warp's list, two per
// lane: loads first, then both FP64 chains, then the appends.  A hit is appended to BOTH rows
// (one atomicAdd per row on its counter; rows are unordered here, k_cell_emit sorts them).
template <int KIND, bool IMAGES>
__device__ __forceinline__ void cell_exact_drain(const BoxParams &bx, const double *__restrict__ fr,
                                                 const int2 *plist, int cnt, double rc, double t2,
                                                 int rowcap, unsigned row0, int *__restrict__ rcnt,
                                                 int *__restrict__ tmp_j, double *__restrict__ tmp_d,
                                                 int *__restrict__ cap_need,
                                                 unsigned long long *__restrict__ ties)
{
    const int lane = threadIdx.x & 31;
    if (lane >= cnt) return;
    const bool two = lane + 32 < cnt;
    const int2 pr[2] = {plist[lane], plist[two ? lane + 32 : lane]};
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
            bool cold = false;
#pragma unroll
            for (int c = 0; c < 3; c++)
                d[c] = wrap_ortho_once(__dadd_rn(pb[q][c], -pa[q][c]), bx.L[c], bx.hL[c], cold);
            if (cold) diff_ortho_exact(bx, pa[q], pb[q], d);
            d2[q] = norm2_exact(d);
        } else {
            if (bx.sparse == 2) diff_general_norm_sp<2>(bx, pa[q], pb[q], d);
            else if (bx.sparse == 1) diff_general_norm_sp<1>(bx, pa[q], pb[q], d);
            else diff_general_norm_exact(bx, pa[q], pb[q], d);
            d2[q] = IMAGES ? min_image_norm2_kept(bx, d) : fmin(1e6, norm2_exact(d));
        }
    }
#pragma unroll
    for (int q = 0; q < 2; q++) dist[q] = convert_distance(bx, sqrt(d2[q]));
    bool hit[2];
    int pi[2] = {0, 0}, pj[2] = {0, 0};
    unsigned nt = 0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const bool live = q == 0 || two;
        hit[q] = live && (bx.conv == CMD_CONV_NONE ? d2[q] <= t2 : dist[q] <= rc) && dist[q] != 0.0;
        nt += live && fabs(dist[q] - rc) <= 1e-11 * rc;
        if (hit[q]) { pi[q] = atomicAdd(rcnt + pr[q].x, 1); pj[q] = atomicAdd(rcnt + pr[q].y, 1); }
    }
    if (nt) atomicAdd(ties, (unsigned long long)nt);
#pragma unroll
    for (int q = 0; q < 2; q++) {
        if (!hit[q]) continue;
        // scratch indices fit 32 bits: a batch is sized to at most 2^30 bytes of 12-byte entries
        if (pi[q] < rowcap) {
            const unsigned at = (row0 + (unsigned)pr[q].x) * (unsigned)rowcap + (unsigned)pi[q];
            tmp_j[at] = pr[q].y; tmp_d[at] = dist[q];
        }
        if (pj[q] < rowcap) {
            const unsigned at = (row0 + (unsigned)pr[q].y) * (unsigned)rowcap + (unsigned)pj[q];
            tmp_j[at] = pr[q].x; tmp_d[at] = dist[q];
        }
        if (pi[q] >= rowcap || pj[q] >= rowcap) atomicMax(cap_need, max(pi[q], pj[q]) + 1);
    }
}

// grid = (columns * segments, frames of the batch), block = CELL_TPB.  One CTA per (x, y) column of
// cells (or per z segment of `zseg` cells of it when the batch is too small to fill the GPU with
// whole columns); its warps take the column's home cells in turn and never meet at a barrier
// after the prologue.
//
// The sorted order runs z fastest, so a column is ONE contiguous run of the sorted atoms and the
// cells z-1 .. z+1 of a column are one run of it (cyclic at the column ends).  The candidates of a
// home cell are the rest of the "half shell": its own column from the home cell on (cells z, z+1;
// inside the home cell only the atoms BEHIND the home atom), and cells z-1 .. z+1 of the columns
// (0,+1), (+1,-1), (+1,0), (+1,+1) -- every unordered pair of adjacent cells exactly once.
//
// Lanes hold CANDIDATES (one per lane and round, in registers); the home atoms of the cell are
// walked in a warp-uniform loop, their coordinates broadcast from shared memory.  Each step runs the
// FP32 / fixed-point filter (pairs_dense.cuh filter_pair) on 32 pairs with no divergence; the
// survivors are compacted with __ballot_sync / __popc into the warp's pair list, and whenever 64
// are waiting they are evaluated in the reference's FP64 arithmetic, two per lane (the memory
// latency of that stage -- coordinate loads, row-counter atomics -- hides under the filter work
// of the SM's other warps; a separate exact kernel was 1.4x slower, profiles/README.md).
template <int KIND, bool IMAGES>
__global__ void __launch_bounds__(CELL_TPB, CELL_MINB)
k_cell_pairs(const __grid_constant__ BoxParams bx, const __grid_constant__ FilterParams fp,
             const __grid_constant__ CellGrid cg, const double *__restrict__ frames,
             const int *__restrict__ ids, const int *__restrict__ n_ids, int first, int n,
             double rc, double t2, int rowcap, int zseg, const int4 *__restrict__ sorted,
             const int *__restrict__ cell_start, int *__restrict__ rowcount,
             int *__restrict__ tmp_j, double *__restrict__ tmp_d, int *__restrict__ cap_need,
             unsigned long long *__restrict__ ties)
{
    __shared__ int4 wbuf_s[CELL_WARPS][32];          // the home atoms in flight
    __shared__ int2 plist_s[CELL_WARPS][CELL_PLIST];
    __shared__ int4 desc_s[CELL_WARPS][6];   // [c] = (ring start, first slot, column begin, column end); [5] = first slots of columns 1..4
    __shared__ int cs_s[5][68];              // cell starts of the five columns (nc[2] <= 64)
    __shared__ int colbase_s[5];
    if (n_ids && first + (int)blockIdx.y >= *n_ids) return;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
    const int nz = cg.nc[2];
    const int nseg = (nz + zseg - 1) / zseg;
    const int column = blockIdx.x / nseg, seg = blockIdx.x - column * nseg;
    const int cy = column % cg.nc[1], cx = column / cg.nc[1];
    const int z0 = seg * zseg, z1 = min(z0 + zseg, nz);
    const int4 *srt = sorted + (int64_t)b * n;
    const int *cs = cell_start + (int64_t)b * (cg.ncell + 1);
    const int64_t f = ids ? ids[first + blockIdx.y] : first + blockIdx.y;
    const double *fr = frames + f * (int64_t)n * 3;
    int *rcnt = rowcount + (int64_t)b * n;
    const unsigned row0 = (unsigned)b * (unsigned)n;

    // the five columns: 0 = own, then the half shell (-1: absent)
    if (tid < 5) {
        const int yw = cg.span[1] == 3 ? 1 : 0, xw = cg.span[0] == 3 ? 1 : 0;
        const int yp = cy + 1 >= cg.nc[1] ? 0 : cy + 1, ym = cy - 1 < 0 ? cg.nc[1] - 1 : cy - 1;
        const int xp = cx + 1 >= cg.nc[0] ? 0 : cx + 1;
        int cb = -1;
        if (tid == 0) cb = (cx * cg.nc[1] + cy) * nz;
        if (tid == 1 && yw) cb = (cx * cg.nc[1] + yp) * nz;
        if (tid == 2 && xw && yw) cb = (xp * cg.nc[1] + ym) * nz;
        if (tid == 3 && xw) cb = (xp * cg.nc[1] + cy) * nz;
        if (tid == 4 && xw && yw) cb = (xp * cg.nc[1] + yp) * nz;
        colbase_s[tid] = cb;
    }
    __syncthreads();
    for (int k = tid; k < 5 * (nz + 1); k += CELL_TPB) {
        const int c = k / (nz + 1), z = k - c * (nz + 1);
        cs_s[c][z] = colbase_s[c] >= 0 ? __ldg(cs + colbase_s[c] + z) : 0;
    }
    __syncthreads();

    int4 *wbuf = wbuf_s[wp];
    int2 *plist = plist_s[wp];
    int4 *desc = desc_s[wp];
    int np = 0;                      // filtered pairs waiting in the list (warp-uniform)
    const int *ccs = cs_s[lane < 5 ? lane : 0];
    const bool have = lane < 5 && colbase_s[lane < 5 ? lane : 0] >= 0;
    const int cbeg = ccs[0], cend = ccs[nz];

    for (int z = z0 + wp; z < z1; z += CELL_WARPS) {
        // run of column `lane` for home cell z: a ring segment [start, start + len) of [cbeg, cend)
        int start = 0, len = 0;
        if (have) {
            if (nz < 3) { start = cbeg; len = cend - cbeg; }
            else if (lane == 0) {
                const int zn = z + 1 < nz ? z + 1 : 0;
                start = ccs[z];
                len = (ccs[z + 1] - start) + (ccs[zn + 1] - ccs[zn]);
            } else {
                const int zm = z - 1 < 0 ? nz - 1 : z - 1, zp = z + 1 < nz ? z + 1 : 0;
                start = ccs[zm];
                len = (ccs[zm + 1] - start) + (ccs[z + 1] - ccs[z]) + (ccs[zp + 1] - ccs[zp]);
            }
        }
        const int h0 = cs_s[0][z], nA = cs_s[0][z + 1] - h0;
        if (nA == 0) continue;       // warp-uniform
        int inc = len;               // lanes >= 5 hold 0
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const int v = __shfl_up_sync(full, inc, o);
            if (lane >= o) inc += v;
        }
        const int C = __shfl_sync(full, inc, 4);
        __syncwarp();
        if (lane < 5) desc[lane] = make_int4(start, inc - len, cbeg, cend);
        if (lane < 4) ((int *)&desc[5])[lane] = inc;
        __syncwarp();
        const int4 bounds = desc[5];
        // the home atoms go to shared memory (one broadcast load per step), 32 at a time
        for (int a0 = 0; a0 < nA; a0 += 32) {
            const int na = min(nA - a0, 32);
            __syncwarp();
            if (lane < na) wbuf[lane] = __ldg(srt + h0 + a0 + lane);
            __syncwarp();
            // one candidate per lane and round, held in registers while the home atoms pass by
            for (int s = lane; s - lane < C; s += 32) {
                int4 cand = make_int4(0, 0, 0, -1);
                int alim = 0;            // home atoms (by position in the cell) the candidate pairs with: those before alim
                if (s < C) {
                    const int c = (s >= bounds.x) + (s >= bounds.y) + (s >= bounds.z) + (s >= bounds.w);
                    const int4 d = desc[c];
                    int p = d.x + (s - d.y);
                    if (p >= d.w) p -= d.w - d.z;
                    cand = __ldg(srt + p);
                    // slot s < nA is home atom s itself: it pairs with the home atoms before it
                    alim = (s < nA ? s : nA) - a0;
                }
                const int amax = __reduce_max_sync(full, alim);   // steps anybody needs (<= na)
#if CELL_UNROLL == 2
#pragma unroll 2
#else
#pragma unroll 1
#endif
                for (int a = 0; a < min(amax, na); a++) {
                    const int4 me = wbuf[a];
                    const bool ok = filter_pair<KIND, IMAGES>(fp, me, cand) & (a < alim);
                    const unsigned hb = __ballot_sync(full, ok);
                    CMD_CHECK(np + 32 <= CELL_PLIST + 1 && (!ok || (me.w >= 0 && me.w < n && cand.w >= 0 && cand.w < n)));
                    if (ok) plist[np + __popc(hb & lt)] = make_int2(me.w, cand.w);
                    np += __popc(hb);
                    if (np >= 64) {
                        __syncwarp();
                        cell_exact_drain<KIND, IMAGES>(bx, fr, plist, 64, rc, t2, rowcap, row0, rcnt, tmp_j,
                                                       tmp_d, cap_need, ties);
                        __syncwarp();
                        np -= 64;
                        if (lane < np) { const int2 e = plist[64 + lane]; plist[lane] = e; }
                        __syncwarp();
                    }
                }
            }
        }
    }
    if (np > 0) {
        __syncwarp();
        cell_exact_drain<KIND, IMAGES>(bx, fr, plist, np, rc, t2, rowcap, row0, rcnt, tmp_j, tmp_d,
                                       cap_need, ties);
    }
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
    // No empty row among the 32 (the rule): the row of entry g follows from ONE warp reduction per
    // trip -- every row lane drops a bit at the entry its row starts with, entry g belongs to the
    // row `rows starting at or before gb` + `bits in (gb, g]`.  An empty row shares its start with
    // the next one (the bits would merge): then a binary search over the offsets.
    const int nxoff = __shfl_down_sync(0xffffffffu, myoff, 1);
    const bool dense_rows = __ballot_sync(0xffffffffu, lane < 31 && r0 + lane < n && nxoff == myoff) == 0u;
    const unsigned le_mask = lane == 31 ? 0xfffffffeu : (((2u << lane) - 1u) & ~1u);
    for (int gb = g0; gb < g1; gb += 32) {   // warp-uniform trip count: the shuffles need all lanes
        const int g = gb + lane;
        int lo;
        if (dense_rows) {
            const int rel = myoff - gb;
            const unsigned starts = __reduce_or_sync(0xffffffffu, rel > 0 && rel < 32 ? 1u << rel : 0u);
            const int cur = __popc(__ballot_sync(0xffffffffu, rel <= 0)) - 1;
            lo = cur + __popc(starts & le_mask);
        } else {
            lo = 0;   // last row of the 32 with offset <= g
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const int probe = __shfl_sync(0xffffffffu, myoff, min(lo + s, 31));
                if (lo + s <= 31 && probe <= g) lo += s;
            }
        }
        const int off_lo = __shfl_sync(0xffffffffu, myoff, lo);
        const int off_nx = __shfl_sync(0xffffffffu, myoff, min(lo + 1, 31));
        if (g < g1) {
            const int r = r0 + lo;
            const int cnt = (lo < 31 ? off_nx : g1) - off_lo;
            const int64_t rbase = ((int64_t)b * n + r) * rowcap;
            const int j = tmp_j[rbase + (g - off_lo)];
            const double dist = tmp_d[rbase + (g - off_lo)];
            // rank = smaller columns in the row; the row is 16-byte aligned (rowcap % 4 == 0)
            const int4 *rj = (const int4 *)(tmp_j + rbase);
            int rank = 0;
            const int c4 = cnt >> 2;
            for (int q = 0; q < c4; q++) {
                const int4 v = __ldg(rj + q);
                rank += (v.x < j) + (v.y < j) + (v.z < j) + (v.w < j);
            }
            if (cnt & 3) {
                const int4 v = __ldg(rj + c4);
                const int rem = cnt & 3;
                rank += (v.x < j) + (rem > 1 && v.y < j) + (rem > 2 && v.z < j);
            }
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
