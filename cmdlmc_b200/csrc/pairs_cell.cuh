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
//   k_cell_pairs   one thread per row atom (in cell order, so a warp walks the same cells):
//                  filter (pairs_dense.cuh filter_pair) over the <= 27 adjacent cells, exact
//                  evaluation of the survivors, per-row sort by column, rows parked in a
//                  fixed-capacity scratch
//   k_cell_scan    per frame: row counts -> row offsets (the LIL->COO order is row-major)
//   k_cell_emit    warp per 32 rows: coalesced write of (start, dest, dist, omega)
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

// grid = (ceil(n / TPB), frames of the batch), block = TPB (128, 64 or 32: long rows take fewer
// threads per CTA); dynamic smem = TPB * rowcap * 4 bytes (the row's column indices only: the kernel
// lives on memory latency, so shared memory per thread decides how many warps hide it).  A row is
// stored in discovery order (tmp_j, tmp_d) together with the order of its columns: tmp_inv[r] = slot
// of the entry with the r-th smallest column.
template <int KIND, bool IMAGES>
__global__ void __launch_bounds__(128)
k_cell_pairs(const __grid_constant__ BoxParams bx, const __grid_constant__ FilterParams fp,
             const __grid_constant__ CellGrid cg, const double *__restrict__ frames,
             const int *__restrict__ ids, const int *__restrict__ n_ids, int first, int n,
             double rc, double t2, int rowcap, const int4 *__restrict__ sorted,
             const int *__restrict__ cell_start, int *__restrict__ rowcount,
             int *__restrict__ tmp_j, double *__restrict__ tmp_d,
             unsigned short *__restrict__ tmp_inv, int *__restrict__ cap_need,
             unsigned long long *__restrict__ ties)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (n_ids && first + (int)blockIdx.y >= *n_ids) return;
    const int64_t f = ids ? ids[first + blockIdx.y] : first + blockIdx.y;
    const int b = blockIdx.y, tid = threadIdx.x, TPB = blockDim.x;
    int *sj = (int *)smem_raw;                             // [rowcap][TPB]
    const int p = blockIdx.x * TPB + tid;
    if (p >= n) return;
    const double *fr = frames + f * (int64_t)n * 3;
    const int4 *srt = sorted + (int64_t)b * n;
    const int *cs = cell_start + (int64_t)b * (cg.ncell + 1);
    const int4 me = __ldg(srt + p);
    const int i = me.w;
    const int ic[3] = {cell_axis(me.x, cg.nc[0]), cell_axis(me.y, cg.nc[1]), cell_axis(me.z, cg.nc[2])};

    // ---- filter over the adjacent cells -------------------------------------------------------
    int ncand = 0;
    for (int dx = 0; dx < cg.span[0]; dx++) {
        int cx = cg.span[0] == 1 ? 0 : ic[0] + dx - 1;
        cx += cx < 0 ? cg.nc[0] : 0; cx -= cx >= cg.nc[0] ? cg.nc[0] : 0;
        for (int dy = 0; dy < cg.span[1]; dy++) {
            int cy = cg.span[1] == 1 ? 0 : ic[1] + dy - 1;
            cy += cy < 0 ? cg.nc[1] : 0; cy -= cy >= cg.nc[1] ? cg.nc[1] : 0;
            // The cells of one (cx, cy) column are consecutive in the sorted order, so the up to three
            // z neighbours are ONE index range -- two where the column wraps around.
            const int col = (cx * cg.nc[1] + cy) * cg.nc[2];
            int zlo[2], zhi[2], nrange = 1;
            if (cg.span[2] == 1) { zlo[0] = 0; zhi[0] = cg.nc[2] - 1; }
            else {
                zlo[0] = ic[2] - 1; zhi[0] = ic[2] + 1;
                if (zlo[0] < 0) { zlo[1] = cg.nc[2] - 1; zhi[1] = cg.nc[2] - 1; zlo[0] = 0; nrange = 2; }
                else if (zhi[0] >= cg.nc[2]) { zlo[1] = 0; zhi[1] = 0; zhi[0] = cg.nc[2] - 1; nrange = 2; }
            }
            for (int rg = 0; rg < nrange; rg++) {
                const int k1 = __ldg(cs + col + zhi[rg] + 1);
                for (int k = __ldg(cs + col + zlo[rg]); k < k1; k++) {
                    const int4 q = __ldg(srt + k);
                    if (q.w != i && filter_pair<KIND, IMAGES>(fp, me, q)) {
                        if (ncand < rowcap) sj[ncand * TPB + tid] = q.w;
                        ncand++;
                    }
                }
            }
        }
    }
    if (ncand > rowcap) {   // scratch row too small: report the need, the host retries
        atomicMax(cap_need, ncand);
        rowcount[(int64_t)b * n + i] = 0;
        return;
    }

    // ---- exact evaluation of the survivors (reference arithmetic), compaction -----------------
    const double pa[3] = {__ldg(fr + 3 * i), __ldg(fr + 3 * i + 1), __ldg(fr + 3 * i + 2)};
    const int64_t row = ((int64_t)b * n + i) * rowcap;
    int nh = 0;
    unsigned long long my_ties = 0;
    for (int c = 0; c < ncand; c++) {
        const int j = sj[c * TPB + tid];
        const double pb[3] = {__ldg(fr + 3 * j), __ldg(fr + 3 * j + 1), __ldg(fr + 3 * j + 2)};
        double d[3], d2;
        if (KIND == 0) {
            diff_ortho_exact(bx, pa, pb, d);
            d2 = norm2_exact(d);
        } else {
            diff_general_exact(bx, pa, pb, d);
            d2 = min_image_norm2_kept(bx, d);
        }
        const double dist = convert_distance(bx, sqrt(d2));
        const bool hit = (bx.conv == CMD_CONV_NONE ? d2 <= t2 : dist <= rc) && dist != 0.0;
        if (i < j && fabs(dist - rc) <= 1e-11 * rc) my_ties++;   // a pair is seen from both rows
        if (hit) {
            sj[nh * TPB + tid] = j;      // nh <= c: never ahead of the read position
            tmp_j[row + nh] = j;
            tmp_d[row + nh] = dist;
            nh++;
        }
    }
    if (my_ties) atomicAdd(ties, my_ties);

    // ---- the row's column order by rank counting (columns are distinct) -------------------------
    for (int a = 0; a < nh; a++) {
        const int kj = sj[a * TPB + tid];
        int rank = 0;
        for (int q = 0; q < nh; q++) rank += sj[q * TPB + tid] < kj;
        tmp_inv[row + rank] = (unsigned short)a;
    }
    rowcount[(int64_t)b * n + i] = nh;
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

// grid = (ceil(n / 256), frames of the batch), block = 256: a warp writes 32 consecutive rows
__global__ void __launch_bounds__(256)
k_cell_emit(const __grid_constant__ RateParams rp, const int *__restrict__ ids,
            const int *__restrict__ n_ids, int first, int n, int64_t stride, int rowcap,
            const int *__restrict__ rowoff, const int *__restrict__ tmp_j,
            const double *__restrict__ tmp_d, const unsigned short *__restrict__ tmp_inv,
            int *__restrict__ out_start,
            int *__restrict__ out_dest, double *__restrict__ out_dist,
            double *__restrict__ out_omega, double *__restrict__ out_rate_sum)
{
    if (n_ids && first + (int)blockIdx.y >= *n_ids) return;
    const int64_t f = ids ? ids[first + blockIdx.y] : first + blockIdx.y;
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int r0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
    if (r0 >= n) return;
    const int *ro = rowoff + (int64_t)b * (n + 1);
    if (ro[n] > stride) return;   // overflow: reported by k_cell_scan
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
        if (g < g1) {
            const int r = r0 + lo;
            const int64_t rbase = ((int64_t)b * n + r) * rowcap;
            const int64_t src = rbase + tmp_inv[rbase + (g - off_lo)];   // columns ascending
            const int j = tmp_j[src];
            const double dist = tmp_d[src];
            const double om = rate_eval(rp, dist, 0.0);
            rsum += om;
            out_start[base + g] = r; out_dest[base + g] = j;
            out_dist[base + g] = dist; out_omega[base + g] = om;
        }
    }
    if (out_rate_sum) {
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_down_sync(0xffffffffu, rsum, o);
        if (lane == 0 && rsum != 0.0) atomicAdd(out_rate_sum + f, rsum);
    }
}
