// pairs_dense.cuh -- the all-pairs neighbour-list kernel (row A7 of SURVEY.md section 8):
// NeighborTopology.get_topology_bruteforce (topology.py:55-72) for one frame per CTA, with the
// jump rate of every listed pair fused in (jumprate_generators.py:33-34).
//
// Three phases, all out of shared memory:
//   1 filter   every unordered pair once.  Coordinates are turned into FRACTIONAL coordinates in
//              32-bit FIXED POINT (one unit = 2^-32 of a cell vector), so the minimum-image wrap
//              of a difference is the two's-complement wrap-around of one integer subtraction --
//              it costs nothing.  The wrapped difference goes through an FP32 upper-triangular
//              cell matrix (the R of h = QR; lengths do not depend on Q) and is compared with a
//              conservatively widened radius.  Each thread owns two consecutive rows and walks
//              the cyclic pairing (i, i+k), k = 1..n/2, so one 16-byte shared-memory load feeds
//              two pairs; hits are collected in register bit masks and appended to the
//              candidate list 32 columns at a time.  The filter NEVER decides a hit.
//   2 exact    the candidates (~4 % of the pairs), densely packed over the CTA, in the reference's
//              FP64 arithmetic (pbc.cuh *_exact): the `dist <= cutoff + buffer` decision, sqrt,
//              adjacency bits.
//   3 emit     row offsets from the adjacency bit matrix (popc + scan) and the ordered write of
//              (start, dest, dist, omega): both directions, row-major, columns ascending.
#pragma once
#include "pbc.cuh"
#include "tma.cuh"

// FP32 side of the filter (host-prepared, cmd_topo_create)
struct FilterParams {
    float R[6];        // upper-triangular cell matrix * 2^-32: r00 r01 r02 r11 r12 r22
    float t2;          // widened squared radius
    int n_img;         // kept periodic images besides the wrapped vector (0 for every sane cell)
    float img[CMD_MAX_IMAGES][3];  // their shifts in the R frame
};

struct DenseSmem {
    double *c;              // [n][3] Cartesian coordinates of the frame as they lie in HBM (the
                            // next frame is prefetched into this place by TMA), phases 1-2
    int4 *fx;               // [2n + 4] fixed-point fractional coordinates, duplicated (cyclic)
    unsigned short *wpre;   // [n][W] exclusive popc prefix per mask word (aliases fx, phase 3)
    double *hit_d;          // [hit_cap] distance of a hit, < 0 otherwise
    unsigned *mask;         // [n][W] adjacency bit matrix
    unsigned *hit_ij;       // [hit_cap] (a << 16) | b
    int *rowoff;            // [n + 1]
    int *misc;              // [0] ncand, [1] total, [2..33] warp sums
    double *red;            // [34] block reductions
};

__host__ __device__ inline size_t dense_smem_bytes(int n, int hit_cap)
{
    int W = (n + 31) / 32;
    size_t b = 0;
    b += (3 * (size_t)n * 8 + 15) / 16 * 16;  // c (padded: fx is read with LDS.128)
    b += (2 * (size_t)n + 4) * 16;        // fx
    b += (size_t)hit_cap * 8;             // hit_d
    b += 40 * 8;                          // red
    b += (size_t)n * W * 4;               // mask
    b += (size_t)hit_cap * 4;             // hit_ij
    b += ((size_t)n + 1) * 4;             // rowoff
    b += 40 * 4;                          // misc
    return b + 16;
}

__host__ __device__ inline bool dense_use_wpre(int n)
{
    int W = (n + 31) / 32;
    return (size_t)n * W * 2 <= (2 * (size_t)n + 4) * 16;
}

__device__ __forceinline__ DenseSmem dense_carve(unsigned char *base, int n, int hit_cap)
{
    DenseSmem s;
    int W = (n + 31) / 32;
    s.c = (double *)base;
    s.fx = (int4 *)(base + (3 * (size_t)n * 8 + 15) / 16 * 16);
    s.wpre = (unsigned short *)s.fx;
    s.hit_d = (double *)(s.fx + 2 * n + 4);
    s.red = s.hit_d + hit_cap;
    s.mask = (unsigned *)(s.red + 40);
    s.hit_ij = s.mask + (size_t)n * W;
    s.rowoff = (int *)(s.hit_ij + hit_cap);
    s.misc = s.rowoff + n + 1;
    return s;
}

// exclusive scan of one int per thread over the CTA; returns the exclusive prefix, total in *tot
__device__ __forceinline__ int block_exclusive_scan(int v, int *warp_sums, int *tot)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int ws = lane < nw ? warp_sums[lane] : 0;
        int winc = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < nw) warp_sums[lane] = winc - ws;
        if (lane == 31) *tot = winc;
    }
    __syncthreads();
    return inc - v + warp_sums[w];
}

// squared reference length over the zero image + the kept images, in the reference's operation
// order ((d + i a) + j b) + k c (numpyatom.pyx:111-118).  Equals the reference's 27-image minimum
// whenever that minimum is <= cutoff + buffer (cmd_box_prune_images, test 2).
__device__ __forceinline__ double min_image_norm2_kept(const BoxParams &bx, const double d[3])
{
    double mind = fmin(1e6, norm2_exact(d));   // image (0,0,0): d + 0*a + 0*b + 0*c == d
    for (int m = 0; m < bx.n_img; m++) {
        const int i = bx.img_ijk[m][0], j = bx.img_ijk[m][1], k = bx.img_ijk[m][2];
        double v[3];
#pragma unroll
        for (int c = 0; c < 3; c++)
            v[c] = __dadd_rn(__dadd_rn(__dadd_rn(d[c], i * bx.h[3 * c]), j * bx.h[3 * c + 1]),
                             k * bx.h[3 * c + 2]);
        double n2 = norm2_exact(v);
        if (n2 < mind) mind = n2;
    }
    return mind;
}

// the filter's verdict on one pair: wrapped fixed-point difference -> FP32 length^2 <= radius^2
template <int KIND, bool IMAGES>
__device__ __forceinline__ bool filter_pair(const FilterParams &fp, const int4 &p, const int4 &q)
{
    const float a = __int2float_rn(q.x - p.x), b = __int2float_rn(q.y - p.y),
                c = __int2float_rn(q.z - p.z);
    float vx, vy, vz;
    if (KIND == 0) {
        vx = fp.R[0] * a; vy = fp.R[3] * b; vz = fp.R[5] * c;
    } else {
        vx = fmaf(fp.R[2], c, fmaf(fp.R[1], b, fp.R[0] * a));
        vy = fmaf(fp.R[4], c, fp.R[3] * b);
        vz = fp.R[5] * c;
    }
    float d2 = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
    if (IMAGES) {
        for (int m = 0; m < fp.n_img; m++) {
            const float ux = vx + fp.img[m][0], uy = vy + fp.img[m][1], uz = vz + fp.img[m][2];
            d2 = fminf(d2, fmaf(uz, uz, fmaf(uy, uy, ux * ux)));
        }
    }
    return d2 <= fp.t2;
}

// One CTA per frame.  grid.x = number of frames to (re)build; frame = ids ? ids[blockIdx.x] :
// blockIdx.x.  blockDim.x = SPLIT * T2 with T2 >= ceil(n / 2) a multiple of 32: SPLIT copies of
// the row set share the column blocks of phase 1 (more warps per frame for phases 2-3).
template <int KIND, bool IMAGES, int SPLIT, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_pairs_dense(const __grid_constant__ BoxParams bx, const __grid_constant__ RateParams rp,
              const __grid_constant__ FilterParams fp, const double *__restrict__ frames,
              const int *__restrict__ ids, const int *__restrict__ n_ids, int n_items, int n,
              double rc,
              double t2, int64_t stride, int hit_cap, int *__restrict__ out_start,
              int *__restrict__ out_dest, double *__restrict__ out_dist,
              double *__restrict__ out_omega, int *__restrict__ out_counts,
              double *__restrict__ out_rate_sum, uint8_t *__restrict__ out_rebuilt,
              int *__restrict__ out_rowoff, int *__restrict__ err,
              unsigned long long *__restrict__ ties)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DenseSmem s = dense_carve(smem_raw, n, hit_cap);
    const int W = (n + 31) / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int total_items = n_ids ? *n_ids : n_items;
    // Persistent CTA: frames item = blockIdx.x, + gridDim.x, ...  The coordinates of a frame are one
    // contiguous 24n-byte run in HBM; TMA (cp.async.bulk) drops the NEXT frame into s.c while
    // phase 3 of the current one runs, so a frame never waits for global memory.  (Bulk copies
    // need 16-byte granules: an odd atom count falls back to plain loads.)
    uint64_t *bar = (uint64_t *)(s.red + 36);
    const bool use_tma = (n & 1) == 0 && (((size_t)frames) & 15) == 0;
    if (tid == 0 && use_tma) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto prefetch = [&](int item) {
        const int64_t fn = ids ? ids[item] : item;
        mbar_expect_tx(bar, (uint32_t)n * 24u);
        tma_load_1d(s.c, frames + fn * (int64_t)n * 3, (uint32_t)n * 24u, bar);
    };
    if (use_tma && tid == 0 && (int)blockIdx.x < total_items) prefetch(blockIdx.x);
    unsigned tma_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
    const int64_t f = ids ? ids[item] : item;
    if (use_tma) {
        mbar_wait(bar, tma_phase & 1u);
        tma_phase++;
    } else {
        const double *fr = frames + f * (int64_t)n * 3;
        for (int k = tid; k < 3 * n; k += blockDim.x) s.c[k] = __ldg(fr + k);
    }
    for (int k = tid; k < n * W; k += blockDim.x) s.mask[k] = 0u;
    if (tid == 0) { s.misc[0] = 0; s.misc[1] = 0; }
    __syncthreads();
    // fractional coordinates in 2^-32 fixed point; the low 32 bits of the rounded product ARE the
    // coordinate modulo one cell vector.  Stored twice so that the cyclic walk needs no modulo.
    for (int a = tid; a < n; a += blockDim.x) {
        const double x = s.c[3 * a], y = s.c[3 * a + 1], z = s.c[3 * a + 2];
        int q[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double v = KIND == 0 ? (c == 0 ? x : c == 1 ? y : z) * bx.hinv[4 * c]
                                 : fma(bx.hinv[3 * c + 2], z, fma(bx.hinv[3 * c + 1], y, bx.hinv[3 * c] * x));
            q[c] = (int)(unsigned)(unsigned long long)__double2ll_rn(v * 4294967296.0);
        }
        const int4 p = make_int4(q[0], q[1], q[2], 0);
        s.fx[a] = p;
        s.fx[a + n] = p;
        if (a < 4) s.fx[a + 2 * n] = p;   // n < 4: entries beyond stay unread (see validity below)
    }
    __syncthreads();

    // ---- phase 1: filter -------------------------------------------------------------------
    // Control flow is uniform over the CTA (rows beyond n only zero their hit masks), so the
    // appends can use full-warp shuffles and no loop ever runs with a split warp.
    {
        const int T2 = blockDim.x / SPLIT;          // threads per copy of the row set
        const int part = tid / T2, t = tid - part * T2;
        const int r0 = 2 * t, r1 = r0 + 1;
        const bool act0 = r0 < n, act1 = r1 < n;
        const int K = (n - 1) >> 1, half = n >> 1;
        const bool even = (n & 1) == 0;
        const int4 p0 = s.fx[act0 ? r0 : 0], p1 = s.fx[act1 ? r1 : 0];
        const int4 *fj = s.fx + (act0 ? r0 : 0) + 1;   // column of step m: atom (r0 + 1 + m) mod n
        const unsigned tag0 = (unsigned)r0 << 16, tag1 = (unsigned)r1 << 16;
        // warp-cooperative append of the hits of steps m0 .. m0+31 (bit u of ha / hb = step m0+u)
        auto flush = [&](unsigned ha, unsigned hb, int m0) {
            const int c = __popc(ha) + __popc(hb);
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            const int tot = __shfl_sync(0xffffffffu, inc, 31);
            if (tot == 0) return;                    // warp-uniform
            int base = 0;
            if (lane == 31) base = atomicAdd(&s.misc[0], tot);
            int pos = __shfl_sync(0xffffffffu, base, 31) + inc - c;
            const int j0 = r0 + 1 + m0;
            while (ha) {
                int j = j0 + __ffs(ha) - 1;
                ha &= ha - 1;
                if (j >= n) j -= n;
                if (pos < hit_cap) s.hit_ij[pos] = tag0 | (unsigned)j;
                pos++;
            }
            while (hb) {
                int j = j0 + __ffs(hb) - 1;
                hb &= hb - 1;
                if (j >= n) j -= n;
                if (pos < hit_cap) s.hit_ij[pos] = tag1 | (unsigned)j;
                pos++;
            }
        };
        // Row r0 meets step m at offset m + 1, row r1 at offset m; offsets 1..K are owned by every
        // row, offset n/2 (even n) by the lower half only -> each unordered pair exactly once.
        // Steps 1 .. K-1 are valid for both rows: `nb` blocks of 32 without any validity test.
        const int nb = K >= 2 ? (K - 1) >> 5 : 0;
        for (int blk = part; blk < nb; blk += SPLIT) {
            const int m = 1 + 32 * blk;
            unsigned ha = 0, hb = 0;
#pragma unroll
            for (int u = 0; u < 32; u++) {
                const int4 q = fj[m + u];
                ha |= (unsigned)filter_pair<KIND, IMAGES>(fp, p0, q) << u;
                hb |= (unsigned)filter_pair<KIND, IMAGES>(fp, p1, q) << u;
            }
            flush(act0 ? ha : 0u, act1 ? hb : 0u, m);
        }
        if (part == SPLIT - 1) {
            // step 0 and the steps behind the last full block, with the validity tests
            auto generic = [&](int mlo, int mhi) {
                unsigned ha = 0, hb = 0;
                for (int m = mlo; m <= mhi; m++) {
                    const bool va = act0 && (m <= K - 1 || (m == K && even && r0 < half));
                    const bool vb = act1 && ((m >= 1 && m <= K) || (m == K + 1 && even && r1 < half));
                    const int4 q = fj[va || vb ? m : 0];
                    ha |= (unsigned)(va && filter_pair<KIND, IMAGES>(fp, p0, q)) << (m - mlo);
                    hb |= (unsigned)(vb && filter_pair<KIND, IMAGES>(fp, p1, q)) << (m - mlo);
                }
                flush(ha, hb, mlo);
            };
            generic(0, 0);
            for (int m = 1 + 32 * nb; m <= K + 1; m += 32) generic(m, min(m + 31, K + 1));
        }
    }
    __syncthreads();
    const int ncand = s.misc[0];
    const bool overflow = ncand > hit_cap;
    if (overflow && tid == 0) {  // capacity probe / overflow: report the (upper bound of the) need
        out_counts[f] = -2 * ncand;
        if (out_rebuilt) out_rebuilt[f] = 1;
        atomicMax(err, 2 * ncand);
    }

    // ---- phase 2: exact evaluation of the candidates ----------------------------------------
    // two candidates per trip: independent FP64 chains for the scheduler to interleave
    unsigned long long my_ties = 0;
    for (int c0 = tid; c0 < (overflow ? 0 : ncand); c0 += 2 * blockDim.x) {
        const int c1 = c0 + blockDim.x;
        const bool two = c1 < ncand;
        const unsigned ij[2] = {s.hit_ij[c0], s.hit_ij[two ? c1 : c0]};
        double d2[2], dist[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int a = ij[q] >> 16, b = ij[q] & 0xffff;
            const double pa[3] = {s.c[3 * a], s.c[3 * a + 1], s.c[3 * a + 2]};
            const double pb[3] = {s.c[3 * b], s.c[3 * b + 1], s.c[3 * b + 2]};
            double d[3];
            if (KIND == 0) {
                // reference: length(frame[hi], frame[lo]) (topology.py:62-66); the arithmetic is
                // sign-symmetric, so the direction does not change a bit
                diff_ortho_exact(bx, pa, pb, d);
                d2[q] = norm2_exact(d);
            } else {
                diff_general_exact(bx, pa, pb, d);
                d2[q] = min_image_norm2_kept(bx, d);
            }
        }
#pragma unroll
        for (int q = 0; q < 2; q++) dist[q] = convert_distance(bx, sqrt(d2[q]));
#pragma unroll
        for (int q = 0; q < 2; q++) {
            if (q == 1 && !two) break;
            const int a = ij[q] >> 16, b = ij[q] & 0xffff;
            const bool hit = (bx.conv == CMD_CONV_NONE ? d2[q] <= t2 : dist[q] <= rc) && dist[q] != 0.0;
            if (fabs(dist[q] - rc) <= 1e-11 * rc) my_ties++;
            s.hit_d[q ? c1 : c0] = hit ? dist[q] : -1.0;
            if (hit) {
                atomicOr(&s.mask[a * W + (b >> 5)], 1u << (b & 31));
                atomicOr(&s.mask[b * W + (a >> 5)], 1u << (a & 31));
            }
        }
    }
    if (my_ties) atomicAdd(ties, my_ties);
    __syncthreads();
    // the coordinates are dead from here on: bring in the next frame of this CTA
    if (use_tma && tid == 0 && item + (int)gridDim.x < total_items) prefetch(item + gridDim.x);

    // ---- phase 3: row counts -> exclusive offsets (LIL->COO order is row-major), write-out ---
    if (!overflow) {
    const bool use_wpre = dense_use_wpre(n);   // the fixed-point coordinates are dead: reuse them
    int cnt = 0, c0 = 0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int i = 2 * tid + q;
        if (i < n) {
            int run = 0;
            for (int w = 0; w < W; w++) {
                if (use_wpre) s.wpre[i * W + w] = (unsigned short)run;
                run += __popc(s.mask[i * W + w]);
            }
            if (q == 0) c0 = run;
            cnt += run;
        }
    }
    int off = block_exclusive_scan(cnt, s.misc + 2, &s.misc[1]);
    if (2 * tid < n) s.rowoff[2 * tid] = off;
    if (2 * tid + 1 < n) s.rowoff[2 * tid + 1] = off + c0;
    __syncthreads();
    const int total = s.misc[1];
    if (tid == 0) {
        out_counts[f] = total > stride ? -total : total;
        if (out_rebuilt) out_rebuilt[f] = 1;
        if (total > stride) atomicMax(err, total);
    }
    if (total <= stride) {
    if (out_rowoff) {   // row index of the frame's list: row i = [rowoff[i], rowoff[i + 1])
        int *ro = out_rowoff + f * (int64_t)cmd_ro_pitch(n);
        for (int k = tid; k < n; k += blockDim.x) ro[k] = s.rowoff[k];
        if (tid == 0) ro[n] = total;
    }

    // position of (a -> b) = rowoff[a] + #set bits of row a below column b
    const int64_t base = f * stride;
    double rsum = 0.0;
    for (int h0 = tid; h0 < ncand; h0 += 2 * blockDim.x) {
        const int h1 = h0 + blockDim.x;
        const bool two = h1 < ncand;
        const double dist[2] = {s.hit_d[h0], two ? s.hit_d[h1] : -1.0};
        const unsigned ij[2] = {s.hit_ij[h0], s.hit_ij[two ? h1 : h0]};
        double om[2];
        rate_eval2(rp, fabs(dist[0]), fabs(dist[1]), om);
#pragma unroll
        for (int q = 0; q < 2; q++) {
            if (dist[q] < 0.0) continue;
            const int a = ij[q] >> 16, b = ij[q] & 0xffff;
            rsum += om[q];
            int pa = s.rowoff[a], pb = s.rowoff[b];
            if (use_wpre) {
                pa += s.wpre[a * W + (b >> 5)];
                pb += s.wpre[b * W + (a >> 5)];
            } else {
                for (int w = 0; w < (b >> 5); w++) pa += __popc(s.mask[a * W + w]);
                for (int w = 0; w < (a >> 5); w++) pb += __popc(s.mask[b * W + w]);
            }
            pa += __popc(s.mask[a * W + (b >> 5)] & ((1u << (b & 31)) - 1u));
            pb += __popc(s.mask[b * W + (a >> 5)] & ((1u << (a & 31)) - 1u));
            out_start[base + pa] = a; out_dest[base + pa] = b;
            out_dist[base + pa] = dist[q]; out_omega[base + pa] = om[q];
            out_start[base + pb] = b; out_dest[base + pb] = a;
            out_dist[base + pb] = dist[q]; out_omega[base + pb] = om[q];
        }
    }
    if (out_rate_sum) {
        // informational per-frame total of all listed rates (both directions)
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_down_sync(0xffffffffu, rsum, o);
        if (lane == 0) s.red[wid] = rsum;
        __syncthreads();
        if (tid == 0) {
            double t = 0;
            for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) t += s.red[w];
            out_rate_sum[f] = 2.0 * t;
        }
    }
    }   // total <= stride
    }   // !overflow
    __syncthreads();   // phase 3 is done with the masks / hit lists before the next frame resets them
    }   // frames of this CTA
}
