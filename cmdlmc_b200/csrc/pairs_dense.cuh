// pairs_dense.cuh -- the all-pairs neighbour-list kernel (row A7 of SURVEY.md section 8):
// NeighborTopology.get_topology_bruteforce (topology.py:55-72) for one frame per trip of a
// persistent CTA, with the jump rate of every listed pair fused in
// (jumprate_generators.py:33-34).
//
// Everything between the frame's coordinates (24 n bytes in) and its list (24 bytes per directed
// pair out) happens in shared memory:
//   1 filter   every unordered pair once, in cheap arithmetic that NEVER decides a hit.
//              FILT_H2 (default): fractional coordinates in 10-bit fixed point, two columns packed
//              per 32-bit word.  One integer add per coordinate forms (x_j - x_i + 512) mod 1024
//              for both columns, one LOP3 turns the two 10-bit fields into the fp16 numbers
//              1024 + u (exponent trick), and from there the wrapped difference, the
//              upper-triangular cell matrix (R of h = QR), the squared length and the compare run
//              as half2 SIMD: ~10 instructions per pair.  The radius is widened by a proven bound
//              on quantisation + fp16 rounding (topo_filter_params).
//              Skin list: consecutive frames of a trajectory (one contiguous run per persistent
//              CTA) share their candidates.  The window filter runs with the radius widened by a
//              skin and its survivors are kept per CTA (global memory, L2 resident); a later
//              frame only re-filters that list, as long as the two largest atomic displacements
//              since the list was built sum to less than the skin (triangle inequality of the
//              periodic metric) -- otherwise the list is rebuilt on the spot.
//              FILT_F32 / FILT_F32_IMG: 32-bit fixed point + FP32 (cells so skewed that periodic
//              images matter, or so large that 10 bits are too coarse).
//   2 exact    the surviving candidates (~1.1x the hits), densely packed over the CTA, in the
//              reference's FP64 arithmetic (pbc.cuh *_exact): the `dist <= cutoff + buffer`
//              decision, sqrt, adjacency bits.
//   3 emit     row offsets from the adjacency bit matrix (popc + scan); every hit drops its
//              candidate index into the two output slots it owns (a->b, b->a); then the CTA walks
//              the output positions IN ORDER, four per thread, and writes (start, dest, dist) and
//              -- after the rates have replaced the distances in shared memory -- omega with
//              16-byte stores: full 32-byte sectors only, no partial-sector traffic in HBM.
#pragma once
#include <cuda_fp16.h>

#include <type_traits>

#include "pbc.cuh"
#include "tma.cuh"

#define FILT_F32 0
#define FILT_F32_IMG 1
#define FILT_H2 2

// filter constants (host-prepared, topo_filter_params)
struct FilterParams {
    float R[6];        // upper-triangular cell matrix * 2^-32: r00 r01 r02 r11 r12 r22
    float t2;          // widened squared radius
    int n_img;         // kept periodic images besides the wrapped vector (0 for every sane cell)
    float img[CMD_MAX_IMAGES][3];  // their shifts in the R frame
    // packed-half filter
    unsigned hR[6];    // R / 1024 as half2 (both halves equal)
    unsigned hT2;      // widened squared radius, half2, rounded up
    int h2_ok;         // the 10-bit filter is valid and tight enough for this cell / radius
    int sort_axis;     // fractional axis the atoms are binned along (-1: no spatial pruning)
    int sort_db;       // window: a pair within the radius is at most this many bins (of 256) apart
    int wpre_ok;       // the popc prefix table fits the dead coordinate + column buffers
    // skin list (frame-to-frame coherence of the candidate set, FILT_H2 only)
    int coh_ok;        // the 10-bit filter is valid at radius + skin as well
    unsigned hT2m;     // widened (radius + skin)^2, half2, rounded up
    int sort_db_m;     // bin window for radius + skin
    float R16[6];      // R * 2^-16: displacement of an atom from 16-bit fractional coordinates
    float skin_eff;    // skin minus the error bound of two such displacements
};

struct DenseSmem {
    double *c;              // [n][3] Cartesian coordinates of the frame as they lie in HBM (the
                            // next frame is prefetched into this place by TMA), phases 1-2
    uint4 *col;             // FILT_H2: [2n + 40] overlapping column pairs of the sorted cyclic
                            // sequence; FILT_F32*: int4 [2n + 4] fixed point
    unsigned short *wpre;   // [n][W] exclusive popc prefix per mask word (aliases c + col, phase 3)
    double *hit_d;          // [cap] distance of a hit, < 0 otherwise; omega after the first emit pass
    double *red;            // [40] block reductions, mbarrier at [36]
    unsigned *mask;         // [n][W] adjacency bit matrix (FILT_H2: 10-bit coordinates before phase 1)
    unsigned *hit_ij;       // [cap] (a << 16) | b
    unsigned short *slot;   // [2 cap] output position -> candidate index | direction << 15
    int *rowoff;            // [n + 1]
    int *misc;              // [0] ncand, [1] total, [2..33] warp sums
    int *bins;              // [258] atoms per sort bin -> exclusive prefix (FILT_H2)
    unsigned short *perm;   // [n] sorted position -> atom (FILT_H2)
    ushort4 *ref16;         // [n] 16-bit fractional coordinates of the frame the skin list was built on
};

__host__ __device__ inline size_t dense_al16(size_t b) { return (b + 15) / 16 * 16; }

__host__ __device__ inline size_t dense_col_bytes(int n, int filt)
{
    return filt == FILT_H2 ? (2 * (size_t)n + 40) * 16 : (2 * (size_t)n + 4) * 16;
}

__host__ __device__ inline size_t dense_mask_bytes(int n)
{
    const int W = (n + 31) / 32;
    size_t m = (size_t)n * W * 4;
    if (m < (size_t)n * 24) m = (size_t)n * 24;   // the 10-bit coordinates are parked here: by sorted
                                                  // position, by atom, and the row constants by atom
    return dense_al16(m);
}

// bytes that do not depend on the candidate capacity / bytes per candidate
__host__ __device__ inline size_t dense_smem_fixed(int n, int filt)
{
    return dense_al16(3 * (size_t)n * 8) + dense_col_bytes(n, filt) + 40 * 8 + dense_mask_bytes(n) +
           dense_al16(((size_t)n + 1) * 4) + 40 * 4 + 264 * 4 + dense_al16((size_t)n * 2) + 16 +
           (filt == FILT_H2 ? (size_t)n * 8 : 0);
}
#define DENSE_BYTES_PER_CAND 16   // hit_d 8 + hit_ij 4 + 2 slots of 2

__host__ __device__ inline bool dense_wpre_fits(int n, int filt)
{
    const int W = (n + 31) / 32;
    return (size_t)n * W * 2 <= dense_al16(3 * (size_t)n * 8) + dense_col_bytes(n, filt);
}

__host__ __device__ inline size_t dense_smem_bytes(int n, int cap, int filt)
{
    return dense_smem_fixed(n, filt) + (size_t)cap * DENSE_BYTES_PER_CAND;
}

__host__ __device__ __forceinline__ DenseSmem dense_carve(unsigned char *base, int n, int cap, int filt)
{
    DenseSmem s;
    const size_t A = dense_al16(3 * (size_t)n * 8), B = dense_col_bytes(n, filt);
    s.c = (double *)base;
    s.col = (uint4 *)(base + A);
    s.wpre = (unsigned short *)base;
    s.hit_d = (double *)(base + A + B);
    s.red = s.hit_d + cap;
    s.mask = (unsigned *)(s.red + 40);
    unsigned char *p = (unsigned char *)s.mask + dense_mask_bytes(n);
    s.hit_ij = (unsigned *)p;
    p += (size_t)cap * 4;
    s.slot = (unsigned short *)p;
    p += (size_t)cap * 4;
    s.rowoff = (int *)p;
    p += dense_al16(((size_t)n + 1) * 4);
    s.misc = (int *)p;
    s.bins = s.misc + 40;
    s.perm = (unsigned short *)(s.bins + 264);
    s.ref16 = (ushort4 *)((unsigned char *)s.perm + dense_al16((size_t)n * 2) + 16);
    return s;
}

// The same carve as byte offsets, prepared on the host and handed to the kernel as a
// __grid_constant__ parameter: a shared-memory address is then one constant-bank operand instead of
// a chain of integer instructions on n and the capacity at every use.
struct DenseLayout {
    unsigned c, col, hit_d, red, mask, hit_ij, slot, rowoff, misc, bins, perm, ref16;
    int W;          // mask words per row
    int om_split;   // doubles that fit the mask's place (the rates of the fused emit pass)
    int fused;      // the rates of all candidates fit the dead mask + column buffers
    int park_hd;    // the quantised coordinates fit the (then dead) distance buffer
};

static inline DenseLayout dense_layout(int n, int cap, int filt)
{
    unsigned char *base = nullptr;
    const DenseSmem s = dense_carve(base, n, cap, filt);
    auto off = [&](const void *p) { return (unsigned)((const unsigned char *)p - base); };
    DenseLayout l;
    l.c = off(s.c); l.col = off(s.col); l.hit_d = off(s.hit_d); l.red = off(s.red);
    l.mask = off(s.mask); l.hit_ij = off(s.hit_ij); l.slot = off(s.slot); l.rowoff = off(s.rowoff);
    l.misc = off(s.misc); l.bins = off(s.bins); l.perm = off(s.perm); l.ref16 = off(s.ref16);
    l.W = (n + 31) / 32;
    l.om_split = (int)(dense_mask_bytes(n) / 8);
    l.fused = (size_t)cap <= dense_mask_bytes(n) / 8 + dense_col_bytes(n, filt) / 8;
    l.park_hd = (size_t)cap * 8 >= (size_t)n * 24;
    return l;
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

// the FP32 filter's verdict on one pair: wrapped fixed-point difference -> length^2 <= radius^2
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

__device__ __forceinline__ __half2 u2h2(unsigned u)
{
    __half2 h;
    *reinterpret_cast<unsigned *>(&h) = u;
    return h;
}

// The packed-half filter on one column pair: 0xffff in the half of every column that passes.
// P = the pair's 10-bit coordinates (low half: even column); C* = (512 - row coordinate) mod 1024
// in both halves.  Each 16-bit lane of P + C is < 2047, so the packed add never carries across.
template <int KIND>
__device__ __forceinline__ unsigned h2_pair_mask(const __half2 (&R)[6], const __half2 T2,
                                                 const uint4 &P, unsigned CX, unsigned CY, unsigned CZ)
{
    // 0x6400 has bit 10 set, so OR-ing it in also reduces the 11-bit lane sum modulo 1024
    const __half2 bias = u2h2(0x66006600u);   // 1536 = 1024 (exponent trick) + 512 (centring)
    const __half2 wx = __hsub2(u2h2((P.x + CX) | 0x64006400u), bias);
    const __half2 wy = __hsub2(u2h2((P.y + CY) | 0x64006400u), bias);
    const __half2 wz = __hsub2(u2h2((P.z + CZ) | 0x64006400u), bias);
    __half2 vx, vy, vz;
    if (KIND == 0) {
        vx = __hmul2(R[0], wx); vy = __hmul2(R[3], wy); vz = __hmul2(R[5], wz);
    } else {
        vx = __hfma2(R[2], wz, __hfma2(R[1], wy, __hmul2(R[0], wx)));
        vy = __hfma2(R[4], wz, __hmul2(R[3], wy));
        vz = __hmul2(R[5], wz);
    }
    const __half2 d2 = __hfma2(vz, vz, __hfma2(vy, vy, __hmul2(vx, vx)));
    return __hle2_mask(d2, T2);
}

// Persistent CTAs: frames item = blockIdx.x, + gridDim.x, ...; frame = ids ? ids[item] : item.
// FILT_H2: blockDim.x >= n, a multiple of 32 -- one (sorted) row per thread.
// FILT_F32*: blockDim.x = SPLIT * T2 with T2 >= ceil(n / 2) a multiple of 32: two rows per thread,
// SPLIT copies of the row set share the column blocks (more warps per frame for phases 2-3).
template <int KIND, int FILT, int SPLIT, int MAXT, int MINB>
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
              unsigned long long *__restrict__ ties, unsigned *__restrict__ lists, int cap_l,
              const __grid_constant__ DenseLayout lay)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DenseSmem s;
    s.c = (double *)(smem_raw + lay.c);
    s.col = (uint4 *)(smem_raw + lay.col);
    s.wpre = (unsigned short *)(smem_raw + lay.c);
    s.hit_d = (double *)(smem_raw + lay.hit_d);
    s.red = (double *)(smem_raw + lay.red);
    s.mask = (unsigned *)(smem_raw + lay.mask);
    s.hit_ij = (unsigned *)(smem_raw + lay.hit_ij);
    s.slot = (unsigned short *)(smem_raw + lay.slot);
    s.rowoff = (int *)(smem_raw + lay.rowoff);
    s.misc = (int *)(smem_raw + lay.misc);
    s.bins = (int *)(smem_raw + lay.bins);
    s.perm = (unsigned short *)(smem_raw + lay.perm);
    s.ref16 = (ushort4 *)(smem_raw + lay.ref16);
    const int W = lay.W;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int total_items = n_ids ? *n_ids : n_items;
    // The coordinates of a frame are one contiguous 24n-byte run in HBM; TMA (cp.async.bulk) drops
    // the NEXT frame into s.c while the emit passes of the current one run, so a frame never waits
    // for global memory.  (Bulk copies need 16-byte granules: an odd atom count falls back to
    // plain loads.)
    uint64_t *bar = (uint64_t *)(s.red + 36);
    const bool use_tma = (n & 1) == 0 && (((size_t)frames) & 15) == 0;
    if (tid == 0 && use_tma) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // skin list state (CTA-uniform, kept in shared memory to spare the registers of the phases):
    // s.misc[34] = entries of the list, s.misc[35] = the list is valid for the frame in s.ref16
    if (tid == 0) { s.misc[34] = 0; s.misc[35] = 0; }
    __syncthreads();
    auto prefetch = [&](int item) {
        const int64_t fn = ids ? ids[item] : item;
        mbar_expect_tx(bar, (uint32_t)n * 24u);
        tma_load_1d(s.c, frames + fn * (int64_t)n * 3, (uint32_t)n * 24u, bar);
    };
    // every CTA takes ONE contiguous run of the items: consecutive frames of a trajectory, so that
    // the skin list of the packed-half filter carries over from frame to frame
    // (without a list the CTAs interleave: the frames in flight are neighbours in HBM)
    const int per_cta = total_items / (int)gridDim.x, extra_items = total_items % (int)gridDim.x;
    const int item_step = lists ? 1 : (int)gridDim.x;
    const int item_lo = lists ? (int)blockIdx.x * per_cta + min((int)blockIdx.x, extra_items) : (int)blockIdx.x;
    const int item_hi = lists ? item_lo + per_cta + ((int)blockIdx.x < extra_items ? 1 : 0) : total_items;
    if (use_tma && tid == 0 && item_lo < item_hi) prefetch(item_lo);
    unsigned tma_phase = 0;
    const int K = (n - 1) >> 1, half = n >> 1;
    const bool even = (n & 1) == 0;
    const bool have_list = lists != nullptr;

    for (int item = item_lo; item < item_hi; item += item_step) {
    const int64_t f = ids ? ids[item] : item;
    if (use_tma) {
        mbar_wait(bar, tma_phase & 1u);
        tma_phase++;
    } else {
        const double *fr = frames + f * (int64_t)n * 3;
        for (int k = tid; k < 3 * n; k += blockDim.x) s.c[k] = __ldg(fr + k);
    }
    if (tid == 0) { s.misc[0] = 0; s.misc[1] = 0; }

    // ---- phase 1: filter -------------------------------------------------------------------
    // Control flow is uniform over every warp (rows beyond n only zero their hit masks), so the
    // appends can use full-warp shuffles and no loop ever runs with a split warp.
    // warp-cooperative append: reserves `cnt` list entries for this lane, -1 if the warp has none
    auto reserve = [&](int cnt) -> int {
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const int tot = __shfl_sync(0xffffffffu, inc, 31);
        if (tot == 0) return -1;                    // warp-uniform
        int base = 0;
        if (lane == 31) base = atomicAdd(&s.misc[0], tot);
        return __shfl_sync(0xffffffffu, base, 31) + inc - cnt;
    };
    if (FILT == FILT_H2) {
        if (!use_tma) __syncthreads();   // the plain loads of s.c
        // (a) fractional coordinates in 10-bit fixed point (the filter) and in 16-bit fixed point
        // (the displacement test of the skin list), one atom per thread
        const bool sorted = fp.sort_axis >= 0;
        // the quantised coordinates are parked in the distance buffer of the exact stage (dead
        // until then) -- then the adjacency mask can be cleared right away; a candidate list too
        // short for that lends the mask's own space and is cleared after the filter
        const bool park_hd = lay.park_hd != 0;
        ushort4 *q16 = park_hd ? (ushort4 *)s.hit_d : (ushort4 *)s.mask;   // by sorted position
        if (park_hd)
            for (int k = tid; k < n * W; k += blockDim.x) s.mask[k] = 0u;
        ushort4 *qa = q16 + n;                     // by atom: the coordinates
        ushort4 *qc = qa + n;                      // by atom: (512 - coordinate) mod 1024
        unsigned q[3] = {0, 0, 0}, r16[3] = {0, 0, 0};
        if (tid < n) {
            const double x = s.c[3 * tid], y = s.c[3 * tid + 1], z = s.c[3 * tid + 2];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double v = KIND == 0 ? (c == 0 ? x : c == 1 ? y : z) * bx.hinv[4 * c]
                                           : fma(bx.hinv[3 * c + 2], z, fma(bx.hinv[3 * c + 1], y, bx.hinv[3 * c] * x));
                q[c] = (unsigned)(unsigned long long)__double2ll_rn(v * 1024.0) & 1023u;
                if (have_list) r16[c] = (unsigned)(unsigned long long)__double2ll_rn(v * 65536.0) & 65535u;
            }
            if (have_list) {
                qa[tid] = make_ushort4((unsigned short)q[0], (unsigned short)q[1], (unsigned short)q[2], 0);
                qc[tid] = make_ushort4((unsigned short)((512u - q[0]) & 1023u), (unsigned short)((512u - q[1]) & 1023u),
                                       (unsigned short)((512u - q[2]) & 1023u), 0);
            }
        }
        // (a') Skin list: the candidates of the frame the list was built on, taken with the radius
        // widened by the skin, hold every pair of THIS frame that is within the radius as long as
        // the two largest displacements since then sum to less than the skin (triangle inequality
        // of the periodic metric; fp.skin_eff carries the error bound of the 16-bit coordinates).
        bool rebuild = true;
        if (have_list && s.misc[35]) {
            unsigned d2b = 0;
            if (tid < n) {
                const ushort4 rf = s.ref16[tid];
                const float wx = (float)(short)(unsigned short)(r16[0] - rf.x);
                const float wy = (float)(short)(unsigned short)(r16[1] - rf.y);
                const float wz = (float)(short)(unsigned short)(r16[2] - rf.z);
                const float vx = fmaf(fp.R16[2], wz, fmaf(fp.R16[1], wy, fp.R16[0] * wx));
                const float vy = fmaf(fp.R16[4], wz, fp.R16[3] * wy);
                const float vz = fp.R16[5] * wz;
                d2b = __float_as_uint(fmaf(vz, vz, fmaf(vy, vy, vx * vx)));   // >= 0: ordered as integers
            }
            // the two largest of the CTA: per warp, then over the warps (every warp redundantly)
            unsigned m1 = __reduce_max_sync(0xffffffffu, d2b);
            const unsigned holder = __ballot_sync(0xffffffffu, d2b == m1);
            unsigned m2 = __reduce_max_sync(0xffffffffu, lane == __ffs(holder) - 1 ? 0u : d2b);
            unsigned *top = (unsigned *)s.red;
            if (lane == 0) { top[2 * wid] = m1; top[2 * wid + 1] = m2; }
            __syncthreads();
            const int nw2 = 2 * (int)(blockDim.x >> 5);
            unsigned v0 = lane < nw2 ? top[lane] : 0u, v1 = lane + 32 < nw2 ? top[lane + 32] : 0u;
            // two values per lane -> top two over 64
            unsigned hi = max(v0, v1), lo = min(v0, v1);
            m1 = __reduce_max_sync(0xffffffffu, hi);
            const unsigned holder2 = __ballot_sync(0xffffffffu, hi == m1);
            m2 = __reduce_max_sync(0xffffffffu, lane == __ffs(holder2) - 1 ? lo : hi);
            rebuild = !(sqrtf(__uint_as_float(m1)) + sqrtf(__uint_as_float(m2)) <= fp.skin_eff);
        }
        // window filter of the sorted rows -> candidate pairs (atom indices) in dst[0 .. cap)
        bool direct = !have_list;              // no skin list: the window filter feeds the exact stage
        while (rebuild) {
        const int db = direct ? fp.sort_db : fp.sort_db_m;
        const __half2 hT2w = u2h2(direct ? fp.hT2 : fp.hT2m);
        // counting sort of the atoms into 256 bins along the axis with the largest cell height.
        // A pair within the radius is at most db bins apart, so a row only meets the columns that
        // FOLLOW it in the sorted cyclic order up to that bin distance: each unordered pair once,
        // ~2 rc / height of them.
        if (sorted) for (int k = tid; k < 258; k += blockDim.x) s.bins[k] = 0;
        __syncthreads();
        int key = 0, rnk = 0;
        if (tid < n && sorted) {
            key = (int)((fp.sort_axis == 0 ? q[0] : fp.sort_axis == 1 ? q[1] : q[2]) >> 2);
            rnk = atomicAdd(&s.bins[key], 1);
        }
        __syncthreads();
        if (sorted && wid == 0) {   // exclusive prefix over the 256 bins, eight per lane
            int v[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) { v[j] = s.bins[8 * lane + j]; sum += v[j]; }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            int run = inc - sum;
#pragma unroll
            for (int j = 0; j < 8; j++) { s.bins[8 * lane + j] = run; run += v[j]; }
            if (lane == 31) { s.bins[256] = run; s.bins[257] = run; }
        }
        if (sorted) __syncthreads();
        if (tid < n) {
            const int pos = sorted ? s.bins[key] + rnk : tid;
            s.perm[pos] = (unsigned short)tid;
            q16[pos] = make_ushort4((unsigned short)q[0], (unsigned short)q[1], (unsigned short)q[2],
                                    (unsigned short)key);
        }
        __syncthreads();
        // (b) overlapping column pairs of the cyclic sorted sequence: entry k = positions (k, k+1)
        const int ncol = 2 * n + 40;
        for (int k = tid; k < ncol; k += blockDim.x) {
            int k0 = k, k1 = k + 1;            // mod n without the integer division
            while (k0 >= n) k0 -= n;
            while (k1 >= n) k1 -= n;
            const ushort4 lo = q16[k0], hi = q16[k1];
            s.col[k] = make_uint4(lo.x | ((unsigned)hi.x << 16), lo.y | ((unsigned)hi.y << 16),
                                  lo.z | ((unsigned)hi.z << 16), 0u);
        }
        // (c) this thread's row: sorted position t, its constants and the length of its window
        const bool act = tid < n;
        const int t = act ? tid : 0;
        unsigned C[3];
        int len;
        {
            const ushort4 me = q16[t];
            C[0] = ((512u - me.x) & 1023u) * 0x00010001u;
            C[1] = ((512u - me.y) & 1023u) * 0x00010001u;
            C[2] = ((512u - me.z) & 1023u) * 0x00010001u;
            if (sorted) {
                const int e = (int)me.w + db;   // last bin of the window (inclusive)
                len = (e < 256 ? s.bins[e + 1] : n + s.bins[e - 255]) - 1 - t;
            } else {
                // offsets 1..K are owned by every row, offset n/2 (even n) by the lower half only
                len = K + ((even && t < half) ? 1 : 0);
            }
            if (!act) len = 0;
        }
        const unsigned my_tag = (unsigned)s.perm[t] << 16;
        __syncthreads();
        __half2 hR[6];
#pragma unroll
        for (int k = 0; k < 6; k++) hR[k] = u2h2(fp.hR[k]);
        // (d) blocks of 32 offsets: entry pk[2u] holds the columns at offsets 1 + 2u (low half ->
        // hit bit u) and 2 + 2u (high half -> hit bit 16 + u) from the row
        const uint4 *pk = s.col + t + 1;
        const int nblk = __reduce_max_sync(0xffffffffu, (len + 31) >> 5);
        for (int blk = 0; blk < nblk; blk++) {
            unsigned ha = 0;
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const uint4 P = pk[32 * blk + 2 * u];
                ha |= h2_pair_mask<KIND>(hR, hT2w, P, C[0], C[1], C[2]) & (0x00010001u << u);
            }
            const int rem = len - 32 * blk;                  // offsets 1..rem of this block are owned
            const int nlo = min(max((rem + 1) >> 1, 0), 16), nhi = min(max(rem >> 1, 0), 16);
            ha &= ((1u << nlo) - 1u) | (((1u << nhi) - 1u) << 16);
            int pos = reserve(__popc(ha));
            if (pos < 0) continue;
            const int c0 = t + 32 * blk + 1;
            while (ha) {
                const int b = __ffs(ha) - 1;
                ha &= ha - 1;
                int c = c0 + 2 * (b & 15) + (b >> 4);
                if (c >= n) c -= n;
                CMD_CHECK(c >= 0 && c < n && t < n);
                const unsigned e = my_tag | (unsigned)s.perm[c];
                if (direct) { if (pos < hit_cap) s.hit_ij[pos] = e; }
                else if (pos < cap_l) lists[(size_t)blockIdx.x * cap_l + pos] = e;
                pos++;
            }
        }
        if (direct) break;
        // the skin list of this frame is complete: it becomes the reference -- unless it does not
        // fit, then this frame goes the direct way and the next one tries again
        __syncthreads();
        const int n_new = s.misc[0];
        __syncthreads();
        if (tid == 0) { s.misc[0] = 0; s.misc[34] = n_new; s.misc[35] = n_new <= cap_l; }
        if (n_new > cap_l) { direct = true; continue; }
        if (tid < n) s.ref16[tid] = make_ushort4((unsigned short)r16[0], (unsigned short)r16[1],
                                                 (unsigned short)r16[2], 0);
        __threadfence_block();
        break;
        }   // window filter
        if (!direct) {
            // (e) the skin list through the same packed-half filter at the radius itself, two
            // candidates per thread (low / high halves); survivors -> the exact stage's list
            // after a rebuild: s.misc[0] == 0, the list and its length visible to the CTA (a reused
            // list: the barrier of the displacement test already stands between this point and the
            // frame's writes of qa / qc / s.misc[0])
            if (rebuild) __syncthreads();
            const int n_list = s.misc[34];
            const unsigned *my_list = lists + (size_t)blockIdx.x * cap_l;
            if (tid == 0) {   // statistics: frames, rebuilds, list entries filtered
                atomicAdd(ties + 2, 1ull);
                if (rebuild) atomicAdd(ties + 1, 1ull);
                atomicAdd(ties + 3, (unsigned long long)n_list);
            }
            __half2 hR[6];
#pragma unroll
            for (int k = 0; k < 6; k++) hR[k] = u2h2(fp.hR[k]);
            const __half2 hT2 = u2h2(fp.hT2);
            const uint2 *qa2 = (const uint2 *)qa, *qc2 = (const uint2 *)qc;
            // LIST_U double trips per round: their list loads (L2) are all in flight before the
            // first one is used
            const int T = (int)blockDim.x;
            // the double trips are spread evenly over rounds of at most four (5 = 3 + 2, not 4 + 1);
            // the round body is compiled for each size, so that its loads are all in flight before
            // the first one is used and no slot of a round is idle
            const int n_dt = (n_list + 2 * T - 1) / (2 * T), n_rounds = (n_dt + 3) / 4;
            const int per_round = n_rounds ? (n_dt + n_rounds - 1) / n_rounds : 1;
            auto rounds = [&](auto uc) {
                constexpr int LIST_U = decltype(uc)::value;
                for (int k0 = 0; k0 < n_list; k0 += 2 * LIST_U * T) {
                    unsigned ija[LIST_U], ijb[LIST_U];
#pragma unroll
                    for (int u = 0; u < LIST_U; u++) {
                        const int ka = k0 + 2 * u * T + tid, kb = ka + T;
                        ija[u] = ka < n_list ? my_list[ka] : 0u;
                        ijb[u] = kb < n_list ? my_list[kb] : 0u;
                    }
                    unsigned pass = 0;   // bit 2u: entry a of trip u, bit 2u + 1: entry b
#pragma unroll
                    for (int u = 0; u < LIST_U; u++) {
                        const int ka = k0 + 2 * u * T + tid, kb = ka + T;
                        CMD_CHECK((int)(ija[u] >> 16) < n && (int)(ija[u] & 0xffffu) < n &&
                                  (int)(ijb[u] >> 16) < n && (int)(ijb[u] & 0xffffu) < n);
                        const uint2 Ca = qc2[ija[u] >> 16], Pa = qa2[ija[u] & 0xffffu];
                        const uint2 Cb = qc2[ijb[u] >> 16], Pb = qa2[ijb[u] & 0xffffu];
                        const uint4 P = make_uint4(__byte_perm(Pa.x, Pb.x, 0x5410), __byte_perm(Pa.x, Pb.x, 0x7632),
                                                   __byte_perm(Pa.y, Pb.y, 0x5410), 0u);
                        const unsigned m = h2_pair_mask<KIND>(hR, hT2, P, __byte_perm(Ca.x, Cb.x, 0x5410),
                                                              __byte_perm(Ca.x, Cb.x, 0x7632),
                                                              __byte_perm(Ca.y, Cb.y, 0x5410));
                        if (ka < n_list && (m & 0xffffu)) pass |= 1u << (2 * u);
                        if (kb < n_list && (m >> 16)) pass |= 2u << (2 * u);
                    }
                    // one reservation per warp and round
                    int pos = reserve(__popc(pass));
                    if (pos < 0) continue;                          // warp-uniform
#pragma unroll
                    for (int u = 0; u < LIST_U; u++) {
                        if (pass & (1u << (2 * u))) { if (pos < hit_cap) s.hit_ij[pos] = ija[u]; pos++; }
                        if (pass & (2u << (2 * u))) { if (pos < hit_cap) s.hit_ij[pos] = ijb[u]; pos++; }
                    }
                }
            };
            if (per_round == 3) rounds(std::integral_constant<int, 3>{});
            else if (per_round == 4) rounds(std::integral_constant<int, 4>{});
            else if (per_round == 2) rounds(std::integral_constant<int, 2>{});
            else rounds(std::integral_constant<int, 1>{});
        }
        if (!park_hd) {
            __syncthreads();
            for (int k = tid; k < n * W; k += blockDim.x) s.mask[k] = 0u;   // the parked coordinates are dead
        }
    } else {
        for (int k = tid; k < n * W; k += blockDim.x) s.mask[k] = 0u;
        const int T2 = blockDim.x / SPLIT;          // threads per copy of the row set
        const int part = tid / T2, t = tid - part * T2;
        const int r0 = 2 * t, r1 = r0 + 1;
        const bool act0 = r0 < n, act1 = r1 < n;
        // 32-bit fixed point (unit 2^-32 of a cell vector): the low 32 bits of the rounded product
        // ARE the coordinate modulo one cell vector, and the minimum-image wrap of a difference is
        // the two's-complement wrap-around of one integer subtraction.  Stored twice so that the
        // cyclic walk needs no modulo.
        int4 *fx = (int4 *)s.col;
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
            fx[a] = p;
            fx[a + n] = p;
            if (a < 4) fx[a + 2 * n] = p;   // n < 4: entries beyond stay unread (see validity below)
        }
        __syncthreads();
        constexpr bool IMAGES = FILT == FILT_F32_IMG;
        const int4 p0 = fx[act0 ? r0 : 0], p1 = fx[act1 ? r1 : 0];
        const int4 *fj = fx + (act0 ? r0 : 0) + 1;   // column of step m: atom (r0 + 1 + m) mod n
        const unsigned tag0 = (unsigned)r0 << 16, tag1 = (unsigned)r1 << 16;
        // append of the hits of steps m0 .. m0+31 (bit u of ha / hb = step m0+u)
        auto flush = [&](unsigned ha, unsigned hb, int m0) {
            const int c = __popc(ha) + __popc(hb);
            int pos = reserve(c);
            if (pos < 0) return;
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
        // SP: structural zeros of the cell matrix (pbc.cuh matvec3_norm_sp), warp-uniform
        auto length2_pair = [&](auto spc) {
            constexpr int SP = decltype(spc)::value;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int a = ij[q] >> 16, b = ij[q] & 0xffff;
                CMD_CHECK(a < n && b < n && a != b);
                const double pa[3] = {s.c[3 * a], s.c[3 * a + 1], s.c[3 * a + 2]};
                const double pb[3] = {s.c[3 * b], s.c[3 * b + 1], s.c[3 * b + 2]};
                double d[3];
                if (KIND == 0) {
                    // reference: length(frame[hi], frame[lo]) (topology.py:62-66); the arithmetic
                    // is sign-symmetric, so the direction does not change a bit
                    diff_ortho_exact(bx, pa, pb, d);
                    d2[q] = norm2_exact(d);
                } else {
                    diff_general_norm_sp<SP>(bx, pa, pb, d);
                    d2[q] = FILT == FILT_F32_IMG ? min_image_norm2_kept(bx, d) : fmin(1e6, norm2_exact(d));
                }
            }
        };
        if (KIND != 0 && bx.sparse == 2) length2_pair(std::integral_constant<int, 2>{});
        else if (KIND != 0 && bx.sparse == 1) length2_pair(std::integral_constant<int, 1>{});
        else length2_pair(std::integral_constant<int, 0>{});
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
    // the coordinates are dead from here on; order the generic-proxy accesses to s.c before the
    // bulk copy that will overwrite it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    // ---- phase 3: row counts -> exclusive offsets (LIL->COO order is row-major), write-out ---
    bool emitted = false;
    int total = 0;
    if (!overflow) {
        const bool use_wpre = fp.wpre_ok != 0;   // coordinates and columns are dead: reuse them
        const int rpt = (n + (int)blockDim.x - 1) / (int)blockDim.x;   // rows per thread: 1 or 2
        int cnt = 0, c0 = 0;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int i = rpt * tid + q;
            if (q < rpt && i < n) {
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
        if (rpt * tid < n) s.rowoff[rpt * tid] = off;
        if (rpt == 2 && 2 * tid + 1 < n) s.rowoff[2 * tid + 1] = off + c0;
        __syncthreads();
        total = s.misc[1];
        if (tid == 0) {
            out_counts[f] = total > stride ? -total : total;
            if (out_rebuilt) out_rebuilt[f] = 1;
            if (total > stride) atomicMax(err, total);
        }
        emitted = total <= stride;
        if (emitted) {
            if (out_rowoff) {   // row index of the frame's list: row i = [rowoff[i], rowoff[i + 1])
                int *ro = out_rowoff + f * (int64_t)cmd_ro_pitch(n);
                for (int k = tid; k < n; k += blockDim.x) ro[k] = s.rowoff[k];
                if (tid == 0) ro[n] = total;
            }
            // position of (a -> b) = rowoff[a] + #set bits of row a below column b; the hit drops
            // its candidate index there (bit 15: reversed direction)
            for (int h = tid; h < ncand; h += blockDim.x) {
                if (s.hit_d[h] < 0.0) continue;
                const unsigned ij = s.hit_ij[h];
                const int a = ij >> 16, b = ij & 0xffff;
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
                CMD_CHECK(pa >= 0 && pa < total && pb >= 0 && pb < total && pa != pb && total <= 2 * hit_cap);
                s.slot[pa] = (unsigned short)h;
                s.slot[pb] = (unsigned short)(h | 0x8000);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    // wpre (aliasing the coordinate buffer) is dead: bring in the next frame of this CTA
    if (use_tma && tid == 0 && item + item_step < item_hi) prefetch(item + item_step);

    if (emitted) {
        const int64_t base = f * stride;   // stride is a multiple of 64: 16-byte aligned rows
        const int total4 = total & ~3;
        // (start, dest) of the directed pair behind slot entry e
        auto pair_of = [&](unsigned e, int &st, int &de) -> int {
            const int h = (int)(e & 0x7fffu);
            CMD_CHECK(h < ncand && h < hit_cap);
            unsigned ij = s.hit_ij[h];
            if (e & 0x8000u) ij = __funnelshift_l(ij, ij, 16);
            st = (int)(ij >> 16);
            de = (int)(ij & 0xffffu);
            return h;
        };
        // The rates of the hits go where the adjacency mask and the column pairs were (both dead),
        // if they fit: then ONE pass over the output positions writes all four arrays.
        const int om_split = lay.om_split;
        const bool fused = lay.fused != 0;
        double *om_a = (double *)s.mask, *om_b = (double *)s.col - om_split;
        auto om_at = [&](int h) -> double * { return (h < om_split ? om_a : om_b) + h; };
        if (!fused) {
            // pass 1 of 2: (start, dest, dist) in output order, four positions per thread
            for (int p = 4 * tid; p < total4; p += 4 * blockDim.x) {
                const uint2 e2 = *(const uint2 *)(s.slot + p);
                int4 st, de;
                const int h0 = pair_of(e2.x & 0xffffu, st.x, de.x), h1 = pair_of(e2.x >> 16, st.y, de.y);
                const int h2 = pair_of(e2.y & 0xffffu, st.z, de.z), h3 = pair_of(e2.y >> 16, st.w, de.w);
                *(int4 *)(out_start + base + p) = st;
                *(int4 *)(out_dest + base + p) = de;
                *(double2 *)(out_dist + base + p) = make_double2(s.hit_d[h0], s.hit_d[h1]);
                *(double2 *)(out_dist + base + p + 2) = make_double2(s.hit_d[h2], s.hit_d[h3]);
            }
            if (tid < total - total4) {
                int st, de;
                const int h = pair_of(s.slot[total4 + tid], st, de);
                out_start[base + total4 + tid] = st;
                out_dest[base + total4 + tid] = de;
                out_dist[base + total4 + tid] = s.hit_d[h];
            }
            __syncthreads();
        }
        // the rates, once per unordered hit (unfused: they replace the distances)
        for (int h0 = tid; h0 < ncand; h0 += 2 * blockDim.x) {
            const int h1 = h0 + blockDim.x;
            const bool two = h1 < ncand;
            const double dist[2] = {s.hit_d[h0], two ? s.hit_d[h1] : -1.0};
            double om[2];
            rate_eval2(rp, fabs(dist[0]), fabs(dist[1]), om);
            if (dist[0] >= 0.0) *(fused ? om_at(h0) : s.hit_d + h0) = om[0];
            if (two && dist[1] >= 0.0) *(fused ? om_at(h1) : s.hit_d + h1) = om[1];
        }
        __syncthreads();
        // output order; the per-frame total of all listed rates is summed in a fixed order
        double rsum = 0.0;
        if (fused) {
#pragma unroll 2
            for (int p = 4 * tid; p < total4; p += 4 * blockDim.x) {
                const uint2 e2 = *(const uint2 *)(s.slot + p);
                int4 st, de;
                const int h0 = pair_of(e2.x & 0xffffu, st.x, de.x), h1 = pair_of(e2.x >> 16, st.y, de.y);
                const int h2 = pair_of(e2.y & 0xffffu, st.z, de.z), h3 = pair_of(e2.y >> 16, st.w, de.w);
                const double o0 = *om_at(h0), o1 = *om_at(h1), o2 = *om_at(h2), o3 = *om_at(h3);
                rsum += (o0 + o1) + (o2 + o3);
                *(int4 *)(out_start + base + p) = st;
                *(int4 *)(out_dest + base + p) = de;
                *(double2 *)(out_dist + base + p) = make_double2(s.hit_d[h0], s.hit_d[h1]);
                *(double2 *)(out_dist + base + p + 2) = make_double2(s.hit_d[h2], s.hit_d[h3]);
                *(double2 *)(out_omega + base + p) = make_double2(o0, o1);
                *(double2 *)(out_omega + base + p + 2) = make_double2(o2, o3);
            }
            if (tid < total - total4) {
                int st, de;
                const int h = pair_of(s.slot[total4 + tid], st, de);
                const double o = *om_at(h);
                rsum += o;
                out_start[base + total4 + tid] = st;
                out_dest[base + total4 + tid] = de;
                out_dist[base + total4 + tid] = s.hit_d[h];
                out_omega[base + total4 + tid] = o;
            }
        } else {
            for (int p = 4 * tid; p < total4; p += 4 * blockDim.x) {
                const uint2 e2 = *(const uint2 *)(s.slot + p);
                const double o0 = s.hit_d[e2.x & 0x7fffu], o1 = s.hit_d[(e2.x >> 16) & 0x7fffu];
                const double o2 = s.hit_d[e2.y & 0x7fffu], o3 = s.hit_d[(e2.y >> 16) & 0x7fffu];
                rsum += (o0 + o1) + (o2 + o3);
                *(double2 *)(out_omega + base + p) = make_double2(o0, o1);
                *(double2 *)(out_omega + base + p + 2) = make_double2(o2, o3);
            }
            if (tid < total - total4) {
                const double o = s.hit_d[s.slot[total4 + tid] & 0x7fffu];
                rsum += o;
                out_omega[base + total4 + tid] = o;
            }
        }
        if (out_rate_sum) {
            for (int o = 16; o > 0; o >>= 1) rsum += __shfl_down_sync(0xffffffffu, rsum, o);
            if (lane == 0) s.red[wid] = rsum;
            __syncthreads();
            if (tid == 0) {
                double tsum = 0;
                for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) tsum += s.red[w];
                out_rate_sum[f] = tsum;
            }
        }
    }
    __syncthreads();   // the emit passes are done with the hit lists before the next frame resets them
    }   // frames of this CTA
}
