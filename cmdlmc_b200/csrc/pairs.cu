// pairs.cu -- neighbour topology on the GPU (rows A7, A8, A9 of SURVEY.md section 8).
//
//   k_pairs_dense   one CTA per frame: all N(N-1)/2 unordered pairs of a frame are evaluated from
//                   shared memory, pairs within cutoff+buffer are emitted in the reference's COO
//                   order (both directions, row-major, columns ascending -- topology.py:55-72)
//                   together with the jump rate of each pair.  Replaces get_topology_bruteforce.
//   k_dr / k_schedule / k_refresh
//                   the Verlet-list generator (topology.py:80-114): per-frame displacements, the
//                   sequential rebuild decision, and the distance refresh of the kept list.
//
// Bit-exactness: every emitted distance and every `dist <= cutoff+buffer` decision is computed
// with the reference's arithmetic (pbc.cuh *_exact).  For general cells a cheap FMA filter with a
// pruned image set rejects far pairs first; it is conservative (relative margin 1e-9) and never
// decides a hit on its own.
#include <math.h>
#include <stdlib.h>

#include "pbc.cuh"

struct cmd_topo {
    BoxParams bx;
    RateParams rate;
    int n;
    double cutoff, buffer, rc;
    double t2;       // largest d2 with sqrt(d2) <= rc  (exact decision on the squared length)
    double lsum;     // sum |h_ij|: scale of the filter's rounding-error allowance
    int mode;
    int64_t stride;  // per-frame pair capacity
    int hit_cap;     // unordered hits per frame that fit the CTA's shared-memory list
    size_t smem_bytes;
    int threads;
    // block results
    int64_t cap_frames, nframes;
    int *d_start, *d_dest, *d_counts, *d_err;
    double *d_dist, *d_omega, *d_rate_sum;
    uint8_t *d_rebuilt;
    unsigned long long *d_ties;
    // Verlet state carried across blocks
    bool have_last;
    double *d_last, *d_displacement, *d_dr;
    int *d_carry_start, *d_carry_dest, *d_carry_count;
    int *d_sched;  // [0] n_rebuild, [1] n_refresh, [2] last head (-1 = carry), then ids
    int *d_rebuild_ids, *d_refresh_ids, *d_head;
    double *d_upload;
    size_t upload_bytes;
    const double *d_frames_last;  // frames of the last block (device)
    int64_t total_frames;
};

// ------------------------------------------------------------------ shared-memory layout ------
struct DenseSmem {
    double *cx, *cy, *cz;   // [n] Cartesian coordinates of the frame (SoA)
    double *sx, *sy, *sz;   // [n] wrapped fractional coordinates (general cells only)
    double *hit_d;          // [hit_cap] candidate d^2 (ortho) -> distance of a hit, < 0 otherwise
    unsigned *mask;         // [n][W] adjacency bit matrix
    unsigned *hit_ij;       // [hit_cap] (i << 16) | j
    int *rowoff;            // [n + 1]
    int *misc;              // [0] ncand, [1] total, [2..33] warp sums
    double *red;            // [34] block reductions
};

__host__ __device__ inline size_t dense_smem_bytes(int n, int hit_cap, int kind)
{
    int W = (n + 31) / 32;
    size_t b = 0;
    b += (kind ? 6 : 3) * (size_t)n * 8;
    b += (size_t)hit_cap * 8;
    b += 40 * 8;
    b += (size_t)n * W * 4;
    b += (size_t)hit_cap * 4;
    b += ((size_t)n + 1) * 4;
    b += 40 * 4;
    return b + 16;
}

__device__ __forceinline__ DenseSmem dense_carve(unsigned char *base, int n, int hit_cap, int kind)
{
    DenseSmem s;
    int W = (n + 31) / 32;
    s.cx = (double *)base;
    s.cy = s.cx + n;
    s.cz = s.cy + n;
    double *nx = s.cz + n;
    s.sx = s.sy = s.sz = nullptr;
    if (kind) { s.sx = nx; s.sy = s.sx + n; s.sz = s.sy + n; nx = s.sz + n; }
    s.hit_d = nx;
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

// ------------------------------------------------------------------ all-pairs kernel ----------
// One CTA per frame.  grid.x = number of frames to (re)build; frame = ids ? ids[blockIdx.x] :
// blockIdx.x.  Three phases, all out of shared memory:
//   1 filter   every unordered pair once (cyclic pairing, thread i <-> row i): FMA arithmetic on
//              pre-wrapped fractional coordinates (general cells) or the exact orthorhombic wrap;
//              pairs within the (conservatively widened) radius are appended to a candidate list
//              with one warp-aggregated shared-memory atomic -- no divergent heavy path;
//   2 exact    the candidates, densely packed over the CTA, in the reference's arithmetic: the
//              `dist <= cutoff + buffer` decision, sqrt, adjacency bits;
//   3 emit     row offsets from the adjacency bit matrix (popc scan) and the ordered write of
//              (start, dest, dist, omega) with the jump rate evaluated per hit.
template <int KIND, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_pairs_dense(const __grid_constant__ BoxParams bx, const __grid_constant__ RateParams rp,
              const double *__restrict__ frames, const int *__restrict__ ids,
              const int *__restrict__ n_ids, int n, double rc, double t2, double lsum,
              int64_t stride, int hit_cap, int *__restrict__ out_start, int *__restrict__ out_dest,
              double *__restrict__ out_dist, double *__restrict__ out_omega,
              int *__restrict__ out_counts, double *__restrict__ out_rate_sum,
              uint8_t *__restrict__ out_rebuilt, int *__restrict__ err,
              unsigned long long *__restrict__ ties)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (n_ids && (int)blockIdx.x >= *n_ids) return;
    const int64_t f = ids ? ids[blockIdx.x] : blockIdx.x;
    DenseSmem s = dense_carve(smem_raw, n, hit_cap, KIND);
    const int W = (n + 31) / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double *fr = frames + f * (int64_t)n * 3;

    // stage the frame: contiguous, coalesced read of 3n doubles, de-interleaved into SoA
    for (int k = tid; k < 3 * n; k += blockDim.x) {
        double v = __ldg(fr + k);
        int a = k / 3, c = k - 3 * a;
        (c == 0 ? s.cx : c == 1 ? s.cy : s.cz)[a] = v;
    }
    for (int k = tid; k < n * W; k += blockDim.x) s.mask[k] = 0u;
    if (tid == 0) { s.misc[0] = 0; s.misc[1] = 0; }
    __syncthreads();

    const int i = tid;
    double t2f = t2;
    if (KIND == 1) {
        // wrapped fractional coordinates, once per atom; e bounds their rounding error
        double e = 0.0;
        if (i < n) {
            const double x = s.cx[i], y = s.cy[i], z = s.cz[i];
            double q[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                double v = fma(bx.hinv[3 * c + 2], z, fma(bx.hinv[3 * c + 1], y, bx.hinv[3 * c] * x));
                double a = fma(fabs(bx.hinv[3 * c + 2]), fabs(z),
                               fma(fabs(bx.hinv[3 * c + 1]), fabs(y), fabs(bx.hinv[3 * c] * x)));
                e = fmax(e, a);
                q[c] = v - rint(v);
            }
            s.sx[i] = q[0]; s.sy[i] = q[1]; s.sz[i] = q[2];
        }
        for (int o = 16; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, o));
        if (lane == 0) s.red[wid] = e;
        __syncthreads();
        e = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) e = fmax(e, s.red[w]);
        // filter radius: rc widened by 1e-9 relative plus the worst Cartesian error of a wrapped
        // difference (8 ulp of the largest fractional magnitude times sum |h|)
        const double thr = rc * (1.0 + 1e-9) + 8.0 * 2.3e-16 * (e + 1.0) * lsum;
        t2f = thr * thr;
    }

    // ---- phase 1: filter -------------------------------------------------------------------
    {
        const int half = n >> 1;
        const bool even = (n & 1) == 0;
        double xi = 0, yi = 0, zi = 0;
        if (i < n) {
            if (KIND == 1) { xi = s.sx[i]; yi = s.sy[i]; zi = s.sz[i]; }
            else { xi = s.cx[i]; yi = s.cy[i]; zi = s.cz[i]; }
        }
        const double pi_[3] = {xi, yi, zi};
        // cyclic pairing: thread i owns pairs (i, i+k mod n), k = 1..floor(n/2); for even n the
        // k = n/2 column is owned by the lower half only -> every unordered pair exactly once
        for (int k = 1; k <= half; k++) {
            const bool valid = i < n && !(even && k == half && i >= half);
            int j = i + k;
            if (j >= n) j -= n;
            if (!valid) j = 0;
            double d2;
            bool cand;
            if (KIND == 0) {
                // reference: length(frame[hi], frame[lo]) (topology.py:62-66); the arithmetic is
                // sign-symmetric, so the direction does not change a bit
                const double pj[3] = {s.cx[j], s.cy[j], s.cz[j]};
                double d[3];
                diff_ortho_exact(bx, pi_, pj, d);
                d2 = norm2_exact(d);
                cand = bx.conv == CMD_CONV_NONE ? d2 <= t2 : convert_distance(bx, sqrt(d2)) <= rc;
            } else {
                double a = s.sx[j] - xi, b = s.sy[j] - yi, c = s.sz[j] - zi;
                a -= rint_magic(a); b -= rint_magic(b); c -= rint_magic(c);
                const double vx = fma(bx.h[2], c, fma(bx.h[1], b, bx.h[0] * a));
                const double vy = fma(bx.h[5], c, fma(bx.h[4], b, bx.h[3] * a));
                const double vz = fma(bx.h[8], c, fma(bx.h[7], b, bx.h[6] * a));
                d2 = fma(vz, vz, fma(vy, vy, vx * vx));
                for (int m = 0; m < bx.n_img; m++) {
                    double ux = vx + bx.img[m][0], uy = vy + bx.img[m][1], uz = vz + bx.img[m][2];
                    d2 = fmin(d2, fma(uz, uz, fma(uy, uy, ux * ux)));
                }
                cand = d2 <= t2f;
            }
            cand = cand && valid;
            const unsigned bal = __ballot_sync(0xffffffffu, cand);
            if (bal) {
                int slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&s.misc[0], __popc(bal));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                const int slot = slot0 + __popc(bal & ((1u << lane) - 1u));
                if (cand && slot < hit_cap) {
                    s.hit_ij[slot] = ((unsigned)i << 16) | (unsigned)j;
                    if (KIND == 0) s.hit_d[slot] = d2;
                }
            }
        }
    }
    __syncthreads();
    const int ncand = s.misc[0];
    if (ncand > hit_cap) {  // capacity probe / overflow: report the (upper bound of the) need
        if (tid == 0) {
            out_counts[f] = -2 * ncand;
            if (out_rebuilt) out_rebuilt[f] = 1;
            atomicMax(err, 2 * ncand);
        }
        return;
    }

    // ---- phase 2: exact evaluation of the candidates ----------------------------------------
    unsigned long long my_ties = 0;
    for (int c = tid; c < ncand; c += blockDim.x) {
        const unsigned ij = s.hit_ij[c];
        const int a = ij >> 16, b = ij & 0xffff;
        double d2;
        if (KIND == 0) d2 = s.hit_d[c];
        else {
            const double pa[3] = {s.cx[a], s.cy[a], s.cz[a]}, pb[3] = {s.cx[b], s.cy[b], s.cz[b]};
            double d[3];
            diff_general_exact(bx, pa, pb, d);
            d2 = min_image_norm2_kept(bx, d);
        }
        const double dist = convert_distance(bx, sqrt(d2));
        const bool hit = (bx.conv == CMD_CONV_NONE ? d2 <= t2 : dist <= rc) && dist != 0.0;
        if (fabs(dist - rc) <= 1e-11 * rc) my_ties++;
        s.hit_d[c] = hit ? dist : -1.0;
        if (hit) {
            atomicOr(&s.mask[a * W + (b >> 5)], 1u << (b & 31));
            atomicOr(&s.mask[b * W + (a >> 5)], 1u << (a & 31));
        }
    }
    if (my_ties) atomicAdd(ties, my_ties);
    __syncthreads();

    // ---- phase 3: row counts -> exclusive offsets (LIL->COO order is row-major), write-out ---
    int cnt = 0;
    if (i < n)
        for (int w = 0; w < W; w++) cnt += __popc(s.mask[i * W + w]);
    int off = block_exclusive_scan(cnt, s.misc + 2, &s.misc[1]);
    if (i < n) s.rowoff[i] = off;
    __syncthreads();
    const int total = s.misc[1];
    if (tid == 0) {
        out_counts[f] = total > stride ? -total : total;
        if (out_rebuilt) out_rebuilt[f] = 1;
        if (total > stride) atomicMax(err, total);
    }
    if (total > stride) return;

    // position of (a -> b) = rowoff[a] + #set bits of row a below column b
    const int64_t base = f * stride;
    double rsum = 0.0;
    for (int h = tid; h < ncand; h += blockDim.x) {
        const double dist = s.hit_d[h];
        if (dist < 0.0) continue;
        const unsigned ij = s.hit_ij[h];
        const int a = ij >> 16, b = ij & 0xffff;
        const double om = rate_eval(rp, dist, 0.0);
        rsum += om;
        int pa = s.rowoff[a], pb = s.rowoff[b];
        for (int w = 0; w < (b >> 5); w++) pa += __popc(s.mask[a * W + w]);
        pa += __popc(s.mask[a * W + (b >> 5)] & ((1u << (b & 31)) - 1u));
        for (int w = 0; w < (a >> 5); w++) pb += __popc(s.mask[b * W + w]);
        pb += __popc(s.mask[b * W + (a >> 5)] & ((1u << (a & 31)) - 1u));
        out_start[base + pa] = a; out_dest[base + pa] = b;
        out_dist[base + pa] = dist; out_omega[base + pa] = om;
        out_start[base + pb] = b; out_dest[base + pb] = a;
        out_dist[base + pb] = dist; out_omega[base + pb] = om;
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
}

// ------------------------------------------------------------------ Verlet pieces -------------
// dr[f][i] = length(frame[f-1][i], frame[f][i])  (topology.py:98); f = 0 uses the carried frame
__global__ void __launch_bounds__(256) k_dr(const __grid_constant__ BoxParams bx,
                                            const double *__restrict__ frames,
                                            const double *__restrict__ last, int have_last, int n,
                                            int64_t nframes, double *__restrict__ dr)
{
    int64_t total = nframes * n;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total;
         g += (int64_t)gridDim.x * blockDim.x) {
        int64_t f = g / n;
        int i = (int)(g - f * n);
        const double *cur = frames + (f * n + i) * 3;
        const double *prev = f > 0 ? cur - (int64_t)n * 3 : (have_last ? last + (int64_t)i * 3 : nullptr);
        double v = 0.0;  // topology.py:95: first frame -> dr = zeros
        if (prev) {
            double a[3] = {prev[0], prev[1], prev[2]}, b[3] = {cur[0], cur[1], cur[2]};
            v = length_exact(bx, a, b);
        }
        dr[g] = v;
    }
}

// Sequential rebuild decision (topology.py:100-107), one CTA, frames in order:
//   displacement += dr;  m1, m2 = two largest;  if m1 + m2 > buffer: rebuild, displacement = 0.
// The very first frame of the trajectory is always built (topology.py:91-93) and then follows the
// same rule.  Emits the compacted id lists the build / refresh kernels consume.
__global__ void __launch_bounds__(1024, 1)
k_schedule(const double *__restrict__ dr, double *__restrict__ displacement, int n,
           int64_t nframes, double buffer, int first_ever, int *__restrict__ sched,
           int *__restrict__ rebuild_ids, int *__restrict__ refresh_ids, int *__restrict__ head,
           uint8_t *__restrict__ rebuilt)
{
    __shared__ double s1[32], s2[32];
    __shared__ int decision;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = (blockDim.x + 31) >> 5;
    // each thread owns atoms tid, tid + blockDim, ... (n may exceed the CTA size)
    const int per = (n + blockDim.x - 1) / blockDim.x;
    int n_rebuild = 0, n_refresh = 0, cur_head = sched[2];
    for (int64_t f = 0; f < nframes; f++) {
        double m1 = -INFINITY, m2 = -INFINITY;
        for (int q = 0; q < per; q++) {
            int a = tid + q * blockDim.x;
            if (a < n) {
                double v = __dadd_rn(displacement[a], dr[f * n + a]);
                displacement[a] = v;
                if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            double o1 = __shfl_down_sync(0xffffffffu, m1, o), o2 = __shfl_down_sync(0xffffffffu, m2, o);
            if (o1 > m1) { m2 = fmax(m1, o2); m1 = o1; } else m2 = fmax(m2, o1);
        }
        if (lane == 0) { s1[w] = m1; s2[w] = m2; }
        __syncthreads();
        if (w == 0) {
            m1 = lane < nw ? s1[lane] : -INFINITY;
            m2 = lane < nw ? s2[lane] : -INFINITY;
            for (int o = 16; o > 0; o >>= 1) {
                double o1 = __shfl_down_sync(0xffffffffu, m1, o), o2 = __shfl_down_sync(0xffffffffu, m2, o);
                if (o1 > m1) { m2 = fmax(m1, o2); m1 = o1; } else m2 = fmax(m2, o1);
            }
            if (lane == 0) {
                // np.sort(displacement)[-2:] -> (m2, m1); displ_max1 + displ_max2 > buffer
                bool cross = n >= 2 && __dadd_rn(m2, m1) > buffer;
                decision = (cross ? 1 : 0) | ((first_ever && f == 0) ? 2 : 0);
            }
        }
        __syncthreads();
        int dec = decision;
        if (dec & 1) {  // rebuild: this frame's list comes from the all-pairs kernel
            for (int q = 0; q < per; q++) {
                int a = tid + q * blockDim.x;
                if (a < n) displacement[a] = 0.0;
            }
            if (tid == 0) { rebuild_ids[n_rebuild] = (int)f; head[f] = (int)f; rebuilt[f] = 1; }
            n_rebuild++;
            cur_head = (int)f;
        } else if (dec & 2) {
            // first frame ever, no crossing: built by brute force, then refreshed with the same
            // values (topology.py:91-93,108-111) -- the build alone yields identical arrays
            if (tid == 0) { rebuild_ids[n_rebuild] = (int)f; head[f] = (int)f; rebuilt[f] = 1; }
            n_rebuild++;
            cur_head = (int)f;
        } else {
            if (tid == 0) { refresh_ids[n_refresh] = (int)f; head[f] = cur_head; rebuilt[f] = 0; }
            n_refresh++;
        }
        __syncthreads();
    }
    if (tid == 0) { sched[0] = n_rebuild; sched[1] = n_refresh; sched[2] = cur_head; }
}

// Refresh of a kept list (topology.py:110): dist = length(frame[row], frame[col]), same pairs.
// One CTA per refreshed frame; the head list is either a frame of this block or the carry.
__global__ void __launch_bounds__(256)
k_refresh(const __grid_constant__ BoxParams bx, const __grid_constant__ RateParams rp,
          const double *__restrict__ frames, const int *__restrict__ ids,
          const int *__restrict__ n_ids, const int *__restrict__ head, int n, int64_t stride,
          const int *__restrict__ carry_start, const int *__restrict__ carry_dest,
          const int *__restrict__ carry_count, int *__restrict__ out_start,
          int *__restrict__ out_dest, double *__restrict__ out_dist,
          double *__restrict__ out_omega, int *__restrict__ out_counts,
          double *__restrict__ out_rate_sum)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if ((int)blockIdx.x >= *n_ids) return;
    double *sp = (double *)smem_raw;  // [3n] AoS copy of the frame
    const int64_t f = ids[blockIdx.x];
    const int hd = head[f];
    const double *fr = frames + f * (int64_t)n * 3;
    for (int k = threadIdx.x; k < 3 * n; k += blockDim.x) sp[k] = __ldg(fr + k);
    // out_counts of a head frame in this block is final: the build kernel ran before us
    const int p = hd < 0 ? *carry_count : out_counts[hd];
    const int *hs = hd < 0 ? carry_start : out_start + hd * stride;
    const int *hdst = hd < 0 ? carry_dest : out_dest + hd * stride;
    __syncthreads();
    const int64_t base = f * stride;
    double rsum = 0.0;
    for (int k = threadIdx.x; k < p; k += blockDim.x) {
        int a = hs[k], b = hdst[k];
        double pa[3] = {sp[3 * a], sp[3 * a + 1], sp[3 * a + 2]};
        double pb[3] = {sp[3 * b], sp[3 * b + 1], sp[3 * b + 2]};
        double dist = length_exact(bx, pa, pb);
        double om = rate_eval(rp, dist, 0.0);
        rsum += om;
        out_start[base + k] = a; out_dest[base + k] = b;
        out_dist[base + k] = dist; out_omega[base + k] = om;
    }
    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_down_sync(0xffffffffu, rsum, o);
    __shared__ double wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = rsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += wsum[w];
        out_rate_sum[f] = t;
        out_counts[f] = p;
    }
}

// keeps the list of the current segment head + the last frame for the next block
__global__ void __launch_bounds__(256)
k_carry(const int *__restrict__ sched, const int *__restrict__ out_start,
        const int *__restrict__ out_dest, const int *__restrict__ out_counts, int64_t stride,
        int *__restrict__ carry_start, int *__restrict__ carry_dest, int *__restrict__ carry_count)
{
    int hd = sched[2];
    if (hd < 0) return;  // the head is still the carried list
    int p = out_counts[hd];
    if (p < 0) p = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < p; k += gridDim.x * blockDim.x) {
        carry_start[k] = out_start[hd * stride + k];
        carry_dest[k] = out_dest[hd * stride + k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *carry_count = p;
}

__global__ void k_upcast_f32(const float *__restrict__ in, double *__restrict__ out, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (double)in[i];
}

// ------------------------------------------------------------------ host side ------------------
static double exact_sq_threshold(double rc)
{
    // largest double t with sqrt(t) <= rc under IEEE round-to-nearest sqrt
    double t = rc * rc;
    while (sqrt(nextafter(t, INFINITY)) <= rc) t = nextafter(t, INFINITY);
    while (sqrt(t) > rc) t = nextafter(t, -INFINITY);
    return t;
}

static void topo_free_block(cmd_topo *t)
{
    cudaFree(t->d_start); cudaFree(t->d_dest); cudaFree(t->d_dist); cudaFree(t->d_omega);
    cudaFree(t->d_counts); cudaFree(t->d_rate_sum); cudaFree(t->d_rebuilt); cudaFree(t->d_dr);
    cudaFree(t->d_rebuild_ids); cudaFree(t->d_refresh_ids); cudaFree(t->d_head);
    t->d_start = t->d_dest = t->d_counts = nullptr;
    t->d_dist = t->d_omega = t->d_rate_sum = t->d_dr = nullptr;
    t->d_rebuilt = nullptr;
    t->d_rebuild_ids = t->d_refresh_ids = t->d_head = nullptr;
    t->cap_frames = 0;
}

extern "C" void cmd_topo_destroy(cmd_topo *t)
{
    if (!t) return;
    cudaStreamSynchronize(cmd_global().stream);
    topo_free_block(t);
    cudaFree(t->d_err); cudaFree(t->d_ties); cudaFree(t->d_last); cudaFree(t->d_displacement);
    cudaFree(t->d_carry_start); cudaFree(t->d_carry_dest); cudaFree(t->d_carry_count);
    cudaFree(t->d_sched); cudaFree(t->d_upload);
    free(t);
}

static int topo_configure(cmd_topo *t, int64_t stride)
{
    // per-frame capacity and the matching shared-memory hit list
    stride = (stride + 63) / 64 * 64;
    int hit_cap = (int)(stride / 2);
    size_t smem = dense_smem_bytes(t->n, hit_cap, t->bx.kind);
    if (smem > 226 * 1024)
        return cmd_set_error(CMD_ECAPACITY,
                             "dense pair kernel needs %zu bytes of shared memory for n=%d, "
                             "capacity %lld (limit 232448): use the cell-list path", smem, t->n,
                             (long long)stride);
    t->stride = stride;
    t->hit_cap = hit_cap;
    t->smem_bytes = smem;
    int th = (t->n + 31) / 32 * 32;
    t->threads = th < 64 ? 64 : th;
    return CMD_OK;
}

extern "C" int cmd_topo_create(const cmd_box *box, int n, double cutoff, double buffer, int mode,
                               int rate_kind, const double par[CMD_RATE_NPAR], int64_t capacity,
                               cmd_topo **out)
{
    CMD_REQUIRE_INIT();
    if (!box || !out || n < 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (n > 1024)
        return cmd_set_error(CMD_EINVAL, "dense topology supports n_atoms <= 1024 (got %d); "
                                         "larger systems need the cell-list path", n);
    if (mode != CMD_TOPO_BRUTEFORCE && mode != CMD_TOPO_VERLET)
        return cmd_set_error(CMD_EINVAL, "bad topology mode %d", mode);
    if (rate_kind < 0 || rate_kind > CMD_RATE_EXP || rate_kind == CMD_RATE_FERMI_ANGLE)
        return cmd_set_error(CMD_EINVAL, "rate kind %d is not a pure distance function", rate_kind);
    if (!(cutoff + buffer >= 0)) return cmd_set_error(CMD_EINVAL, "cutoff + buffer must be >= 0");
    cmd_topo *t = (cmd_topo *)calloc(1, sizeof(cmd_topo));
    if (!t) return cmd_set_error(CMD_ENOMEM, "out of host memory");
    t->bx = box->p;
    t->rate.kind = rate_kind;
    if (par) memcpy(t->rate.par, par, sizeof(t->rate.par));
    t->n = n;
    t->cutoff = cutoff;
    t->buffer = buffer;
    t->rc = cutoff + buffer;  // topology.py:67 adds them in double exactly like this
    t->t2 = exact_sq_threshold(t->rc);
    // the filter only has to look at the images that can come within rc of the origin
    cmd_box_prune_images(t->bx, t->rc);
    t->lsum = 0;
    for (int c = 0; c < 9; c++) t->lsum += fabs(t->bx.h[c]);
    t->mode = mode;
    int rc = topo_configure(t, capacity > 0 ? capacity : 0);
    if (rc) { free(t); return rc; }
    if (capacity <= 0) t->stride = 0;  // sized from the first frame
#define TALLOC(ptr, bytes)                                                       \
    if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) {                   \
        cudaGetLastError();                                                      \
        cmd_topo_destroy(t);                                                     \
        return cmd_set_error(CMD_ENOMEM, "cudaMalloc failed for %s", #ptr);      \
    }
    TALLOC(t->d_err, sizeof(int));
    TALLOC(t->d_ties, sizeof(unsigned long long));
    TALLOC(t->d_last, (size_t)n * 24);
    TALLOC(t->d_displacement, (size_t)n * 8);
    TALLOC(t->d_carry_count, sizeof(int));
    TALLOC(t->d_sched, 4 * sizeof(int));
    cudaStream_t st = cmd_global().stream;
    CMD_CUDA(cudaMemsetAsync(t->d_err, 0, sizeof(int), st));
    CMD_CUDA(cudaMemsetAsync(t->d_ties, 0, sizeof(unsigned long long), st));
    CMD_CUDA(cudaMemsetAsync(t->d_displacement, 0, (size_t)n * 8, st));
    CMD_CUDA(cudaMemsetAsync(t->d_carry_count, 0, sizeof(int), st));
    int sched0[4] = {0, 0, -1, 0};
    CMD_CUDA(cudaMemcpyAsync(t->d_sched, sched0, sizeof(sched0), cudaMemcpyHostToDevice, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    *out = t;
    return CMD_OK;
}

static int launch_dense(cmd_topo *t, const double *d_frames, const int *ids, const int *n_ids,
                        int64_t grid, int *start, int *dest, double *dist, double *omega,
                        int *counts, double *rate_sum, uint8_t *rebuilt, int64_t stride,
                        int hit_cap, size_t smem)
{
    cudaStream_t st = cmd_global().stream;
    const bool ortho = t->bx.kind == 0;
#define DENSE_LAUNCH(K, MT, MB)                                                                  \
    do {                                                                                         \
        CMD_CUDA(cudaFuncSetAttribute(k_pairs_dense<K, MT, MB>,                                  \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        k_pairs_dense<K, MT, MB><<<(unsigned)grid, t->threads, smem, st>>>(                      \
            t->bx, t->rate, d_frames, ids, n_ids, t->n, t->rc, t->t2, t->lsum, stride,           \
            hit_cap, start, dest, dist, omega, counts, rate_sum, rebuilt, t->d_err, t->d_ties);  \
    } while (0)
    if (t->threads <= 256) {
        if (ortho) DENSE_LAUNCH(0, 256, 3); else DENSE_LAUNCH(1, 256, 3);
    } else if (t->threads <= 512) {
        if (ortho) DENSE_LAUNCH(0, 512, 2); else DENSE_LAUNCH(1, 512, 2);
    } else {
        if (ortho) DENSE_LAUNCH(0, 1024, 1); else DENSE_LAUNCH(1, 1024, 1);
    }
#undef DENSE_LAUNCH
    CMD_LAUNCHED();
    return CMD_OK;
}

// sizes the per-frame capacity from a probe of one frame (count-only: nothing fits, so the
// kernel reports -P through out_counts and err)
static int topo_autosize(cmd_topo *t, const double *d_frame)
{
    cudaStream_t st = cmd_global().stream;
    int *d_cnt;
    int rc = cmd_scratch(4, 64, (void **)&d_cnt);
    if (rc) return rc;
    t->threads = ((t->n + 31) / 32 * 32) < 64 ? 64 : (t->n + 31) / 32 * 32;
    size_t smem = dense_smem_bytes(t->n, 0, t->bx.kind);
    CMD_CUDA(cudaMemsetAsync(t->d_err, 0, sizeof(int), st));
    rc = launch_dense(t, d_frame, nullptr, nullptr, 1, nullptr, nullptr, nullptr, nullptr, d_cnt,
                      nullptr, nullptr, 0, 0, smem);
    if (rc) return rc;
    int cnt = 0;
    CMD_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    CMD_CUDA(cudaMemsetAsync(t->d_err, 0, sizeof(int), st));
    CMD_CUDA(cudaMemsetAsync(t->d_ties, 0, sizeof(unsigned long long), st));
    int64_t p0 = cnt < 0 ? -cnt : cnt;
    int64_t want = p0 + p0 / 2 + 128;
    // shrink to what the shared-memory hit list can hold
    while (want > p0 + 64 && dense_smem_bytes(t->n, (int)((want + 63) / 64 * 64 / 2), t->bx.kind) > 226 * 1024)
        want -= 64;
    return topo_configure(t, want);
}

static int topo_reserve(cmd_topo *t, int64_t nframes)
{
    if (nframes <= t->cap_frames) return CMD_OK;
    CMD_CUDA(cudaStreamSynchronize(cmd_global().stream));
    topo_free_block(t);
    size_t np = (size_t)nframes * t->stride;
#define BALLOC(ptr, bytes)                                                                       \
    if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) {                                   \
        cudaGetLastError();                                                                      \
        topo_free_block(t);                                                                      \
        return cmd_set_error(CMD_ENOMEM, "cudaMalloc of %zu bytes failed for %s (%lld frames x " \
                                         "%lld pairs)", (size_t)(bytes), #ptr, (long long)nframes, \
                             (long long)t->stride);                                              \
    }
    BALLOC(t->d_start, np * 4);
    BALLOC(t->d_dest, np * 4);
    BALLOC(t->d_dist, np * 8);
    BALLOC(t->d_omega, np * 8);
    BALLOC(t->d_counts, (size_t)nframes * 4);
    BALLOC(t->d_rate_sum, (size_t)nframes * 8);
    BALLOC(t->d_rebuilt, (size_t)nframes);
    if (t->mode == CMD_TOPO_VERLET) {
        BALLOC(t->d_dr, (size_t)nframes * t->n * 8);
        BALLOC(t->d_rebuild_ids, (size_t)nframes * 4);
        BALLOC(t->d_refresh_ids, (size_t)nframes * 4);
        BALLOC(t->d_head, (size_t)nframes * 4);
    }
    t->cap_frames = nframes;
    return CMD_OK;
}

extern "C" int cmd_topo_build_dev(cmd_topo *t, const double *d_frames, int64_t nframes)
{
    CMD_REQUIRE_INIT();
    if (!t || !d_frames || nframes < 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (nframes > 0x7fffffff / 2) return cmd_set_error(CMD_EINVAL, "block too large");
    CmdGlobal &g = cmd_global();
    cudaStream_t st = g.stream;
    int rc;
    if (t->stride == 0 && (rc = topo_autosize(t, d_frames))) return rc;
    if ((rc = topo_reserve(t, nframes))) return rc;
    if (t->mode == CMD_TOPO_VERLET && !t->d_carry_start) {
        if (cudaMalloc((void **)&t->d_carry_start, t->stride * 4) != cudaSuccess ||
            cudaMalloc((void **)&t->d_carry_dest, t->stride * 4) != cudaSuccess) {
            cudaGetLastError();
            return cmd_set_error(CMD_ENOMEM, "cudaMalloc failed for the carried pair list");
        }
    }
    t->nframes = nframes;
    t->d_frames_last = d_frames;
    if (t->mode == CMD_TOPO_BRUTEFORCE) {
        rc = launch_dense(t, d_frames, nullptr, nullptr, nframes, t->d_start, t->d_dest, t->d_dist,
                          t->d_omega, t->d_counts, t->d_rate_sum, t->d_rebuilt, t->stride,
                          t->hit_cap, t->smem_bytes);
        if (rc) return rc;
    } else {
        int blocks = cmd_div_up(nframes * t->n, 256);
        if (blocks > g.sm_count * 16) blocks = g.sm_count * 16;
        k_dr<<<blocks, 256, 0, st>>>(t->bx, d_frames, t->d_last, t->have_last ? 1 : 0, t->n, nframes,
                                     t->d_dr);
        CMD_LAUNCHED();
        int sth = (t->n + 31) / 32 * 32;
        if (sth > 1024) sth = 1024;
        k_schedule<<<1, sth, 0, st>>>(t->d_dr, t->d_displacement, t->n, nframes, t->buffer,
                                      t->total_frames == 0 ? 1 : 0, t->d_sched, t->d_rebuild_ids,
                                      t->d_refresh_ids, t->d_head, t->d_rebuilt);
        CMD_LAUNCHED();
        rc = launch_dense(t, d_frames, t->d_rebuild_ids, t->d_sched, nframes, t->d_start, t->d_dest,
                          t->d_dist, t->d_omega, t->d_counts, t->d_rate_sum, nullptr, t->stride,
                          t->hit_cap, t->smem_bytes);
        if (rc) return rc;
        size_t rsmem = (size_t)t->n * 24;
        k_refresh<<<(unsigned)nframes, 256, rsmem, st>>>(
            t->bx, t->rate, d_frames, t->d_refresh_ids, t->d_sched + 1, t->d_head, t->n, t->stride,
            t->d_carry_start, t->d_carry_dest, t->d_carry_count, t->d_start, t->d_dest, t->d_dist,
            t->d_omega, t->d_counts, t->d_rate_sum);
        CMD_LAUNCHED();
        k_carry<<<8, 256, 0, st>>>(t->d_sched, t->d_start, t->d_dest, t->d_counts, t->stride,
                                   t->d_carry_start, t->d_carry_dest, t->d_carry_count);
        CMD_LAUNCHED();
        CMD_CUDA(cudaMemcpyAsync(t->d_last, d_frames + (nframes - 1) * (int64_t)t->n * 3,
                                 (size_t)t->n * 24, cudaMemcpyDeviceToDevice, st));
        // after this block the head, if any, lives in the carry
        CMD_CUDA(cudaMemsetAsync(t->d_sched + 2, 0xff, sizeof(int), st));
        t->have_last = true;
    }
    t->total_frames += nframes;
    // capacity check (one 4-byte read-back per block)
    int err = 0;
    CMD_CUDA(cudaMemcpyAsync(&err, t->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    if (err > 0) {
        CMD_CUDA(cudaMemsetAsync(t->d_err, 0, sizeof(int), st));
        return cmd_set_error(CMD_ECAPACITY, "a frame has %d directed pairs but the per-frame "
                             "capacity is %lld: re-create the topology with a larger "
                             "capacity_per_frame", err, (long long)t->stride);
    }
    return CMD_OK;
}

extern "C" int cmd_topo_build(cmd_topo *t, const void *h_frames, int dtype_bytes, int64_t nframes)
{
    CMD_REQUIRE_INIT();
    if (!t || !h_frames || nframes < 1 || (dtype_bytes != 4 && dtype_bytes != 8))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    size_t elems = (size_t)nframes * t->n * 3;
    size_t need = elems * 8 + (dtype_bytes == 4 ? elems * 4 : 0);
    if (t->upload_bytes < need) {
        CMD_CUDA(cudaStreamSynchronize(st));
        cudaFree(t->d_upload);
        t->d_upload = nullptr;
        t->upload_bytes = 0;
        if (cudaMalloc((void **)&t->d_upload, need) != cudaSuccess) {
            cudaGetLastError();
            return cmd_set_error(CMD_ENOMEM, "cudaMalloc of %zu staging bytes failed", need);
        }
        t->upload_bytes = need;
    }
    if (dtype_bytes == 8) {
        CMD_CUDA(cudaMemcpyAsync(t->d_upload, h_frames, elems * 8, cudaMemcpyHostToDevice, st));
    } else {
        float *d32 = (float *)(t->d_upload + elems);
        CMD_CUDA(cudaMemcpyAsync(d32, h_frames, elems * 4, cudaMemcpyHostToDevice, st));
        int blocks = cmd_div_up(elems, 256);
        if (blocks > cmd_global().sm_count * 16) blocks = cmd_global().sm_count * 16;
        k_upcast_f32<<<blocks, 256, 0, st>>>(d32, t->d_upload, (int64_t)elems);
        CMD_LAUNCHED();
    }
    return cmd_topo_build_dev(t, t->d_upload, nframes);
}

extern "C" int cmd_topo_frame_info(const cmd_topo *t, int64_t *counts, uint8_t *rebuilt,
                                   double *rate_sum)
{
    CMD_REQUIRE_INIT();
    if (!t || t->nframes < 1) return cmd_set_error(CMD_ESTATE, "no block has been built");
    cudaStream_t st = cmd_global().stream;
    if (counts) {
        int *tmp = (int *)malloc(t->nframes * sizeof(int));
        if (!tmp) return cmd_set_error(CMD_ENOMEM, "out of host memory");
        cudaError_t e = cudaMemcpyAsync(tmp, t->d_counts, t->nframes * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        for (int64_t f = 0; f < t->nframes; f++) counts[f] = tmp[f];
        free(tmp);
        CMD_CUDA(e);
    }
    if (rebuilt) CMD_CUDA(cudaMemcpyAsync(rebuilt, t->d_rebuilt, t->nframes, cudaMemcpyDeviceToHost, st));
    if (rate_sum)
        CMD_CUDA(cudaMemcpyAsync(rate_sum, t->d_rate_sum, t->nframes * 8, cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    return CMD_OK;
}

extern "C" int64_t cmd_topo_stride(const cmd_topo *t) { return t ? t->stride : -1; }
extern "C" int cmd_topo_n_images(const cmd_topo *t) { return t ? t->bx.n_img : -1; }
extern "C" int64_t cmd_topo_nframes(const cmd_topo *t) { return t ? t->nframes : -1; }

extern "C" int cmd_topo_get_frame(const cmd_topo *t, int64_t f, int *start, int *dest, double *dist,
                                  double *omega)
{
    CMD_REQUIRE_INIT();
    if (!t || f < 0 || f >= t->nframes) return cmd_set_error(CMD_EINVAL, "frame out of range");
    cudaStream_t st = cmd_global().stream;
    int p = 0;
    CMD_CUDA(cudaMemcpyAsync(&p, t->d_counts + f, 4, cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    if (p < 0) return cmd_set_error(CMD_ECAPACITY, "frame %lld overflowed its capacity", (long long)f);
    int64_t base = f * t->stride;
    if (start) CMD_CUDA(cudaMemcpyAsync(start, t->d_start + base, (size_t)p * 4, cudaMemcpyDeviceToHost, st));
    if (dest) CMD_CUDA(cudaMemcpyAsync(dest, t->d_dest + base, (size_t)p * 4, cudaMemcpyDeviceToHost, st));
    if (dist) CMD_CUDA(cudaMemcpyAsync(dist, t->d_dist + base, (size_t)p * 8, cudaMemcpyDeviceToHost, st));
    if (omega) CMD_CUDA(cudaMemcpyAsync(omega, t->d_omega + base, (size_t)p * 8, cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    return CMD_OK;
}

extern "C" int cmd_topo_device_arrays(const cmd_topo *t, const int **start, const int **dest,
                                      const double **dist, const double **omega, const int **counts)
{
    if (!t || t->nframes < 1) return cmd_set_error(CMD_ESTATE, "no block has been built");
    if (start) *start = t->d_start;
    if (dest) *dest = t->d_dest;
    if (dist) *dist = t->d_dist;
    if (omega) *omega = t->d_omega;
    if (counts) *counts = t->d_counts;
    return CMD_OK;
}

extern "C" int cmd_topo_positions(const cmd_topo *t, const double **d_frames)
{
    if (!t || t->nframes < 1 || !d_frames) return cmd_set_error(CMD_ESTATE, "no block has been built");
    *d_frames = t->d_frames_last;
    return CMD_OK;
}

extern "C" int64_t cmd_topo_tie_count(const cmd_topo *t)
{
    if (!t || !cmd_global().inited) return -1;
    unsigned long long v = 0;
    cudaStream_t st = cmd_global().stream;
    if (cudaMemcpyAsync(&v, t->d_ties, sizeof(v), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    cudaStreamSynchronize(st);
    return (int64_t)v;
}
