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
#include <cooperative_groups.h>

#include "pairs_cell.cuh"

#ifndef TOPO_UPLOAD_CHUNKS
#define TOPO_UPLOAD_CHUNKS 8    // host->device pipeline depth of cmd_topo_build (measured: 4 -> 3.71, 6 -> 3.55, 8 -> 3.54, 16 -> 3.89 ms per 16384 C2 frames)
#endif

#ifndef DENSE_SPLIT_MID
#define DENSE_SPLIT_MID 2
#endif

struct cmd_topo {
    BoxParams bx;
    BoxParams bx_refresh;   // same box, image set pruned for the refresh radius rc + buffer
    RateParams rate;
    int n;
    double cutoff, buffer, rc;
    double t2;       // largest d2 with sqrt(d2) <= rc  (exact decision on the squared length)
    FilterParams fp; // filter constants of the dense kernel (packed-half and FP32 forms)
    int filt;        // FILT_* variant the dense kernel runs with
    int mode;
    int64_t stride;  // per-frame pair capacity
    int64_t capacity_needed;   // directed pairs of the frame that overflowed (last CMD_ECAPACITY)
    int hit_cap;     // filter candidates per frame that fit the CTA's shared-memory list
    size_t smem_bytes;
    int threads;
    // cell-list path (boxes the one-CTA-per-frame kernel cannot hold)
    int path;        // 0 dense, 1 cell list
    int force_path;  // -1 automatic (cmd_topo_set_path)
    CellGrid cg;
    int rowcap, cell_batch;
    int4 *d_fxu, *d_sorted;
    int *d_slot, *d_cell_start, *d_rowcount, *d_rowoff_tmp, *d_tmp_j, *d_cap_need;
    double *d_tmp_d;
    double *d_cell_part;   // [batch][warps per frame] partial rate sums of k_cell_emit
    // block results
    int64_t cap_frames, nframes;
    int *d_start, *d_dest, *d_counts, *d_err;
    int *d_rowoff;   // [frames][n + 1] row index of every frame's list
    // AngleTopology (topology.py:124-167): angle colvar of the listed pairs
    int n_extra;
    int *d_group;      // [n] donor -> index of the extra atom it is bonded to
    double *d_theta;   // block array [frames * stride]
    double *d_extra_upload;
    size_t extra_upload_bytes;
    bool rate_is_fermi_angle;
    // HydroniumTopology (topology.py:234-253): per site the k nearest listed neighbours
    int near_k;
    int64_t near_stride;
    int *d_near_start, *d_near_dest, *d_near_counts;
    double *d_near_dist;
    double *d_dist, *d_omega, *d_rate_sum;
    uint8_t *d_rebuilt;
    unsigned long long *d_ties;
    unsigned *d_lists;     // skin lists of the dense kernel's persistent CTAs (pairs_dense.cuh): two
                           // sets, so that two launches on different streams never share one
    size_t lists_words;    // words per set
    int cap_l;             // entries per CTA
    int list_set;          // the set the next launch takes
    // Verlet state carried across blocks
    bool have_last;
    double *d_last, *d_displacement, *d_dr;
    int *d_carry_start, *d_carry_dest, *d_carry_count, *d_carry_rowoff;
    int *d_sched;  // [0] n_rebuild, [1] n_refresh, [2] last head (-1 = carry), then ids
    int *d_rebuild_ids, *d_refresh_ids, *d_head, *d_next;
    double *d_part;      // rate sums of a refresh split over several CTAs per frame
    size_t part_cap;
    double *d_upload;
    size_t upload_bytes;
    // donor selection on the device (cmd_topo_set_selection): host blocks hold n_total atoms per
    // frame, the donors are rows d_sel[0 .. n) of each
    int *d_sel;
    int n_total;
    // asynchronous blocks (cmd_topo_build_async): `done` fires when the block's last kernel has
    // finished; `pending` until cmd_topo_wait has checked the capacity
    cudaEvent_t done;
    bool done_valid, pending;
    const double *d_frames_last;  // frames of the last block (device)
    int64_t total_frames;
};

// ------------------------------------------------------------------ Verlet pieces -------------
// dr[f][i] = length(frame[f-1][i], frame[f][i])  (topology.py:98); f = 0 uses the carried frame
__global__ void __launch_bounds__(256) k_dr(const __grid_constant__ BoxParams bx,
                                            const double *__restrict__ frames,
                                            const double *__restrict__ last, int have_last, int n,
                                            int64_t nframes, double *__restrict__ dr)
{
    // grid = (atom blocks, frame lanes): no 64-bit division per element
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int64_t f = blockIdx.y; f < nframes; f += gridDim.y) {
        const double *cur = frames + (f * n + i) * 3;
        const double *prev = f > 0 ? cur - (int64_t)n * 3 : (have_last ? last + (int64_t)i * 3 : nullptr);
        double v = 0.0;  // topology.py:95: first frame -> dr = zeros
        if (prev) {
            double a[3] = {prev[0], prev[1], prev[2]}, b[3] = {cur[0], cur[1], cur[2]};
            v = length_exact(bx, a, b);
        }
        dr[f * n + i] = v;
    }
}


// ---- parallel rebuild schedule ---------------------------------------------------------------------
// The rebuild decision looks sequential (displacement accumulates frame after frame), but after a
// rebuild at frame r the displacement is exactly zero, so the NEXT rebuild frame is a function of r
// alone.  k_sched_walk evaluates that function for every possible r at once -- one small CTA per
// start frame walks forward, accumulating in the reference's order (bit-identical sums), until the
// two largest displacements cross the buffer -- and k_sched_chase then just follows the chain
// r -> next[r] from the block's start state.  Walks are capped (SCHED_CAP frames); a chain link
// that hit the cap is walked to its end by the chase kernel itself.
#define SCHED_THREADS 128
#define SCHED_CAP 384

// One walk: displacement (shared memory, n doubles, initialised by the caller) += dr[f] for
// f = first, first+1, ...; returns the first frame whose two largest displacements exceed the
// buffer, `nframes` if the block ends first, -1 if `cap` frames went by without a crossing.
__device__ int sched_walk(const double *__restrict__ dr, double *disp, double *red, int n,
                          int64_t nframes, double buffer, int64_t first, int64_t cap)
{
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    for (int64_t f = first; f < nframes; f++) {
        if (f - first >= cap) return -1;
        double m1 = -INFINITY, m2 = -INFINITY;
        for (int a = tid; a < n; a += blockDim.x) {
            const double v = __dadd_rn(disp[a], dr[f * n + a]);
            disp[a] = v;
            if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double o1 = __shfl_down_sync(0xffffffffu, m1, o), o2 = __shfl_down_sync(0xffffffffu, m2, o);
            if (o1 > m1) { m2 = fmax(m1, o2); m1 = o1; } else m2 = fmax(m2, o1);
        }
        if (lane == 0) { red[2 * w] = m1; red[2 * w + 1] = m2; }
        __syncthreads();
        m1 = red[0]; m2 = red[1];
        for (int q = 1; q < nw; q++) {
            const double o1 = red[2 * q], o2 = red[2 * q + 1];
            if (o1 > m1) { m2 = fmax(m1, o2); m1 = o1; } else m2 = fmax(m2, o1);
        }
        __syncthreads();
        // np.sort(displacement)[-2:] -> (m2, m1); displ_max1 + displ_max2 > buffer (topology.py:101-104)
        if (n >= 2 && __dadd_rn(m2, m1) > buffer) return (int)f;
    }
    return (int)nframes;
}

// walker w >= 1: the list was rebuilt at frame w - 1, displacement starts at zero, first frame w
__global__ void __launch_bounds__(SCHED_THREADS)
k_sched_walk(const double *__restrict__ dr, int n, int64_t nframes, double buffer,
             int *__restrict__ next)
{
    extern __shared__ double sched_smem[];
    double *disp = sched_smem, *red = sched_smem + n;
    const int64_t w = (int64_t)blockIdx.x + 1;
    for (int a = threadIdx.x; a < n; a += blockDim.x) disp[a] = 0.0;
    __syncthreads();
    const int r = sched_walk(dr, disp, red, n, nframes, buffer, w, SCHED_CAP);
    if (threadIdx.x == 0) next[w] = r;
}

// The same walkers, one WARP each (n <= 32 * PER): the displacements live in registers, the next
// frame's dr is requested before this frame's two largest values are reduced, no shared memory and
// no CTA barrier.  Same per-atom additions, and the two largest values do not depend on the order
// of the reduction: identical next[].
template <int PER>
__global__ void __launch_bounds__(128)
k_sched_walk_warp(const double *__restrict__ dr, int n, int64_t nframes, double buffer,
                  int *__restrict__ next)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5) + 1;
    if (w > nframes) return;
    double disp[PER], nx[PER];
#pragma unroll
    for (int q = 0; q < PER; q++) {
        disp[q] = 0.0;
        const int a = lane + 32 * q;
        nx[q] = (a < n && w < nframes) ? dr[w * n + a] : 0.0;
    }
    int r = (int)nframes;
    for (int64_t f = w; f < nframes; f++) {
        if (f - w >= SCHED_CAP) { r = -1; break; }
        double m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            if (lane + 32 * q < n) {
                const double v = __dadd_rn(disp[q], nx[q]);
                disp[q] = v;
                if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v;
            }
        }
        if (f + 1 < nframes) {
#pragma unroll
            for (int q = 0; q < PER; q++) {
                const int a = lane + 32 * q;
                if (a < n) nx[q] = dr[(f + 1) * n + a];
            }
        }
        for (int o = 16; o > 0; o >>= 1) {   // butterfly: every lane ends with the warp's two largest
            const double o1 = __shfl_xor_sync(0xffffffffu, m1, o), o2 = __shfl_xor_sync(0xffffffffu, m2, o);
            if (o1 > m1) { m2 = fmax(m1, o2); m1 = o1; } else m2 = fmax(m2, o1);
        }
        if (n >= 2 && __dadd_rn(m2, m1) > buffer) { r = (int)f; break; }
    }
    if (lane == 0) next[w] = r;
}

// follows the chain from the block's start state, marks the rebuild frames, leaves the displacement
// of the block's end in `displacement`
__global__ void __launch_bounds__(SCHED_THREADS)
k_sched_chase(const double *__restrict__ dr, double *__restrict__ displacement, int n,
              int64_t nframes, double buffer, int first_ever, const int *__restrict__ next,
              uint8_t *__restrict__ rebuilt, int next_smem)
{
    extern __shared__ double sched_smem[];
    double *disp = sched_smem, *red = sched_smem + n;
    // the chain is pointer chasing: stage next[] in shared memory when it fits (next_smem > 0)
    int *snext = (int *)(red + 2 * (SCHED_THREADS / 32));
    if (next_smem)
        for (int64_t q = threadIdx.x; q <= nframes; q += blockDim.x) snext[q] = q >= 1 ? next[q] : 0;
    __syncthreads();
    int64_t cur;      // walker index: 0 = carried displacement from frame 0, w = zero state from w
    if (first_ever) {   // the very first frame is always built (topology.py:91-93); dr[0] = 0
        if (threadIdx.x == 0) rebuilt[0] = 1;
        cur = 1;
    } else {
        cur = 0;
    }
    for (;;) {
        int nx;
        if (cur == 0) {
            for (int a = threadIdx.x; a < n; a += blockDim.x) disp[a] = displacement[a];
            __syncthreads();
            nx = sched_walk(dr, disp, red, n, nframes, buffer, 0, nframes + 1);
        } else if (cur >= nframes) {
            nx = (int)nframes;
        } else {
            nx = next_smem ? snext[cur] : next[cur];
            if (nx < 0) {   // the capped walk did not find the crossing: walk it here
                for (int a = threadIdx.x; a < n; a += blockDim.x) disp[a] = 0.0;
                __syncthreads();
                nx = sched_walk(dr, disp, red, n, nframes, buffer, cur, nframes + 1);
            }
        }
        if (nx >= nframes) break;
        if (threadIdx.x == 0) rebuilt[nx] = 1;
        cur = nx + 1;
    }
    // displacement at the end of the block: the last segment, walked without a crossing
    __syncthreads();
    if (cur == 0) {
        // `disp` already holds carried + all frames of the block (the walk above ran to the end)
    } else {
        for (int a = threadIdx.x; a < n; a += blockDim.x) disp[a] = 0.0;
        __syncthreads();
        if (cur < nframes) sched_walk(dr, disp, red, n, nframes, INFINITY, cur, nframes + 1);
    }
    __syncthreads();
    for (int a = threadIdx.x; a < n; a += blockDim.x) displacement[a] = disp[a];
}

// rebuilt[] flags -> compacted id lists, head frame of every refresh frame, counts
__global__ void __launch_bounds__(1024, 1)
k_sched_fill(const uint8_t *__restrict__ rebuilt, int64_t nframes, int *__restrict__ sched,
             int *__restrict__ rebuild_ids, int *__restrict__ refresh_ids, int *__restrict__ head)
{
    __shared__ int scan[40], lastreb[1024];
    const int tid = threadIdx.x;
    const int64_t per = (nframes + blockDim.x - 1) / blockDim.x;
    const int64_t f0 = tid * per, f1 = f0 + per < nframes ? f0 + per : nframes;
    int cnt = 0, last = -1;
    for (int64_t f = f0; f < f1; f++)
        if (rebuilt[f]) { cnt++; last = (int)f; }
    lastreb[tid] = last;
    const int ex = block_exclusive_scan(cnt, scan, &scan[33]);
    const int total = scan[33];
    // head of the frames in front of this thread's range: last rebuild of an earlier range, else
    // the list carried from the previous block (sched[2], -1)
    int hd = sched[2];
    for (int q = tid - 1; q >= 0; q--)
        if (lastreb[q] >= 0) { hd = lastreb[q]; break; }
    int nreb = ex;
    for (int64_t f = f0; f < f1; f++) {
        if (rebuilt[f]) {
            rebuild_ids[nreb++] = (int)f;
            hd = (int)f;
            head[f] = (int)f;
        } else {
            refresh_ids[f - nreb] = (int)f;   // refresh frames in front of f: f - #rebuilds so far
            head[f] = hd;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int last_all = sched[2];   // unchanged (the carried list) when the block has no rebuild
        for (int q = blockDim.x - 1; q >= 0; q--)
            if (lastreb[q] >= 0) { last_all = lastreb[q]; break; }
        sched[0] = total;
        sched[1] = (int)(nframes - total);
        sched[2] = last_all;
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

// k_schedule for systems whose displacement vector does not fit one CTA: a CLUSTER of 8 CTAs shares
// the atoms (displacements in shared memory, the next frame's dr in registers while this frame's
// two largest displacements are reduced); the CTAs exchange their (max1, max2) through distributed
// shared memory, one cluster barrier per frame.  Same decisions as k_schedule: the per-atom sums are
// the same sequential additions and the sum of the two largest values does not depend on the order
// of the reduction.
#define SCHED_CLUSTER 8
#define SCHED_CL_THREADS 1024
#define SCHED_CL_PER 8    // atoms per thread held in registers for the look-ahead

__device__ __forceinline__ void top2_merge(double &m1, double &m2, double o1, double o2)
{
    if (o1 > m1) { m2 = fmax(m1, o2); m1 = o1; } else m2 = fmax(m2, o1);
}

__global__ void __launch_bounds__(SCHED_CL_THREADS, 1)
k_schedule_cluster(const double *__restrict__ dr, double *__restrict__ displacement, int n,
                   int64_t nframes, double buffer, int first_ever, int *__restrict__ sched,
                   int *__restrict__ rebuild_ids, int *__restrict__ refresh_ids,
                   int *__restrict__ head, uint8_t *__restrict__ rebuilt, int chunk)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *disp = (double *)smem_raw;                       // [chunk] displacements of this CTA's atoms
    double2 *slots = (double2 *)(disp + chunk);              // [2][SCHED_CLUSTER] (max1, max2) per CTA
    double2 *wred = slots + 2 * SCHED_CLUSTER;               // [32] warp results
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int a0 = rank * chunk;
    const int per = (chunk + SCHED_CL_THREADS - 1) / SCHED_CL_THREADS;   // <= SCHED_CL_PER (host)
    for (int q = 0; q < per; q++) {
        const int i = tid + q * SCHED_CL_THREADS;
        if (i < chunk) disp[i] = a0 + i < n ? displacement[a0 + i] : 0.0;
    }
    double nx[SCHED_CL_PER];
#pragma unroll
    for (int q = 0; q < SCHED_CL_PER; q++) {
        const int i = tid + q * SCHED_CL_THREADS;
        nx[q] = (q < per && i < chunk && a0 + i < n) ? dr[a0 + i] : 0.0;
    }
    __syncthreads();
    cluster.sync();
    int n_rebuild = 0, n_refresh = 0, cur_head = sched[2];
    for (int64_t f = 0; f < nframes; f++) {
        double m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
        for (int q = 0; q < SCHED_CL_PER; q++) {
            const int i = tid + q * SCHED_CL_THREADS;
            if (q < per && i < chunk && a0 + i < n) {
                const double v = __dadd_rn(disp[i], nx[q]);
                disp[i] = v;
                if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v;
            }
        }
        if (f + 1 < nframes) {   // next frame's dr: in flight during the reduction
#pragma unroll
            for (int q = 0; q < SCHED_CL_PER; q++) {
                const int i = tid + q * SCHED_CL_THREADS;
                if (q < per && i < chunk && a0 + i < n) nx[q] = dr[(f + 1) * n + a0 + i];
            }
        }
        for (int o = 16; o > 0; o >>= 1)
            top2_merge(m1, m2, __shfl_down_sync(0xffffffffu, m1, o), __shfl_down_sync(0xffffffffu, m2, o));
        if (lane == 0) wred[w] = make_double2(m1, m2);
        __syncthreads();
        if (w == 0) {
            const double2 v = wred[lane];
            m1 = v.x; m2 = v.y;
            for (int o = 16; o > 0; o >>= 1)
                top2_merge(m1, m2, __shfl_down_sync(0xffffffffu, m1, o), __shfl_down_sync(0xffffffffu, m2, o));
            // every CTA of the cluster gets this CTA's pair
            if (lane < SCHED_CLUSTER) {
                const double b1 = __shfl_sync(0xffu, m1, 0), b2 = __shfl_sync(0xffu, m2, 0);
                double2 *remote = cluster.map_shared_rank(slots, lane);
                remote[(f & 1) * SCHED_CLUSTER + rank] = make_double2(b1, b2);
            }
        }
        cluster.sync();
        m1 = -INFINITY; m2 = -INFINITY;
#pragma unroll
        for (int c = 0; c < SCHED_CLUSTER; c++) {
            const double2 v = slots[(f & 1) * SCHED_CLUSTER + c];
            top2_merge(m1, m2, v.x, v.y);
        }
        // np.sort(displacement)[-2:] -> (m2, m1); displ_max1 + displ_max2 > buffer
        const bool cross = n >= 2 && __dadd_rn(m2, m1) > buffer;
        const bool first = first_ever && f == 0;
        if (cross) {
            for (int q = 0; q < per; q++) {
                const int i = tid + q * SCHED_CL_THREADS;
                if (i < chunk) disp[i] = 0.0;
            }
        }
        if (cross || first) {
            if (rank == 0 && tid == 0) { rebuild_ids[n_rebuild] = (int)f; head[f] = (int)f; rebuilt[f] = 1; }
            n_rebuild++;
            cur_head = (int)f;
        } else {
            if (rank == 0 && tid == 0) { refresh_ids[n_refresh] = (int)f; head[f] = cur_head; rebuilt[f] = 0; }
            n_refresh++;
        }
    }
    for (int q = 0; q < per; q++) {
        const int i = tid + q * SCHED_CL_THREADS;
        if (i < chunk && a0 + i < n) displacement[a0 + i] = disp[i];
    }
    if (rank == 0 && tid == 0) { sched[0] = n_rebuild; sched[1] = n_refresh; sched[2] = cur_head; }
    cluster.sync();   // no CTA leaves while its shared memory may still be written
}

#ifndef REFRESH_MINB
#define REFRESH_MINB 5   // 48 registers.  With two pairs per trip and the next trip's pairs prefetched
                         // (below): C2 Verlet pipeline 1.63 -> 1.50 ms per 16 384 frames (one pair per trip
                         // at six CTAs of 40 registers: 1.52 ms; eight CTAs of 32 registers spill: 1.57 ms)
#endif
// Refresh of a kept list (topology.py:110): dist = length(frame[row], frame[col]), same pairs.
// One CTA per refreshed frame; the head list is either a frame of this block or the carry.
template <bool STAGE>
__global__ void __launch_bounds__(256, REFRESH_MINB)
k_refresh(const __grid_constant__ BoxParams bx, const __grid_constant__ RateParams rp,
          const double *__restrict__ frames, const int *__restrict__ ids,
          const int *__restrict__ n_ids, const int *__restrict__ head, int n, int64_t stride,
          const int *__restrict__ carry_start, const int *__restrict__ carry_dest,
          const int *__restrict__ carry_count, int *__restrict__ out_start,
          int *__restrict__ out_dest, double *__restrict__ out_dist,
          double *__restrict__ out_omega, int *__restrict__ out_counts,
          double *__restrict__ out_rate_sum, const int *__restrict__ carry_rowoff,
          int *__restrict__ out_rowoff, double *__restrict__ part)
{
    // gridDim.y CTAs share the pairs of one frame (large systems); their rate sums meet in `part`
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if ((int)blockIdx.x >= *n_ids) return;
    const int64_t f = ids[blockIdx.x];
    const int hd = head[f];
    const double *fr = frames + f * (int64_t)n * 3;
    // [3n] AoS copy of the frame; frames too large for shared memory are read through L1/L2
    const double *sp = STAGE ? (const double *)smem_raw : fr;
    if (STAGE)
        for (int k = threadIdx.x; k < 3 * n; k += blockDim.x) ((double *)smem_raw)[k] = __ldg(fr + k);
    // out_counts of a head frame in this block is final: the build kernel ran before us
    const int p = hd < 0 ? *carry_count : out_counts[hd];
    const int *hs = hd < 0 ? carry_start : out_start + hd * stride;
    const int *hdst = hd < 0 ? carry_dest : out_dest + hd * stride;
    __syncthreads();
    if (blockIdx.y == 0) {   // the list is the head's list: so is its row index
        const int *hro = hd < 0 ? carry_rowoff : out_rowoff + hd * (int64_t)cmd_ro_pitch(n);
        int *ro = out_rowoff + f * (int64_t)cmd_ro_pitch(n);
        for (int k = threadIdx.x; k <= n; k += blockDim.x) ro[k] = hro[k];
    }
    const int64_t base = f * stride;
    double rsum = 0.0;
    const int share = (p + (int)gridDim.y - 1) / (int)gridDim.y;
    const int k_lo = min((int)blockIdx.y * share, p), k_hi = min(k_lo + share, p);
    // the pair of the NEXT trip is fetched before this trip's arithmetic: the head list may be a
    // frame of this very block (hs aliases out_start), so the compiler cannot hoist the loads over
    // the stores itself
    // two pairs per trip (independent FP64 chains), the pairs of the next trip fetched first
    const int T = blockDim.x;
    int k = k_lo + threadIdx.x;
    int a_nx[2], b_nx[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        a_nx[q] = k + q * T < k_hi ? hs[k + q * T] : 0;
        b_nx[q] = k + q * T < k_hi ? hdst[k + q * T] : 0;
    }
    for (; k < k_hi; k += 2 * T) {
        const int a[2] = {a_nx[0], a_nx[1]}, b[2] = {b_nx[0], b_nx[1]};
        const bool two = k + T < k_hi;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int kn = k + (2 + q) * T;
            if (kn < k_hi) { a_nx[q] = hs[kn]; b_nx[q] = hdst[kn]; }
        }
        double dist[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const double pa[3] = {sp[3 * a[q]], sp[3 * a[q] + 1], sp[3 * a[q] + 2]};
            const double pb[3] = {sp[3 * b[q]], sp[3 * b[q] + 1], sp[3 * b[q] + 2]};
            if (bx.kind == 0) dist[q] = length_exact(bx, pa, pb);
            else {   // zero image + the images kept for the refresh radius == the 27-image minimum here
                double d[3];
                if (bx.sparse == 2) diff_general_norm_sp<2>(bx, pa, pb, d);
                else if (bx.sparse == 1) diff_general_norm_sp<1>(bx, pa, pb, d);
                else diff_general_norm_exact(bx, pa, pb, d);
                dist[q] = sqrt(min_image_norm2_kept(bx, d));
            }
        }
        double om[2];
        rate_eval2(rp, dist[0], dist[1], om);
        rsum += om[0];
        out_start[base + k] = a[0]; out_dest[base + k] = b[0];
        out_dist[base + k] = dist[0]; out_omega[base + k] = om[0];
        if (two) {
            rsum += om[1];
            out_start[base + k + T] = a[1]; out_dest[base + k + T] = b[1];
            out_dist[base + k + T] = dist[1]; out_omega[base + k + T] = om[1];
        }
    }
    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_down_sync(0xffffffffu, rsum, o);
    __shared__ double wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = rsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += wsum[w];
        if (gridDim.y == 1) out_rate_sum[f] = t;
        else part[(int64_t)blockIdx.x * gridDim.y + blockIdx.y] = t;
        if (blockIdx.y == 0) out_counts[f] = p;
    }
}

// per-frame rate sums of a refresh that was split over several CTAs per frame, in CTA order
__global__ void __launch_bounds__(256)
k_refresh_sum(const int *__restrict__ ids, const int *__restrict__ n_ids, const double *__restrict__ part,
              int chunks, double *__restrict__ out_rate_sum)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_ids) return;
    double t = 0.0;
    for (int y = 0; y < chunks; y++) t += part[(int64_t)i * chunks + y];
    out_rate_sum[ids[i]] = t;
}

// keeps the list of the current segment head + the last frame for the next block
__global__ void __launch_bounds__(256)
k_carry(const int *__restrict__ sched, const int *__restrict__ out_start,
        const int *__restrict__ out_dest, const int *__restrict__ out_counts, int64_t stride,
        int *__restrict__ carry_start, int *__restrict__ carry_dest, int *__restrict__ carry_count,
        const int *__restrict__ out_rowoff, int *__restrict__ carry_rowoff, int n)
{
    int hd = sched[2];
    if (hd < 0) return;  // the head is still the carried list
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= n; k += gridDim.x * blockDim.x)
        carry_rowoff[k] = out_rowoff[hd * (int64_t)cmd_ro_pitch(n) + k];
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

// donor rows of whole frames, up-cast on the way: out[f][i][:] = in[f][sel[i]][:]
template <typename T>
__global__ void k_gather_cast(const T *__restrict__ in, const int *__restrict__ sel, int n_total, int n,
                              int64_t nframes, double *__restrict__ out)
{
    const int64_t per = (int64_t)n * 3, total = nframes * per;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i / per;
        const int r = (int)(i - f * per), a = r / 3, c = r - 3 * a;
        out[i] = (double)in[(f * n_total + sel[a]) * 3 + c];
    }
}

// the conversion behind a host copy: raw (float32 / float64, all atoms or the donors only) ->
// float64 donor frames
static int topo_convert(cmd_topo *t, const void *d_raw, int dtype_bytes, int64_t nframes, double *d_out,
                        cudaStream_t st)
{
    const int64_t elems = nframes * (int64_t)t->n * 3;
    int blocks = cmd_div_up(elems, 256);
    if (blocks > cmd_global().sm_count * 16) blocks = cmd_global().sm_count * 16;
    if (t->d_sel) {
        if (dtype_bytes == 4)
            k_gather_cast<float><<<blocks, 256, 0, st>>>((const float *)d_raw, t->d_sel, t->n_total, t->n, nframes, d_out);
        else
            k_gather_cast<double><<<blocks, 256, 0, st>>>((const double *)d_raw, t->d_sel, t->n_total, t->n, nframes, d_out);
    } else if (dtype_bytes == 4) {
        k_upcast_f32<<<blocks, 256, 0, st>>>((const float *)d_raw, d_out, elems);
    } else {
        return CMD_OK;   // float64 donors: copied in place
    }
    CMD_LAUNCHED();
    return CMD_OK;
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
    cudaFree(t->d_rowoff); cudaFree(t->d_theta);
    cudaFree(t->d_near_start); cudaFree(t->d_near_dest); cudaFree(t->d_near_counts); cudaFree(t->d_near_dist);
    t->d_rowoff = nullptr;
    t->d_theta = nullptr;
    t->d_near_start = t->d_near_dest = t->d_near_counts = nullptr;
    t->d_near_dist = nullptr;
    cudaFree(t->d_counts); cudaFree(t->d_rate_sum); cudaFree(t->d_rebuilt); cudaFree(t->d_dr);
    cudaFree(t->d_rebuild_ids); cudaFree(t->d_refresh_ids); cudaFree(t->d_head); cudaFree(t->d_next);
    cudaFree(t->d_part);
    t->d_part = nullptr;
    t->part_cap = 0;
    t->d_start = t->d_dest = t->d_counts = nullptr;
    t->d_dist = t->d_omega = t->d_rate_sum = t->d_dr = nullptr;
    t->d_rebuilt = nullptr;
    t->d_rebuild_ids = t->d_refresh_ids = t->d_head = t->d_next = nullptr;
    t->cap_frames = 0;
}

static void cell_free(cmd_topo *t)
{
    cudaFree(t->d_fxu); cudaFree(t->d_sorted); cudaFree(t->d_slot); cudaFree(t->d_cell_start);
    cudaFree(t->d_rowcount); cudaFree(t->d_rowoff_tmp); cudaFree(t->d_tmp_j); cudaFree(t->d_tmp_d);
    cudaFree(t->d_cell_part);
    t->d_cell_part = nullptr;
    t->d_fxu = t->d_sorted = nullptr;
    t->d_slot = t->d_cell_start = t->d_rowcount = t->d_rowoff_tmp = t->d_tmp_j = nullptr;
    t->d_tmp_d = nullptr;
    t->cell_batch = 0;
}

extern "C" void cmd_topo_destroy(cmd_topo *t)
{
    if (!t) return;
    cudaStreamSynchronize(cmd_global().stream);
    topo_free_block(t);
    cudaFree(t->d_err); cudaFree(t->d_ties); cudaFree(t->d_last); cudaFree(t->d_displacement);
    cudaFree(t->d_carry_start); cudaFree(t->d_carry_dest); cudaFree(t->d_carry_count);
    cudaFree(t->d_carry_rowoff);
    cudaFree(t->d_sched); cudaFree(t->d_upload); cudaFree(t->d_cap_need); cudaFree(t->d_lists);
    cudaFree(t->d_sel);
    if (t->done) cudaEventDestroy(t->done);
    cudaFree(t->d_group); cudaFree(t->d_extra_upload);
    cell_free(t);
    free(t);
}

// FP32 side of the dense filter: h = Q R (Gram-Schmidt on the cell vectors = columns of h);
// |h s| = |R s|, so the filter works in the rotated frame where the cell matrix is upper
// triangular (6 multiply-adds instead of 9).  R carries the 2^-32 of the fixed-point unit.
// Radius: rc (or the water-conversion window's right edge, beyond which convert_distance is the
// identity) widened by far more than the FP32 error of a wrapped difference (~3e-7 * sum|R|).
static void topo_filter_params(cmd_topo *t)
{
    FilterParams &fp = t->fp;
    const BoxParams &bx = t->bx;
    double c[3][3], q[3][3], R[3][3] = {{0}};
    for (int k = 0; k < 3; k++)
        for (int r = 0; r < 3; r++) c[k][r] = bx.h[3 * r + k];   // column k of h
    for (int k = 0; k < 3; k++) {
        double u[3] = {c[k][0], c[k][1], c[k][2]};
        for (int m = 0; m < k; m++) {
            R[m][k] = q[m][0] * c[k][0] + q[m][1] * c[k][1] + q[m][2] * c[k][2];
            for (int r = 0; r < 3; r++) u[r] -= R[m][k] * q[m][r];
        }
        R[k][k] = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        for (int r = 0; r < 3; r++) q[k][r] = u[r] / R[k][k];
    }
    const double unit = 1.0 / 4294967296.0;
    const double rr[6] = {R[0][0], R[0][1], R[0][2], R[1][1], R[1][2], R[2][2]};
    double rsum = 0;
    for (int k = 0; k < 6; k++) { fp.R[k] = (float)(rr[k] * unit); rsum += fabs(rr[k]); }
    double radius = t->rc;
    if (bx.conv != CMD_CONV_NONE && bx.conv_par[4] > radius) radius = bx.conv_par[4];
    const double thr = radius * (1.0 + 1e-5) + 1e-5 * rsum;
    fp.t2 = nextafterf((float)(thr * thr * (1.0 + 1e-6)), INFINITY);
    fp.n_img = bx.n_img;
    for (int m = 0; m < bx.n_img; m++)   // shift of image m in the R frame: Q^T T
        for (int k = 0; k < 3; k++)
            fp.img[m][k] = (float)(q[k][0] * bx.img[m][0] + q[k][1] * bx.img[m][1] +
                                   q[k][2] * bx.img[m][2]);

    // ---- packed-half filter (FILT_H2): 10-bit fractional coordinates, fp16 arithmetic ---------
    // Error of the computed vector against the true wrapped vector v = R s (s = wrapped fractional
    // difference), per component r with q_r = sum_c |R_rc| / 1024:
    //   quantisation   |w_c - 1024 s_c| <= 1 (two roundings of half a unit)          -> q_r
    //   fp16 matrix    |R16_rc - R_rc/1024| <= 2^-11 |R_rc/1024|, |w_c| <= 512       -> q_r / 4
    //   fp16 rounding  <= 3 results per component, each <= 2^-11 * 512 q_r (1+2^-11) -> 3 q_r / 4
    // so |dv| <= 2.1 sqrt(sum q_r^2) =: E, and the squared length carries three more roundings.
    // A pair within `radius` therefore has a computed d2 <= (radius + E)^2 (1 + 2^-11)^3 + tiny.
    // The integer wrap may pick the other periodic image when a fractional difference lies within
    // 2 units of +-1/2; such a pair is farther apart than radius + E as long as every cell height
    // times (1/2 - 2/1024) exceeds it, which is required below.
    {
        const double u10 = 1.0 / 1024.0;
        double qs = 0, hmin = 1e300;
        const int rows[3][3] = {{0, 1, 2}, {-1, 3, 4}, {-1, -1, 5}};
        for (int r = 0; r < 3; r++) {
            double qr = 0;
            for (int c = 0; c < 3; c++) if (rows[r][c] >= 0) qr += fabs(rr[rows[r][c]]) * u10;
            qs += qr * qr;
        }
        for (int c = 0; c < 3; c++) {
            const double height = 1.0 / sqrt(bx.hinv[3 * c] * bx.hinv[3 * c] + bx.hinv[3 * c + 1] * bx.hinv[3 * c + 1] +
                                             bx.hinv[3 * c + 2] * bx.hinv[3 * c + 2]);
            if (height < hmin) hmin = height;
        }
        const double E = 2.1 * sqrt(qs);
        const double lim = (radius + E) * (radius + E) * 1.0015 + 1e-6;
        for (int k = 0; k < 6; k++) {
            const __half hv = __float2half_rn((float)(rr[k] * u10));
            const unsigned short bits = *reinterpret_cast<const unsigned short *>(&hv);
            fp.hR[k] = bits * 0x00010001u;
        }
        __half ht = __float2half_ru((float)lim);
        const unsigned short tb = *reinterpret_cast<const unsigned short *>(&ht);
        fp.hT2 = tb * 0x00010001u;
        fp.h2_ok = bx.n_img == 0 && radius > 0 && E <= 0.05 * radius + 0.05 && lim < 60000.0 &&
                   hmin * (0.5 - 2.0 * u10) > radius + E && getenv("CMDLMC_B200_DENSE_F32") == nullptr;
        // spatial pruning: atoms binned (256 bins) along the fractional axis with the largest
        // height; a pair within `radius` differs by <= radius / height in that coordinate, i.e.
        // by <= radius / height * 1024 + 1 units of the 10-bit coordinates (two roundings) and
        // by at most ceil(that / 4) + 1 bins.  Forward windows of two rows cannot both hold the
        // other row while 2 db < 256; it pays while the window is well below half the atoms.
        // skin list (pairs_dense.cuh): the same filter at radius + skin, valid under the same
        // conditions -- including that no periodic image besides the wrapped vector matters out
        // to that radius -- plus the FP32 displacement test on 16-bit fractional coordinates
        // (error per displacement <= 2 units of every coordinate: two roundings of half a unit,
        // FP32 noise far below).
        {
            const char *sk = getenv("CMDLMC_B200_DENSE_SKIN");
            const double skin = sk ? atof(sk) : 0.6;
            const double rm = radius + skin;
            const double lim_m = (rm + E) * (rm + E) * 1.0015 + 1e-6;
            __half htm = __float2half_ru((float)lim_m);
            fp.hT2m = *reinterpret_cast<const unsigned short *>(&htm) * 0x00010001u;
            BoxParams wide = bx;
            cmd_box_prune_images(wide, rm * (1.0 + 1e-9) + 1e-9);
            double e16 = 0;
            for (int r = 0; r < 3; r++) {
                double qr = 0;
                for (int c = 0; c < 3; c++) if (rows[r][c] >= 0) qr += fabs(rr[rows[r][c]]) / 65536.0;
                e16 += qr * qr;
            }
            for (int k = 0; k < 6; k++) fp.R16[k] = (float)(rr[k] / 65536.0);
            fp.skin_eff = (float)(skin - 2.0 * (2.0 * sqrt(e16)) * 1.001 - 1e-4 * (1.0 + rsum * 1e-2));
            fp.coh_ok = fp.h2_ok && skin > 0 && wide.n_img == 0 && lim_m < 60000.0 && fp.skin_eff > 0 &&
                        hmin * (0.5 - 2.0 * u10) > rm + E;
        }
        int axis = 0;
        double hbest = 0;
        for (int c = 0; c < 3; c++) {
            const double height = 1.0 / sqrt(bx.hinv[3 * c] * bx.hinv[3 * c] + bx.hinv[3 * c + 1] * bx.hinv[3 * c + 1] +
                                             bx.hinv[3 * c + 2] * bx.hinv[3 * c + 2]);
            if (height > hbest) { hbest = height; axis = c; }
        }
        const int db = (int)ceil((radius / hbest * 1024.0 + 1.0) / 4.0) + 1;
        fp.sort_axis = -1;
        fp.sort_db = 0;
        fp.sort_db_m = 0;
        if (db <= 100 && t->n >= 96 && getenv("CMDLMC_B200_DENSE_NOSORT") == nullptr) {
            fp.sort_axis = axis;
            fp.sort_db = db;
            const char *sk = getenv("CMDLMC_B200_DENSE_SKIN");
            const double rm = radius + (sk ? atof(sk) : 0.6);
            fp.sort_db_m = (int)ceil((rm / hbest * 1024.0 + 1.0) / 4.0) + 1;
            if (fp.sort_db_m > 100) fp.coh_ok = 0;
        }
    }
    t->filt = fp.h2_ok ? FILT_H2 : bx.n_img == 0 ? FILT_F32 : FILT_F32_IMG;
    fp.wpre_ok = dense_wpre_fits(t->n, t->filt) ? 1 : 0;
}

static int dense_threads(int n)
{
    int th = ((n + 1) / 2 + 31) / 32 * 32;   // FP32 filters: two rows per thread
    return th < 32 ? 32 : th;
}

static int dense_threads_h2(int n)
{
    int th = (n + 31) / 32 * 32;             // packed-half filter: one row per thread
    if (th < 64) th = 64;
    // rows leave room below the next launch-bound shape: spend it on warps for the exact / emit
    // phases (their work is spread over the CTA, not over the rows)
    static const char *ex = getenv("CMDLMC_B200_DENSE_EXTRA_WARPS");
    const int extra = ex ? atoi(ex) : 1;
    const int lim = th <= 256 ? 256 : th <= 512 ? 512 : 1024;
    th += 32 * extra;
    // an even number of warps per CTA: the resident CTAs then load the four schedulers of an SM
    // evenly (C4, 384 rows: 13 warps 1.63 ms, 14 warps 1.53 ms per 16384 frames)
    if (!ex && ((th / 32) & 1)) th += 32;
    return th > lim ? lim : th;
}

// Cells per fractional axis: the cell must be at least one filter radius thick.
static void topo_cell_grid(cmd_topo *t)
{
    const BoxParams &bx = t->bx;
    double radius = t->rc;
    if (bx.conv != CMD_CONV_NONE && bx.conv_par[4] > radius) radius = bx.conv_par[4];
    radius *= 1.0 + 1e-6;
    CellGrid &cg = t->cg;
    for (int c = 0; c < 3; c++) {
        const double height = 1.0 / sqrt(bx.hinv[3 * c] * bx.hinv[3 * c] + bx.hinv[3 * c + 1] * bx.hinv[3 * c + 1] +
                                         bx.hinv[3 * c + 2] * bx.hinv[3 * c + 2]);
        double q = radius > 0 ? floor(height / radius) : 64.0;
        int nc = q > 64.0 ? 64 : (int)q;
        if (nc < 3) nc = 1;
        cg.nc[c] = nc;
    }
    // more cells than atoms buys nothing; keep the per-frame count table inside shared memory
    while ((int64_t)cg.nc[0] * cg.nc[1] * cg.nc[2] > 32768 ||
           ((int64_t)cg.nc[0] * cg.nc[1] * cg.nc[2] > 4 * (int64_t)t->n + 64 && cg.nc[0] * cg.nc[1] * cg.nc[2] > 27)) {
        int big = 0;
        for (int c = 1; c < 3; c++) if (cg.nc[c] > cg.nc[big]) big = c;
        if (cg.nc[big] <= 3) break;
        cg.nc[big]--;
    }
    for (int c = 0; c < 3; c++) cg.span[c] = cg.nc[c] >= 3 ? 3 : 1;
    cg.ncell = cg.nc[0] * cg.nc[1] * cg.nc[2];
}

// candidate capacity of the dense kernel's shared-memory lists for a per-frame capacity `stride`:
// stride / 2 unordered hits fit the HBM rows; the filter passes ~1.1x the hits, and the rows are
// sized 1.5x the first frame, so a list that holds 0.72 * stride / 2 candidates loses nothing in
// practice -- if two CTAs per SM fit with that, take them (the phases of the two overlap).
#define DENSE_SMEM_ONE (226 * 1024)
#define DENSE_SMEM_TWO 115200
static int dense_candidate_cap(int n, int64_t stride, int filt)
{
    int64_t cap = stride / 2;
    const size_t fixed = dense_smem_fixed(n, filt);
    if (fixed < DENSE_SMEM_TWO) {
        int64_t cap2 = (int64_t)((DENSE_SMEM_TWO - fixed) / DENSE_BYTES_PER_CAND) / 4 * 4;
        if (cap2 >= cap) return (int)(cap > 32764 ? 32764 : cap);
        if (cap2 * 100 >= cap * 72) return (int)(cap2 > 32764 ? 32764 : cap2);
    }
    if (fixed < DENSE_SMEM_ONE) {
        int64_t cap1 = (int64_t)((DENSE_SMEM_ONE - fixed) / DENSE_BYTES_PER_CAND) / 4 * 4;
        if (cap1 < cap) cap = cap1;
    }
    return (int)(cap > 32764 ? 32764 : cap);   // 15-bit candidate index in the slot map
}

static int topo_configure(cmd_topo *t, int64_t stride)
{
    // per-frame capacity and the matching shared-memory candidate list
    stride = (stride + 63) / 64 * 64;
    int hit_cap = dense_candidate_cap(t->n, stride, t->filt);
    size_t smem = dense_smem_bytes(t->n, hit_cap, t->filt);
    t->stride = stride;
    t->threads = dense_threads(t->n);
    // the list must hold at least 0.72 * stride / 2 candidates, else the cell list takes over
    const bool fits = t->n <= 1024 && smem <= DENSE_SMEM_ONE && (int64_t)hit_cap * 100 >= stride / 2 * 72;
    if (!fits && t->force_path == 0)
        return cmd_set_error(CMD_ECAPACITY, "dense pair kernel needs %zu bytes of shared memory for "
                             "n=%d, capacity %lld (limit 231424)",
                             dense_smem_bytes(t->n, (int)(stride / 2), t->filt), t->n, (long long)stride);
    if (!fits || t->force_path == 1) {   // too large for one CTA per frame: cell list
        t->path = 1;
        t->hit_cap = 0;
        t->smem_bytes = 0;
        if (t->rowcap < 8) t->rowcap = 40;
        return CMD_OK;
    }
    t->path = 0;
    t->hit_cap = hit_cap;
    t->smem_bytes = smem;
    return CMD_OK;
}

extern "C" int cmd_topo_create(const cmd_box *box, int n, double cutoff, double buffer, int mode,
                               int rate_kind, const double par[CMD_RATE_NPAR], int64_t capacity,
                               cmd_topo **out)
{
    CMD_REQUIRE_INIT();
    if (!box || !out || n < 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (mode != CMD_TOPO_BRUTEFORCE && mode != CMD_TOPO_VERLET)
        return cmd_set_error(CMD_EINVAL, "bad topology mode %d", mode);
    if (rate_kind < 0 || rate_kind > CMD_RATE_EXP)
        return cmd_set_error(CMD_EINVAL, "bad rate kind %d", rate_kind);
    if (!(cutoff + buffer >= 0)) return cmd_set_error(CMD_EINVAL, "cutoff + buffer must be >= 0");
    cmd_topo *t = (cmd_topo *)calloc(1, sizeof(cmd_topo));
    if (!t) return cmd_set_error(CMD_ENOMEM, "out of host memory");
    t->bx = box->p;
    // structural zeros of the cell matrix shorten the exact stage (CMDLMC_B200_NO_SPARSE=1: the
    // full products, for A/B tests of the bit-identity)
    t->bx.sparse = getenv("CMDLMC_B200_NO_SPARSE") ? 0 : cmd_box_sparsity(t->bx);
    // FermiAngle = Fermi masked by the angle colvar: the list kernels evaluate the Fermi part, the
    // mask is applied by cmd_topo_apply_angles once the angles of the block are known
    t->rate_is_fermi_angle = rate_kind == CMD_RATE_FERMI_ANGLE;
    t->rate.kind = t->rate_is_fermi_angle ? CMD_RATE_FERMI : rate_kind;
    if (par) memcpy(t->rate.par, par, sizeof(t->rate.par));
    t->n = n;
    t->cutoff = cutoff;
    t->buffer = buffer;
    t->rc = cutoff + buffer;  // topology.py:67 adds them in double exactly like this
    t->t2 = exact_sq_threshold(t->rc);
    // the filter only has to look at the images that can come within rc of the origin
    cmd_box_prune_images(t->bx, t->rc);
    // Between rebuilds the two largest path lengths sum to <= buffer, so a listed pair (<= rc at
    // its rebuild) is never farther apart than rc + buffer when it is refreshed: the images that
    // cannot come within that radius cannot hold the reference's 27-image minimum either.
    t->bx_refresh = box->p;
    t->bx_refresh.sparse = t->bx.sparse;
    cmd_box_prune_images(t->bx_refresh, (t->rc + buffer) * (1.0 + 1e-9) + 1e-9);
    topo_filter_params(t);
    topo_cell_grid(t);
    t->mode = mode;
    t->force_path = -1;
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
    TALLOC(t->d_ties, 4 * sizeof(unsigned long long));   // [0] ties, [1..3] skin-list statistics
    TALLOC(t->d_last, (size_t)n * 24);
    TALLOC(t->d_displacement, (size_t)n * 8);
    TALLOC(t->d_carry_count, sizeof(int));
    TALLOC(t->d_sched, 4 * sizeof(int));
    TALLOC(t->d_cap_need, sizeof(int));
    cudaStream_t st = cmd_global().stream;
    CMD_CUDA(cudaMemsetAsync(t->d_err, 0, sizeof(int), st));
    CMD_CUDA(cudaMemsetAsync(t->d_ties, 0, 4 * sizeof(unsigned long long), st));
    CMD_CUDA(cudaMemsetAsync(t->d_displacement, 0, (size_t)n * 8, st));
    CMD_CUDA(cudaMemsetAsync(t->d_carry_count, 0, sizeof(int), st));
    CMD_CUDA(cudaMemsetAsync(t->d_cap_need, 0, sizeof(int), st));
    int sched0[4] = {0, 0, -1, 0};
    CMD_CUDA(cudaMemcpyAsync(t->d_sched, sched0, sizeof(sched0), cudaMemcpyHostToDevice, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    *out = t;
    return CMD_OK;
}

// Skin lists: one per persistent CTA, 3x the exact stage's candidate capacity (the list is taken at
// radius + skin: (1 + skin / radius)^3 times the candidates; a list that does not fit only costs the
// frame the direct path).
static int dense_lists_reserve(cmd_topo *t, int64_t ctas, int hit_cap)
{
    // (CMDLMC_B200_DENSE_LIST_CAP: entries per list, for tests of the overflow path)
    const char *lc = getenv("CMDLMC_B200_DENSE_LIST_CAP");
    const int cap_l = lc ? (atoi(lc) > 64 ? atoi(lc) : 64) : 3 * hit_cap;
    const size_t words = (size_t)ctas * cap_l;
    if (t->d_lists && t->lists_words >= words && t->cap_l == cap_l) return CMD_OK;
    CMD_CUDA(cudaStreamSynchronize(cmd_global().stream));
    cudaFree(t->d_lists);
    t->d_lists = nullptr;
    t->lists_words = 0;
    if (cudaMalloc((void **)&t->d_lists, 2 * words * 4) != cudaSuccess) {
        cudaGetLastError();
        return cmd_set_error(CMD_ENOMEM, "cudaMalloc of %zu skin-list bytes failed", words * 4);
    }
    t->lists_words = words;
    t->cap_l = cap_l;
    return CMD_OK;
}

static int launch_dense(cmd_topo *t, const double *d_frames, const int *ids, const int *n_ids,
                        int64_t grid, int *start, int *dest, double *dist, double *omega,
                        int *counts, double *rate_sum, uint8_t *rebuilt, int *rowoff,
                        int64_t stride, int hit_cap, size_t smem)
{
    cudaStream_t st = cmd_global().stream;
    const bool ortho = t->bx.kind == 0;
    // the skin list needs consecutive frames of one trajectory: contiguous items, at least a few
    // frames per CTA (CMDLMC_B200_DENSE_SKIN=0 switches it off)
    const bool use_skin = t->fp.coh_ok && ids == nullptr && grid >= 4 * (int64_t)cmd_global().sm_count;
    // persistent CTAs: as many as are resident at once, each walking its share of the frames
#define DENSE_LAUNCH(K, IM, SP, MT, MB)                                                          \
    do {                                                                                         \
        CMD_CUDA(cudaFuncSetAttribute(k_pairs_dense<K, IM, SP, MT, MB>,                          \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        int nb_ = 1;                                                                             \
        CMD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                  \
            &nb_, k_pairs_dense<K, IM, SP, MT, MB>, t->threads * SP, smem));                     \
        int64_t pgrid = (int64_t)cmd_global().sm_count * (nb_ > 0 ? nb_ : 1);                    \
        if (pgrid > grid) pgrid = grid;                                                          \
        unsigned *lists_ = nullptr;                                                              \
        if (IM == FILT_H2 && use_skin) {                                                         \
            int rc_ = dense_lists_reserve(t, pgrid, hit_cap);                                    \
            if (rc_) return rc_;                                                                 \
            lists_ = t->d_lists + (size_t)(t->list_set & 1) * t->lists_words;                    \
        }                                                                                        \
        k_pairs_dense<K, IM, SP, MT, MB><<<(unsigned)pgrid, t->threads * SP, smem, st>>>(        \
            t->bx, t->rate, t->fp, d_frames, ids, n_ids, (int)grid, t->n, t->rc, t->t2, stride,  \
            hit_cap,                                                                             \
            start, dest, dist, omega, counts, rate_sum, rebuilt, rowoff, t->d_err, t->d_ties,    \
            lists_, t->cap_l, dense_layout(t->n, hit_cap, IM));                                  \
    } while (0)
#define DENSE_PICK(SP, MT, MB)                                                                   \
    do {                                                                                         \
        if (ortho) DENSE_LAUNCH(0, FILT_F32, SP, MT, MB);                                        \
        else if (t->fp.n_img == 0) DENSE_LAUNCH(1, FILT_F32, SP, MT, MB);                        \
        else DENSE_LAUNCH(1, FILT_F32_IMG, SP, MT, MB);                                          \
    } while (0)
    if (t->filt == FILT_H2) {
        // one row per thread; register budget 64 in every shape
        const int th = dense_threads_h2(t->n);
        const int keep = t->threads;
        t->threads = th;
        if (th <= 256) { if (ortho) DENSE_LAUNCH(0, FILT_H2, 1, 256, 4); else DENSE_LAUNCH(1, FILT_H2, 1, 256, 4); }
        else if (th <= 512) { if (ortho) DENSE_LAUNCH(0, FILT_H2, 1, 512, 2); else DENSE_LAUNCH(1, FILT_H2, 1, 512, 2); }
        else { if (ortho) DENSE_LAUNCH(0, FILT_H2, 1, 1024, 1); else DENSE_LAUNCH(1, FILT_H2, 1, 1024, 1); }
        t->threads = keep;
    }
    // t->threads = one thread per two rows; small frames run two copies of the row set
    else if (t->threads <= 128) DENSE_PICK(2, 256, 2);
    else if (t->threads <= 256) DENSE_PICK(DENSE_SPLIT_MID, 256 * DENSE_SPLIT_MID, 2);
    else DENSE_PICK(1, 512, 1);
#undef DENSE_PICK
#undef DENSE_LAUNCH
    CMD_LAUNCHED();
    return CMD_OK;
}


// ---- cell-list path ----------------------------------------------------------------------------
static int cell_reserve(cmd_topo *t, int batch)
{
    if (batch <= t->cell_batch) return CMD_OK;
    CMD_CUDA(cudaStreamSynchronize(cmd_global().stream));
    cell_free(t);
    const size_t n = (size_t)t->n, B = (size_t)batch;
#define CALLOC(ptr, bytes)                                                                       \
    if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) {                                   \
        cudaGetLastError();                                                                      \
        cell_free(t);                                                                            \
        return cmd_set_error(CMD_ENOMEM, "cudaMalloc of %zu bytes failed for %s", (size_t)(bytes), #ptr); \
    }
    CALLOC(t->d_fxu, B * n * 16);
    CALLOC(t->d_sorted, B * n * 16);
    CALLOC(t->d_slot, B * n * 4);
    CALLOC(t->d_cell_start, B * (t->cg.ncell + 1) * 4);
    CALLOC(t->d_rowcount, B * n * 4);
    CALLOC(t->d_rowoff_tmp, B * (n + 1) * 4);
    CALLOC(t->d_tmp_j, B * n * t->rowcap * 4);
    CALLOC(t->d_tmp_d, B * n * t->rowcap * 8);
    CALLOC(t->d_cell_part, B * (size_t)cmd_div_up(t->n, 256) * 8 * 8);
#undef CALLOC
    t->cell_batch = batch;
    return CMD_OK;
}

static int cell_batch_size(const cmd_topo *t, int64_t want)
{
    const size_t per = (size_t)t->n * (42 + 12 * (size_t)t->rowcap) + (size_t)(t->cg.ncell + 1) * 4 + 8;
    int64_t b = (int64_t)((size_t)1 << 30) / (int64_t)per;
    if (b < 1) b = 1;
    if (b > 8192) b = 8192;
    return (int)(b < want ? b : want);
}

// Builds the lists of `count` frames (ids[first..] or frames first..first+count) through the cell
// list.  emit = false stops after the per-frame totals (capacity probe).
static int launch_cell(cmd_topo *t, const double *d_frames, const int *ids, const int *n_ids,
                       int64_t count, int *start, int *dest, double *dist, double *omega,
                       int *counts, double *rate_sum, uint8_t *rebuilt, int *rowoff, int64_t stride,
                       bool emit)
{
    cudaStream_t st = cmd_global().stream;
    const bool ortho = t->bx.kind == 0;
    const int n = t->n;
    for (int64_t first = 0; first < count;) {
        int batch = cell_batch_size(t, count - first);
        int rc = cell_reserve(t, batch);
        if (rc) return rc;
        batch = t->cell_batch < count - first ? t->cell_batch : (int)(count - first);
        const size_t bsm = (size_t)(t->cg.ncell + 1 + 40) * 4;
        if (ortho) {
            CMD_CUDA(cudaFuncSetAttribute(k_cell_build<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
            k_cell_build<0><<<batch, 1024, bsm, st>>>(t->bx, t->cg, d_frames, ids, n_ids, (int)first, n,
                                                     t->d_fxu, t->d_slot, t->d_sorted, t->d_cell_start);
        } else {
            CMD_CUDA(cudaFuncSetAttribute(k_cell_build<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
            k_cell_build<1><<<batch, 1024, bsm, st>>>(t->bx, t->cg, d_frames, ids, n_ids, (int)first, n,
                                                     t->d_fxu, t->d_slot, t->d_sorted, t->d_cell_start);
        }
        CMD_LAUNCHED();
        CMD_CUDA(cudaMemsetAsync(t->d_rowcount, 0, (size_t)batch * n * 4, st));
        // whole (x, y) columns per CTA when the batch fills the GPU with them, z segments otherwise
        const int ncolumn = t->cg.nc[0] * t->cg.nc[1];
        int zseg = t->cg.nc[2];
        {
            const int64_t want = (int64_t)cmd_global().sm_count * 14;
            if ((int64_t)ncolumn * batch < want) {
                zseg = (int)(((int64_t)ncolumn * batch * t->cg.nc[2] + want - 1) / want);
                if (zseg < 1) zseg = 1;
                if (zseg > t->cg.nc[2]) zseg = t->cg.nc[2];
            }
        }
        dim3 pgrid((unsigned)(ncolumn * cmd_div_up(t->cg.nc[2], zseg)), (unsigned)batch);
#define CELL_PAIRS(K, IM)                                                                        \
    k_cell_pairs<K, IM><<<pgrid, CELL_TPB, 0, st>>>(                                             \
        t->bx, t->fp, t->cg, d_frames, ids, n_ids, (int)first, n, t->rc, t->t2, t->rowcap, zseg, \
        t->d_sorted, t->d_cell_start, t->d_rowcount, t->d_tmp_j, t->d_tmp_d, t->d_cap_need,      \
        t->d_ties)
        if (ortho) CELL_PAIRS(0, false);
        else if (t->fp.n_img == 0) CELL_PAIRS(1, false);
        else CELL_PAIRS(1, true);
#undef CELL_PAIRS
        CMD_LAUNCHED();
        // scratch rows too short?  (one 4-byte read-back per batch)
        int need = 0;
        CMD_CUDA(cudaMemcpyAsync(&need, t->d_cap_need, sizeof(int), cudaMemcpyDeviceToHost, st));
        CMD_CUDA(cudaStreamSynchronize(st));
        if (need > t->rowcap) {
            CMD_CUDA(cudaMemsetAsync(t->d_cap_need, 0, sizeof(int), st));
            if ((size_t)(need + 8) * (size_t)n * 12 > ((size_t)1 << 32))
                return cmd_set_error(CMD_ECAPACITY, "an atom has %d neighbours: more than the "
                                     "cell-list scratch rows can hold", need);
            t->rowcap = (need + 8 + 3) / 4 * 4;   // k_cell_emit reads the rows 16 bytes at a time
            cell_free(t);
            continue;   // redo this batch with longer rows
        }
        k_cell_scan<<<batch, 1024, 0, st>>>(ids, n_ids, (int)first, n, stride, t->d_rowcount,
                                            t->d_rowoff_tmp, counts, rebuilt, rate_sum,
                                            emit ? rowoff : nullptr, t->d_err);
        CMD_LAUNCHED();
        if (emit) {
            dim3 egrid((unsigned)((n + 255) / 256), (unsigned)batch);
            k_cell_emit<<<egrid, 256, 0, st>>>(t->rate, ids, n_ids, (int)first, n, stride, t->rowcap,
                                               t->d_rowoff_tmp, t->d_tmp_j, t->d_tmp_d, start, dest, dist,
                                               omega, rate_sum ? t->d_cell_part : nullptr);
            CMD_LAUNCHED();
            if (rate_sum) {
                k_cell_rsum<<<batch, 256, 0, st>>>(ids, n_ids, (int)first, (int)egrid.x * 8, t->d_cell_part,
                                                   rate_sum);
                CMD_LAUNCHED();
            }
        }
        first += batch;
    }
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
    t->threads = dense_threads(t->n);
    size_t smem = dense_smem_bytes(t->n, 0, t->filt);
    CMD_CUDA(cudaMemsetAsync(t->d_err, 0, sizeof(int), st));
    if (t->n <= 1024 && smem <= DENSE_SMEM_ONE && t->force_path != 1) {
        rc = launch_dense(t, d_frame, nullptr, nullptr, 1, nullptr, nullptr, nullptr, nullptr, d_cnt,
                          nullptr, nullptr, nullptr, 0, 0, smem);
    } else {
        if (t->rowcap < 8) t->rowcap = 40;
        rc = launch_cell(t, d_frame, nullptr, nullptr, 1, nullptr, nullptr, nullptr, nullptr, d_cnt,
                         nullptr, nullptr, nullptr, (int64_t)1 << 40, false);
    }
    if (rc) return rc;
    int cnt = 0;
    CMD_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    CMD_CUDA(cudaMemsetAsync(t->d_err, 0, sizeof(int), st));
    CMD_CUDA(cudaMemsetAsync(t->d_ties, 0, 4 * sizeof(unsigned long long), st));
    int64_t p0 = cnt < 0 ? -cnt : cnt;
    int64_t want = p0 + p0 / 2 + 128;
    // prefer the dense kernel: shrink the head-room to what its shared-memory lists can hold
    while (t->n <= 1024 && t->force_path != 1 && want > p0 + p0 / 4 + 64 &&
           (int64_t)dense_candidate_cap(t->n, (want + 63) / 64 * 64, t->filt) * 100 < (want + 63) / 64 * 64 / 2 * 72)
        want -= 64;
    return topo_configure(t, want);
}

// all-pairs lists of `grid` frames (or of the ids[0 .. *n_ids) frames) by the configured path
// all-pairs lists of `grid` frames (or of the ids[0 .. *n_ids) frames) by the configured path, into
// the block arrays.  Frame indices are relative to d_frames; results land at frame f0 + index.
static int launch_pairs(cmd_topo *t, const double *d_frames, const int *ids, const int *n_ids,
                        int64_t grid, int64_t f0, bool with_rebuilt)
{
    int *start = t->d_start + f0 * t->stride, *dest = t->d_dest + f0 * t->stride;
    double *dist = t->d_dist + f0 * t->stride, *omega = t->d_omega + f0 * t->stride;
    int *counts = t->d_counts + f0;
    double *rate_sum = t->d_rate_sum + f0;
    uint8_t *rebuilt = with_rebuilt ? t->d_rebuilt + f0 : nullptr;
    int *rowoff = t->d_rowoff + f0 * cmd_ro_pitch(t->n);
    if (t->path == 0)
        return launch_dense(t, d_frames, ids, n_ids, grid, start, dest, dist, omega, counts,
                            rate_sum, rebuilt, rowoff, t->stride, t->hit_cap, t->smem_bytes);
    int64_t count = grid;
    if (n_ids) {   // Verlet: the number of rebuild frames lives on the device
        int h = 0;
        cudaStream_t st = cmd_global().stream;
        CMD_CUDA(cudaMemcpyAsync(&h, n_ids, sizeof(int), cudaMemcpyDeviceToHost, st));
        CMD_CUDA(cudaStreamSynchronize(st));
        count = h;
        if (count <= 0) return CMD_OK;
    }
    return launch_cell(t, d_frames, ids, n_ids, count, start, dest, dist, omega, counts, rate_sum,
                       rebuilt, rowoff, t->stride, true);
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
    BALLOC(t->d_rowoff, (size_t)nframes * cmd_ro_pitch(t->n) * 4);
    BALLOC(t->d_rate_sum, (size_t)nframes * 8);
    BALLOC(t->d_rebuilt, (size_t)nframes);
    if (t->mode == CMD_TOPO_VERLET) {
        BALLOC(t->d_dr, (size_t)nframes * t->n * 8);
        BALLOC(t->d_rebuild_ids, (size_t)nframes * 4);
        BALLOC(t->d_refresh_ids, (size_t)nframes * 4);
        BALLOC(t->d_head, (size_t)nframes * 4);
        BALLOC(t->d_next, ((size_t)nframes + 2) * 4);
    }
    t->cap_frames = nframes;
    return CMD_OK;
}

static int topo_build_impl(cmd_topo *t, const double *d_frames, int64_t nframes, bool skip);

// capacity check (one 4-byte read-back per block)
static int topo_check_capacity(cmd_topo *t)
{
    cudaStream_t st = cmd_global().stream;
    int err = 0;
    CMD_CUDA(cudaMemcpyAsync(&err, t->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    if (err > 0) {
        CMD_CUDA(cudaMemsetAsync(t->d_err, 0, sizeof(int), st));
        t->capacity_needed = err;
        return cmd_set_error(CMD_ECAPACITY, "a frame has %d directed pairs but the per-frame "
                             "capacity is %lld: re-create the topology with a larger "
                             "capacity_per_frame", err, (long long)t->stride);
    }
    return CMD_OK;
}

extern "C" int cmd_topo_build_dev(cmd_topo *t, const double *d_frames, int64_t nframes)
{
    return topo_build_impl(t, d_frames, nframes, false);
}

// Frame-block sharding of a Verlet run: a rank whose block starts at frame s walks frames [0, s)
// through the (cheap, sequential) displacement / rebuild-decision pass only and builds just the
// list of the last rebuild frame, which leaves exactly the state a sequential run has at s.
extern "C" int cmd_topo_skip_dev(cmd_topo *t, const double *d_frames, int64_t nframes)
{
    if (t && t->mode != CMD_TOPO_VERLET) {   // brute force keeps no state between frames
        t->total_frames += nframes;
        return CMD_OK;
    }
    return topo_build_impl(t, d_frames, nframes, true);
}

// Rebuild schedule of `nframes` frames from their step lengths dr[nframes][n] (topology.py:96-107),
// continuing from the carried displacement: fills d_rebuilt / d_rebuild_ids / d_refresh_ids / d_head
// / d_sched and leaves the displacement after the last frame in d_displacement.
static int topo_schedule(cmd_topo *t, const double *dr, int64_t nframes)
{
    CmdGlobal &g = cmd_global();
    cudaStream_t st = g.stream;
    (void)g;
    const size_t ssm = ((size_t)t->n + 2 * (SCHED_THREADS / 32)) * 8;
    if (ssm <= 48 * 1024 && nframes >= 64) {
        // parallel schedule: next-rebuild function for every start frame, then chain following
        CMD_CUDA(cudaMemsetAsync(t->d_rebuilt, 0, (size_t)nframes, st));
        const unsigned wblocks = (unsigned)((nframes + 3) / 4);
        if (t->n <= 128) k_sched_walk_warp<4><<<wblocks, 128, 0, st>>>(dr, t->n, nframes, t->buffer, t->d_next);
        else if (t->n <= 256) k_sched_walk_warp<8><<<wblocks, 128, 0, st>>>(dr, t->n, nframes, t->buffer, t->d_next);
        else if (t->n <= 384) k_sched_walk_warp<12><<<wblocks, 128, 0, st>>>(dr, t->n, nframes, t->buffer, t->d_next);
        else if (t->n <= 512) k_sched_walk_warp<16><<<wblocks, 128, 0, st>>>(dr, t->n, nframes, t->buffer, t->d_next);
        else
        k_sched_walk<<<(unsigned)nframes, SCHED_THREADS, ssm, st>>>(dr, t->n, nframes, t->buffer,
                                                                    t->d_next);
        CMD_LAUNCHED();
        const size_t nsm = ((size_t)nframes + 2) * 4;
        const int next_smem = ssm + nsm <= 200 * 1024 ? 1 : 0;
        const size_t csm = ssm + (next_smem ? nsm : 0);
        CMD_CUDA(cudaFuncSetAttribute(k_sched_chase, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
        k_sched_chase<<<1, SCHED_THREADS, csm, st>>>(dr, t->d_displacement, t->n, nframes,
                                                    t->buffer, t->total_frames == 0 ? 1 : 0,
                                                    t->d_next, t->d_rebuilt, next_smem);
        CMD_LAUNCHED();
        k_sched_fill<<<1, 1024, 0, st>>>(t->d_rebuilt, nframes, t->d_sched, t->d_rebuild_ids,
                                         t->d_refresh_ids, t->d_head);
        CMD_LAUNCHED();
    } else if (t->n > 4096 && t->n <= SCHED_CLUSTER * SCHED_CL_THREADS * SCHED_CL_PER) {
        // large systems: one cluster of 8 CTAs walks the frames
        int chunk = (t->n + SCHED_CLUSTER - 1) / SCHED_CLUSTER;
        chunk = (chunk + 31) / 32 * 32;
        const size_t csm = (size_t)chunk * 8 + (2 * SCHED_CLUSTER + 32) * sizeof(double2);
        CMD_CUDA(cudaFuncSetAttribute(k_schedule_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(SCHED_CLUSTER);
        cfg.blockDim = dim3(SCHED_CL_THREADS);
        cfg.dynamicSmemBytes = csm;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = SCHED_CLUSTER;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CMD_CUDA(cudaLaunchKernelEx(&cfg, k_schedule_cluster, (const double *)dr, t->d_displacement,
                                    t->n, (int64_t)nframes, t->buffer, t->total_frames == 0 ? 1 : 0,
                                    t->d_sched, t->d_rebuild_ids, t->d_refresh_ids, t->d_head,
                                    t->d_rebuilt, chunk));
        CMD_LAUNCHED();
    } else {
        int sth = (t->n + 31) / 32 * 32;
        if (sth > 1024) sth = 1024;
        k_schedule<<<1, sth, 0, st>>>(dr, t->d_displacement, t->n, nframes, t->buffer,
                                      t->total_frames == 0 ? 1 : 0, t->d_sched, t->d_rebuild_ids,
                                      t->d_refresh_ids, t->d_head, t->d_rebuilt);
        CMD_LAUNCHED();
    }
    return CMD_OK;
}

static int topo_build_impl(cmd_topo *t, const double *d_frames, int64_t nframes, bool skip)
{
    CMD_REQUIRE_INIT();
    if (!t || !d_frames || nframes < 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (nframes > 0x7fffffff / 2) return cmd_set_error(CMD_EINVAL, "block too large");
    CmdGlobal &g = cmd_global();
    cudaStream_t st = g.stream;
    int rc;
    t->done_valid = false;   // this block completes in stream order, not behind the last pipelined one
    if (t->stride == 0 && (rc = topo_autosize(t, d_frames))) return rc;
    if ((rc = topo_reserve(t, nframes))) return rc;
    if (t->mode == CMD_TOPO_VERLET && !t->d_carry_start) {
        if (cudaMalloc((void **)&t->d_carry_start, t->stride * 4) != cudaSuccess ||
            cudaMalloc((void **)&t->d_carry_dest, t->stride * 4) != cudaSuccess ||
            cudaMalloc((void **)&t->d_carry_rowoff, (size_t)(t->n + 1) * 4) != cudaSuccess) {
            cudaGetLastError();
            return cmd_set_error(CMD_ENOMEM, "cudaMalloc failed for the carried pair list");
        }
    }
    t->nframes = skip ? 0 : nframes;
    t->d_frames_last = d_frames;
    if (t->mode == CMD_TOPO_BRUTEFORCE) {
        rc = launch_pairs(t, d_frames, nullptr, nullptr, nframes, 0, true);
        if (rc) return rc;
    } else {
        {
            const int ablocks = cmd_div_up(t->n, 256);
            int64_t flanes = (int64_t)g.sm_count * 16 / ablocks;
            if (flanes < 1) flanes = 1;
            if (flanes > nframes) flanes = nframes;
            if (flanes > 65535) flanes = 65535;
            k_dr<<<dim3(ablocks, (unsigned)flanes), 256, 0, st>>>(t->bx, d_frames, t->d_last,
                                                                  t->have_last ? 1 : 0, t->n, nframes, t->d_dr);
        }
        CMD_LAUNCHED();
        if ((rc = topo_schedule(t, t->d_dr, nframes))) return rc;
        if (skip) {
            int sched[3] = {0, 0, -1};
            CMD_CUDA(cudaMemcpyAsync(sched, t->d_sched, sizeof(sched), cudaMemcpyDeviceToHost, st));
            CMD_CUDA(cudaStreamSynchronize(st));
            if (sched[0] > 0) {
                rc = launch_pairs(t, d_frames, t->d_rebuild_ids + sched[0] - 1, nullptr, 1, 0, false);
                if (rc) return rc;
            }
        } else {
        rc = launch_pairs(t, d_frames, t->d_rebuild_ids, t->d_sched, nframes, 0, false);
        if (rc) return rc;
        size_t rsmem = (size_t)t->n * 24;
        // up to 100 KB (4266 atoms) the frame is staged in shared memory: two or more CTAs per SM
        // (staging frames of up to 100 KB through opt-in shared memory was measured on C3, 2048
        // atoms: 0.70 ms per 512 frames against 0.64 ms for the split path -- not taken)
        if (rsmem > 40 * 1024) {
            // frames too large to stage: several CTAs per frame, coordinates through L1 / L2
            int chunks = (int)((t->stride + 8191) / 8192);
            if (chunks < 1) chunks = 1;
            if (chunks > 1024) chunks = 1024;
            if (chunks > 1 && t->part_cap < (size_t)nframes * chunks) {
                CMD_CUDA(cudaStreamSynchronize(st));
                cudaFree(t->d_part);
                t->d_part = nullptr;
                t->part_cap = 0;
                if (cudaMalloc((void **)&t->d_part, (size_t)nframes * chunks * 8) != cudaSuccess) {
                    cudaGetLastError();
                    return cmd_set_error(CMD_ENOMEM, "cudaMalloc failed for the refresh partial sums");
                }
                t->part_cap = (size_t)nframes * chunks;
            }
            k_refresh<false><<<dim3((unsigned)nframes, chunks), 256, 0, st>>>(
                t->bx_refresh, t->rate, d_frames, t->d_refresh_ids, t->d_sched + 1, t->d_head, t->n, t->stride,
                t->d_carry_start, t->d_carry_dest, t->d_carry_count, t->d_start, t->d_dest, t->d_dist,
                t->d_omega, t->d_counts, t->d_rate_sum, t->d_carry_rowoff, t->d_rowoff, t->d_part);
            if (chunks > 1) {
                CMD_LAUNCHED();
                k_refresh_sum<<<(unsigned)((nframes + 255) / 256), 256, 0, st>>>(
                    t->d_refresh_ids, t->d_sched + 1, t->d_part, chunks, t->d_rate_sum);
            }
        } else
        k_refresh<true><<<(unsigned)nframes, 256, rsmem, st>>>(
            t->bx_refresh, t->rate, d_frames, t->d_refresh_ids, t->d_sched + 1, t->d_head, t->n, t->stride,
            t->d_carry_start, t->d_carry_dest, t->d_carry_count, t->d_start, t->d_dest, t->d_dist,
            t->d_omega, t->d_counts, t->d_rate_sum, t->d_carry_rowoff, t->d_rowoff, nullptr);
        CMD_LAUNCHED();
        }
        int carry_blocks = (int)((t->stride + 2047) / 2048);   // a copy of up to `stride` pairs
        if (carry_blocks < 8) carry_blocks = 8;
        if (carry_blocks > 4 * g.sm_count) carry_blocks = 4 * g.sm_count;
        k_carry<<<carry_blocks, 256, 0, st>>>(t->d_sched, t->d_start, t->d_dest, t->d_counts, t->stride,
                                   t->d_carry_start, t->d_carry_dest, t->d_carry_count,
                                   t->d_rowoff, t->d_carry_rowoff, t->n);
        CMD_LAUNCHED();
        CMD_CUDA(cudaMemcpyAsync(t->d_last, d_frames + (nframes - 1) * (int64_t)t->n * 3,
                                 (size_t)t->n * 24, cudaMemcpyDeviceToDevice, st));
        // after this block the head, if any, lives in the carry
        CMD_CUDA(cudaMemsetAsync(t->d_sched + 2, 0xff, sizeof(int), st));
        t->have_last = true;
    }
    t->total_frames += nframes;
    return topo_check_capacity(t);
}

// ---- frame-block sharding from all-gathered step lengths --------------------------------------
// dr[f][i] = length(frame[f-1][i], frame[f][i]) of a block on the device (topology.py:98);
// frame 0 of the block is measured against d_prev (the frame before the block; NULL: zeros, the
// very first frame of a trajectory, topology.py:95).  Touches no state of the topology.
extern "C" int cmd_topo_dr_dev(const cmd_topo *t, const double *d_frames, int64_t nframes,
                               const double *d_prev, double *d_dr)
{
    CMD_REQUIRE_INIT();
    if (!t || !d_frames || !d_dr || nframes < 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    CmdGlobal &g = cmd_global();
    const int ablocks = cmd_div_up(t->n, 256);
    int64_t flanes = (int64_t)g.sm_count * 16 / ablocks;
    if (flanes < 1) flanes = 1;
    if (flanes > nframes) flanes = nframes;
    if (flanes > 65535) flanes = 65535;
    k_dr<<<dim3(ablocks, (unsigned)flanes), 256, 0, g.stream>>>(t->bx, d_frames, d_prev, d_prev ? 1 : 0,
                                                              t->n, nframes, d_dr);
    CMD_LAUNCHED();
    return CMD_OK;
}

// A Verlet topology that has built no list yet walks the rebuild schedule of frames that PRECEDE
// this rank's block from their step lengths alone (every rank's dr, all-gathered: N x 8 bytes per
// frame instead of the coordinates).  May be called chunk after chunk (the displacement carries
// over).  *h_last_rebuild = index, relative to this chunk, of its last rebuild frame (the list in
// force afterwards was built there), -1 if the chunk holds none.  Follow with cmd_topo_seed_dev.
extern "C" int cmd_topo_skip_dr_dev(cmd_topo *t, const double *d_dr, int64_t nframes,
                                    int64_t *h_last_rebuild)
{
    CMD_REQUIRE_INIT();
    if (!t || !h_last_rebuild || nframes < 0 || (nframes && !d_dr))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    if (t->mode != CMD_TOPO_VERLET) return cmd_set_error(CMD_ESTATE, "not a Verlet-mode topology");
    if (t->have_last) return cmd_set_error(CMD_ESTATE, "the topology has built lists already");
    if (t->stride == 0) return cmd_set_error(CMD_ESTATE, "create the topology with a capacity first");
    *h_last_rebuild = -1;
    if (nframes == 0) return CMD_OK;
    if (nframes > 0x7fffffff / 2) return cmd_set_error(CMD_EINVAL, "block too large");
    cudaStream_t st = cmd_global().stream;
    int rc;
    if ((rc = topo_reserve(t, nframes))) return rc;
    if ((rc = topo_schedule(t, d_dr, nframes))) return rc;
    int sched[3] = {0, 0, -1};
    CMD_CUDA(cudaMemcpyAsync(sched, t->d_sched, sizeof(sched), cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    if (sched[0] > 0) {
        int last = -1;
        CMD_CUDA(cudaMemcpyAsync(&last, t->d_rebuild_ids + sched[0] - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        CMD_CUDA(cudaStreamSynchronize(st));
        *h_last_rebuild = last;
    }
    // the chunk's heads are never materialised: the list in force is seeded by cmd_topo_seed_dev
    CMD_CUDA(cudaMemsetAsync(t->d_sched + 2, 0xff, sizeof(int), st));
    t->total_frames += nframes;
    t->nframes = 0;
    return CMD_OK;
}

// Completes cmd_topo_skip_dr_dev: builds the list of the last rebuild frame from its coordinates
// (d_frame_rebuild, one frame) and carries it, and remembers d_frame_prev (the frame right before
// the block) for the first step length of the block.  After this the topology is in exactly the
// state a sequential run has at the block start.
extern "C" int cmd_topo_seed_dev(cmd_topo *t, const double *d_frame_rebuild, const double *d_frame_prev)
{
    CMD_REQUIRE_INIT();
    if (!t || !d_frame_rebuild || !d_frame_prev) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (t->mode != CMD_TOPO_VERLET || t->total_frames == 0 || t->stride == 0)
        return cmd_set_error(CMD_ESTATE, "call cmd_topo_skip_dr_dev first");
    t->done_valid = false;
    CmdGlobal &g = cmd_global();
    cudaStream_t st = g.stream;
    int rc;
    if ((rc = topo_reserve(t, 1))) return rc;
    if (!t->d_carry_start) {
        if (cudaMalloc((void **)&t->d_carry_start, t->stride * 4) != cudaSuccess ||
            cudaMalloc((void **)&t->d_carry_dest, t->stride * 4) != cudaSuccess ||
            cudaMalloc((void **)&t->d_carry_rowoff, (size_t)(t->n + 1) * 4) != cudaSuccess) {
            cudaGetLastError();
            return cmd_set_error(CMD_ENOMEM, "cudaMalloc failed for the carried pair list");
        }
    }
    if ((rc = launch_pairs(t, d_frame_rebuild, nullptr, nullptr, 1, 0, false))) return rc;
    CMD_CUDA(cudaMemsetAsync(t->d_sched + 2, 0, sizeof(int), st));      // head = slot 0
    int carry_blocks = (int)((t->stride + 2047) / 2048);
    if (carry_blocks < 8) carry_blocks = 8;
    if (carry_blocks > 4 * g.sm_count) carry_blocks = 4 * g.sm_count;
    k_carry<<<carry_blocks, 256, 0, st>>>(t->d_sched, t->d_start, t->d_dest, t->d_counts, t->stride,
                                          t->d_carry_start, t->d_carry_dest, t->d_carry_count,
                                          t->d_rowoff, t->d_carry_rowoff, t->n);
    CMD_LAUNCHED();
    CMD_CUDA(cudaMemsetAsync(t->d_sched + 2, 0xff, sizeof(int), st));   // the head lives in the carry
    CMD_CUDA(cudaMemcpyAsync(t->d_last, d_frame_prev, (size_t)t->n * 24, cudaMemcpyDeviceToDevice, st));
    t->have_last = true;
    t->nframes = 0;
    return topo_check_capacity(t);
}

// host frames -> the topology's staging buffer in HBM (float32 blocks are up-cast on the device)
static int topo_stage(cmd_topo *t, const void *h_frames, int dtype_bytes, int64_t nframes)
{
    cudaStream_t st = cmd_global().stream;
    const size_t elems = (size_t)nframes * t->n * 3;
    const bool raw = t->d_sel != nullptr || dtype_bytes == 4;   // the copy lands behind the frames
    const size_t raw_bytes = (size_t)nframes * (t->d_sel ? t->n_total : t->n) * 3 * dtype_bytes;
    const size_t need = elems * 8 + (raw ? raw_bytes : 0);
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
    void *dst = raw ? (void *)(t->d_upload + elems) : (void *)t->d_upload;
    int rc;
    if ((rc = cmd_h2d_staged(dst, h_frames, raw ? raw_bytes : elems * 8, st))) return rc;
    return topo_convert(t, dst, dtype_bytes, nframes, t->d_upload, st);
}

// Brute-force blocks from host memory: frames are independent, so the block is cut into chunks and
// the host->device copy of chunk i+1 (copy stream) overlaps the pair kernel of chunk i.
static int topo_build_pipelined(cmd_topo *t, const void *h_frames, int dtype_bytes, int64_t nframes,
                                bool async = false)
{
    CmdGlobal &g = cmd_global();
    cudaStream_t st = g.stream;
    const size_t per_frame = (size_t)t->n * 3;
    const size_t elems = (size_t)nframes * per_frame;
    const bool raw = t->d_sel != nullptr || dtype_bytes == 4;
    const size_t raw_per_frame = (size_t)(t->d_sel ? t->n_total : t->n) * 3 * dtype_bytes;   // bytes
    const size_t need = elems * 8 + (raw ? (size_t)nframes * raw_per_frame : 0);
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
    if (!g.copy_stream) {
        CMD_CUDA(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) CMD_CUDA(cudaEventCreateWithFlags(&g.copy_event[i], cudaEventDisableTiming));
    }
    if (!g.aux_stream) {
        CMD_CUDA(cudaStreamCreateWithFlags(&g.aux_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) CMD_CUDA(cudaEventCreateWithFlags(&g.aux_event[i], cudaEventDisableTiming));
    }
    static const bool one_stream = getenv("CMDLMC_B200_UPLOAD_ONE_STREAM") != nullptr;
    // Equal chunks.  (Measured and dropped in round 2: a small first chunk followed by large ones --
    // the kernel and the copy of a chunk take about the same time, so the second chunk's copy is
    // not done when the first chunk's kernel ends and the SMs idle: float32 2.04 -> 2.2-2.6 ms per
    // 16 384 C2 frames; tapered last chunks for float64 blocks gain 3 %, within the noise of the
    // host.)
    static const char *nch_env = getenv("CMDLMC_B200_UPLOAD_CHUNKS");
    const int nch = nch_env && atoi(nch_env) > 0 ? atoi(nch_env) : TOPO_UPLOAD_CHUNKS;
    int64_t chunk = (nframes + nch - 1) / nch;
    if (chunk < 1024) chunk = 1024;   // several consecutive frames per persistent CTA (skin list)
    // the staging buffer may still be read by kernels of the previous block: everything queued on
    // the stream (blocking call), or this topology's own previous block (asynchronous call -- the
    // copies must not wait for the blocks of OTHER topologies still queued on the stream)
    if (!t->done) CMD_CUDA(cudaEventCreateWithFlags(&t->done, cudaEventDisableTiming));
    if (async) {
        if (t->done_valid) CMD_CUDA(cudaStreamWaitEvent(g.copy_stream, t->done, 0));
    } else {
        CMD_CUDA(cudaEventRecord(g.copy_event[0], st));
        CMD_CUDA(cudaStreamWaitEvent(g.copy_stream, g.copy_event[0], 0));
    }
    unsigned char *d_raw = (unsigned char *)(t->d_upload + elems);
    int rc;
    bool sized = t->stride != 0;
    t->nframes = nframes;
    t->d_frames_last = t->d_upload;
    // Chunks alternate between the library's stream and a second one: a kernel on one stream does
    // not wait for the last CTAs of the chunk before it, whose SMs it fills as they drain (each
    // launch has its own set of skin lists).  The second stream starts behind whatever is queued
    // on the first and is joined into it at the end.
    CMD_CUDA(cudaEventRecord(g.aux_event[0], st));
    CMD_CUDA(cudaStreamWaitEvent(g.aux_stream, g.aux_event[0], 0));
    int ci = 0;
    struct StreamGuard {   // launch_pairs and the macros launch on cmd_global().stream
        CmdGlobal &g; cudaStream_t keep;
        ~StreamGuard() { g.stream = keep; }
    } guard{g, st};
    for (int64_t c0 = 0; c0 < nframes; ci++) {
        const int64_t cn = nframes - c0 < chunk ? nframes - c0 : chunk;
        const size_t off = (size_t)c0 * per_frame, ce = (size_t)cn * per_frame;
        const bool on_aux = !one_stream && (ci & 1) && sized;
        g.stream = st;
        if (!raw)
            rc = cmd_h2d_staged(t->d_upload + off, (const double *)h_frames + off, ce * 8, g.copy_stream);
        else
            rc = cmd_h2d_staged(d_raw + (size_t)c0 * raw_per_frame,
                                (const unsigned char *)h_frames + (size_t)c0 * raw_per_frame,
                                (size_t)cn * raw_per_frame, g.copy_stream);
        if (rc) return rc;
        CMD_CUDA(cudaEventRecord(g.copy_event[1], g.copy_stream));
        cudaStream_t cs = on_aux ? g.aux_stream : st;
        CMD_CUDA(cudaStreamWaitEvent(cs, g.copy_event[1], 0));
        g.stream = cs;
        if (raw && (rc = topo_convert(t, d_raw + (size_t)c0 * raw_per_frame, dtype_bytes, cn, t->d_upload + off, cs)))
            return rc;
        if (!sized) {   // first block ever: probe the capacity on the first frame, then allocate
            if ((rc = topo_autosize(t, t->d_upload))) return rc;
            sized = true;
        }
        if (c0 == 0 && (rc = topo_reserve(t, nframes))) return rc;
        t->list_set = on_aux ? 1 : 0;
        if ((rc = launch_pairs(t, t->d_upload + off, nullptr, nullptr, cn, c0, true))) return rc;
        c0 += cn;
    }
    g.stream = st;
    t->list_set = 0;
    CMD_CUDA(cudaEventRecord(g.aux_event[1], g.aux_stream));
    CMD_CUDA(cudaStreamWaitEvent(st, g.aux_event[1], 0));
    t->total_frames += nframes;
    CMD_CUDA(cudaEventRecord(t->done, st));
    t->done_valid = true;
    if (async) { t->pending = true; return CMD_OK; }
    return topo_check_capacity(t);
}

// the stream for read-backs of a finished block: behind the block's `done` event only
static int topo_ctl_stream(const cmd_topo *t, cudaStream_t *out)
{
    CmdGlobal &g = cmd_global();
    if (!t->done_valid) { *out = g.stream; return CMD_OK; }
    if (!g.ctl_stream) CMD_CUDA(cudaStreamCreateWithFlags(&g.ctl_stream, cudaStreamNonBlocking));
    CMD_CUDA(cudaStreamWaitEvent(g.ctl_stream, t->done, 0));
    *out = g.ctl_stream;
    return CMD_OK;
}

extern "C" int cmd_topo_build_async(cmd_topo *t, const void *h_frames, int dtype_bytes, int64_t nframes)
{
    CMD_REQUIRE_INIT();
    if (!t || !h_frames || nframes < 1 || (dtype_bytes != 4 && dtype_bytes != 8))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    if (nframes > 0x7fffffff / 2) return cmd_set_error(CMD_EINVAL, "block too large");
    if (t->pending) return cmd_set_error(CMD_ESTATE, "cmd_topo_wait has not been called for the last block");
    if (t->mode != CMD_TOPO_BRUTEFORCE || t->stride == 0 || nframes < 512)
        return cmd_set_error(CMD_ESTATE, "asynchronous blocks: brute-force mode, at least 512 frames, and "
                                         "a topology that has built one block already (capacity known)");
    return topo_build_pipelined(t, h_frames, dtype_bytes, nframes, true);
}

extern "C" int cmd_topo_wait(cmd_topo *t)
{
    CMD_REQUIRE_INIT();
    if (!t) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (!t->pending) return CMD_OK;
    t->pending = false;
    cudaStream_t cs;
    int rc = topo_ctl_stream(t, &cs);
    if (rc) return rc;
    int err = 0;
    CMD_CUDA(cudaMemcpyAsync(&err, t->d_err, sizeof(int), cudaMemcpyDeviceToHost, cs));
    CMD_CUDA(cudaStreamSynchronize(cs));
    if (err > 0) {
        CMD_CUDA(cudaMemsetAsync(t->d_err, 0, sizeof(int), cs));
        CMD_CUDA(cudaStreamSynchronize(cs));
        t->capacity_needed = err;
        return cmd_set_error(CMD_ECAPACITY, "a frame has %d directed pairs but the per-frame "
                             "capacity is %lld: re-create the topology with a larger "
                             "capacity_per_frame", err, (long long)t->stride);
    }
    return CMD_OK;
}

extern "C" int cmd_topo_set_selection(cmd_topo *t, int n_total, const int *h_index)
{
    CMD_REQUIRE_INIT();
    if (!t || n_total < 0 || (n_total > 0 && (!h_index || n_total < t->n)))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    CMD_CUDA(cudaStreamSynchronize(st));
    cudaFree(t->d_sel);
    t->d_sel = nullptr;
    t->n_total = 0;
    if (n_total == 0) return CMD_OK;
    for (int i = 0; i < t->n; i++)
        if (h_index[i] < 0 || h_index[i] >= n_total)
            return cmd_set_error(CMD_EINVAL, "selection index %d out of range (%d atoms)", h_index[i], n_total);
    if (cudaMalloc((void **)&t->d_sel, (size_t)t->n * 4) != cudaSuccess) {
        cudaGetLastError();
        return cmd_set_error(CMD_ENOMEM, "cudaMalloc failed for the selection");
    }
    CMD_CUDA(cudaMemcpyAsync(t->d_sel, h_index, (size_t)t->n * 4, cudaMemcpyHostToDevice, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    t->n_total = n_total;
    return CMD_OK;
}

extern "C" int cmd_topo_build(cmd_topo *t, const void *h_frames, int dtype_bytes, int64_t nframes)
{
    CMD_REQUIRE_INIT();
    if (!t || !h_frames || nframes < 1 || (dtype_bytes != 4 && dtype_bytes != 8))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    if (nframes > 0x7fffffff / 2) return cmd_set_error(CMD_EINVAL, "block too large");
    if (t->mode == CMD_TOPO_BRUTEFORCE && nframes >= 512)
        return topo_build_pipelined(t, h_frames, dtype_bytes, nframes);
    int rc = topo_stage(t, h_frames, dtype_bytes, nframes);
    if (rc) return rc;
    return cmd_topo_build_dev(t, t->d_upload, nframes);
}

extern "C" int cmd_topo_skip(cmd_topo *t, const void *h_frames, int dtype_bytes, int64_t nframes)
{
    CMD_REQUIRE_INIT();
    if (!t || !h_frames || nframes < 1 || (dtype_bytes != 4 && dtype_bytes != 8))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    if (t->mode != CMD_TOPO_VERLET) return cmd_topo_skip_dev(t, nullptr, nframes);
    int rc = topo_stage(t, h_frames, dtype_bytes, nframes);
    if (rc) return rc;
    return cmd_topo_skip_dev(t, t->d_upload, nframes);
}

extern "C" int cmd_topo_frame_info(const cmd_topo *t, int64_t *counts, uint8_t *rebuilt,
                                   double *rate_sum)
{
    CMD_REQUIRE_INIT();
    if (!t || t->nframes < 1) return cmd_set_error(CMD_ESTATE, "no block has been built");
    // a block built by the pipelined path is read back behind its own completion, not behind
    // whatever else is queued on the stream (asynchronous blocks of other topologies)
    cudaStream_t st;
    {
        int rc = topo_ctl_stream(t, &st);
        if (rc) return rc;
    }
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

int cmd_block_stats_launch(const int *d_counts, const double *d_rate_sum, int64_t nframes, double *d_out);

extern "C" int cmd_topo_block_stats_dev(const cmd_topo *t, double *d_out)
{
    CMD_REQUIRE_INIT();
    if (!t || !d_out) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (t->nframes < 1) return cmd_set_error(CMD_ESTATE, "no block has been built");
    return cmd_block_stats_launch(t->d_counts, t->d_rate_sum, t->nframes, d_out);
}

extern "C" int cmd_topo_set_path(cmd_topo *t, int path)
{
    if (!t || path < -1 || path > 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (t->nframes > 0 || t->total_frames > 0)
        return cmd_set_error(CMD_ESTATE, "the search path is fixed once a block has been built");
    t->force_path = path;
    if (t->stride > 0) return topo_configure(t, t->stride);
    return CMD_OK;
}

extern "C" int cmd_topo_path(const cmd_topo *t) { return t ? t->path : -1; }

extern "C" int64_t cmd_topo_stride(const cmd_topo *t) { return t ? t->stride : -1; }
extern "C" int64_t cmd_topo_capacity_needed(const cmd_topo *t) { return t ? t->capacity_needed : -1; }
extern "C" int cmd_topo_n_images(const cmd_topo *t) { return t ? t->bx.n_img : -1; }
extern "C" int64_t cmd_topo_nframes(const cmd_topo *t) { return t ? t->nframes : -1; }

// The lists of frames [f0, f0 + nf) of the last block in ONE strided copy per array: host rows of
// `width` entries (>= the largest count among them; the counts come from cmd_topo_frame_info).
extern "C" int cmd_topo_get_block(const cmd_topo *t, int64_t f0, int64_t nf, int64_t width, int *start,
                                  int *dest, double *dist, double *omega)
{
    CMD_REQUIRE_INIT();
    if (!t || f0 < 0 || nf < 1 || f0 + nf > t->nframes || width < 1 || width > t->stride)
        return cmd_set_error(CMD_EINVAL, "bad frame range or row width");
    cudaStream_t st = cmd_global().stream;
    const int64_t base = f0 * t->stride;
    const size_t sp = (size_t)t->stride;
#define GET2D(h, d, T)                                                                            \
    if (h) CMD_CUDA(cudaMemcpy2DAsync(h, (size_t)width * sizeof(T), (d) + base, sp * sizeof(T),   \
                                      (size_t)width * sizeof(T), (size_t)nf, cudaMemcpyDeviceToHost, st))
    GET2D(start, t->d_start, int);
    GET2D(dest, t->d_dest, int);
    GET2D(dist, t->d_dist, double);
    GET2D(omega, t->d_omega, double);
#undef GET2D
    CMD_CUDA(cudaStreamSynchronize(st));
    return CMD_OK;
}

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

extern "C" int cmd_topo_row_offsets(const cmd_topo *t, const int **d_rowoff)
{
    if (!t || t->nframes < 1 || !d_rowoff) return cmd_set_error(CMD_ESTATE, "no block has been built");
    *d_rowoff = t->d_rowoff;
    return CMD_OK;
}

extern "C" int cmd_topo_n_atoms(const cmd_topo *t) { return t ? t->n : -1; }

extern "C" int cmd_topo_positions(const cmd_topo *t, const double **d_frames)
{
    if (!t || t->nframes < 1 || !d_frames) return cmd_set_error(CMD_ESTATE, "no block has been built");
    *d_frames = t->d_frames_last;
    return CMD_OK;
}


// ---- AngleTopology._determine_colvars (topology.py:158-167) --------------------------------------
// theta[k] = atombox.angle(extra[group[start]], donor[start], donor[dest]) for every listed pair of
// the block; with a FermiAngle rate omega becomes 0 where theta < theta0
// (jumprate_generators.py:42-43).  One CTA per frame.
__global__ void __launch_bounds__(256)
k_pair_angles(const __grid_constant__ BoxParams bx, const double *__restrict__ donors,
              const double *__restrict__ extras, const int *__restrict__ group, int n, int n_extra,
              int64_t stride, const int *__restrict__ counts, const int *__restrict__ start,
              const int *__restrict__ dest, double *__restrict__ theta, double *__restrict__ omega,
              double *__restrict__ rate_sum, int mask_rate, double theta0)
{
    const int64_t f = blockIdx.x;
    const int p = counts[f];
    const double *don = donors + f * (int64_t)n * 3, *ext = extras + f * (int64_t)n_extra * 3;
    const int64_t base = f * stride;
    double rsum = 0.0;
    for (int k = threadIdx.x; k < p; k += blockDim.x) {
        const int i = start[base + k], j = dest[base + k], g = group[i];
        const double p1[3] = {ext[3 * g], ext[3 * g + 1], ext[3 * g + 2]};
        const double p2[3] = {don[3 * i], don[3 * i + 1], don[3 * i + 2]};
        const double p3[3] = {don[3 * j], don[3 * j + 1], don[3 * j + 2]};
        const double th = angle_exact(bx, p1, p2, p3);
        theta[base + k] = th;
        double om = omega[base + k];
        if (mask_rate && th < theta0) { om = 0.0; omega[base + k] = 0.0; }
        rsum += om;
    }
    __shared__ double wsum[8];
    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_down_sync(0xffffffffu, rsum, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = rsum;
    __syncthreads();
    if (threadIdx.x == 0 && mask_rate) {
        double t = 0;
        for (int w = 0; w < 8; w++) t += wsum[w];
        rate_sum[f] = t;
    }
}

extern "C" int cmd_topo_set_groups(cmd_topo *t, const int *h_group, int n_extra)
{
    CMD_REQUIRE_INIT();
    if (!t || !h_group || n_extra < 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    for (int i = 0; i < t->n; i++)
        if (h_group[i] < 0 || h_group[i] >= n_extra)
            return cmd_set_error(CMD_EINVAL, "donor %d has no extra atom (the reference raises "
                                             "KeyError there)", i);
    if (!t->d_group && cudaMalloc((void **)&t->d_group, (size_t)t->n * 4) != cudaSuccess) {
        cudaGetLastError();
        return cmd_set_error(CMD_ENOMEM, "cudaMalloc failed for the group map");
    }
    cudaStream_t st = cmd_global().stream;
    CMD_CUDA(cudaMemcpyAsync(t->d_group, h_group, (size_t)t->n * 4, cudaMemcpyHostToDevice, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    t->n_extra = n_extra;
    return CMD_OK;
}

extern "C" int cmd_topo_apply_angles_dev(cmd_topo *t, const double *d_extra_frames)
{
    CMD_REQUIRE_INIT();
    if (!t || !d_extra_frames) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (t->nframes < 1) return cmd_set_error(CMD_ESTATE, "no block has been built");
    if (!t->d_group) return cmd_set_error(CMD_ESTATE, "cmd_topo_set_groups has not been called");
    if (!t->d_theta) {
        size_t bytes = (size_t)t->cap_frames * t->stride * 8;
        if (cudaMalloc((void **)&t->d_theta, bytes) != cudaSuccess) {
            cudaGetLastError();
            return cmd_set_error(CMD_ENOMEM, "cudaMalloc of %zu bytes failed for the angles", bytes);
        }
    }
    k_pair_angles<<<(unsigned)t->nframes, 256, 0, cmd_global().stream>>>(
        t->bx, t->d_frames_last, d_extra_frames, t->d_group, t->n, t->n_extra, t->stride, t->d_counts,
        t->d_start, t->d_dest, t->d_theta, t->d_omega, t->d_rate_sum, t->rate_is_fermi_angle ? 1 : 0,
        t->rate.par[3]);
    CMD_LAUNCHED();
    return CMD_OK;
}

extern "C" int cmd_topo_apply_angles(cmd_topo *t, const void *h_extra_frames, int dtype_bytes)
{
    CMD_REQUIRE_INIT();
    if (!t || !h_extra_frames || (dtype_bytes != 4 && dtype_bytes != 8))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    if (t->nframes < 1 || t->n_extra < 1) return cmd_set_error(CMD_ESTATE, "no block / no groups");
    cudaStream_t st = cmd_global().stream;
    const size_t elems = (size_t)t->nframes * t->n_extra * 3;
    const size_t need = elems * 8 + (dtype_bytes == 4 ? elems * 4 : 0);
    if (t->extra_upload_bytes < need) {
        CMD_CUDA(cudaStreamSynchronize(st));
        cudaFree(t->d_extra_upload);
        t->d_extra_upload = nullptr;
        t->extra_upload_bytes = 0;
        if (cudaMalloc((void **)&t->d_extra_upload, need) != cudaSuccess) {
            cudaGetLastError();
            return cmd_set_error(CMD_ENOMEM, "cudaMalloc of %zu staging bytes failed", need);
        }
        t->extra_upload_bytes = need;
    }
    if (dtype_bytes == 8) {
        int rc = cmd_h2d_staged(t->d_extra_upload, h_extra_frames, elems * 8, st);
        if (rc) return rc;
    } else {
        float *d32 = (float *)(t->d_extra_upload + elems);
        int rc = cmd_h2d_staged(d32, h_extra_frames, elems * 4, st);
        if (rc) return rc;
        int blocks = cmd_div_up(elems, 256);
        if (blocks > cmd_global().sm_count * 16) blocks = cmd_global().sm_count * 16;
        k_upcast_f32<<<blocks, 256, 0, st>>>(d32, t->d_extra_upload, (int64_t)elems);
        CMD_LAUNCHED();
    }
    return cmd_topo_apply_angles_dev(t, t->d_extra_upload);
}

extern "C" int cmd_topo_get_frame_angles(const cmd_topo *t, int64_t f, double *h_theta)
{
    CMD_REQUIRE_INIT();
    if (!t || f < 0 || f >= t->nframes || !h_theta) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (!t->d_theta) return cmd_set_error(CMD_ESTATE, "cmd_topo_apply_angles has not been called");
    cudaStream_t st = cmd_global().stream;
    int p = 0;
    CMD_CUDA(cudaMemcpyAsync(&p, t->d_counts + f, 4, cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    if (p > 0) {
        CMD_CUDA(cudaMemcpyAsync(h_theta, t->d_theta + f * t->stride, (size_t)p * 8, cudaMemcpyDeviceToHost, st));
        CMD_CUDA(cudaStreamSynchronize(st));
    }
    return CMD_OK;
}


// ---- HydroniumTopology._determine_colvars, lattice-independent part (topology.py:234-249) ----------
// Per frame and site: the k listed neighbours with the smallest distance, ascending (np.argsort of
// the row's distances; equal distances keep list order).  One warp per (frame, site).  Output in
// the layout of a pair list -- entry e = site * k + q: start = site, dest, dist -- so the KMC kernels
// read it like any other list (stride = near_stride, count = n * k).
__global__ void __launch_bounds__(256)
k_nearest(const int *__restrict__ counts, const int *__restrict__ rowoff, int ro_pitch,
          const int *__restrict__ dest, const double *__restrict__ dist, int64_t stride, int n, int k,
          int64_t nframes, int64_t near_stride, int *__restrict__ ns, int *__restrict__ nd,
          double *__restrict__ ndist, int *__restrict__ ncounts, int *__restrict__ err)
{
    const int lane = threadIdx.x & 31;
    const int64_t wglobal = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (wglobal >= nframes * n) return;
    const int64_t f = wglobal / n;
    const int s = (int)(wglobal - f * n);
    const int *ro = rowoff + f * ro_pitch;
    const int r0 = ro[s], r1 = ro[s + 1];
    const int64_t base = f * stride, obase = f * near_stride + (int64_t)s * k;
    if (s == 0 && lane == 0) ncounts[f] = n * k;
    if (r1 - r0 < k) {   // the reference fails here (ValueError, topology.py:250)
        if (lane == 0) atomicMax(err, s + 1);
        for (int q = lane; q < k; q += 32) { ns[obase + q] = s; nd[obase + q] = s; ndist[obase + q] = INFINITY; }
        return;
    }
    unsigned long long taken = 0ull;   // entries of the row already emitted (rows hold < 64 * 32)
    for (int q = 0; q < k; q++) {
        double best = INFINITY;
        int bi = 0x7fffffff;
        for (int e = r0 + lane; e < r1; e += 32) {
            const int slot = (e - r0) >> 5;
            if (slot < 64 && ((taken >> slot) & 1ull)) continue;
            const double d = __ldg(dist + base + e);
            if (d < best || (d == best && e < bi)) { best = d; bi = e; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (((bi - r0) & 31) == lane) taken |= 1ull << ((bi - r0) >> 5);
        if (lane == 0) { ns[obase + q] = s; nd[obase + q] = __ldg(dest + base + bi); ndist[obase + q] = best; }
    }
}

extern "C" int cmd_topo_nearest(cmd_topo *t, int k)
{
    CMD_REQUIRE_INIT();
    if (!t || k < 1 || k > 16) return cmd_set_error(CMD_EINVAL, "bad argument (1 <= k <= 16)");
    if (t->nframes < 1) return cmd_set_error(CMD_ESTATE, "no block has been built");
    cudaStream_t st = cmd_global().stream;
    const int64_t ns = ((int64_t)t->n * k + 63) / 64 * 64;
    if (!t->d_near_dest || t->near_k != k) {
        CMD_CUDA(cudaStreamSynchronize(st));
        cudaFree(t->d_near_start); cudaFree(t->d_near_dest); cudaFree(t->d_near_counts); cudaFree(t->d_near_dist);
        t->d_near_start = t->d_near_dest = t->d_near_counts = nullptr;
        t->d_near_dist = nullptr;
        const size_t np = (size_t)t->cap_frames * ns;
        if (cudaMalloc((void **)&t->d_near_start, np * 4) != cudaSuccess ||
            cudaMalloc((void **)&t->d_near_dest, np * 4) != cudaSuccess ||
            cudaMalloc((void **)&t->d_near_dist, np * 8) != cudaSuccess ||
            cudaMalloc((void **)&t->d_near_counts, (size_t)t->cap_frames * 4) != cudaSuccess) {
            cudaGetLastError();
            return cmd_set_error(CMD_ENOMEM, "cudaMalloc failed for the nearest-neighbour arrays");
        }
        t->near_k = k;
        t->near_stride = ns;
    }
    const int64_t warps = t->nframes * t->n;
    k_nearest<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(
        t->d_counts, t->d_rowoff, cmd_ro_pitch(t->n), t->d_dest, t->d_dist, t->stride, t->n, k, t->nframes,
        t->near_stride, t->d_near_start, t->d_near_dest, t->d_near_dist, t->d_near_counts, t->d_err);
    CMD_LAUNCHED();
    int err = 0;
    CMD_CUDA(cudaMemcpyAsync(&err, t->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    if (err > 0) {
        CMD_CUDA(cudaMemsetAsync(t->d_err, 0, sizeof(int), st));
        return cmd_set_error(CMD_EINVAL, "site %d has fewer than %d listed neighbours (the reference "
                             "raises ValueError there, topology.py:250)", err - 1, k);
    }
    return CMD_OK;
}

extern "C" int cmd_topo_near_arrays(const cmd_topo *t, const int **d_start, const int **d_dest,
                                    const double **d_dist, const int **d_counts, int64_t *stride)
{
    if (!t || !t->d_near_dest || t->nframes < 1)
        return cmd_set_error(CMD_ESTATE, "cmd_topo_nearest has not been called for this block");
    if (d_start) *d_start = t->d_near_start;
    if (d_dest) *d_dest = t->d_near_dest;
    if (d_dist) *d_dist = t->d_near_dist;
    if (d_counts) *d_counts = t->d_near_counts;
    if (stride) *stride = t->near_stride;
    return CMD_OK;
}

extern "C" int cmd_topo_get_frame_nearest(const cmd_topo *t, int64_t f, int *h_dest, double *h_dist)
{
    CMD_REQUIRE_INIT();
    if (!t || f < 0 || f >= t->nframes) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (!t->d_near_dest) return cmd_set_error(CMD_ESTATE, "cmd_topo_nearest has not been called");
    cudaStream_t st = cmd_global().stream;
    const size_t cnt = (size_t)t->n * t->near_k;
    if (h_dest) CMD_CUDA(cudaMemcpyAsync(h_dest, t->d_near_dest + f * t->near_stride, cnt * 4, cudaMemcpyDeviceToHost, st));
    if (h_dist) CMD_CUDA(cudaMemcpyAsync(h_dist, t->d_near_dist + f * t->near_stride, cnt * 8, cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    return CMD_OK;
}

// ---- K5: histogram of the listed pair distances of the last block -------------
// One CTA per frame slice; shared-memory bins, one global atomic per non-empty bin per CTA.
__global__ void __launch_bounds__(256)
k_pair_hist(const double *__restrict__ dist, const int *__restrict__ counts, int64_t stride,
            int64_t nframes, double lo, double inv_width, int nbins,
            unsigned long long *__restrict__ hist)
{
    extern __shared__ unsigned int bins[];
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) bins[b] = 0u;
    __syncthreads();
    for (int64_t f = blockIdx.x; f < nframes; f += gridDim.x) {
        const int p = counts[f];
        const double *d = dist + f * stride;
        for (int k = threadIdx.x; k < p; k += blockDim.x) {
            const double b = floor((d[k] - lo) * inv_width);
            if (b >= 0 && b < nbins) atomicAdd(&bins[(int)b], 1u);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x)
        if (bins[b]) atomicAdd(hist + b, (unsigned long long)bins[b]);
}

extern "C" int cmd_topo_distance_histogram_dev(const cmd_topo *t, double lo, double hi, int nbins,
                                               unsigned long long *d_hist)
{
    CMD_REQUIRE_INIT();
    if (!t || !d_hist || nbins < 1 || nbins > 8192 || !(hi > lo))
        return cmd_set_error(CMD_EINVAL, "bad argument (1 <= nbins <= 8192, hi > lo)");
    if (t->nframes < 1) return cmd_set_error(CMD_ESTATE, "no block has been built");
    int64_t blocks = t->nframes < (int64_t)cmd_global().sm_count * 8 ? t->nframes
                                                                     : (int64_t)cmd_global().sm_count * 8;
    k_pair_hist<<<(unsigned)blocks, 256, (size_t)nbins * 4, cmd_global().stream>>>(
        t->d_dist, t->d_counts, t->stride, t->nframes, lo, nbins / (hi - lo), nbins, d_hist);
    CMD_LAUNCHED();
    return CMD_OK;
}

extern "C" int cmd_topo_distance_histogram(const cmd_topo *t, double lo, double hi, int nbins,
                                           int64_t *h_hist)
{
    CMD_REQUIRE_INIT();
    if (!h_hist || nbins < 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    void *d;
    int rc = cmd_scratch(3, (size_t)nbins * 8, &d);
    if (rc) return rc;
    cudaStream_t st = cmd_global().stream;
    CMD_CUDA(cudaMemsetAsync(d, 0, (size_t)nbins * 8, st));
    if ((rc = cmd_topo_distance_histogram_dev(t, lo, hi, nbins, (unsigned long long *)d))) return rc;
    int64_t *tmp = (int64_t *)malloc((size_t)nbins * 8);
    if (!tmp) return cmd_set_error(CMD_ENOMEM, "out of host memory");
    cudaError_t e = cudaMemcpyAsync(tmp, d, (size_t)nbins * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) for (int b = 0; b < nbins; b++) h_hist[b] += tmp[b];
    free(tmp);
    CMD_CUDA(e);
    return CMD_OK;
}

extern "C" int cmd_topo_skin_stats(const cmd_topo *t, int64_t *frames, int64_t *rebuilds, int64_t *list_entries)
{
    CMD_REQUIRE_INIT();
    if (!t) return cmd_set_error(CMD_EINVAL, "bad argument");
    unsigned long long v[4] = {0, 0, 0, 0};
    cudaStream_t st = cmd_global().stream;
    CMD_CUDA(cudaMemcpyAsync(v, t->d_ties, sizeof(v), cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    if (frames) *frames = (int64_t)v[2];
    if (rebuilds) *rebuilds = (int64_t)v[1];
    if (list_entries) *list_entries = (int64_t)v[3];
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
