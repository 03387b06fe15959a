// lmc.cu -- the legacy LMC sweep (row A14 of SURVEY.md section 8).  PARITY UNPINNED: the engine
// (LMCHelper.pyx / LMCRoutine.sweep) is not in the reference tree; this restates the in-tree
// specification -- "A sweep is the number of single proton jump attempts, after which (on
// average) each oxygen bond has been selected once" (mdlmc/IO/config_parser.py:182-189), per-frame
// jump probabilities omega(d) * dt (KMC/excess_kmc.py:398-403), sweep / sweep_with_jumpmatrix
// (tests/LMC/test_MDMC.py.bak:61-84) -- and is checked against oracle/cmdlmc_oracle.c
// orc_lmc_sweep, which restates the same text.
//
// One sweep on frame f = P_f attempts.  Attempt a picks a listed pair k uniformly in [0, P_f) and
// a uniform u in [0, 1); the proton hops start[k] -> dest[k] iff the start site is occupied, the
// destination is empty and u < omega[k] * prob_scale.  Attempts of a sweep depend on each other
// through the lattice, so a WARP per replica evaluates 32 attempts against the current lattice,
// commits the first accepted one (ballot + ffs) and re-evaluates only the attempts behind it --
// exact, and cheap because acceptance is rare.
//
//   replay mode  pick / u come from host-pregenerated streams (e.g. a GSL-style MT19937
//                gsl_rng_uniform_int / gsl_rng_uniform sequence): occupancy trajectories are
//                bit-identical to the CPU restatement on the same streams;
//   Philox mode  Philox4x32-10, counter = (attempt pair, sweep, replica), key = seed; one call
//                yields the pick and a 32-bit acceptance uniform for two attempts.
#include <math.h>
#include <stdlib.h>

#include "pbc.cuh"
#include "philox.cuh"
#include "tma.cuh"

struct cmd_lmc {
    int n_sites, n_replicas, rng_mode;
    uint64_t seed;
    int *d_lattice;                 // [R][n_sites]
    long long *d_jumps, *d_attempts, *d_cursor, *d_sweeps;   // [R]
    int *d_pick;                    // replay: [R][n_stream]
    double *d_acc;
    int64_t n_stream;
    unsigned long long *d_jumpmatrix;   // [n_sites][n_sites] summed over replicas (optional)
    int *d_halt;                    // [R] 1 = replay stream exhausted
};

struct LmcArgs {
    int n_sites, n_replicas, rng_mode, replicas_per_cta, sweeps_per_frame;
    uint64_t seed;
    int64_t stride, nframes, n_stream;
    double prob_scale;
    const int *start, *dest, *counts;
    const double *omega;
    int *lattice;
    long long *jumps, *attempts, *cursor, *sweeps;
    const int *pick;
    const double *acc;
    unsigned long long *jumpmatrix;
    int *halt;
    int resident;        // the frame's arrays are staged in shared memory (TMA) for all replicas
    int64_t buf_pairs;   // capacity of that stage, a multiple of 64 pairs
};

// One sweep-block of the frame for one replica (warp): `p` attempts against `lat`.
// RES = true: start / dest / omega point into shared memory (the frame was staged by TMA).
template <bool RES>
__device__ __forceinline__ void lmc_frame(const LmcArgs &a, int r, int lane, int *lat, int p,
                                          const int *__restrict__ fstart, const int *__restrict__ fdest,
                                          const double *__restrict__ fomega, long long &jumps,
                                          long long &attempts, long long &cursor, long long &sweeps,
                                          bool &halted)
{
    for (int sw = 0; sw < a.sweeps_per_frame && !halted; sw++) {
        if (a.rng_mode == CMD_RNG_REPLAY && cursor + p > a.n_stream) {
            halted = true;   // not enough pregenerated numbers for a whole sweep
            break;
        }
        // 128 attempts per trip: the four groups draw their numbers and gather their pairs first
        // (independent chains), then commit in attempt order group by group
        for (int a0 = 0; a0 < p; a0 += 128) {
            int si[4], di[4];
            bool cand[4];
            uint32_t rnd[2][4];
            if (a.rng_mode != CMD_RNG_REPLAY) {
                // one Philox call serves two attempts: (pick, 32-bit acceptance uniform) each
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    rnd[h][0] = (uint32_t)(a0 + 64 * h + lane); rnd[h][1] = (uint32_t)sweeps;
                    rnd[h][2] = (uint32_t)r; rnd[h][3] = (uint32_t)((uint64_t)sweeps >> 32);
                    philox4x32_10(rnd[h], (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                }
            }
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int at = a0 + 32 * g + lane;
                const bool valid = at < p;
                int k = 0;
                double u = 2.0;
                if (valid) {
                    if (a.rng_mode == CMD_RNG_REPLAY) {
                        k = a.pick[(int64_t)r * a.n_stream + cursor + at];
                        u = a.acc[(int64_t)r * a.n_stream + cursor + at];
                    } else {
                        k = (int)__umulhi(rnd[g >> 1][2 * (g & 1)], (uint32_t)p);
                        u = ((double)rnd[g >> 1][2 * (g & 1) + 1] + 0.5) * 2.3283064365386963e-10;
                    }
                }
                const bool inrange = valid && k >= 0 && k < p;
                si[g] = inrange ? (RES ? fstart[k] : __ldg(fstart + k)) : 0;
                di[g] = inrange ? (RES ? fdest[k] : __ldg(fdest + k)) : 0;
                const double om = inrange ? (RES ? fomega[k] : __ldg(fomega + k)) : 0.0;
                // acceptance against the probability does not depend on the lattice
                cand[g] = inrange && u < __dmul_rn(om, a.prob_scale);
            }
#pragma unroll
            for (int g = 0; g < 4; g++) {
                unsigned pending = __ballot_sync(0xffffffffu, cand[g]);
                while (pending) {
                    const bool ok = ((pending >> lane) & 1u) && lat[si[g]] != 0 && lat[di[g]] == 0;
                    const unsigned bal = __ballot_sync(0xffffffffu, ok);
                    if (!bal) break;
                    const int l = __ffs(bal) - 1;
                    if (lane == l) {
                        lat[di[g]] = lat[si[g]];
                        lat[si[g]] = 0;
                        if (a.jumpmatrix)
                            atomicAdd(a.jumpmatrix + (int64_t)si[g] * a.n_sites + di[g], 1ull);
                    }
                    jumps++;
                    pending &= l == 31 ? 0u : ~((2u << l) - 1u);   // attempts behind the hop
                    __syncwarp();
                }
            }
        }
        attempts += p;
        cursor += p;
        sweeps++;
    }
}

__global__ void __launch_bounds__(512, 1) k_lmc_sweep(const __grid_constant__ LmcArgs a)
{
    extern __shared__ __align__(16) unsigned char lmc_smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * a.replicas_per_cta + w;
    const bool active = r < a.n_replicas;
    // resident mode: [omega f64 | start i32 | dest i32] of one frame, one mbarrier, the lattices
    double *s_omega = (double *)lmc_smem;
    int *s_start = (int *)(s_omega + (a.resident ? a.buf_pairs : 0));
    int *s_dest = s_start + (a.resident ? a.buf_pairs : 0);
    uint64_t *bar = (uint64_t *)(s_dest + (a.resident ? a.buf_pairs : 0));
    int *lat = (int *)(bar + (a.resident ? 2 : 0)) + (size_t)w * a.n_sites;
    long long jumps = 0, attempts = 0, cursor = 0, sweeps = 0;
    bool halted = false;
    if (active) {
        for (int s = lane; s < a.n_sites; s += 32) lat[s] = a.lattice[(int64_t)r * a.n_sites + s];
        jumps = a.jumps[r]; attempts = a.attempts[r]; cursor = a.cursor[r]; sweeps = a.sweeps[r];
        halted = a.halt[r] != 0;
        __syncwarp();
    }
    if (a.resident && threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned nload = 0;   // completed TMA stages: the mbarrier phase
    for (int64_t f = 0; f < a.nframes; f++) {
        const int p = a.counts[f];
        const int64_t base = f * a.stride;
        if (a.resident) {
            // the picks of a sweep are random: the whole frame has to be on chip.  One thread
            // issues three TMA bulk copies, every replica of the CTA then gathers from shared memory
            if (threadIdx.x == 0 && p > 0) {
                const uint32_t e = (uint32_t)((p + 63) / 64 * 64);
                mbar_expect_tx(bar, e * 16u);
                tma_load_1d(s_omega, a.omega + base, e * 8u, bar);
                tma_load_1d(s_start, a.start + base, e * 4u, bar);
                tma_load_1d(s_dest, a.dest + base, e * 4u, bar);
            }
            if (p > 0) { mbar_wait(bar, nload & 1u); nload++; }
            if (active && !halted && p > 0)
                lmc_frame<true>(a, r, lane, lat, p, s_start, s_dest, s_omega, jumps, attempts, cursor,
                                sweeps, halted);
        } else if (active && !halted && p > 0) {
            lmc_frame<false>(a, r, lane, lat, p, a.start + base, a.dest + base, a.omega + base, jumps,
                             attempts, cursor, sweeps, halted);
        }
        __syncthreads();   // every replica is done with the frame before it is replaced
    }
    if (active) {
        __syncwarp();
        for (int s = lane; s < a.n_sites; s += 32) a.lattice[(int64_t)r * a.n_sites + s] = lat[s];
        if (lane == 0) {
            a.jumps[r] = jumps; a.attempts[r] = attempts; a.cursor[r] = cursor; a.sweeps[r] = sweeps;
            a.halt[r] = halted ? 1 : 0;
        }
    }
}

// ------------------------------------------------------------------ host side ------------------
extern "C" void cmd_lmc_destroy(cmd_lmc *k)
{
    if (!k) return;
    cudaStreamSynchronize(cmd_global().stream);
    cudaFree(k->d_lattice); cudaFree(k->d_jumps); cudaFree(k->d_attempts); cudaFree(k->d_cursor);
    cudaFree(k->d_sweeps); cudaFree(k->d_pick); cudaFree(k->d_acc); cudaFree(k->d_jumpmatrix);
    cudaFree(k->d_halt);
    free(k);
}

#define LALLOC(ptr, bytes)                                                                   \
    if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) {                               \
        cudaGetLastError();                                                                  \
        cmd_lmc_destroy(k);                                                                  \
        return cmd_set_error(CMD_ENOMEM, "cudaMalloc failed for %s (%zu bytes)", #ptr,       \
                             (size_t)(bytes));                                               \
    }

extern "C" int cmd_lmc_create(int n_sites, int n_replicas, const int *lattices, int rng_mode,
                              uint64_t seed, cmd_lmc **out)
{
    CMD_REQUIRE_INIT();
    if (!out || !lattices || n_sites < 1 || n_replicas < 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (rng_mode != CMD_RNG_REPLAY && rng_mode != CMD_RNG_PHILOX)
        return cmd_set_error(CMD_EINVAL, "bad rng mode");
    if ((size_t)n_sites * 4 > 200 * 1024)
        return cmd_set_error(CMD_ECAPACITY, "lattice of %d sites exceeds shared memory", n_sites);
    cmd_lmc *k = (cmd_lmc *)calloc(1, sizeof(cmd_lmc));
    if (!k) return cmd_set_error(CMD_ENOMEM, "out of host memory");
    k->n_sites = n_sites;
    k->n_replicas = n_replicas;
    k->rng_mode = rng_mode;
    k->seed = seed;
    cudaStream_t st = cmd_global().stream;
    size_t R = (size_t)n_replicas;
    LALLOC(k->d_lattice, R * n_sites * 4);
    LALLOC(k->d_jumps, R * 8);
    LALLOC(k->d_attempts, R * 8);
    LALLOC(k->d_cursor, R * 8);
    LALLOC(k->d_sweeps, R * 8);
    LALLOC(k->d_halt, R * 4);
    CMD_CUDA(cudaMemcpyAsync(k->d_lattice, lattices, R * n_sites * 4, cudaMemcpyHostToDevice, st));
    CMD_CUDA(cudaMemsetAsync(k->d_jumps, 0, R * 8, st));
    CMD_CUDA(cudaMemsetAsync(k->d_attempts, 0, R * 8, st));
    CMD_CUDA(cudaMemsetAsync(k->d_cursor, 0, R * 8, st));
    CMD_CUDA(cudaMemsetAsync(k->d_sweeps, 0, R * 8, st));
    CMD_CUDA(cudaMemsetAsync(k->d_halt, 0, R * 4, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    *out = k;
    return CMD_OK;
}

extern "C" int cmd_lmc_set_replay_stream(cmd_lmc *k, const int *pick, const double *acc, int64_t n)
{
    CMD_REQUIRE_INIT();
    if (!k || !pick || !acc || n < 1) return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    CMD_CUDA(cudaStreamSynchronize(st));
    cudaFree(k->d_pick); cudaFree(k->d_acc);
    k->d_pick = nullptr; k->d_acc = nullptr;
    size_t R = (size_t)k->n_replicas;
    LALLOC(k->d_pick, R * n * 4);
    LALLOC(k->d_acc, R * n * 8);
    CMD_CUDA(cudaMemcpyAsync(k->d_pick, pick, R * n * 4, cudaMemcpyHostToDevice, st));
    CMD_CUDA(cudaMemcpyAsync(k->d_acc, acc, R * n * 8, cudaMemcpyHostToDevice, st));
    CMD_CUDA(cudaMemsetAsync(k->d_cursor, 0, R * 8, st));
    CMD_CUDA(cudaMemsetAsync(k->d_halt, 0, R * 4, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    k->n_stream = n;
    return CMD_OK;
}

extern "C" int cmd_lmc_enable_jump_matrix(cmd_lmc *k, int enable)
{
    CMD_REQUIRE_INIT();
    if (!k) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (enable && !k->d_jumpmatrix) {
        size_t bytes = (size_t)k->n_sites * k->n_sites * 8;
        LALLOC(k->d_jumpmatrix, bytes);
        CMD_CUDA(cudaMemsetAsync(k->d_jumpmatrix, 0, bytes, cmd_global().stream));
    } else if (!enable && k->d_jumpmatrix) {
        CMD_CUDA(cudaStreamSynchronize(cmd_global().stream));
        cudaFree(k->d_jumpmatrix);
        k->d_jumpmatrix = nullptr;
    }
    return CMD_OK;
}

extern "C" int cmd_lmc_advance(cmd_lmc *k, const cmd_topo *t, double prob_scale, int sweeps_per_frame)
{
    CMD_REQUIRE_INIT();
    if (!k || !t || sweeps_per_frame < 1 || !(prob_scale >= 0))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    if (k->rng_mode == CMD_RNG_REPLAY && !k->d_pick)
        return cmd_set_error(CMD_ESTATE, "replay mode needs cmd_lmc_set_replay_stream first");
    const int *d_start, *d_dest, *d_counts;
    const double *d_dist, *d_omega;
    int rc = cmd_topo_device_arrays(t, &d_start, &d_dest, &d_dist, &d_omega, &d_counts);
    if (rc) return rc;
    CmdGlobal &g = cmd_global();
    LmcArgs a;
    memset(&a, 0, sizeof(a));
    a.n_sites = k->n_sites; a.n_replicas = k->n_replicas; a.rng_mode = k->rng_mode;
    a.sweeps_per_frame = sweeps_per_frame; a.seed = k->seed;
    a.stride = cmd_topo_stride(t); a.nframes = cmd_topo_nframes(t); a.n_stream = k->n_stream;
    a.prob_scale = prob_scale;
    a.start = d_start; a.dest = d_dest; a.counts = d_counts; a.omega = d_omega;
    a.lattice = k->d_lattice; a.jumps = k->d_jumps; a.attempts = k->d_attempts;
    a.cursor = k->d_cursor; a.sweeps = k->d_sweeps; a.pick = k->d_pick; a.acc = k->d_acc;
    a.jumpmatrix = k->d_jumpmatrix; a.halt = k->d_halt;
    size_t per_warp = (size_t)k->n_sites * 4;
    int rpc = (k->n_replicas + g.sm_count - 1) / g.sm_count;
    if (rpc < 1) rpc = 1;
    if (rpc > 16) rpc = 16;
    while (rpc > 1 && per_warp * rpc > 200 * 1024) rpc--;
    // resident mode: the largest frame of the block (16 B per pair) next to the lattices
    int maxp = 0;
    {
        int64_t nf = a.nframes;
        int *hc = (int *)malloc((size_t)nf * sizeof(int));
        if (!hc) return cmd_set_error(CMD_ENOMEM, "out of host memory");
        cudaError_t e = cudaMemcpyAsync(hc, d_counts, (size_t)nf * 4, cudaMemcpyDeviceToHost, g.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
        for (int64_t f = 0; e == cudaSuccess && f < nf; f++) if (hc[f] > maxp) maxp = hc[f];
        free(hc);
        CMD_CUDA(e);
    }
    a.buf_pairs = ((int64_t)maxp + 63) / 64 * 64;
    if (a.buf_pairs > a.stride) a.buf_pairs = a.stride;
    a.resident = 0;
    size_t smem = per_warp * rpc;
    if (a.buf_pairs > 0 && (size_t)a.buf_pairs * 16 + 16 + per_warp * rpc <= 220 * 1024) {
        a.resident = 1;
        smem = (size_t)a.buf_pairs * 16 + 16 + per_warp * rpc;
    }
    a.replicas_per_cta = rpc;
    CMD_CUDA(cudaFuncSetAttribute(k_lmc_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (k->n_replicas + rpc - 1) / rpc;
    k_lmc_sweep<<<blocks, rpc * 32, smem, g.stream>>>(a);
    CMD_LAUNCHED();
    return CMD_OK;
}

extern "C" int cmd_lmc_get_state(const cmd_lmc *k, int *lattices, int64_t *jumps, int64_t *attempts,
                                 int64_t *sweeps, int *halted)
{
    CMD_REQUIRE_INIT();
    if (!k) return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    size_t R = (size_t)k->n_replicas;
    if (lattices) CMD_CUDA(cudaMemcpyAsync(lattices, k->d_lattice, R * k->n_sites * 4, cudaMemcpyDeviceToHost, st));
    if (jumps) CMD_CUDA(cudaMemcpyAsync(jumps, k->d_jumps, R * 8, cudaMemcpyDeviceToHost, st));
    if (attempts) CMD_CUDA(cudaMemcpyAsync(attempts, k->d_attempts, R * 8, cudaMemcpyDeviceToHost, st));
    if (sweeps) CMD_CUDA(cudaMemcpyAsync(sweeps, k->d_sweeps, R * 8, cudaMemcpyDeviceToHost, st));
    if (halted) CMD_CUDA(cudaMemcpyAsync(halted, k->d_halt, R * 4, cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    return CMD_OK;
}

extern "C" int cmd_lmc_get_jump_matrix(const cmd_lmc *k, int64_t *h_matrix)
{
    CMD_REQUIRE_INIT();
    if (!k || !h_matrix) return cmd_set_error(CMD_EINVAL, "bad argument");
    if (!k->d_jumpmatrix) return cmd_set_error(CMD_ESTATE, "the jump matrix is not enabled");
    cudaStream_t st = cmd_global().stream;
    CMD_CUDA(cudaMemcpyAsync(h_matrix, k->d_jumpmatrix, (size_t)k->n_sites * k->n_sites * 8,
                             cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    return CMD_OK;
}
