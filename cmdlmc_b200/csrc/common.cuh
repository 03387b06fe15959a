// common.cuh -- shared host/device declarations of libcmdlmc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cmdlmc_b200.h"

#define CMD_MAX_IMAGES 26

// Box parameters as the kernels see them (passed by value as a __grid_constant__ argument).
// Mirrors the state of the reference's AtomBox family (PBCHelper.pyx:30-32,216-226,248-260).
struct BoxParams {
    int kind;        // 0 orthorhombic (AtomBoxCubic), 1 general cell (AtomBoxMonoclinic)
    int conv;        // CMD_CONV_*
    int sparse;      // structural zeros of h and hinv (pbc.cuh cmd_box_sparsity); set per topology
    double L[3];     // periodic_boundaries_extended (ortho)
    double hL[3];    // L/2
    double h[9];     // row-major, columns = cell vectors (extended)
    double hinv[9];  // row-major inverse of h
    double conv_par[5];
    // images (i,j,k) != 0 that can beat the fractionally wrapped vector by more than rounding
    // noise somewhere in the wrap cube; used by the FAST filter only (the exact stage always
    // walks all 27 images in the reference order, numpyatom.pyx:111-123).
    int n_img;
    double img[CMD_MAX_IMAGES][3];  // Cartesian shift i*a + j*b + k*c
    int img_ijk[CMD_MAX_IMAGES][3]; // the (i, j, k) of each kept image
};

struct cmd_box {
    BoxParams p;
    int n_values;             // 3 or 9
    double pbc[9];            // periodic_boundaries as given
    double pbc_extended[9];   // periodic_boundaries_extended
    double pbc_matrix[9];     // rows = cell vectors (NOT extended, PBCHelper.pyx:219-222,260)
    int mult[3];
};

struct CmdGlobal {
    bool inited = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = 0;
    // second stream for host->device copies that overlap kernels (cmd_topo_build)
    cudaStream_t copy_stream = 0;
    cudaEvent_t copy_event[2] = {0, 0};
    // every other chunk of a host block runs on this stream: the CTAs of chunk i+1 move onto the SMs
    // as those of chunk i drain (cmd_topo_build)
    cudaStream_t aux_stream = 0;
    cudaEvent_t aux_event[2] = {0, 0};
    // read-backs of a block that is complete while later blocks are still queued on `stream`
    // (cmd_topo_build_async / cmd_topo_wait / cmd_topo_frame_info)
    cudaStream_t ctl_stream = 0;
    int64_t launches = 0;
    // stream-ordered scratch for the host-pointer entry points
    void *scratch[6] = {0, 0, 0, 0, 0, 0};
    size_t scratch_bytes[6] = {0, 0, 0, 0, 0, 0};
};

CmdGlobal &cmd_global();
int cmd_set_error(int code, const char *fmt, ...);
int cmd_scratch(int slot, size_t bytes, void **out);  // grows slot to >= bytes
// staging.cu: dst (device) <- src (any host pointer), ordered on `stream`; pageable sources pass
// through the page-locked ring
int cmd_h2d_staged(void *dst, const void *src, size_t bytes, cudaStream_t stream);
void cmd_staging_shutdown();
// Rebuilds p.img / p.img_ijk: the periodic images a pair filter with radius `rc` has to look at
// besides the fractionally wrapped vector (rc < 0: no radius, every image that can beat it).
void cmd_box_prune_images(BoxParams &p, double rc);

#define CMD_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return cmd_set_error(CMD_ECUDA, "%s failed: %s (%s:%d)", #expr,                 \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)

#define CMD_REQUIRE_INIT()                                                                  \
    do {                                                                                    \
        if (!cmd_global().inited)                                                           \
            return cmd_set_error(CMD_ENODEV, "cmd_init() has not succeeded: no CUDA device " \
                                             "bound (there is no CPU fallback)");           \
    } while (0)

// CMDLMC_B200_SYNC_LAUNCHES=1 (debugging): wait for every kernel, so that a device fault is reported
// at the launch that caused it
inline bool cmd_sync_launches()
{
    static const bool on = getenv("CMDLMC_B200_SYNC_LAUNCHES") != nullptr;
    return on;
}

#define CMD_LAUNCHED()                                                                      \
    do {                                                                                    \
        cmd_global().launches++;                                                            \
        CMD_CUDA(cudaGetLastError());                                                       \
        if (cmd_sync_launches()) CMD_CUDA(cudaStreamSynchronize(cmd_global().stream));      \
    } while (0)

// -DCMD_BOUNDS_CHECK: index checks inside the kernels that address shared memory and scratch lists
// by computed positions (the pool's compute-sanitizer is closed: profiles/r2h_sanitizer_closed.txt);
// a violated check traps, which the next API call reports as a CUDA error.
#ifdef CMD_BOUNDS_CHECK
#define CMD_CHECK(cond) do { if (!(cond)) { printf("CMD_CHECK failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define CMD_CHECK(cond) do { } while (0)
#endif

// row pitch (ints) of the per-frame row index: n + 1 entries padded to 16 bytes (TMA bulk copies)
__host__ __device__ inline int cmd_ro_pitch(int n) { return (n + 4) & ~3; }

static inline int cmd_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
