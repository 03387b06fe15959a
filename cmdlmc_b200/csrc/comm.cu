// comm.cu -- the one collective of the hot path: sums of the jump statistics, MSD sums and
// histograms over the GPUs of a box (the quantities of
// mdlmc/LMC/output.py:17-49 and the jumpstat histograms), and the all-gather of per-frame step
// lengths that lets every rank walk the Verlet rebuild schedule of the WHOLE trajectory
// (mdlmc/topo/topology.py:96-107) while uploading only its own frame block.
//
// One process per GPU.  NCCL (over NVLink / NVSwitch) is bound at run time with dlopen -- the
// library has no link-time dependency on it, a single-GPU caller never loads it.  The unique id is
// created on rank 0 (cmd_comm_unique_id) and carried to the other ranks by whatever the host
// already has (torch.distributed, MPI, a file); cmd_comm_init then builds the communicator on
// the library's device and stream.
#include <dlfcn.h>
#include <stdlib.h>

#include "../../include/cmdlmc_b200.h"
#include "common.cuh"

namespace {

// the slice of nccl.h this file uses (stable since NCCL 2.0)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt64 = 4, ncclFloat64 = 8, ncclUint8 = 1 };   // ncclDataType_t values
enum { ncclSum = 0 };                                       // ncclRedOp_t

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};

NcclApi g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_world = 1;

int nccl_load()
{
    if (g_nccl.handle) return CMD_OK;
    // a libnccl the process already holds (PyTorch's bundled copy has the same soname) is reused
    const char *names[] = {getenv("CMDLMC_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return cmd_set_error(CMD_ENODEV, "libnccl.so.2 not found (set CMDLMC_B200_NCCL_LIB): %s", dlerror());
#define NCCL_SYM(field, name)                                                            \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                          \
    if (!g_nccl.field) { dlclose(h); return cmd_set_error(CMD_ENODEV, "%s missing from libnccl", name); }
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    NCCL_SYM(CommInitRank, "ncclCommInitRank");
    NCCL_SYM(CommDestroy, "ncclCommDestroy");
    NCCL_SYM(AllReduce, "ncclAllReduce");
    NCCL_SYM(AllGather, "ncclAllGather");
    NCCL_SYM(GetErrorString, "ncclGetErrorString");
    NCCL_SYM(GetVersion, "ncclGetVersion");
#undef NCCL_SYM
    g_nccl.handle = h;
    return CMD_OK;
}

#define CMD_NCCL(expr)                                                                   \
    do {                                                                                 \
        int _r = (expr);                                                                 \
        if (_r != ncclSuccess)                                                           \
            return cmd_set_error(CMD_ECUDA, "%s failed: %s", #expr, g_nccl.GetErrorString(_r)); \
    } while (0)

// sum of the per-frame pair counts and rate sums of a block, in frame order (deterministic)
__global__ void k_block_stats(const int *__restrict__ counts, const double *__restrict__ rate_sum,
                              int64_t nframes, double *__restrict__ out)
{
    __shared__ double s_cnt[256], s_rate[256];
    // thread t sums the frames of its contiguous slice, the slices are combined by a fixed tree
    const int64_t per = (nframes + 255) / 256;
    const int64_t lo = threadIdx.x * per, hi = lo + per < nframes ? lo + per : nframes;
    double c = 0.0, r = 0.0;
    for (int64_t f = lo; f < hi; f++) {
        const int k = counts[f];
        c += k < 0 ? 0.0 : (double)k;
        r += rate_sum[f];
    }
    s_cnt[threadIdx.x] = c;
    s_rate[threadIdx.x] = r;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            s_cnt[threadIdx.x] += s_cnt[threadIdx.x + s];
            s_rate[threadIdx.x] += s_rate[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] += s_cnt[0]; out[1] += s_rate[0]; }
}

}  // namespace

extern "C" int cmd_comm_unique_id(unsigned char h_id[128])
{
    if (!h_id) return cmd_set_error(CMD_EINVAL, "bad argument");
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    CMD_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(h_id, id.internal, 128);
    return CMD_OK;
}

extern "C" int cmd_comm_init(int rank, int world, const unsigned char h_id[128])
{
    CMD_REQUIRE_INIT();
    if (world < 1 || rank < 0 || rank >= world) return cmd_set_error(CMD_EINVAL, "bad rank / world");
    if (g_comm) return cmd_set_error(CMD_ESTATE, "the communicator exists already");
    g_rank = rank;
    g_world = world;
    if (world == 1) return CMD_OK;   // nothing to talk to: the collectives are identities
    if (!h_id) return cmd_set_error(CMD_EINVAL, "bad argument");
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(id.internal, h_id, 128);
    CMD_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
    return CMD_OK;
}

extern "C" int cmd_comm_destroy(void)
{
    if (g_comm) {
        cudaStreamSynchronize(cmd_global().stream);
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    g_rank = 0;
    g_world = 1;
    return CMD_OK;
}

extern "C" int cmd_comm_rank(void) { return g_rank; }
extern "C" int cmd_comm_world(void) { return g_world; }

extern "C" int cmd_comm_nccl_version(void)
{
    if (nccl_load()) return -1;
    int v = 0;
    return g_nccl.GetVersion(&v) == ncclSuccess ? v : -1;
}

extern "C" int cmd_stats_allreduce_dev(double *d_f64, int64_t n_f64, int64_t *d_i64, int64_t n_i64)
{
    CMD_REQUIRE_INIT();
    if (n_f64 < 0 || n_i64 < 0 || (n_f64 && !d_f64) || (n_i64 && !d_i64))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    if (g_world == 1) return CMD_OK;
    if (!g_comm) return cmd_set_error(CMD_ESTATE, "cmd_comm_init has not been called");
    cudaStream_t st = cmd_global().stream;
    if (n_f64) CMD_NCCL(g_nccl.AllReduce(d_f64, d_f64, (size_t)n_f64, ncclFloat64, ncclSum, g_comm, st));
    if (n_i64) CMD_NCCL(g_nccl.AllReduce(d_i64, d_i64, (size_t)n_i64, ncclInt64, ncclSum, g_comm, st));
    return CMD_OK;
}

extern "C" int cmd_stats_allreduce(double *h_f64, int64_t n_f64, int64_t *h_i64, int64_t n_i64)
{
    CMD_REQUIRE_INIT();
    if (n_f64 < 0 || n_i64 < 0 || (n_f64 && !h_f64) || (n_i64 && !h_i64))
        return cmd_set_error(CMD_EINVAL, "bad argument");
    if (g_world == 1) return CMD_OK;
    cudaStream_t st = cmd_global().stream;
    void *buf;
    int rc = cmd_scratch(3, (size_t)(n_f64 + n_i64) * 8, &buf);
    if (rc) return rc;
    double *d_f = (double *)buf;
    int64_t *d_i = (int64_t *)buf + n_f64;
    if (n_f64) CMD_CUDA(cudaMemcpyAsync(d_f, h_f64, (size_t)n_f64 * 8, cudaMemcpyHostToDevice, st));
    if (n_i64) CMD_CUDA(cudaMemcpyAsync(d_i, h_i64, (size_t)n_i64 * 8, cudaMemcpyHostToDevice, st));
    rc = cmd_stats_allreduce_dev(n_f64 ? d_f : nullptr, n_f64, n_i64 ? d_i : nullptr, n_i64);
    if (rc) return rc;
    if (n_f64) CMD_CUDA(cudaMemcpyAsync(h_f64, d_f, (size_t)n_f64 * 8, cudaMemcpyDeviceToHost, st));
    if (n_i64) CMD_CUDA(cudaMemcpyAsync(h_i64, d_i, (size_t)n_i64 * 8, cudaMemcpyDeviceToHost, st));
    CMD_CUDA(cudaStreamSynchronize(st));
    return CMD_OK;
}

extern "C" int cmd_allgather_dev(const void *d_send, void *d_recv, int64_t bytes_per_rank)
{
    CMD_REQUIRE_INIT();
    if (!d_send || !d_recv || bytes_per_rank < 0) return cmd_set_error(CMD_EINVAL, "bad argument");
    cudaStream_t st = cmd_global().stream;
    if (g_world == 1) {
        if (d_send != d_recv)
            CMD_CUDA(cudaMemcpyAsync(d_recv, d_send, (size_t)bytes_per_rank, cudaMemcpyDeviceToDevice, st));
        return CMD_OK;
    }
    if (!g_comm) return cmd_set_error(CMD_ESTATE, "cmd_comm_init has not been called");
    CMD_NCCL(g_nccl.AllGather(d_send, d_recv, (size_t)bytes_per_rank, ncclUint8, g_comm, st));
    return CMD_OK;
}

// accumulates (listed directed pairs, sum of the listed rates) of a block into d_out[0..1]
int cmd_block_stats_launch(const int *d_counts, const double *d_rate_sum, int64_t nframes, double *d_out)
{
    k_block_stats<<<1, 256, 0, cmd_global().stream>>>(d_counts, d_rate_sum, nframes, d_out);
    CMD_LAUNCHED();
    return CMD_OK;
}
