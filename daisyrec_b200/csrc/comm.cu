// comm.cu -- native NCCL path of the user-sharded multi-GPU step.
//
// The reference has no multi-device path.  Here one process drives one GPU; the per-step exchange
//   phase 1 (local triples) -> all-reduce {gQ fp32, cntI u64, acc fp64} -> phase 2 (local P rows + replicated Q)
// is enqueued by the library itself: the three reductions are ONE grouped NCCL launch (ncclGroupStart/End) on the same
// stream as the kernels, so a whole epoch segment is queued without a host round trip per step.
// NCCL is resolved at run time from the already-loaded libnccl.so.2 (the copy torch.distributed uses); only its
// public C API is used.  The communicator is bootstrapped by the host (rank 0's unique id is broadcast with
// torch.distributed, which stays plumbing).
#include <dlfcn.h>
#include <nccl.h>

#include "step.cuh"

namespace drb {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
    bool ok;
};

static NcclApi g_nccl = {};
static ncclComm_t g_comm = nullptr;
static int g_world = 0, g_rank = -1;

static int load_nccl()
{
    if (g_nccl.ok) return DRB_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    DRB_REQUIRE(h != nullptr, "libnccl.so.2 not found: %s", dlerror());
#define DRB_SYM(field, name)                                         \
    *(void **)(&g_nccl.field) = dlsym(h, name);                      \
    DRB_REQUIRE(g_nccl.field != nullptr, "NCCL symbol %s missing", name);
    DRB_SYM(GetUniqueId, "ncclGetUniqueId")
    DRB_SYM(CommInitRank, "ncclCommInitRank")
    DRB_SYM(CommDestroy, "ncclCommDestroy")
    DRB_SYM(AllReduce, "ncclAllReduce")
    DRB_SYM(GroupStart, "ncclGroupStart")
    DRB_SYM(GroupEnd, "ncclGroupEnd")
    DRB_SYM(GetErrorString, "ncclGetErrorString")
#undef DRB_SYM
    g_nccl.ok = true;
    return DRB_OK;
}

#define DRB_NCCL(call)                                                                        \
    do {                                                                                      \
        ncclResult_t _r = (call);                                                             \
        if (_r != ncclSuccess) {                                                              \
            set_error("NCCL error %d (%s) in %s", (int)_r, g_nccl.GetErrorString(_r), #call); \
            return DRB_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

}  // namespace drb

using namespace drb;

extern "C" int drb_comm_unique_id(uint8_t *out128)
{
    DRB_REQUIRE(out128 != nullptr, "comm_unique_id: null buffer");
    int rc = load_nccl();
    if (rc != DRB_OK) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    DRB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return DRB_OK;
}

extern "C" int drb_comm_init(const uint8_t *id128, int32_t rank, int32_t world)
{
    DRB_REQUIRE(id128 && world >= 1 && rank >= 0 && rank < world, "comm_init: bad arguments");
    int rc = load_nccl();
    if (rc != DRB_OK) return rc;
    if (g_comm) {
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    DRB_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
    g_world = world;
    g_rank = rank;
    return DRB_OK;
}

extern "C" int drb_comm_destroy(void)
{
    if (g_comm && g_nccl.ok) g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
    g_world = 0;
    g_rank = -1;
    return DRB_OK;
}

// One synchronous GLOBAL step on this rank's share [d_b* .. d_b* + count) of the global batch:
//   phase 1 (local triples) -> ONE grouped all-reduce of {gQ, cntI, acc} -> phase 2 (local P rows + the replicated Q).
static int sharded_step(float *d_P_local, float *d_Q, void *d_ws, int32_t U_local, int32_t I, int32_t F, const int32_t *d_bu,
                        const int32_t *d_bi, const int32_t *d_bj, int64_t count, const drb_hyper *hyper, int64_t adam_step,
                        double *d_loss, const int64_t *lay, cudaStream_t st)
{
    char *ws = (char *)d_ws;
    double *acc = (double *)(ws + lay[0]);
    float *gq = (float *)(ws + lay[2]);
    unsigned long long *cnt_i = (unsigned long long *)(ws + lay[4]);
    int rc = drb_mf_bpr_phase(d_P_local, d_Q, d_ws, U_local, I, F, d_bu, d_bi, d_bj, 0, count, 1, hyper, adam_step, d_loss,
                              (void *)st);
    if (rc != DRB_OK) return rc;
    DRB_NCCL(g_nccl.GroupStart());
    DRB_NCCL(g_nccl.AllReduce(gq, gq, (size_t)I * F, ncclFloat32, ncclSum, g_comm, st));
    DRB_NCCL(g_nccl.AllReduce(cnt_i, cnt_i, (size_t)I, ncclUint64, ncclSum, g_comm, st));
    DRB_NCCL(g_nccl.AllReduce(acc, acc, 8, ncclFloat64, ncclSum, g_comm, st));
    DRB_NCCL(g_nccl.GroupEnd());
    return drb_mf_bpr_phase(d_P_local, d_Q, d_ws, U_local, I, F, d_bu, d_bi, d_bj, 0, count, 2, hyper, adam_step, d_loss,
                            (void *)st);
}

// n_steps synchronous GLOBAL steps on this rank's shard: step s trains local triples [h_step_offsets[s], h_step_offsets[s+1]).
extern "C" int drb_mf_bpr_train_steps_sharded(float *d_P_local, float *d_Q, void *d_ws, int32_t U_local, int32_t I, int32_t F,
                                              const int32_t *d_bu, const int32_t *d_bi, const int32_t *d_bj,
                                              const int64_t *h_step_offsets, int64_t first_step, int64_t n_steps,
                                              const drb_hyper *hyper, int64_t adam_step0, double *d_step_loss, void *stream)
{
    DRB_REQUIRE(g_comm != nullptr, "train_steps_sharded: call drb_comm_init first");
    DRB_REQUIRE(d_P_local && d_Q && d_ws && d_bu && d_bi && d_bj && h_step_offsets && hyper && d_step_loss && n_steps >= 0,
                "train_steps_sharded: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t lay[8];
    int rc = drb_mf_workspace_layout(U_local, I, F, hyper->opt, lay);
    if (rc != DRB_OK) return rc;
    for (int64_t s = 0; s < n_steps; ++s) {
        const int64_t b = h_step_offsets[first_step + s], e = h_step_offsets[first_step + s + 1];
        rc = sharded_step(d_P_local, d_Q, d_ws, U_local, I, F, d_bu + b, d_bi + b, d_bj + b, e - b, hyper, adam_step0 + s,
                          d_step_loss + s, lay, st);
        if (rc != DRB_OK) return rc;
    }
    return DRB_OK;
}

// The same global steps fed from HOST (pinned) planes holding this rank's share of every global batch: the H2D copy of the
// share of step s+1 (copy stream) overlaps the kernels and the collective of step s; every step's global loss is read
// back to the host.  Mirrors drb_mf_bpr_train_steps_host for N > 1.  d_stage: 2 slots x 3 planes x stride int32, where
// stride = max local share rounded up to 4.
extern "C" int drb_mf_bpr_train_steps_sharded_host(float *d_P_local, float *d_Q, void *d_ws, int32_t U_local, int32_t I,
                                                   int32_t F, const int32_t *h_bu, const int32_t *h_bi, const int32_t *h_bj,
                                                   const int64_t *h_step_offsets, int64_t first_step, int64_t n_steps,
                                                   const drb_hyper *hyper, int64_t adam_step0, int32_t *d_stage,
                                                   int64_t stage_stride, double *d_step_loss, double *h_step_loss,
                                                   void *stream)
{
    DRB_REQUIRE(g_comm != nullptr, "train_steps_sharded_host: call drb_comm_init first");
    DRB_REQUIRE(d_P_local && d_Q && d_ws && h_bu && h_bi && h_bj && h_step_offsets && hyper && d_stage && d_step_loss &&
                    h_step_loss && n_steps >= 0 && stage_stride > 0 && stage_stride % 4 == 0,
                "train_steps_sharded_host: bad arguments");
    if (n_steps == 0) return DRB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t lay[8];
    int rc = drb_mf_workspace_layout(U_local, I, F, hyper->opt, lay);
    if (rc != DRB_OK) return rc;
    static thread_local cudaStream_t copy_st = nullptr;
    if (!copy_st) DRB_CUDA(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
    cudaEvent_t ready[2], freed[2], start;
    for (int k = 0; k < 2; ++k) {
        DRB_CUDA(cudaEventCreateWithFlags(&ready[k], cudaEventDisableTiming));
        DRB_CUDA(cudaEventCreateWithFlags(&freed[k], cudaEventDisableTiming));
    }
    DRB_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    DRB_CUDA(cudaEventRecord(start, st));
    DRB_CUDA(cudaStreamWaitEvent(copy_st, start, 0));
    for (int64_t s = 0; s < n_steps && rc == DRB_OK; ++s) {
        const int slot = (int)(s & 1);
        int32_t *sb = d_stage + (size_t)slot * 3 * stage_stride;
        const int64_t b = h_step_offsets[first_step + s], cnt = h_step_offsets[first_step + s + 1] - b;
        if (cnt > stage_stride) {
            set_error("train_steps_sharded_host: local share %lld exceeds the staging stride %lld", (long long)cnt,
                      (long long)stage_stride);
            rc = DRB_ERR_INVALID;
            break;
        }
        const size_t bytes = sizeof(int32_t) * (size_t)cnt;
        if (s >= 2) DRB_CUDA(cudaStreamWaitEvent(copy_st, freed[slot], 0));
        if (cnt > 0) {
            DRB_CUDA(cudaMemcpyAsync(sb, h_bu + b, bytes, cudaMemcpyHostToDevice, copy_st));
            DRB_CUDA(cudaMemcpyAsync(sb + stage_stride, h_bi + b, bytes, cudaMemcpyHostToDevice, copy_st));
            DRB_CUDA(cudaMemcpyAsync(sb + 2 * stage_stride, h_bj + b, bytes, cudaMemcpyHostToDevice, copy_st));
        }
        DRB_CUDA(cudaEventRecord(ready[slot], copy_st));
        DRB_CUDA(cudaStreamWaitEvent(st, ready[slot], 0));
        rc = sharded_step(d_P_local, d_Q, d_ws, U_local, I, F, sb, sb + stage_stride, sb + 2 * stage_stride, cnt, hyper,
                          adam_step0 + s, d_step_loss + s, lay, st);
        if (rc != DRB_OK) break;
        DRB_CUDA(cudaEventRecord(freed[slot], st));
        DRB_CUDA(cudaMemcpyAsync(h_step_loss + s, d_step_loss + s, sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    cudaStreamSynchronize(copy_st);
    if (rc == DRB_OK) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    }
    for (int k = 0; k < 2; ++k) {
        cudaEventDestroy(ready[k]);
        cudaEventDestroy(freed[k]);
    }
    cudaEventDestroy(start);
    return rc;
}
