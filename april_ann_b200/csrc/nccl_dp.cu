// Data-parallel plumbing: one NCCL communicator per context, gradient all-reduce on a
// communication stream ordered against the compute stream with events.  New relative to the
// reference, which hard-codes device 0 (mathcore/c_src/gpu_helper.h:65-68) and has no
// communication layer.  NCCL is dlopen'ed so the library loads on boxes without it.
#include <dlfcn.h>

#include "common.cuh"

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

NcclApi *nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char *names[] = {getenv("B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm) continue;
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    b200_set_error("NCCL not found (set B200_NCCL_LIB to libnccl.so.2): %s", dlerror());
    return nullptr;
  }
#define LOAD(field, sym) *(void **)(&api.field) = dlsym(api.handle, sym)
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(AllReduce, "ncclAllReduce");
  LOAD(Broadcast, "ncclBroadcast");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.Broadcast) {
    b200_set_error("NCCL library lacks required symbols");
    dlclose(api.handle);
    api.handle = nullptr;
    return nullptr;
  }
  return &api;
}

int nccl_check(int r, const char *what) {
  if (r == ncclSuccess) return B200_OK;
  NcclApi *a = nccl();
  b200_set_error("NCCL error %d in %s: %s", r, what, (a && a->GetErrorString) ? a->GetErrorString(r) : "?");
  return B200_ERR_NCCL;
}

// order the comm stream after everything enqueued on the compute stream, and back
int fence_in(b200_ctx *ctx) {
  CUDA_TRY(cudaEventRecord(ctx->ev_compute, ctx->stream));
  CUDA_TRY(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_compute, 0));
  // a bucket may hold gradients produced on several side branches of the step (and on the main stream)
  if (ctx->stream != ctx->main_stream) {
    CUDA_TRY(cudaEventRecord(ctx->ev_comm, ctx->main_stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_comm, 0));
  }
  for (int i = 0; i < b200_ctx::kBranches; ++i) {
    if (!ctx->side_open[i] || ctx->side_stream[i] == ctx->stream) continue;
    CUDA_TRY(cudaEventRecord(ctx->ev_side[i], ctx->side_stream[i]));
    CUDA_TRY(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_side[i], 0));
  }
  return B200_OK;
}
int fence_out(b200_ctx *ctx) {
  CUDA_TRY(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
  return B200_OK;
}

}  // namespace

extern "C" int b200_comm_unique_id(void *id128) {
  ARG_CHECK(id128, "NULL pointer");
  NcclApi *a = nccl();
  if (!a) return B200_ERR_NCCL;
  ncclUniqueId id;
  int st = nccl_check(a->GetUniqueId(&id), "ncclGetUniqueId");
  if (st) return st;
  memcpy(id128, &id, sizeof(id));
  return B200_OK;
}

extern "C" int b200_comm_init(b200_ctx *ctx, int nranks, int rank, const void *id128) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && id128, "NULL pointer");
  ARG_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank");
  NcclApi *a = nccl();
  if (!a) return B200_ERR_NCCL;
  CUDA_TRY(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  int st = nccl_check(a->CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
  if (st) return st;
  ctx->nccl_comm = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  return B200_OK;
}

extern "C" int b200_comm_destroy(b200_ctx *ctx) {
  B200_ENTER(ctx);
  if (!ctx || !ctx->nccl_comm) return B200_OK;
  NcclApi *a = nccl();
  if (a && a->CommDestroy) a->CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->nranks = 1;
  ctx->rank = 0;
  return B200_OK;
}

static int allreduce_impl(b200_ctx *ctx, void *buf, size_t n, int dtype) {
  ARG_CHECK(ctx && buf, "NULL pointer");
  if (ctx->nranks <= 1 || n == 0) return B200_OK;
  ARG_CHECK(ctx->nccl_comm, "communicator not initialised (b200_comm_init)");
  NcclApi *a = nccl();
  if (!a) return B200_ERR_NCCL;
  // Same-stream ordering keeps the whole step capturable in one CUDA graph; NCCL kernels
  // still overlap with nothing here -- bucketed overlap is driven by the trainer.
  int st = nccl_check(a->AllReduce(buf, buf, n, dtype, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream),
                      "ncclAllReduce");
  if (st) return st;
  ctx->launches++;
  return B200_OK;
}
extern "C" int b200_allreduce_sum_async(b200_ctx *ctx, float *buf, size_t n, int slot) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && buf, "NULL pointer");
  ARG_CHECK(slot >= 0 && slot < 16, "slot out of range");
  if (ctx->nranks <= 1 || n == 0) return B200_OK;
  ARG_CHECK(ctx->nccl_comm, "communicator not initialised (b200_comm_init)");
  NcclApi *a = nccl();
  if (!a) return B200_ERR_NCCL;
  int st = fence_in(ctx);   // the bucket's gradients are complete on the compute stream
  if (st) return st;
  st = nccl_check(a->AllReduce(buf, buf, n, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->comm_stream),
                  "ncclAllReduce");
  if (st) return st;
  ctx->launches++;
  CUDA_TRY(cudaEventRecord(ctx->ev_bucket[slot], ctx->comm_stream));
  return B200_OK;
}
extern "C" int b200_comm_wait(b200_ctx *ctx, int slot) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "NULL pointer");
  ARG_CHECK(slot >= 0 && slot < 16, "slot out of range");
  if (ctx->nranks <= 1) return B200_OK;
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_bucket[slot], 0));
  return B200_OK;
}
extern "C" int b200_allreduce_sum(b200_ctx *ctx, float *buf, size_t n) {
  B200_ENTER(ctx); return allreduce_impl(ctx, buf, n, ncclFloat32); }
extern "C" int b200_allreduce_sum_f64(b200_ctx *ctx, double *buf, size_t n) {
  B200_ENTER(ctx);
  return allreduce_impl(ctx, buf, n, ncclFloat64);
}
extern "C" int b200_broadcast(b200_ctx *ctx, float *buf, size_t n, int root) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && buf, "NULL pointer");
  if (ctx->nranks <= 1 || n == 0) return B200_OK;
  ARG_CHECK(ctx->nccl_comm, "communicator not initialised (b200_comm_init)");
  NcclApi *a = nccl();
  if (!a) return B200_ERR_NCCL;
  int st = fence_in(ctx);
  if (st) return st;
  st = nccl_check(a->Broadcast(buf, buf, n, ncclFloat32, root, (ncclComm_t)ctx->nccl_comm, ctx->comm_stream),
                  "ncclBroadcast");
  if (st) return st;
  return fence_out(ctx);
}
