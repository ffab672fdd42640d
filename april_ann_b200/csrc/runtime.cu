// Runtime: per-device context, stream-ordered caching pool, copies, events, errors.
// Replaces GPUHelper (mathcore/c_src/gpu_helper.h:42-148) and the device side of
// GPUMirroredMemoryBlock (mathcore/c_src/gpu_mirrored_memory_block.h:178-304): the
// reference cuMemAlloc/cuMemFree's per activation per step; here blocks are recycled.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void b200_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int b200_check_cuda(cudaError_t e, const char *what, const char *file, int line) {
  if (e == cudaSuccess) return B200_OK;
  b200_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return B200_ERR_CUDA;
}

extern "C" const char *b200_last_error_string(void) { return g_err; }

// One host thread may drive several contexts (devices): launches, allocations and function attributes
// apply to the CURRENT device, so every entry point switches to its context's device when it differs.
int b200_make_current(b200_ctx *ctx) {
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess && cur == ctx->device) return B200_OK;
  return b200_check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice", __FILE__, __LINE__);
}

extern "C" int b200_device_count(int *count) {
  ARG_CHECK(count, "count is NULL");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    return b200_check_cuda(e, "cudaGetDeviceCount", __FILE__, __LINE__);
  }
  return B200_OK;
}

extern "C" int b200_create(int device, b200_ctx **out) {
  ARG_CHECK(out, "out is NULL");
  *out = nullptr;
  int n = 0;
  CUDA_TRY(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) {
    b200_set_error("b200_create: device %d not present (%d devices); this build has no CPU fallback",
                   device, n);
    return B200_ERR_CUDA;
  }
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    b200_set_error("b200_create: device %d is sm_%d%d; this library is built for sm_100a only",
                   device, prop.major, prop.minor);
    return B200_ERR_UNSUPPORTED;
  }
  b200_ctx *ctx = new b200_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  // the main stream carries the critical path of a step (forward, loss, data gradients): when its kernels
  // and a side branch's are ready together, the block scheduler serves the main stream first
  int prio_least = 0, prio_greatest = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  CUDA_TRY(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest));
  ctx->main_stream = ctx->stream;
  for (int i = 0; i < b200_ctx::kBranches; ++i) {
    // branch 2 (the big weight-gradient contractions) outranks branches 0 / 1 (updates, light gradients)
    const int prio = (i == 2 && prio_greatest < prio_least - 1) ? prio_greatest + 1 : prio_least;
    CUDA_TRY(cudaStreamCreateWithPriority(&ctx->side_stream[i], cudaStreamNonBlocking, prio));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fork[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_side[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_compute, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
  for (auto &e : ctx->ev_bucket) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : ctx->ev_fence) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *out = ctx;
  return B200_OK;
}

extern "C" int b200_pool_trim(b200_ctx *ctx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaStreamSynchronize(ctx->main_stream));
  for (auto &kv : ctx->free_blocks) cudaFree(kv.second);
  ctx->free_blocks.clear();
  return B200_OK;
}

extern "C" int b200_comm_destroy(b200_ctx *ctx);

extern "C" int b200_destroy(b200_ctx *ctx) {
  B200_ENTER(ctx);
  if (!ctx) return B200_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  b200_comm_destroy(ctx);
  gemm_tc_destroy(ctx);
  b200_pool_trim(ctx);
  for (auto &kv : ctx->live_blocks) cudaFree(kv.first);
  for (auto &p : ctx->scratch) if (p) cudaFree(p);
  for (void *p : ctx->retired_scratch) cudaFree(p);
  for (auto &kv : ctx->deferred_free) cudaFree(kv.second);
  for (int i = 0; i < b200_ctx::kBranches; ++i) {
    cudaEventDestroy(ctx->ev_fork[i]);
    cudaEventDestroy(ctx->ev_side[i]);
    cudaStreamDestroy(ctx->side_stream[i]);
  }
  cudaEventDestroy(ctx->ev_compute);
  cudaEventDestroy(ctx->ev_comm);
  for (auto &e : ctx->ev_bucket) if (e) cudaEventDestroy(e);
  for (auto &e : ctx->ev_fence) if (e) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->main_stream);
  cudaStreamDestroy(ctx->comm_stream);
  delete ctx;
  return B200_OK;
}

extern "C" int b200_set_math_mode(b200_ctx *ctx, int mode) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  ARG_CHECK(mode == B200_MATH_FP32 || mode == B200_MATH_TF32, "unknown math mode");
  ctx->math_mode = mode;
  return B200_OK;
}
extern "C" int b200_get_math_mode(b200_ctx *ctx, int *mode) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && mode, "NULL");
  *mode = ctx->math_mode;
  return B200_OK;
}
extern "C" int b200_sm_count(b200_ctx *ctx, int *count) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && count, "NULL");
  *count = ctx->sm_count;
  return B200_OK;
}
extern "C" void *b200_stream(b200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int b200_sync(b200_ctx *ctx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  CUDA_TRY(cudaStreamSynchronize(ctx->main_stream));
  for (int i = 0; i < b200_ctx::kBranches; ++i) CUDA_TRY(cudaStreamSynchronize(ctx->side_stream[i]));
  CUDA_TRY(cudaStreamSynchronize(ctx->comm_stream));
  return B200_OK;
}

// ---------------------------------------------------------------- side branches
// b200_branch_begin(i): everything issued until b200_branch_end goes to side stream i, ordered after
// what the main stream holds at this point.  b200_branch_wait(i): the current stream waits for what
// side stream i holds.  b200_branch_join_all: the main stream waits for every open branch.  Inside
// a stream capture these calls become the fork / join edges of the graph.
extern "C" int b200_branch_begin(b200_ctx *ctx, int branch) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  ARG_CHECK(branch >= 0 && branch < b200_ctx::kBranches, "no such branch");
  ARG_CHECK(ctx->cur_branch < 0, "branches do not nest");
  CUDA_TRY(cudaEventRecord(ctx->ev_fork[branch], ctx->main_stream));
  CUDA_TRY(cudaStreamWaitEvent(ctx->side_stream[branch], ctx->ev_fork[branch], 0));
  ctx->side_open[branch] = true;
  ctx->cur_branch = branch;
  ctx->stream = ctx->side_stream[branch];
  return B200_OK;
}
extern "C" int b200_branch_end(b200_ctx *ctx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  ctx->cur_branch = -1;
  ctx->stream = ctx->main_stream;
  return B200_OK;
}
extern "C" int b200_branch_wait(b200_ctx *ctx, int branch) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  ARG_CHECK(branch >= 0 && branch < b200_ctx::kBranches, "no such branch");
  if (!ctx->side_open[branch] || ctx->cur_branch == branch) return B200_OK;
  CUDA_TRY(cudaEventRecord(ctx->ev_side[branch], ctx->side_stream[branch]));
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_side[branch], 0));
  return B200_OK;
}
extern "C" int b200_branch_join_all(b200_ctx *ctx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  ctx->cur_branch = -1;
  ctx->stream = ctx->main_stream;
  for (int i = 0; i < b200_ctx::kBranches; ++i) {
    if (!ctx->side_open[i]) continue;
    CUDA_TRY(cudaEventRecord(ctx->ev_side[i], ctx->side_stream[i]));
    CUDA_TRY(cudaStreamWaitEvent(ctx->main_stream, ctx->ev_side[i], 0));
    ctx->side_open[i] = false;
  }
  std::lock_guard<std::mutex> lk(ctx->mu);
  for (auto &kv : ctx->deferred_free) ctx->free_blocks.emplace(kv.first, kv.second);
  ctx->deferred_free.clear();
  return B200_OK;
}
// A point-to-point edge between the streams of a step: b200_fence_record(id) marks what the CURRENT stream (main or
// the open branch) holds now, b200_fence_wait(id) makes the current stream wait for exactly that -- unlike
// b200_branch_wait, which waits for everything the branch holds at the time of the call.
extern "C" int b200_fence_record(b200_ctx *ctx, int id) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && id >= 0 && id < 8, "fence id in [0, 8)");
  CUDA_TRY(cudaEventRecord(ctx->ev_fence[id], ctx->stream));
  return B200_OK;
}
extern "C" int b200_fence_wait(b200_ctx *ctx, int id) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && id >= 0 && id < 8, "fence id in [0, 8)");
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_fence[id], 0));
  return B200_OK;
}
extern "C" int b200_set_sm_budget(b200_ctx *ctx, int sms) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  ctx->sm_budget = (sms > 0 && sms < ctx->sm_count) ? sms : 0;
  return B200_OK;
}

static size_t round_size(size_t bytes) {
  if (bytes < 512) return 512;
  if (bytes < (1u << 20)) return (bytes + 511) & ~size_t(511);
  return (bytes + (1u << 20) - 1) & ~size_t((1u << 20) - 1);  // 1 MiB granules for big blocks
}

extern "C" int b200_malloc(b200_ctx *ctx, void **dptr, size_t bytes) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && dptr, "NULL");
  size_t sz = round_size(bytes ? bytes : 1);
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->free_blocks.lower_bound(sz);
  // reuse a cached block unless it wastes more than 2x (all work is on one stream, so
  // reuse is stream-ordered and needs no event)
  if (it != ctx->free_blocks.end() && it->first <= 2 * sz) {
    *dptr = it->second;
    ctx->live_blocks[*dptr] = it->first;
    ctx->free_blocks.erase(it);
    return B200_OK;
  }
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaError_t e = cudaMalloc(dptr, sz + 512);  // +512: TMA boxes may touch the tail of the last row
  if (e != cudaSuccess) {
    // give cached blocks back and retry once
    for (auto &kv : ctx->free_blocks) cudaFree(kv.second);
    ctx->free_blocks.clear();
    (void)cudaGetLastError();
    e = cudaMalloc(dptr, sz + 512);
  }
  if (e != cudaSuccess) {
    b200_set_error("b200_malloc: cudaMalloc(%zu) failed: %s", sz, cudaGetErrorString(e));
    (void)cudaGetLastError();
    *dptr = nullptr;
    return B200_ERR_ALLOC;
  }
  ctx->live_blocks[*dptr] = sz;
  return B200_OK;
}

extern "C" int b200_free(b200_ctx *ctx, void *dptr) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  if (!dptr) return B200_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->live_blocks.find(dptr);
  ARG_CHECK(it != ctx->live_blocks.end(), "pointer was not allocated by b200_malloc on this context");
  bool open = false;
  for (bool o : ctx->side_open) open = open || o;
  // while side branches are open a block may still be in use on another stream: recycle it at the join
  if (open) ctx->deferred_free.emplace_back(it->second, dptr);
  else ctx->free_blocks.emplace(it->second, dptr);
  ctx->live_blocks.erase(it);
  return B200_OK;
}

void *b200_scratch(b200_ctx *ctx, size_t bytes) {
  const int slot = ctx->cur_branch + 1;
  if (bytes <= ctx->scratch_bytes[slot]) return ctx->scratch[slot];
  // Growing the scratch: kernels in flight and, above all, already CAPTURED graphs (split reductions, column
  // sums, convolution weight gradients) hold the old pointer, and graphs are cached per bunch size -- a small
  // bunch captured first, a larger one afterwards, then the first graph replayed.  The old block is therefore
  // retired, not freed: it stays allocated until b200_destroy.  (Scratch blocks are small and grow
  // geometrically in practice: a handful of retired blocks per context.)
  if (ctx->scratch[slot]) ctx->retired_scratch.push_back(ctx->scratch[slot]);
  size_t sz = round_size(bytes);
  if (cudaMalloc(&ctx->scratch[slot], sz) != cudaSuccess) {
    ctx->scratch[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
    (void)cudaGetLastError();
    return nullptr;
  }
  ctx->scratch_bytes[slot] = sz;
  return ctx->scratch[slot];
}

extern "C" int b200_host_alloc(void **hptr, size_t bytes) {
  ARG_CHECK(hptr, "NULL");
  CUDA_TRY(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return B200_OK;
}
extern "C" int b200_host_free(void *hptr) {
  if (hptr) CUDA_TRY(cudaFreeHost(hptr));
  return B200_OK;
}
extern "C" int b200_memcpy_h2d(b200_ctx *ctx, void *dst, const void *src, size_t bytes) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  if (bytes) CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return B200_OK;
}
extern "C" int b200_memcpy_d2h(b200_ctx *ctx, void *dst, const void *src, size_t bytes) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  if (bytes) CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return B200_OK;
}
extern "C" int b200_memcpy_d2d(b200_ctx *ctx, void *dst, const void *src, size_t bytes) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  if (bytes) CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return B200_OK;
}
extern "C" int b200_memset_zero(b200_ctx *ctx, void *dst, size_t bytes) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "ctx is NULL");
  if (bytes) CUDA_TRY(cudaMemsetAsync(dst, 0, bytes, ctx->stream));
  return B200_OK;
}

extern "C" int b200_event_create(void **ev) {
  ARG_CHECK(ev, "NULL");
  cudaEvent_t e;
  CUDA_TRY(cudaEventCreate(&e));
  *ev = (void *)e;
  return B200_OK;
}
extern "C" int b200_event_destroy(void *ev) {
  if (ev) CUDA_TRY(cudaEventDestroy((cudaEvent_t)ev));
  return B200_OK;
}
extern "C" int b200_event_record(b200_ctx *ctx, void *ev) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && ev, "NULL");
  CUDA_TRY(cudaEventRecord((cudaEvent_t)ev, ctx->stream));
  return B200_OK;
}
extern "C" int b200_event_elapsed_ms(void *start, void *stop, float *ms) {
  ARG_CHECK(start && stop && ms, "NULL");
  CUDA_TRY(cudaEventSynchronize((cudaEvent_t)stop));
  CUDA_TRY(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return B200_OK;
}
extern "C" int b200_launch_count(b200_ctx *ctx, uint64_t *count) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && count, "NULL");
  *count = ctx->launches;
  return B200_OK;
}
extern "C" int b200_add_launches(b200_ctx *ctx, uint64_t n) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "NULL");
  ctx->launches += n;
  return B200_OK;
}
