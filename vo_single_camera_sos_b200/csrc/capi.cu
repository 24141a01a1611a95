// Context / memory half of the C-ABI (include/sosfront.h).  No reference counterpart: the reference is
// single-process NumPy; this is the plumbing a host needs to drive the kernels without PyTorch.
#include <stdarg.h>

#include "sos_common.cuh"

static thread_local char g_err[512] = "";

void sos_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sos_arena_get(sos_ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->arena_bytes) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
      sos_set_error("scratch arena too small (%zu < %zu bytes) while the stream is being captured; call sos_ctx_reserve first",
                    ctx->arena_bytes, bytes);
      return SOS_ERR_CAPTURE;
    }
    size_t want = sos_align_up(bytes + bytes / 4, (size_t)1 << 20);
    SOS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->arena) SOS_CUDA(cudaFree(ctx->arena));
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
    cudaError_t e = cudaMalloc(&ctx->arena, want);
    if (e != cudaSuccess) {
      sos_set_error("scratch arena: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return SOS_ERR_NOMEM;
    }
    ctx->arena_bytes = want;
  }
  *out = ctx->arena;
  return SOS_OK;
}

void sos_prof_mark_launch(sos_ctx* ctx, const char* name) {
  sos_prof_mark m;
  m.name = name;
  if (cudaEventCreate(&m.ev) != cudaSuccess) return;
  cudaEventRecord(m.ev, ctx->stream);
  ctx->marks.push_back(m);
}

extern "C" {

int sos_ctx_profile_begin(sos_ctx* ctx) {
  SOS_CHECK_ARG(ctx, "ctx is NULL");
  for (auto& m : ctx->marks) cudaEventDestroy(m.ev);
  ctx->marks.clear();
  ctx->prof = true;
  sos_prof_mark_launch(ctx, "begin");
  return SOS_OK;
}

int sos_ctx_profile_end(sos_ctx* ctx, char* names, size_t names_cap, float* ms, int max_n, int* n_out) {
  SOS_CHECK_ARG(ctx && n_out, "NULL argument");
  ctx->prof = false;
  SOS_CUDA(cudaStreamSynchronize(ctx->stream));
  int n = 0;
  size_t used = 0;
  if (names && names_cap) names[0] = 0;
  for (size_t i = 1; i < ctx->marks.size() && n < max_n; ++i, ++n) {
    float t = 0.f;
    SOS_CUDA(cudaEventElapsedTime(&t, ctx->marks[i - 1].ev, ctx->marks[i].ev));
    if (ms) ms[n] = t;
    if (names) {
      const int w = snprintf(names + used, used < names_cap ? names_cap - used : 0, "%s\n", ctx->marks[i].name);
      if (w > 0 && used + (size_t)w < names_cap) used += (size_t)w;
    }
  }
  for (auto& m : ctx->marks) cudaEventDestroy(m.ev);
  ctx->marks.clear();
  *n_out = n;
  return SOS_OK;
}

int sos_abi_version(void) { return SOS_ABI_VERSION; }
const char* sos_last_error(void) { return g_err; }

int sos_device_count(int* count) {
  SOS_CHECK_ARG(count, "count is NULL");
  SOS_CUDA(cudaGetDeviceCount(count));
  return SOS_OK;
}

int sos_ctx_create(int device, sos_ctx** out) {
  SOS_CHECK_ARG(out, "out is NULL");
  int n = 0;
  SOS_CUDA(cudaGetDeviceCount(&n));
  SOS_CHECK_ARG(device >= 0 && device < n, "device out of range");
  SOS_CUDA(cudaSetDevice(device));
  sos_ctx* c = new sos_ctx();
  c->device = device;
  cudaDeviceProp prop;
  SOS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    delete c;
    sos_set_error("sos_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return SOS_ERR_UNSUPPORTED;
  }
  c->sm_count = prop.multiProcessorCount;
  SOS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  *out = c;
  return SOS_OK;
}

int sos_ctx_destroy(sos_ctx* ctx) {
  if (!ctx) return SOS_OK;
  cudaSetDevice(ctx->device);
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SOS_OK;
}

int sos_ctx_set_stream(sos_ctx* ctx, void* cuda_stream) {
  SOS_CHECK_ARG(ctx, "ctx is NULL");
  SOS_CUDA(cudaSetDevice(ctx->device));
  if (ctx->own_stream && ctx->stream) {
    SOS_CUDA(cudaStreamSynchronize(ctx->stream));
    SOS_CUDA(cudaStreamDestroy(ctx->stream));
    ctx->own_stream = false;
  }
  ctx->stream = (cudaStream_t)cuda_stream;  // NULL = the legacy default stream
  return SOS_OK;
}

void* sos_ctx_get_stream(sos_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int sos_ctx_sync(sos_ctx* ctx) {
  SOS_CHECK_ARG(ctx, "ctx is NULL");
  SOS_CUDA(cudaStreamSynchronize(ctx->stream));
  return SOS_OK;
}

int sos_ctx_reserve(sos_ctx* ctx, size_t bytes) {
  SOS_CHECK_ARG(ctx, "ctx is NULL");
  SOS_CUDA(cudaSetDevice(ctx->device));
  void* p;
  return sos_arena_get(ctx, bytes, &p);
}

int64_t sos_ctx_launch_count(sos_ctx* ctx) { return ctx ? ctx->launches : 0; }

int sos_malloc(sos_ctx* ctx, size_t bytes, void** dptr) {
  SOS_CHECK_ARG(ctx && dptr, "NULL argument");
  SOS_CUDA(cudaSetDevice(ctx->device));
  cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    sos_set_error("sos_malloc(%zu): %s", bytes, cudaGetErrorString(e));
    return SOS_ERR_NOMEM;
  }
  return SOS_OK;
}

int sos_free(sos_ctx* ctx, void* dptr) {
  SOS_CHECK_ARG(ctx, "ctx is NULL");
  if (dptr) SOS_CUDA(cudaFree(dptr));
  return SOS_OK;
}

int sos_malloc_host(size_t bytes, void** hptr) {
  SOS_CHECK_ARG(hptr, "hptr is NULL");
  cudaError_t e = cudaMallocHost(hptr, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    sos_set_error("sos_malloc_host(%zu): %s", bytes, cudaGetErrorString(e));
    return SOS_ERR_NOMEM;
  }
  return SOS_OK;
}

int sos_free_host(void* hptr) {
  if (hptr) SOS_CUDA(cudaFreeHost(hptr));
  return SOS_OK;
}

int sos_memcpy_h2d(sos_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  SOS_CHECK_ARG(ctx && (bytes == 0 || (dst_dev && src_host)), "NULL argument");
  if (bytes) SOS_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return SOS_OK;
}

int sos_memcpy_d2h(sos_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  SOS_CHECK_ARG(ctx && (bytes == 0 || (dst_host && src_dev)), "NULL argument");
  if (bytes) SOS_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return SOS_OK;
}

int sos_memset(sos_ctx* ctx, void* dst_dev, int value, size_t bytes) {
  SOS_CHECK_ARG(ctx && (bytes == 0 || dst_dev), "NULL argument");
  if (bytes) SOS_CUDA(cudaMemsetAsync(dst_dev, value, bytes, ctx->stream));
  return SOS_OK;
}

}  // extern "C"
