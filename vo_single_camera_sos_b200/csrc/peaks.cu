// Roofline denominators that MEASURED_PEAKS.json does not carry: the integer POPC pipe (Hamming matching) and the
// FP32 FMA pipe (RANSAC scoring).  Pure register microbenchmarks, timed with CUDA events on the context stream.
#include "sos_common.cuh"

namespace {

constexpr int PK_ITERS = 4096;
constexpr int PK_ILP = 8;

__global__ void __launch_bounds__(256) popc_peak_kernel(uint32_t seed, uint32_t* out) {
  uint32_t x[PK_ILP], acc[PK_ILP];
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) {
    x[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u + blockIdx.x;
    acc[i] = 0;
  }
  for (int it = 0; it < PK_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) {
      acc[i] += __popc(x[i]);  // 1 POPC + 1 IADD
      x[i] ^= acc[i];          // keeps the chain data dependent so nothing folds away
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) s += acc[i];
  if (s == 0xFFFFFFFFu) out[0] = s;  // never true in practice; defeats dead-code elimination
}

__global__ void __launch_bounds__(256) ffma_peak_kernel(float a, float b, float* out) {
  float x[PK_ILP];
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) x[i] = a * (float)(threadIdx.x + i);
  for (int it = 0; it < PK_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) x[i] = __fmaf_rn(x[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) s += x[i];
  if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double a, double b, double* out) {
  double x[PK_ILP];
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) x[i] = a * (double)(threadIdx.x + i);
  for (int it = 0; it < PK_ITERS / 4; ++it) {
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
int time_kernel(sos_ctx* ctx, F launch, float* best_ms) {
  cudaEvent_t e0, e1;
  SOS_CUDA(cudaEventCreate(&e0));
  SOS_CUDA(cudaEventCreate(&e1));
  *best_ms = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    SOS_CUDA(cudaEventRecord(e0, ctx->stream));
    launch();
    SOS_CUDA(cudaEventRecord(e1, ctx->stream));
    SOS_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    SOS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < *best_ms) *best_ms = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return SOS_OK;
}

}  // namespace

extern "C" int sos_peak_popc(sos_ctx* ctx, double* tera_popc_per_s) {
  SOS_CHECK_ARG(ctx && tera_popc_per_s, "NULL argument");
  SOS_CUDA(cudaSetDevice(ctx->device));
  void* scratch;
  int rc = sos_arena_get(ctx, 256, &scratch);
  if (rc != SOS_OK) return rc;
  const int blocks = ctx->sm_count * 16;
  float ms;
  rc = time_kernel(ctx, [&] { popc_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(12345u, (uint32_t*)scratch); ctx->launches++; }, &ms);
  if (rc != SOS_OK) return rc;
  SOS_CUDA(cudaGetLastError());
  *tera_popc_per_s = (double)blocks * 256.0 * PK_ITERS * PK_ILP / (ms * 1e-3) / 1e12;
  return SOS_OK;
}

extern "C" int sos_peak_ffma(sos_ctx* ctx, double* tflops) {
  SOS_CHECK_ARG(ctx && tflops, "NULL argument");
  SOS_CUDA(cudaSetDevice(ctx->device));
  void* scratch;
  int rc = sos_arena_get(ctx, 256, &scratch);
  if (rc != SOS_OK) return rc;
  const int blocks = ctx->sm_count * 16;
  float ms;
  rc = time_kernel(ctx, [&] { ffma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(1.0000001f, 0.5f, (float*)scratch); ctx->launches++; }, &ms);
  if (rc != SOS_OK) return rc;
  SOS_CUDA(cudaGetLastError());
  *tflops = 2.0 * (double)blocks * 256.0 * PK_ITERS * PK_ILP / (ms * 1e-3) / 1e12;
  return SOS_OK;
}

extern "C" int sos_peak_dfma(sos_ctx* ctx, double* tflops) {
  SOS_CHECK_ARG(ctx && tflops, "NULL argument");
  SOS_CUDA(cudaSetDevice(ctx->device));
  void* scratch;
  int rc = sos_arena_get(ctx, 256, &scratch);
  if (rc != SOS_OK) return rc;
  const int blocks = ctx->sm_count * 16;
  float ms;
  rc = time_kernel(ctx, [&] { dfma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(1.0000001, 0.5, (double*)scratch); ctx->launches++; }, &ms);
  if (rc != SOS_OK) return rc;
  SOS_CUDA(cudaGetLastError());
  *tflops = 2.0 * (double)blocks * 256.0 * (PK_ITERS / 4) * PK_ILP / (ms * 1e-3) / 1e12;
  return SOS_OK;
}
