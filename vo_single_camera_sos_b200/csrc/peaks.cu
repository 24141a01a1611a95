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

// packed FP32 FMA (fma.rn.f32x2, SASS FFMA2): two FMAs per instruction — the form the RANSAC score kernel issues
__global__ void __launch_bounds__(256) ffma2_peak_kernel(float a, float b, float* out) {
  unsigned long long x[PK_ILP];
  const unsigned long long av = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
  const unsigned long long bv = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) {
    const float v = a * (float)(threadIdx.x + i);
    x[i] = ((unsigned long long)__float_as_uint(v) << 32) | __float_as_uint(v + 1.f);
  }
  for (int it = 0; it < PK_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(av), "l"(bv));
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) s ^= x[i];
  if (s == 0x123456789ABCDEFull) out[0] = 1.f;
}

// Tensor-memory read bandwidth (tcgen05.ld 32x32b.x32 from all four lane quarters, two warps per quarter): the bound of the
// tensor-core Hamming engine's epilogue.  One CTA per SM (the CTA owns all 512 TMEM columns); the contents are not initialised.
constexpr int TM_ITERS = 256;
__global__ void __launch_bounds__(256, 1) tmem_read_peak_kernel(uint32_t* out) {
  extern __shared__ __align__(16) uint8_t dyn[];   // sized by the host so that only one CTA fits on an SM
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 256u;
  uint32_t acc = dyn[0];
  for (int it = 0; it < TM_ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
          "tcgen05.wait::ld.sync.aligned;"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(base + (uint32_t)c * 32u)
          : "memory");
      acc ^= v[0] ^ v[31];
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
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

extern "C" int sos_peak_ffma2(sos_ctx* ctx, double* tflops) {
  SOS_CHECK_ARG(ctx && tflops, "NULL argument");
  SOS_CUDA(cudaSetDevice(ctx->device));
  void* scratch;
  int rc = sos_arena_get(ctx, 256, &scratch);
  if (rc != SOS_OK) return rc;
  const int blocks = ctx->sm_count * 16;
  float ms;
  rc = time_kernel(ctx, [&] { ffma2_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(1.0000001f, 0.5f, (float*)scratch); ctx->launches++; }, &ms);
  if (rc != SOS_OK) return rc;
  SOS_CUDA(cudaGetLastError());
  *tflops = 4.0 * (double)blocks * 256.0 * PK_ITERS * PK_ILP / (ms * 1e-3) / 1e12;
  return SOS_OK;
}

extern "C" int sos_peak_tmem_read(sos_ctx* ctx, double* tera_bytes_per_s) {
  SOS_CHECK_ARG(ctx && tera_bytes_per_s, "NULL argument");
  SOS_CUDA(cudaSetDevice(ctx->device));
  void* scratch;
  int rc = sos_arena_get(ctx, 256, &scratch);
  if (rc != SOS_OK) return rc;
  const int smem = 120 * 1024;   // more than half an SM's shared memory: one CTA per SM
  SOS_CUDA(cudaFuncSetAttribute(tmem_read_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int blocks = ctx->sm_count;
  float ms;
  rc = time_kernel(ctx, [&] { tmem_read_peak_kernel<<<blocks, 256, smem, ctx->stream>>>((uint32_t*)scratch); ctx->launches++; }, &ms);
  if (rc != SOS_OK) return rc;
  SOS_CUDA(cudaGetLastError());
  *tera_bytes_per_s = (double)blocks * 256.0 * TM_ITERS * 8.0 * 32.0 * 4.0 / (ms * 1e-3) / 1e12;
  return SOS_OK;
}
