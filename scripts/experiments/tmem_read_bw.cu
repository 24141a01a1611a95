// Experiment: sustained tcgen05.ld throughput per SM (bytes / clock), by warps per CTA and by load shape.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_read_bw tmem_read_bw.cu ; ./tmem_read_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define LD32(v, addr)                                                                                                        \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                     \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                      \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                      \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),   \
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),        \
                 "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),       \
                 "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                     \
               : "r"(addr)                                                                                                   \
               : "memory")
#define WAIT32(v)                                                                                                            \
  asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                              \
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),   \
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),        \
                 "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),       \
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])                     \
               :                                                                                                             \
               : "memory")

// MODE 0: load, wait, consume all 32 values (xor).  MODE 1: two loads in flight (second issued before the first is consumed).
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, uint32_t* out, long long* cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)((it * 32 + (warp >> 2) * 64) & 511) & ~31u;
    if (MODE == 0) {
      uint32_t v[32];
      LD32(v, base + col);
      WAIT32(v);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i];
    } else {
      uint32_t v[32], w[32];
      LD32(v, base + col);
      LD32(w, base + ((col + 32) & 511));
      WAIT32(v);
      WAIT32(w);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i] + w[i];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(iters, out, cyc);
        else k<1><<<148, warps * 32>>>(iters, out, cyc);
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * iters * 32 * 32 * 4 * (mode ? 2 : 1);
      printf("mode %d warps %2d: %.1f bytes/clk/SM (%lld cycles)  err=%s\n", mode, warps, bytes / (double)h[0], h[0], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
