// Experiment: can the texture unit's bilinear filter reproduce cv2.remap's fixed-point blend bit for bit?
// cv2: coordinates in 1/32 pixel (ax, ay in 0..31), weights (32-ax)(32-ay)... * 32 in Q15, result (sum + 16384) >> 15.
// The texture unit filters with 8-bit fractions, in which k/32 is exact.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o tex_study tex_bilinear_study.cu ; run: ./tex_study
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

__global__ void study(cudaTextureObject_t tex, const uchar4* img, size_t pitch_px, int W, int H, int n_per_thread,
                      unsigned long long* mism /* [4]: floor(x+.5), floor(x+.5+eps), rintf, max |diff| > 1 */) {
  uint32_t s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned long long m0 = 0, m1 = 0, m2 = 0, m3 = 0;
  for (int it = 0; it < n_per_thread; ++it) {
    const int x0 = rng(s) % (W - 1), y0 = rng(s) % (H - 1), ax = rng(s) & 31, ay = rng(s) & 31;
    const float u = (float)x0 + (float)ax * (1.0f / 32.0f) + 0.5f, v = (float)y0 + (float)ay * (1.0f / 32.0f) + 0.5f;
    const float4 t = tex2D<float4>(tex, u, v);
    const uchar4 p00 = img[(size_t)y0 * pitch_px + x0], p01 = img[(size_t)y0 * pitch_px + x0 + 1];
    const uchar4 p10 = img[(size_t)(y0 + 1) * pitch_px + x0], p11 = img[(size_t)(y0 + 1) * pitch_px + x0 + 1];
    const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
    const float tv[3] = {t.x, t.y, t.z};
    const int c00[3] = {p00.x, p00.y, p00.z}, c01[3] = {p01.x, p01.y, p01.z}, c10[3] = {p10.x, p10.y, p10.z},
              c11[3] = {p11.x, p11.y, p11.z};
    for (int c = 0; c < 3; ++c) {
      const int want = (c00[c] * w00 + c01[c] * w01 + c10[c] * w10 + c11[c] * w11 + 16384) >> 15;
      const float f = tv[c] * 255.0f;
      const int g0 = (int)floorf(f + 0.5f), g1 = (int)floorf(f + 0.5f + 2e-4f), g2 = (int)rintf(f);
      m0 += g0 != want;
      m1 += g1 != want;
      m2 += g2 != want;
      m3 += abs(g1 - want) > 1;
    }
  }
  atomicAdd(&mism[0], m0); atomicAdd(&mism[1], m1); atomicAdd(&mism[2], m2); atomicAdd(&mism[3], m3);
}

int main() {
  const int W = 512, H = 512;
  std::vector<uchar4> h((size_t)W * H);
  uint32_t s = 12345;
  for (auto& p : h) { s = s * 1664525u + 1013904223u; p = make_uchar4(s >> 24, (s >> 16) & 255, (s >> 8) & 255, 255); }
  uchar4* d; size_t pitch;
  cudaMallocPitch(&d, &pitch, W * sizeof(uchar4), H);
  cudaMemcpy2D(d, pitch, h.data(), W * sizeof(uchar4), W * sizeof(uchar4), H, cudaMemcpyHostToDevice);
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = d; rd.res.pitch2D.width = W;
  rd.res.pitch2D.height = H; rd.res.pitch2D.pitchInBytes = pitch; rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
  cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  unsigned long long* mism; cudaMalloc(&mism, 32); cudaMemset(mism, 0, 32);
  const int blocks = 592, threads = 256, per = 256;
  study<<<blocks, threads>>>(tex, d, pitch / sizeof(uchar4), W, H, per, mism);
  unsigned long long r[4]; cudaMemcpy(r, mism, 32, cudaMemcpyDeviceToHost);
  const double n = 3.0 * blocks * threads * per;
  printf("{\"samples\": %.0f, \"mismatch_floor_half\": %llu, \"mismatch_floor_half_eps\": %llu, \"mismatch_rint\": %llu, "
         "\"off_by_more_than_1\": %llu, \"err\": \"%s\"}\n", n, r[0], r[1], r[2], r[3], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
