// 11 x 11 median filter of the panoramas (pre-processing).
//
// replaces: cv2.medianBlur(pano_img, median_win_size) with median_win_size = 11 (pose_est_tools.py:297;
//           camera_models.py:1708-1709): exact median of the 121 window values per channel, BORDER_REPLICATE.
//
// Sliding-window histograms (Huang).  Every lane owns one (column, channel) histogram of byte counters and walks down its
// column by itself: per row 11 values leave and 11 enter the window with plain byte read-modify-writes — no atomics, no
// ballots, no idle lanes.  (The first version — one warp per pair of columns with shared-memory atomics and ballots — took
// 2.3x longer: 219 instructions per two-pixel step with 18 of 32 lanes active, 5.8 wavefronts per atomic; removed in round 2.)
// Layout (8 KB per warp): the counter of bin b of lane l is byte (b & 3) of word (b >> 2) * 32 + l, so a lane only ever touches
// its own bank.  (First layout: byte b * 32 + l, i.e. four lanes per word - ncu: LSU data-pipe wavefronts at 93 % of peak, 167
// per step for 68 requests, because lanes of one word group collide whenever their bins differ by a multiple of 4.)
#include "sos_common.cuh"

namespace {

constexpr int MB_R = 5;             // window radius (11 x 11)
constexpr int MB_HALF = 60;         // 0-based rank of the median among 121 values
constexpr int ML_WARPS = 4;

__device__ __forceinline__ int ml_slot(int b) { return ((b & 0xFC) << 5) | (b & 3); }

// Move the median candidate m (with ltm = number of window values below m) to the median: ltm <= MB_HALF < ltm + count(m).
// The four bins 4k .. 4k+3 of a lane share one word, so long walks (an edge entering the window moves the median by tens of
// levels, and a warp waits for its slowest lane) take four bins per shared-memory read.
__device__ __forceinline__ void ml_walk(const uint8_t* hb, int& m, int& ltm) {
  const uint32_t* hw = (const uint32_t*)hb;                // word k of this lane: hw[k * 32]
  if (ltm > MB_HALF) {                                     // down; the first step is the usual, single one
    --m;
    ltm -= hb[ml_slot(m)];
    while (ltm > MB_HALF) {
      if ((m & 3) == 0) {
        const int sw = (int)__dp4a(hw[((m >> 2) - 1) * 32], 0x01010101u, 0u);
        if (ltm - sw > MB_HALF) { ltm -= sw; m -= 4; continue; }
      }
      --m;
      ltm -= hb[ml_slot(m)];
    }
  }
  int c = hb[ml_slot(m)];
  if (ltm + c > MB_HALF) return;                           // up; most steps end here
  ltm += c;
  ++m;
  while (true) {
    if ((m & 3) == 0) {
      const int sw = (int)__dp4a(hw[(m >> 2) * 32], 0x01010101u, 0u);
      if (ltm + sw <= MB_HALF) { ltm += sw; m += 4; continue; }
    }
    c = hb[ml_slot(m)];
    if (ltm + c > MB_HALF) break;
    ltm += c;
    ++m;
  }
}

// cv::cvtColor(BGR2GRAY) on the three medians of a column (15-bit fixed point, orb.cu): the lanes of one column are
// neighbours, so the lane of channel 0 collects the other two with shuffles.  `amask` = the warp's lanes that walk columns.
template <int C>
__device__ __forceinline__ void ml_store(uint8_t* __restrict__ out, uint8_t* __restrict__ gray, size_t pix, int ch, int m,
                                         unsigned amask) {
  if (out) out[pix * C + ch] = (uint8_t)m;
  if (C == 3 && gray) {
    const int lane = threadIdx.x & 31;
    const uint32_t g = __shfl_sync(amask, m, lane - ch + 1), r = __shfl_sync(amask, m, lane - ch + 2);
    if (ch == 0) gray[pix] = (uint8_t)(((uint32_t)m * 3735u + g * 19235u + r * 9798u + 16384u) >> 15);
  }
}

template <int C, bool INTERIOR>
__device__ __forceinline__ void median11_lane_column(const uint8_t* __restrict__ img, uint8_t* __restrict__ out,
                                                     uint8_t* __restrict__ gray, unsigned amask, uint8_t* hb, int H, int W,
                                                     int x, int ch, int y0, int y1) {
  int xo[2 * MB_R + 1];                                  // byte offsets of the window columns inside a row (replicated border)
#pragma unroll
  for (int k = 0; k <= 2 * MB_R; ++k) xo[k] = INTERIOR ? (x + k - MB_R) * C + ch : min(max(x + k - MB_R, 0), W - 1) * C + ch;
  const size_t pitch = (size_t)W * C;
  int m = 0, ltm = 0;                                    // median candidate and the number of window values below it

  for (int r = y0 - MB_R; r <= y0 + MB_R; ++r) {
    const uint8_t* row = img + (size_t)min(max(r, 0), H - 1) * pitch + (INTERIOR ? xo[0] : 0);
#pragma unroll
    for (int k = 0; k <= 2 * MB_R; ++k) hb[ml_slot(INTERIOR ? row[k * C] : row[xo[k]])] += 1;
  }
  ml_walk(hb, m, ltm);
  ml_store<C>(out, gray, (size_t)y0 * W + x, ch, m, amask);

  for (int y = y0 + 1; y < y1; ++y) {
    const uint8_t* rold = img + (size_t)min(max(y - MB_R - 1, 0), H - 1) * pitch + (INTERIOR ? xo[0] : 0);
    const uint8_t* rnew = img + (size_t)min(max(y + MB_R, 0), H - 1) * pitch + (INTERIOR ? xo[0] : 0);
    int vo[2 * MB_R + 1], vn[2 * MB_R + 1];
#pragma unroll
    for (int k = 0; k <= 2 * MB_R; ++k) {                // all 22 loads first: they are independent
      vo[k] = __ldg(INTERIOR ? rold + k * C : rold + xo[k]);
      vn[k] = __ldg(INTERIOR ? rnew + k * C : rnew + xo[k]);
    }
#pragma unroll
    for (int k = 0; k <= 2 * MB_R; ++k) {
      hb[ml_slot(vo[k])] -= 1;
      hb[ml_slot(vn[k])] += 1;
      ltm += (vn[k] < m ? 1 : 0) - (vo[k] < m ? 1 : 0);
    }
    ml_walk(hb, m, ltm);
    ml_store<C>(out, gray, (size_t)y * W + x, ch, m, amask);
  }
}

template <int C>
__global__ void __launch_bounds__(ML_WARPS * 32)
median11_lane_kernel(const uint8_t* __restrict__ src, int H, int W, int strip_rows, uint8_t* __restrict__ dst,
                     uint8_t* __restrict__ gray_dst) {
  constexpr int COLS = 32 / C;                           // columns per warp: 10 for BGR, 32 for gray
  __shared__ uint32_t hist_all[ML_WARPS][256 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = (blockIdx.x * ML_WARPS + warp) * COLS;
  if (x0 >= W) return;                                   // whole warps leave together; only __syncwarp below
  const int y0 = blockIdx.y * strip_rows, y1 = min(H, y0 + strip_rows);
  const uint8_t* img = src + (size_t)blockIdx.z * H * W * C;
  uint8_t* out = dst ? dst + (size_t)blockIdx.z * H * W * C : nullptr;
  uint8_t* gray = gray_dst ? gray_dst + (size_t)blockIdx.z * H * W : nullptr;
  uint32_t* hw = hist_all[warp];
  for (int i = lane; i < 256 * 8; i += 32) hw[i] = 0u;
  __syncwarp();
  const int col = lane / C, ch = lane - col * C;
  const bool active = col < COLS && x0 + col < W;
  const unsigned amask = __ballot_sync(0xFFFFFFFFu, active);
  if (!active) return;
  uint8_t* hb = (uint8_t*)(hw + lane);                   // this lane's words: hw[(b >> 2) * 32 + lane]
  // warp-uniform: every window column of every lane lies inside the image -> constant offsets between the 11 loads of a row
  if (x0 - MB_R >= 0 && x0 + COLS - 1 + MB_R <= W - 1) median11_lane_column<C, true>(img, out, gray, amask, hb, H, W, x0 + col, ch, y0, y1);
  else median11_lane_column<C, false>(img, out, gray, amask, hb, H, W, x0 + col, ch, y0, y1);
}

}  // namespace

extern "C" int sos_median_blur_11(sos_ctx* ctx, const uint8_t* src, int n_images, int height, int width, int channels,
                                  uint8_t* dst) {
  SOS_CHECK_ARG(ctx, "ctx is NULL");
  SOS_CHECK_ARG(n_images >= 0 && height >= 0 && width >= 0, "negative size");
  SOS_CHECK_ARG(channels == 1 || channels == 3, "channels must be 1 or 3");
  if (n_images == 0 || height == 0 || width == 0) return SOS_OK;
  SOS_CHECK_ARG(src && dst && src != dst, "NULL array or in-place call");
  SOS_CHECK_ARG(n_images <= 65535, "too many images");
  SOS_CUDA(cudaSetDevice(ctx->device));
  const int strip = 128;
  const int cols_per_block = ML_WARPS * (32 / channels);
  dim3 grid(sos_div_up(width, cols_per_block), sos_div_up(height, strip), n_images);
  if (channels == 3) median11_lane_kernel<3><<<grid, ML_WARPS * 32, 0, ctx->stream>>>(src, height, width, strip, dst, nullptr);
  else median11_lane_kernel<1><<<grid, ML_WARPS * 32, 0, ctx->stream>>>(src, height, width, strip, dst, nullptr);
  SOS_LAUNCHED_AS(ctx, "median11_lane_kernel");
  return SOS_OK;
}

extern "C" int sos_median_blur_11_gray(sos_ctx* ctx, const uint8_t* src, int n_images, int height, int width, uint8_t* dst_bgr,
                                       uint8_t* gray) {
  SOS_CHECK_ARG(ctx, "ctx is NULL");
  SOS_CHECK_ARG(n_images >= 0 && height >= 0 && width >= 0, "negative size");
  if (n_images == 0 || height == 0 || width == 0) return SOS_OK;
  SOS_CHECK_ARG(src && gray && src != dst_bgr, "NULL array or in-place call");
  SOS_CHECK_ARG(n_images <= 65535, "too many images");
  SOS_CUDA(cudaSetDevice(ctx->device));
  const int strip = 128;
  dim3 grid(sos_div_up(width, ML_WARPS * (32 / 3)), sos_div_up(height, strip), n_images);
  median11_lane_kernel<3><<<grid, ML_WARPS * 32, 0, ctx->stream>>>(src, height, width, strip, dst_bgr, gray);
  SOS_LAUNCHED_AS(ctx, "median11_lane_kernel");
  return SOS_OK;
}
