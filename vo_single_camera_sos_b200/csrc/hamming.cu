// Step 2 of the SOS front-end: brute-force Hamming matching of 256-bit ORB descriptors.
//
// Replaces cv2.BFMatcher(NORM_HAMMING).match / knnMatch(k=2) as used by FeatureMatcher.match
// (reference omnistereo/camera_models.py:402-446), the post-sort at camera_models.py:444 and the pixel gate
// filter_pixel_correspondences (common_cv.py:167-188).
//
// Kernel design (integer-pipe bound):
//   * a thread keeps QPT query descriptors in registers (8 x u32 each);
//   * the block stages a tile of train descriptors in shared memory; every lane reads the SAME train
//     descriptor (two broadcast LDS.128), so shared-memory traffic is 32 B per 32*QPT descriptor pairs;
//   * per pair: 8 LOP3(xor) + 8 POPC + 4 IADD3, then a branch-free top-2 update on the packed key
//     (distance << 22 | train index) — min/max on the key gives OpenCV's (distance, lowest index) order;
//   * the train range of a segment is split over blockIdx.y so that small problems still fill 148 SMs;
//     the per-split partial keys are merged by a second, tiny kernel.
#include "sos_common.cuh"

namespace {

constexpr int HB_THREADS = 128;  // threads per block
constexpr int HB_QPT = 2;        // query descriptors per thread
constexpr int HB_TILE_Q = HB_THREADS * HB_QPT;
constexpr int HB_TILE_T = 256;   // train descriptors per shared-memory tile (8 KB)
constexpr int KEY_IDX_BITS = 22; // up to 4M train rows per segment; distance <= 256 needs 9 bits
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t hamming256(const uint4& qa, const uint4& qb, const uint4& ta, const uint4& tb) {
  uint32_t s0 = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z);
  uint32_t s1 = __popc(qa.w ^ ta.w) + __popc(qb.x ^ tb.x) + __popc(qb.y ^ tb.y);
  uint32_t s2 = __popc(qb.z ^ tb.z) + __popc(qb.w ^ tb.w) + s0;
  return s1 + s2;
}

__global__ void __launch_bounds__(HB_THREADS)
hamming_partial_kernel(const uint4* __restrict__ q, const uint4* __restrict__ t, const int32_t* __restrict__ q_start,
                       const int32_t* __restrict__ q_len, const int32_t* __restrict__ t_start,
                       const int32_t* __restrict__ t_len, int max_nq, int splits, uint2* __restrict__ partial) {
  __shared__ uint4 tile[HB_TILE_T * 2];

  const int seg = blockIdx.z;
  const int q0 = q_start[seg], nq = min(q_len[seg], max_nq);
  const int q_tile = blockIdx.x * HB_TILE_Q;
  if (q_tile >= nq) return;
  const int t0 = t_start[seg], nt = t_len[seg];
  const int split = blockIdx.y;
  const int chunk = (nt + splits - 1) / splits;
  const int t_begin = min(nt, split * chunk);
  const int t_end = min(nt, t_begin + chunk);

  uint4 qa[HB_QPT], qb[HB_QPT];
  uint32_t k0[HB_QPT], k1[HB_QPT];
#pragma unroll
  for (int r = 0; r < HB_QPT; ++r) {
    int row = q_tile + r * HB_THREADS + threadIdx.x;
    if (row >= nq) row = nq - 1;  // clamp: computes a duplicate that is never stored
    const uint4* p = q + (size_t)(q0 + row) * 2;
    qa[r] = __ldg(p);
    qb[r] = __ldg(p + 1);
    k0[r] = KEY_NONE;
    k1[r] = KEY_NONE;
  }

  for (int tb = t_begin; tb < t_end; tb += HB_TILE_T) {
    const int n_tile = min(HB_TILE_T, t_end - tb);
    __syncthreads();
    const uint4* src = t + (size_t)(t0 + tb) * 2;
    for (int i = threadIdx.x; i < n_tile * 2; i += HB_THREADS) tile[i] = __ldg(src + i);
    __syncthreads();

    int j = 0;
    for (; j + 4 <= n_tile; j += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint4 ta = tile[(j + u) * 2], tbv = tile[(j + u) * 2 + 1];
        const uint32_t tj = (uint32_t)(tb + j + u);
#pragma unroll
        for (int r = 0; r < HB_QPT; ++r) {
          const uint32_t key = (hamming256(qa[r], qb[r], ta, tbv) << KEY_IDX_BITS) | tj;
          k1[r] = min(k1[r], max(k0[r], key));
          k0[r] = min(k0[r], key);
        }
      }
    }
    for (; j < n_tile; ++j) {
      const uint4 ta = tile[j * 2], tbv = tile[j * 2 + 1];
      const uint32_t tj = (uint32_t)(tb + j);
#pragma unroll
      for (int r = 0; r < HB_QPT; ++r) {
        const uint32_t key = (hamming256(qa[r], qb[r], ta, tbv) << KEY_IDX_BITS) | tj;
        k1[r] = min(k1[r], max(k0[r], key));
        k0[r] = min(k0[r], key);
      }
    }
  }

#pragma unroll
  for (int r = 0; r < HB_QPT; ++r) {
    const int row = q_tile + r * HB_THREADS + threadIdx.x;
    if (row < nq) partial[((size_t)seg * max_nq + row) * splits + split] = make_uint2(k0[r], k1[r]);
  }
}

__global__ void __launch_bounds__(256)
hamming_merge_kernel(const uint2* __restrict__ partial, const int32_t* __restrict__ q_start,
                     const int32_t* __restrict__ q_len, int max_nq, int splits, int32_t* __restrict__ idx0,
                     int32_t* __restrict__ d0, int32_t* __restrict__ idx1, int32_t* __restrict__ d1) {
  const int seg = blockIdx.y;
  const int q0 = q_start[seg], nq = min(q_len[seg], max_nq);
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nq) return;
  const uint2* p = partial + ((size_t)seg * max_nq + row) * splits;
  uint32_t k0 = KEY_NONE, k1 = KEY_NONE;
  for (int s = 0; s < splits; ++s) {
    const uint2 v = p[s];
    // merge two sorted pairs: v.x <= v.y
    k1 = min(k1, max(k0, v.x));
    k0 = min(k0, v.x);
    k1 = min(k1, v.y);  // v.y >= v.x, so it can only compete for second place
  }
  const uint32_t mask = (1u << KEY_IDX_BITS) - 1u;
  idx0[q0 + row] = (k0 == KEY_NONE) ? -1 : (int32_t)(k0 & mask);
  d0[q0 + row] = (k0 == KEY_NONE) ? -1 : (int32_t)(k0 >> KEY_IDX_BITS);
  if (idx1) idx1[q0 + row] = (k1 == KEY_NONE) ? -1 : (int32_t)(k1 & mask);
  if (d1) d1[q0 + row] = (k1 == KEY_NONE) ? -1 : (int32_t)(k1 >> KEY_IDX_BITS);
}

// ---------------------------------------------------------------------------------------------------
// match_select: mode filter + pixel gate + stable counting sort by (distance, query index).
// One block per segment.
// ---------------------------------------------------------------------------------------------------
constexpr int MS_THREADS = 1024;
constexpr int MS_WARPS = MS_THREADS / 32;
constexpr int MS_BINS = 257;

struct SelectArgs {
  int mode;
  double ratio;
  const int32_t *idx0, *d0, *d1, *rev_idx0, *q_start, *q_len, *t_start;
  const float2 *px_q, *px_t;
  double max_du, min_dv;
  int32_t *out_q, *out_t, *out_d, *out_count;
};

__device__ __forceinline__ bool select_keep(const SelectArgs& a, int q0, int t0, int q, int& t_local, int& dist) {
  t_local = a.idx0[q0 + q];
  dist = a.d0[q0 + q];
  if (t_local < 0) return false;
  if (a.mode == SOS_MATCH_RATIO) {
    const int dd1 = a.d1[q0 + q];
    if (dd1 < 0) return false;                                  // len(m) == 2 fails (camera_models.py:423)
    if (!((double)dist < (double)dd1 * a.ratio)) return false;
  } else if (a.mode == SOS_MATCH_CROSS) {
    if (a.rev_idx0[t0 + t_local] != q) return false;
  }
  if (a.px_q != nullptr) {
    const float2 pq = a.px_q[q0 + q];
    const float2 pt = a.px_t[t0 + t_local];
    if (a.max_du > 0.0 && !(fabs((double)pt.x - (double)pq.x) <= a.max_du)) return false;  // common_cv.py:177-178
    if (a.min_dv >= 0.0 && !((double)pt.y - (double)pq.y >= a.min_dv)) return false;       // common_cv.py:182-184
  }
  return true;
}

__global__ void __launch_bounds__(MS_THREADS) match_select_kernel(SelectArgs a) {
  __shared__ int hist[MS_BINS + 1];
  __shared__ int cursor[MS_BINS];
  __shared__ unsigned char warpcnt[MS_WARPS][MS_BINS + 3];

  const int seg = blockIdx.x;
  const int q0 = a.q_start[seg], nq = a.q_len[seg];
  const int t0 = a.t_start[seg];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i <= MS_BINS; i += MS_THREADS) hist[i] = 0;
  for (int i = tid; i < MS_WARPS * (MS_BINS + 3); i += MS_THREADS) (&warpcnt[0][0])[i] = 0;
  __syncthreads();

  for (int q = tid; q < nq; q += MS_THREADS) {
    int tl, d;
    if (select_keep(a, q0, t0, q, tl, d)) atomicAdd(&hist[d], 1);
  }
  __syncthreads();
  // exclusive scan of 257 bins by one warp (9 bins per lane)
  if (warp == 0) {
    int local[9];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int b = lane * 9 + k;
      local[k] = (b < MS_BINS) ? hist[b] : 0;
      sum += local[k];
    }
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
      if (lane >= off) incl += v;
    }
    int run = incl - sum;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int b = lane * 9 + k;
      if (b < MS_BINS) cursor[b] = run;
      run += local[k];
    }
    if (lane == 31) a.out_count[seg] = incl;
  }
  __syncthreads();

  const int n_chunks = (nq + MS_THREADS - 1) / MS_THREADS;
  for (int c = 0; c < n_chunks; ++c) {
    const int q = c * MS_THREADS + tid;
    int tl = -1, d = 0;
    const bool keep = (q < nq) && select_keep(a, q0, t0, q, tl, d);
    // rank among lanes of this warp with the same distance
    const unsigned vote = __ballot_sync(0xFFFFFFFFu, keep);
    unsigned same = 0;
    if (keep) same = __match_any_sync(vote, d);
    const int rank_in_warp = __popc(same & ((1u << lane) - 1u));
    const bool leader = keep && rank_in_warp == 0;
    if (leader) warpcnt[warp][d] = (unsigned char)__popc(same);
    __syncthreads();
    if (keep) {
      int before = 0;
      for (int w = 0; w < warp; ++w) before += warpcnt[w][d];
      const int pos = cursor[d] + before + rank_in_warp;
      a.out_q[q0 + pos] = q0 + q;
      a.out_t[q0 + pos] = t0 + tl;
      a.out_d[q0 + pos] = d;
    }
    __syncthreads();
    if (leader) {
      atomicAdd(&cursor[d], (int)warpcnt[warp][d]);
      warpcnt[warp][d] = 0;
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int sos_hamming_top2(sos_ctx* ctx, const uint32_t* q, const uint32_t* t, const int32_t* q_start,
                                const int32_t* q_len, const int32_t* t_start, const int32_t* t_len, int n_seg,
                                int max_nq, int max_nt, int32_t* idx0, int32_t* d0, int32_t* idx1, int32_t* d1) {
  SOS_CHECK_ARG(ctx, "ctx is NULL");
  SOS_CHECK_ARG(n_seg >= 0 && max_nq >= 0 && max_nt >= 0, "negative size");
  SOS_CHECK_ARG(n_seg <= 65535, "at most 65535 segments per call");
  SOS_CHECK_ARG(max_nt < (1 << KEY_IDX_BITS), "segment has too many train rows (limit 4194303)");
  if (n_seg == 0 || max_nq == 0) return SOS_OK;
  SOS_CHECK_ARG(q && t && q_start && q_len && t_start && t_len && idx0 && d0, "NULL array");
  SOS_CHECK_ARG(((uintptr_t)q & 15) == 0 && ((uintptr_t)t & 15) == 0, "descriptor arrays must be 16-byte aligned");
  SOS_CUDA(cudaSetDevice(ctx->device));

  const int q_tiles = sos_div_up(max_nq, HB_TILE_Q);
  const int t_tiles = sos_div_up(max_nt > 0 ? max_nt : 1, HB_TILE_T);
  // enough blocks for ~4 per SM, but never split finer than one shared-memory tile
  int splits = sos_div_up(4 * ctx->sm_count, q_tiles * n_seg);
  if (splits > t_tiles) splits = t_tiles;
  if (splits > 64) splits = 64;
  if (splits < 1) splits = 1;

  void* ws = nullptr;
  const size_t rows_bound = (size_t)max_nq * (size_t)n_seg;
  int rc = sos_arena_get(ctx, rows_bound * splits * sizeof(uint2), &ws);
  if (rc != SOS_OK) return rc;

  dim3 grid(q_tiles, splits, n_seg);
  hamming_partial_kernel<<<grid, HB_THREADS, 0, ctx->stream>>>((const uint4*)q, (const uint4*)t, q_start, q_len, t_start,
                                                               t_len, max_nq, splits, (uint2*)ws);
  SOS_LAUNCHED(ctx);
  dim3 mgrid(sos_div_up(max_nq, 256), n_seg);
  hamming_merge_kernel<<<mgrid, 256, 0, ctx->stream>>>((const uint2*)ws, q_start, q_len, max_nq, splits, idx0, d0, idx1, d1);
  SOS_LAUNCHED(ctx);
  return SOS_OK;
}

extern "C" int sos_match_select(sos_ctx* ctx, int mode, double ratio, const int32_t* idx0, const int32_t* d0,
                                const int32_t* d1, const int32_t* rev_idx0, const int32_t* q_start,
                                const int32_t* q_len, const int32_t* t_start, int n_seg, const float* px_q,
                                const float* px_t, double max_du, double min_dv, int32_t* out_q,
                                int32_t* out_t, int32_t* out_d, int32_t* out_count) {
  SOS_CHECK_ARG(ctx, "ctx is NULL");
  SOS_CHECK_ARG(mode == SOS_MATCH_NN || mode == SOS_MATCH_RATIO || mode == SOS_MATCH_CROSS, "unknown mode");
  SOS_CHECK_ARG(n_seg >= 0, "negative size");
  if (n_seg == 0) return SOS_OK;
  SOS_CHECK_ARG(idx0 && d0 && q_start && q_len && t_start && out_q && out_t && out_d && out_count, "NULL array");
  SOS_CHECK_ARG(mode != SOS_MATCH_RATIO || d1, "ratio mode needs d1");
  SOS_CHECK_ARG(mode != SOS_MATCH_CROSS || rev_idx0, "cross mode needs rev_idx0");
  SOS_CHECK_ARG((px_q == nullptr) == (px_t == nullptr), "px_q and px_t must be given together");
  SOS_CUDA(cudaSetDevice(ctx->device));
  SelectArgs a;
  a.mode = mode;
  a.ratio = ratio;
  a.idx0 = idx0; a.d0 = d0; a.d1 = d1; a.rev_idx0 = rev_idx0; a.q_start = q_start; a.q_len = q_len; a.t_start = t_start;
  a.px_q = (const float2*)px_q; a.px_t = (const float2*)px_t;
  a.max_du = max_du; a.min_dv = min_dv;
  a.out_q = out_q; a.out_t = out_t; a.out_d = out_d; a.out_count = out_count;
  match_select_kernel<<<n_seg, MS_THREADS, 0, ctx->stream>>>(a);
  SOS_LAUNCHED(ctx);
  return SOS_OK;
}
