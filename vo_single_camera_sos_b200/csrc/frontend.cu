// The batched, GPU-resident SOS front-end: one "step" takes B new frames (omni image + per-view ORB features) and
// produces B frame-pair poses (frame i-1 -> frame i; the last frame of a step is carried over as the reference of the
// next step).  It chains the five hot-path steps exactly as StereoPanoramicFrame.establish_stereo_correspondences
// (reference omnistereo/pose_est_tools.py:320-402) and TrackerStereoSE3.track_frame (pose_est_tools.py:736-847) do,
// but for all frames of the batch at once and without ever returning to the host: every data-dependent size (matches
// per bucket, triangulated points per frame, correspondences per pair) stays in device memory and is read by the next
// kernel.  The kernel chain is captured once into a CUDA graph and replayed per step.
//
//   remap (2 views)                                           panorama.py:258-321, camera_models.py:3107-3120
//   stereo Hamming match per azimuthal bucket, bottom = query  camera_models.py:3027-3101
//   sort by distance + pixel gate (|du| <= 2.5, dv >= 1)        camera_models.py:444, common_cv.py:167-188
//   lift + midpoint triangulation + range gate + compaction    pose_est_tools.py:344-397
//   temporal Hamming match top<->top, bottom<->bottom           pose_est_tools.py:741-749, 211-269
//   stack correspondences (top first, then bottom)             pose_est_tools.py:752-778
//   RANSAC over seeded Arun hypotheses (+ inlier refit)        pose_est_tools.py:785, 830
#include <string.h>

#include <vector>

#include <algorithm>
#include <climits>

#include "sos_common.cuh"

namespace {

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
};

// One thread per frame: the (query = bottom, train = top) segment of every azimuthal bucket.  Offsets are clamped to the
// frame's feature capacity F and a bucket's rows to max_bucket (the per-segment capacity of the matching scratch and of
// the candidate-pair arrays); what the clamps drop is counted in overflow[b] so that the caller can see it (the
// reference has no such cap: size max_feat_per_bucket for the detector's per-bucket budget, pose_est_tools.py:862).
__global__ void stereo_segments_kernel(const int32_t* __restrict__ boff_top, const int32_t* __restrict__ boff_bot, int B,
                                       int nb, int F, int max_bucket, int32_t* __restrict__ q_start,
                                       int32_t* __restrict__ q_len, int32_t* __restrict__ t_start,
                                       int32_t* __restrict__ t_len, int32_t* __restrict__ overflow) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int32_t* ot = boff_top + (size_t)b * (nb + 1);
  const int32_t* ob = boff_bot + (size_t)b * (nb + 1);
  int dropped = 0;
  for (int k = 0; k < nb; ++k) {
    const int s = b * nb + k;
    const int q0 = min(max(ob[k], 0), F), q1 = min(max(ob[k + 1], q0), F);
    const int t0 = min(max(ot[k], 0), F), t1 = min(max(ot[k + 1], t0), F);
    const int nq = min(q1 - q0, max_bucket), nt = min(t1 - t0, max_bucket);
    dropped += max(0, ob[k + 1] - ob[k]) - nq + max(0, ot[k + 1] - ot[k]) - nt;
    q_start[s] = b * F + q0;  // bottom view = query (camera_models.py:3042)
    q_len[s] = nq;
    t_start[s] = b * F + t0;  // top view = train
    t_len[s] = nt;
  }
  overflow[b] = dropped;
}

// desc_c[(view*(B+1) + slot)*cap + k] = desc_view[src_view[slot*cap + k]] for the B new slots (1..B)
__global__ void gather_desc_kernel(const uint4* __restrict__ desc_top, const uint4* __restrict__ desc_bot,
                                   const int32_t* __restrict__ src_top, const int32_t* __restrict__ src_bot,
                                   const int32_t* __restrict__ n, int B, int cap, uint4* __restrict__ desc_c) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = blockIdx.y + 1;
  const int view = blockIdx.z;
  if (k >= n[slot]) return;
  const size_t row = (size_t)slot * cap + k;
  const int src = view == 0 ? src_top[row] : src_bot[row];
  const uint4* d = (view == 0 ? desc_top : desc_bot) + (size_t)src * 2;
  uint4* o = desc_c + ((size_t)(view * (B + 1) + slot) * cap + k) * 2;
  o[0] = __ldg(d);
  o[1] = __ldg(d + 1);
}

// Pair i tracks store slot i + 1 against its reference slot: slot i (consecutive frames) or ref_slot[i] (keyframe mode:
// the slot of the tracking reference, pose_est_tools.py:1489; -1 = pair switched off).
__global__ void temporal_segments_kernel(const int32_t* __restrict__ n, const int32_t* __restrict__ ref_slot, int B, int cap,
                                         int32_t* __restrict__ q_start, int32_t* __restrict__ q_len,
                                         int32_t* __restrict__ t_start, int32_t* __restrict__ t_len) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= 2 * B) return;
  const int view = s / B, i = s % B;
  const int r = ref_slot ? ref_slot[i] : i;
  const bool on = r >= 0 && r <= B && r != i + 1;
  q_start[s] = (view * (B + 1) + i + 1) * cap;  // query = current frame (pose_est_tools.py:215)
  q_len[s] = on ? n[i + 1] : 0;
  t_start[s] = (view * (B + 1) + (on ? r : 0)) * cap;  // train = reference frame
  t_len[s] = on ? n[r] : 0;
}

// Stack the temporal matches of both views into one correspondence list per frame pair (pose_est_tools.py:752-778).
__global__ void __launch_bounds__(256)
assemble_kernel(const int32_t* __restrict__ m_q, const int32_t* __restrict__ m_t, const int32_t* __restrict__ m_count,
                const int32_t* __restrict__ q_start, int B, int cap, const float* __restrict__ xyz,
                const float* __restrict__ b_top, const float* __restrict__ b_bot, float* __restrict__ p_ref,
                float* __restrict__ p_cur, float* __restrict__ f_cur, uint8_t* __restrict__ cam,
                int32_t* __restrict__ n_corr, int32_t* __restrict__ n_corr_top) {
  const int i = blockIdx.y;  // frame pair; blockIdx.x: 256 rows of its list (one block per pair left 116 of 148 SMs idle)
  const int c_top = m_count[i], c_bot = m_count[B + i];
  const int total = c_top + c_bot;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    n_corr[i] = total;
    n_corr_top[i] = c_top;
  }
  const int cap2 = 2 * cap;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    const int view = k < c_top ? 0 : 1;
    const int s = view * B + i;
    const int kk = view == 0 ? k : k - c_top;
    const int base = view * (B + 1) * cap;        // descriptor-store rows of this view start here
    const int rq = m_q[q_start[s] + kk] - base;   // store row of the current-frame correspondence
    const int rt = m_t[q_start[s] + kk] - base;   // store row of the reference-frame correspondence
    const size_t o = ((size_t)i * cap2 + k) * 3;
    const float* bq = (view == 0 ? b_top : b_bot) + (size_t)rq * 3;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      p_ref[o + d] = xyz[(size_t)rt * 3 + d];
      p_cur[o + d] = xyz[(size_t)rq * 3 + d];
      f_cur[o + d] = bq[d];
    }
    cam[(size_t)i * cap2 + k] = (uint8_t)view;
  }
}

// Store slot `from` -> slot 0 (the tracking reference of what follows).
__global__ void carry_over_kernel(int B, int from, int cap, float* __restrict__ uv_top, float* __restrict__ uv_bot,
                                  float* __restrict__ b_top, float* __restrict__ b_bot, float* __restrict__ xyz,
                                  uint32_t* __restrict__ desc_c, int32_t* __restrict__ n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int cnt = n[from];
  if (k == 0) n[0] = cnt;
  if (k >= cnt) return;
  const size_t src = (size_t)from * cap + k, dst = k;
  ((float2*)uv_top)[dst] = ((const float2*)uv_top)[src];
  ((float2*)uv_bot)[dst] = ((const float2*)uv_bot)[src];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    b_top[3 * dst + c] = b_top[3 * src + c];
    b_bot[3 * dst + c] = b_bot[3 * src + c];
    xyz[3 * dst + c] = xyz[3 * src + c];
  }
  for (int view = 0; view < 2; ++view) {
    uint4* base = (uint4*)(desc_c + (size_t)view * (B + 1) * cap * 8);
    base[2 * dst] = base[2 * src];
    base[2 * dst + 1] = base[2 * src + 1];
  }
}

// Per source row: which 64-byte blocks any live panorama pixel may read (both tap rows, plus the slack of the wide loads).
// blocks[row * words + w] is a bitmap over blocks 64 w .. 64 w + 63 of that row.
__global__ void lut_row_blocks_kernel(const sos_lut_entry* __restrict__ lut, size_t n, int src_h, int src_w, int ch, int words,
                                      unsigned long long* __restrict__ blocks) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t e = lut[i];
  if (((e >> 48) & 0xF) == 0) return;   // no tap inside the image: nothing is read
  const int x0 = (int)(int16_t)(e & 0xFFFF), y0 = (int)(int16_t)((e >> 16) & 0xFFFF);
  const int lo = max(0, x0 * ch - 16), hi = min(src_w * ch, (x0 + 2) * ch + 24);
  if (hi <= lo) return;
  for (int y = y0; y <= y0 + 1; ++y)
    if (y >= 0 && y < src_h)
      for (int b = lo >> 6; b <= (hi - 1) >> 6; ++b) atomicOr(&blocks[(size_t)y * words + (b >> 6)], 1ull << (b & 63));
}

__global__ void stats_kernel(const int32_t* __restrict__ n, const int32_t* __restrict__ n_corr,
                             const int32_t* __restrict__ best_count, const int32_t* __restrict__ best_hyp, int B,
                             int32_t* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  stats[4 * i + 0] = n[i + 1];       // triangulated stereo correspondences of the new frame
  stats[4 * i + 1] = n_corr[i];      // temporal correspondences handed to RANSAC
  stats[4 * i + 2] = best_count[i];  // RANSAC inliers
  stats[4 * i + 3] = best_hyp[i];
}

}  // namespace

struct sos_frontend {
  sos_ctx* ctx = nullptr;     // private (own scratch arena)
  sos_ctx* parent = nullptr;  // the caller's context (launch accounting)
  sos_frontend_config cfg;
  const sos_lut_entry* lut = nullptr;
  const uint32_t* hyp = nullptr;
  // device state
  sos_frontend_buffers d;
  std::vector<void*> owned;
  // staging for the host API (double buffered)
  static constexpr int DEPTH = 2;
  struct Stage {
    uint8_t* omni = nullptr;
    float *px_top = nullptr, *px_bot = nullptr;
    uint32_t *desc_top = nullptr, *desc_bot = nullptr;
    int32_t *boff_top = nullptr, *boff_bot = nullptr;
    float* h_poses = nullptr;     // pinned
    int32_t* h_stats = nullptr;   // pinned
    cudaEvent_t copied = nullptr, done = nullptr;
    bool in_flight = false;
  } stage[DEPTH];
  cudaStream_t copy_stream = nullptr;
  // host API: only the bytes of the omni image that some LUT entry can read are uploaded, in bands of rows
  struct Band {
    int row0, nrows, x0_bytes, width_bytes;
  };
  std::vector<Band> bands;
  int64_t omni_bytes_per_frame = 0;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;  // fences against the caller's stream
  // the remap is a branch of its own in the step (nothing later in the step reads the panoramas): it fills the SMs while the
  // many small, latency-bound kernels of the matching / RANSAC chain run
  cudaStream_t remap_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool remap_forked = false;
  int next_stage = 0;
  // graph cache keyed on the input pointers
  struct GraphKey {
    const void* p[7];
  };
  struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec = nullptr;
  };
  std::vector<GraphEntry> graphs;
  bool use_graph = true;
};

namespace {

template <typename T>
int fe_alloc(sos_frontend* fe, T** out, size_t count) {
  void* p = nullptr;
  const size_t bytes = sos_align_up(count * sizeof(T) + 256, 256);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    sos_set_error("sos_frontend: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return SOS_ERR_NOMEM;
  }
  cudaMemsetAsync(p, 0, bytes, fe->ctx->stream);
  fe->owned.push_back(p);
  *out = (T*)p;
  return SOS_OK;
}

// Stage A of a step (remap, stereo matching, lifting + triangulation into store slots 1..B), enqueued on ctx->stream.
int enqueue_stage_a(sos_frontend* fe, const uint8_t* omni, const float* px_top, const uint32_t* desc_top,
                 const int32_t* boff_top, const float* px_bot, const uint32_t* desc_bot, const int32_t* boff_bot) {
  sos_ctx* ctx = fe->ctx;
  const sos_frontend_config& c = fe->cfg;
  sos_frontend_buffers& d = fe->d;
  const int B = c.batch, nb = c.n_buckets, F = c.max_feat_per_view, cap = c.cap;
  int rc;
  // step 1, as a parallel branch (fork here, join at the end of the step).  Round 1 measured no gain from this — then the
  // chain was three long kernels that filled the GPU; now a fifth of the step is small launches that leave most SMs idle.
  // Per-kernel profiling runs keep it in line so that every kernel is timed alone.
  fe->remap_forked = fe->remap_stream != nullptr && !ctx->prof;
  if (fe->remap_forked) {
    SOS_CUDA(cudaEventRecord(fe->ev_fork, ctx->stream));
    SOS_CUDA(cudaStreamWaitEvent(fe->remap_stream, fe->ev_fork, 0));
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = fe->remap_stream;
    rc = sos_remap_u8(ctx, omni, B, c.src_h, c.src_w, c.channels, fe->lut, 2, c.pano_rows, c.pano_cols, c.border,
                      c.background, d.pano);
    ctx->stream = main_stream;
    if (rc) return rc;
    SOS_CUDA(cudaEventRecord(fe->ev_join, fe->remap_stream));
  } else {
    rc = sos_remap_u8(ctx, omni, B, c.src_h, c.src_w, c.channels, fe->lut, 2, c.pano_rows, c.pano_cols, c.border,
                      c.background, d.pano);
    if (rc) return rc;
  }
  // step 2a: stereo matching per bucket
  const int S = B * nb;
  stereo_segments_kernel<<<sos_div_up(B, 128), 128, 0, ctx->stream>>>(boff_top, boff_bot, B, nb, F, c.max_feat_per_bucket,
                                                                       d.st_q_start, d.st_q_len, d.st_t_start, d.st_t_len,
                                                                       d.overflow);
  SOS_LAUNCHED_AS(ctx, "stereo_segments_kernel");
  rc = sos_hamming_top2(ctx, desc_bot, desc_top, d.st_q_start, d.st_q_len, d.st_t_start, d.st_t_len, S,
                        c.max_feat_per_bucket, c.max_feat_per_bucket, d.st_idx0, d.st_d0, nullptr, nullptr);
  if (rc) return rc;
  rc = sos_match_select(ctx, SOS_MATCH_NN, 0.75, d.st_idx0, d.st_d0, nullptr, nullptr, d.st_q_start, d.st_q_len,
                        d.st_t_start, S, px_bot, px_top, c.stereo_max_du, c.stereo_min_dv, d.st_pair_q, d.st_pair_t,
                        d.st_pair_d, d.st_pair_count);
  if (rc) return rc;
  // steps 3+4 into store slots 1..B
  rc = sos_stereo_lift_triangulate(ctx, c.pano_top, c.pano_bot, px_top, px_bot, d.st_pair_q, d.st_pair_t, d.st_pair_count,
                                   d.st_q_start, B, nb, c.max_feat_per_bucket, B * F, c.f_top, c.f_bot, c.min_range,
                                   c.max_range, c.homogeneous_norm, cap,
                                   d.uv_top + (size_t)cap * 2, d.uv_bot + (size_t)cap * 2, d.b_top + (size_t)cap * 3,
                                   d.b_bot + (size_t)cap * 3, d.xyz + (size_t)cap * 3, d.src_top + cap, d.src_bot + cap,
                                   d.n + 1);
  if (rc) return rc;
  {
    dim3 grid(sos_div_up(cap, 128), B, 2);
    gather_desc_kernel<<<grid, 128, 0, ctx->stream>>>((const uint4*)desc_top, (const uint4*)desc_bot, d.src_top, d.src_bot,
                                                      d.n, B, cap, (uint4*)d.desc_c);
    SOS_LAUNCHED_AS(ctx, "gather_desc_kernel");
  }
  return SOS_OK;
}

// Stage B: temporal matching against the reference slots, RANSAC, refinement.  Re-runnable on its own (sos_frontend_retrack).
int enqueue_stage_b(sos_frontend* fe) {
  sos_ctx* ctx = fe->ctx;
  const sos_frontend_config& c = fe->cfg;
  sos_frontend_buffers& d = fe->d;
  const int B = c.batch, cap = c.cap;
  int rc;
  // step 2b: temporal matching, 2 views x B pairs
  temporal_segments_kernel<<<sos_div_up(2 * B, 128), 128, 0, ctx->stream>>>(d.n, c.keyframe_mode ? d.ref_slot : nullptr, B, cap,
                                                                            d.tm_q_start, d.tm_q_len, d.tm_t_start, d.tm_t_len);
  SOS_LAUNCHED_AS(ctx, "temporal_segments_kernel");
  rc = sos_hamming_top2(ctx, d.desc_c, d.desc_c, d.tm_q_start, d.tm_q_len, d.tm_t_start, d.tm_t_len, 2 * B, cap, cap,
                        d.tm_idx0, d.tm_d0, nullptr, nullptr);
  if (rc) return rc;
  // pixel gate on u only (pose_est_tools.py:245-247): coordinates of the compacted store, both views in one array
  rc = sos_match_select(ctx, SOS_MATCH_NN, 0.75, d.tm_idx0, d.tm_d0, nullptr, nullptr, d.tm_q_start, d.tm_q_len,
                        d.tm_t_start, 2 * B, d.uv_c, d.uv_c, c.temporal_max_du, -1.0, d.tm_pair_q, d.tm_pair_t, d.tm_pair_d,
                        d.tm_pair_count);
  if (rc) return rc;
  assemble_kernel<<<dim3(sos_div_up(2 * cap, 256), B), 256, 0, ctx->stream>>>(d.tm_pair_q, d.tm_pair_t, d.tm_pair_count, d.tm_q_start, B, cap, d.xyz, d.b_top,
                                              d.b_bot, d.p_ref, d.p_cur, d.f_cur, d.cam, d.n_corr, d.n_corr_top);
  SOS_LAUNCHED_AS(ctx, "assemble_kernel");
  // step 5
  rc = c.solver == SOS_SOLVER_P3P
           ? sos_ransac_p3p(ctx, d.p_ref, d.f_cur, d.cam, d.n_corr, B, 2 * cap, c.rig, 2, fe->hyp, c.n_hyp, 0, c.ransac_threshold,
                            d.ransac_pose, d.best_hyp, d.best_count, d.inlier_mask, nullptr, nullptr)
           : sos_ransac_p3d(ctx, d.p_ref, d.p_cur, d.f_cur, d.cam, d.n_corr, B, 2 * cap, c.rig, 2, fe->hyp, c.n_hyp, 0,
                            c.score_mode, c.ransac_threshold, d.ransac_pose, d.best_hyp, d.best_count, d.inlier_mask, nullptr,
                            nullptr);
  if (rc) return rc;
  if (c.refit == SOS_REFINE_LM) {
    rc = sos_refine_pose(ctx, d.p_ref, d.f_cur, d.cam, d.inlier_mask, d.n_corr, B, 2 * cap, c.rig, 2, d.ransac_pose,
                         c.refine_iters > 0 ? c.refine_iters : 20, 0, d.pose, nullptr, d.refine_stats);
    if (rc) return rc;
  } else if (c.refit == SOS_REFINE_ARUN) {
    rc = sos_refit_inliers(ctx, d.p_ref, d.p_cur, d.inlier_mask, d.n_corr, B, 2 * cap, d.pose, d.n_refit);
    if (rc) return rc;
  } else {
    SOS_CUDA(cudaMemcpyAsync(d.pose, d.ransac_pose, (size_t)B * 12 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  stats_kernel<<<sos_div_up(B, 128), 128, 0, ctx->stream>>>(d.n, d.n_corr, d.best_count, d.best_hyp, B, d.stats);
  SOS_LAUNCHED_AS(ctx, "stats_kernel");
  return SOS_OK;
}

int enqueue_carry(sos_frontend* fe, int from) {
  sos_ctx* ctx = fe->ctx;
  sos_frontend_buffers& d = fe->d;
  carry_over_kernel<<<sos_div_up(fe->cfg.cap, 256), 256, 0, ctx->stream>>>(fe->cfg.batch, from, fe->cfg.cap, d.uv_top, d.uv_bot,
                                                                            d.b_top, d.b_bot, d.xyz, d.desc_c, d.n);
  SOS_LAUNCHED_AS(ctx, "carry_over_kernel");
  return SOS_OK;
}

// The kernel chain of one step.  Consecutive mode ends by carrying the last frame over as the reference of the next
// step (slot B -> slot 0, one kernel instead of 8 copy nodes); in keyframe mode the caller promotes a slot itself.
int enqueue_step(sos_frontend* fe, const uint8_t* omni, const float* px_top, const uint32_t* desc_top,
                 const int32_t* boff_top, const float* px_bot, const uint32_t* desc_bot, const int32_t* boff_bot) {
  int rc = enqueue_stage_a(fe, omni, px_top, desc_top, boff_top, px_bot, desc_bot, boff_bot);
  if (rc == SOS_OK) rc = enqueue_stage_b(fe);
  if (rc == SOS_OK && !fe->cfg.keyframe_mode) rc = enqueue_carry(fe, fe->cfg.batch);
  if (fe->remap_forked) {   // join the remap branch (also on an error path: a capture must not end with an open fork)
    fe->remap_forked = false;
    SOS_CUDA(cudaStreamWaitEvent(fe->ctx->stream, fe->ev_join, 0));
  }
  return rc;
}

int run_step(sos_frontend* fe, const uint8_t* omni, const float* px_top, const uint32_t* desc_top, const int32_t* boff_top,
             const float* px_bot, const uint32_t* desc_bot, const int32_t* boff_bot) {
  sos_ctx* ctx = fe->ctx;
  if (!fe->use_graph || ctx->prof) return enqueue_step(fe, omni, px_top, desc_top, boff_top, px_bot, desc_bot, boff_bot);
  sos_frontend::GraphKey key = {{omni, px_top, desc_top, boff_top, px_bot, desc_bot, boff_bot}};
  for (auto& g : fe->graphs) {
    if (memcmp(&g.key, &key, sizeof(key)) == 0) {
      SOS_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
      ctx->launches += fe->d.launches_per_step;
      return SOS_OK;
    }
  }
  // first time with these buffers: run once eagerly (sizes the scratch arena), then capture
  const int64_t before = ctx->launches;
  int rc = enqueue_step(fe, omni, px_top, desc_top, boff_top, px_bot, desc_bot, boff_bot);
  if (rc) return rc;
  fe->d.launches_per_step = (int)(ctx->launches - before);
  SOS_CUDA(cudaStreamSynchronize(ctx->stream));
  // the eager run advanced the carried state; capturing does not execute anything
  cudaGraph_t graph = nullptr;
  SOS_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  const int64_t l0 = ctx->launches;
  rc = enqueue_step(fe, omni, px_top, desc_top, boff_top, px_bot, desc_bot, boff_bot);
  ctx->launches = l0;
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) {
    sos_set_error("sos_frontend: graph capture failed: %s", cudaGetErrorString(e));
    return SOS_ERR_CUDA;
  }
  sos_frontend::GraphEntry ge;
  ge.key = key;
  SOS_CUDA(cudaGraphInstantiate(&ge.exec, graph, 0));
  cudaGraphDestroy(graph);
  if (fe->graphs.size() >= 8) {
    cudaGraphExecDestroy(fe->graphs.front().exec);
    fe->graphs.erase(fe->graphs.begin());
  }
  fe->graphs.push_back(ge);
  return SOS_OK;
}

}  // namespace

extern "C" int sos_frontend_create(sos_ctx* ctx, const sos_frontend_config* cfg, const sos_lut_entry* lut,
                                   const uint32_t* hyp, sos_frontend** out) {
  SOS_CHECK_ARG(ctx && cfg && lut && hyp && out, "NULL argument");
  SOS_CHECK_ARG(cfg->batch > 0 && cfg->batch <= 4096, "batch out of range");
  SOS_CHECK_ARG(cfg->channels == 1 || cfg->channels == 3 || cfg->channels == 4, "channels must be 1, 3 or 4");
  SOS_CHECK_ARG(cfg->n_buckets > 0 && cfg->max_feat_per_view > 0 && cfg->max_feat_per_bucket > 0 && cfg->cap > 0, "bad capacity");
  SOS_CHECK_ARG(cfg->n_hyp >= 0, "negative n_hyp");
  SOS_CHECK_ARG(cfg->solver == SOS_SOLVER_ARUN || cfg->solver == SOS_SOLVER_P3P, "unknown solver");
  SOS_CHECK_ARG(cfg->solver != SOS_SOLVER_P3P || cfg->score_mode == SOS_SCORE_BEARING, "the bearing-only solver scores bearings");
  SOS_CUDA(cudaSetDevice(ctx->device));
  sos_frontend* fe = new sos_frontend();
  // private context: same device, but its own scratch arena — the arena address is baked into the captured graphs, so
  // no other caller may regrow it
  fe->ctx = new sos_ctx();
  fe->ctx->device = ctx->device;
  // every failure below goes through sos_frontend_destroy (streams, events, allocations made so far)
#define FE_CUDA(call)                                                                               \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess) {                                                                       \
      sos_set_error("sos_frontend_create: %s failed: %s", #call, cudaGetErrorString(e__));          \
      sos_frontend_destroy(fe);                                                                     \
      return SOS_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)
  // ... and its own compute stream (the caller's may be the legacy default stream, which cannot be captured); every
  // step is fenced against the caller's current stream with events, so the caller sees ordinary stream semantics
  int prio_lo = 0, prio_hi = 0;   // (numerically lower = higher priority)
  FE_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  FE_CUDA(cudaStreamCreateWithPriority(&fe->ctx->stream, cudaStreamNonBlocking, prio_hi));
  fe->ctx->own_stream = true;
  FE_CUDA(cudaEventCreateWithFlags(&fe->ev_in, cudaEventDisableTiming));
  FE_CUDA(cudaEventCreateWithFlags(&fe->ev_out, cudaEventDisableTiming));
  // lowest priority for the remap branch: the chain's kernels take the SMs they can use first, the remap's blocks fill
  // what is left (measured: 1.321 -> 1.303 ms per C2 step; shorter remap blocks to free SMs sooner were slower)
  FE_CUDA(cudaStreamCreateWithPriority(&fe->remap_stream, cudaStreamNonBlocking, prio_lo));
  FE_CUDA(cudaEventCreateWithFlags(&fe->ev_fork, cudaEventDisableTiming));
  FE_CUDA(cudaEventCreateWithFlags(&fe->ev_join, cudaEventDisableTiming));
  fe->ctx->sm_count = ctx->sm_count;
  fe->parent = ctx;
  fe->cfg = *cfg;
  fe->lut = lut;
  fe->hyp = hyp;
  memset(&fe->d, 0, sizeof(fe->d));
  const sos_frontend_config& c = fe->cfg;
  const size_t B = c.batch, F = c.max_feat_per_view, cap = c.cap, S = B * c.n_buckets, slots = B + 1;
  sos_frontend_buffers& d = fe->d;
  int rc = SOS_OK;
#define FE_ALLOC(field, count) \
  if (rc == SOS_OK) rc = fe_alloc(fe, &d.field, (count))
  FE_ALLOC(pano, B * 2 * c.pano_rows * c.pano_cols * c.channels);
  FE_ALLOC(st_q_start, S); FE_ALLOC(st_q_len, S); FE_ALLOC(st_t_start, S); FE_ALLOC(st_t_len, S);
  FE_ALLOC(st_idx0, B * F); FE_ALLOC(st_d0, B * F);
  FE_ALLOC(st_pair_q, B * F); FE_ALLOC(st_pair_t, B * F); FE_ALLOC(st_pair_d, B * F); FE_ALLOC(st_pair_count, S);
  // compacted per-frame store: uv_c holds both views back to back so that descriptor-store rows index it directly
  FE_ALLOC(uv_c, 2 * slots * cap * 2);
  d.uv_top = d.uv_c;
  d.uv_bot = d.uv_c + slots * cap * 2;
  FE_ALLOC(b_top, slots * cap * 3); FE_ALLOC(b_bot, slots * cap * 3); FE_ALLOC(xyz, slots * cap * 3);
  FE_ALLOC(src_top, slots * cap); FE_ALLOC(src_bot, slots * cap); FE_ALLOC(n, slots);
  FE_ALLOC(desc_c, 2 * slots * cap * 8);
  FE_ALLOC(tm_q_start, 2 * B); FE_ALLOC(tm_q_len, 2 * B); FE_ALLOC(tm_t_start, 2 * B); FE_ALLOC(tm_t_len, 2 * B);
  FE_ALLOC(tm_idx0, 2 * slots * cap); FE_ALLOC(tm_d0, 2 * slots * cap);
  FE_ALLOC(tm_pair_q, 2 * slots * cap); FE_ALLOC(tm_pair_t, 2 * slots * cap); FE_ALLOC(tm_pair_d, 2 * slots * cap);
  FE_ALLOC(tm_pair_count, 2 * B);
  FE_ALLOC(p_ref, B * 2 * cap * 3); FE_ALLOC(p_cur, B * 2 * cap * 3); FE_ALLOC(f_cur, B * 2 * cap * 3);
  FE_ALLOC(cam, B * 2 * cap); FE_ALLOC(n_corr, B); FE_ALLOC(n_corr_top, B);
  FE_ALLOC(ransac_pose, B * 12); FE_ALLOC(pose, B * 12); FE_ALLOC(best_hyp, B); FE_ALLOC(best_count, B);
  FE_ALLOC(n_refit, B); FE_ALLOC(inlier_mask, B * 2 * cap); FE_ALLOC(stats, B * 4); FE_ALLOC(refine_stats, B * 4); FE_ALLOC(ref_slot, B); FE_ALLOC(overflow, B);
#undef FE_ALLOC
  if (rc != SOS_OK) {
    sos_frontend_destroy(fe);
    return rc;
  }
  d.batch = c.batch;
  d.cap = c.cap;
  // the zero fills above were queued on the private stream: finish them before anybody can look at the buffers
  FE_CUDA(cudaStreamSynchronize(fe->ctx->stream));
  FE_CUDA(cudaStreamSynchronize(ctx->stream));
#undef FE_CUDA
  *out = fe;
  return SOS_OK;
}

extern "C" int sos_frontend_destroy(sos_frontend* fe) {
  if (!fe) return SOS_OK;
  cudaSetDevice(fe->ctx->device);
  if (fe->ctx->stream) cudaStreamSynchronize(fe->ctx->stream);
  if (fe->copy_stream) cudaStreamSynchronize(fe->copy_stream);
  if (fe->remap_stream) cudaStreamSynchronize(fe->remap_stream);
  for (auto& g : fe->graphs) cudaGraphExecDestroy(g.exec);
  for (void* p : fe->owned) cudaFree(p);
  for (auto& s : fe->stage) {
    if (s.h_poses) cudaFreeHost(s.h_poses);
    if (s.h_stats) cudaFreeHost(s.h_stats);
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.done) cudaEventDestroy(s.done);
  }
  if (fe->copy_stream) cudaStreamDestroy(fe->copy_stream);
  if (fe->ev_in) cudaEventDestroy(fe->ev_in);
  if (fe->ev_out) cudaEventDestroy(fe->ev_out);
  if (fe->ev_fork) cudaEventDestroy(fe->ev_fork);
  if (fe->ev_join) cudaEventDestroy(fe->ev_join);
  if (fe->remap_stream) cudaStreamDestroy(fe->remap_stream);
  if (fe->ctx->own_stream && fe->ctx->stream) cudaStreamDestroy(fe->ctx->stream);
  if (fe->ctx->arena) cudaFree(fe->ctx->arena);
  delete fe->ctx;
  delete fe;
  return SOS_OK;
}

extern "C" int sos_frontend_reset(sos_frontend* fe) {
  SOS_CHECK_ARG(fe, "fe is NULL");
  SOS_CUDA(cudaMemsetAsync(fe->d.n, 0, sizeof(int32_t) * (fe->cfg.batch + 1), fe->ctx->stream));
  SOS_CUDA(cudaStreamSynchronize(fe->ctx->stream));
  return SOS_OK;
}

extern "C" int sos_frontend_set_graph(sos_frontend* fe, int enabled) {
  SOS_CHECK_ARG(fe, "fe is NULL");
  fe->use_graph = enabled != 0;
  return SOS_OK;
}

namespace {
struct Fence {  // order the front-end's own stream after / before the caller's stream
  sos_frontend* fe;
  explicit Fence(sos_frontend* f) : fe(f) {
    cudaEventRecord(fe->ev_in, fe->parent->stream);
    cudaStreamWaitEvent(fe->ctx->stream, fe->ev_in, 0);
  }
  ~Fence() {
    cudaEventRecord(fe->ev_out, fe->ctx->stream);
    cudaStreamWaitEvent(fe->parent->stream, fe->ev_out, 0);
  }
};
}  // namespace

extern "C" int sos_frontend_set_ref_slots(sos_frontend* fe, const int32_t* ref_slots) {
  SOS_CHECK_ARG(fe && ref_slots, "NULL argument");
  SOS_CHECK_ARG(fe->cfg.keyframe_mode, "reference slots are only used in keyframe mode");
  for (int i = 0; i < fe->cfg.batch; ++i)
    SOS_CHECK_ARG(ref_slots[i] >= -1 && ref_slots[i] <= fe->cfg.batch && ref_slots[i] != i + 1, "reference slot out of range");
  SOS_CUDA(cudaSetDevice(fe->ctx->device));
  Fence fence(fe);
  SOS_CUDA(cudaMemcpyAsync(fe->d.ref_slot, ref_slots, sizeof(int32_t) * fe->cfg.batch, cudaMemcpyHostToDevice, fe->ctx->stream));
  SOS_CUDA(cudaStreamSynchronize(fe->ctx->stream));  // ref_slots may be pageable / reused by the caller
  return SOS_OK;
}

extern "C" int sos_frontend_promote(sos_frontend* fe, int slot) {
  SOS_CHECK_ARG(fe, "fe is NULL");
  SOS_CHECK_ARG(slot >= 1 && slot <= fe->cfg.batch, "slot out of range");
  SOS_CUDA(cudaSetDevice(fe->ctx->device));
  Fence fence(fe);
  const int64_t before = fe->ctx->launches;
  const int rc = enqueue_carry(fe, slot);
  fe->parent->launches += fe->ctx->launches - before;
  return rc;
}

extern "C" int sos_frontend_retrack(sos_frontend* fe) {
  SOS_CHECK_ARG(fe, "fe is NULL");
  SOS_CUDA(cudaSetDevice(fe->ctx->device));
  Fence fence(fe);
  const int64_t before = fe->ctx->launches;
  const int rc = enqueue_stage_b(fe);
  fe->parent->launches += fe->ctx->launches - before;
  return rc;
}

extern "C" int sos_frontend_profile_begin(sos_frontend* fe) {
  SOS_CHECK_ARG(fe, "fe is NULL");
  return sos_ctx_profile_begin(fe->ctx);  // run_step launches eagerly while profiling: events cannot be captured
}

extern "C" int sos_frontend_profile_end(sos_frontend* fe, char* names, size_t names_cap, float* ms, int max_n, int* n_out) {
  SOS_CHECK_ARG(fe, "fe is NULL");
  return sos_ctx_profile_end(fe->ctx, names, names_cap, ms, max_n, n_out);
}

extern "C" int sos_frontend_get_buffers(sos_frontend* fe, sos_frontend_buffers* out) {
  SOS_CHECK_ARG(fe && out, "NULL argument");
  *out = fe->d;
  return SOS_OK;
}

extern "C" int sos_frontend_step(sos_frontend* fe, const uint8_t* omni, const float* px_top, const uint32_t* desc_top,
                                 const int32_t* bucket_off_top, const float* px_bot, const uint32_t* desc_bot,
                                 const int32_t* bucket_off_bot) {
  SOS_CHECK_ARG(fe, "fe is NULL");
  SOS_CHECK_ARG(omni && px_top && desc_top && bucket_off_top && px_bot && desc_bot && bucket_off_bot, "NULL input");
  SOS_CUDA(cudaSetDevice(fe->ctx->device));
  // inputs were produced on the caller's stream; results must be visible to it afterwards
  SOS_CUDA(cudaEventRecord(fe->ev_in, fe->parent->stream));
  SOS_CUDA(cudaStreamWaitEvent(fe->ctx->stream, fe->ev_in, 0));
  const int64_t before = fe->ctx->launches;
  const int rc = run_step(fe, omni, px_top, desc_top, bucket_off_top, px_bot, desc_bot, bucket_off_bot);
  fe->parent->launches += fe->ctx->launches - before;
  if (rc) return rc;
  SOS_CUDA(cudaEventRecord(fe->ev_out, fe->ctx->stream));
  SOS_CUDA(cudaStreamWaitEvent(fe->parent->stream, fe->ev_out, 0));
  return SOS_OK;
}

static int ensure_staging(sos_frontend* fe) {
  if (fe->copy_stream) return SOS_OK;
  const sos_frontend_config& c = fe->cfg;
  const size_t B = c.batch, F = c.max_feat_per_view;
  SOS_CUDA(cudaStreamCreateWithFlags(&fe->copy_stream, cudaStreamNonBlocking));
  for (auto& s : fe->stage) {
    int rc = fe_alloc(fe, &s.omni, B * c.src_h * c.src_w * c.channels);
    if (!rc) rc = fe_alloc(fe, &s.px_top, B * F * 2);
    if (!rc) rc = fe_alloc(fe, &s.px_bot, B * F * 2);
    if (!rc) rc = fe_alloc(fe, &s.desc_top, B * F * 8);
    if (!rc) rc = fe_alloc(fe, &s.desc_bot, B * F * 8);
    if (!rc) rc = fe_alloc(fe, &s.boff_top, B * (c.n_buckets + 1));
    if (!rc) rc = fe_alloc(fe, &s.boff_bot, B * (c.n_buckets + 1));
    if (rc) return rc;
    SOS_CUDA(cudaMallocHost((void**)&s.h_poses, B * 12 * sizeof(float)));
    SOS_CUDA(cudaMallocHost((void**)&s.h_stats, B * 4 * sizeof(int32_t)));
    SOS_CUDA(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    SOS_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  }
  SOS_CUDA(cudaStreamSynchronize(fe->ctx->stream));
  // which bytes of an omni image can the remap read?  Per source row a bitmap of the 64-byte blocks some LUT entry touches;
  // rows are grouped into bands of 64, and every run of used blocks of a band (runs closer than 1 KB are merged) becomes one
  // strided copy: the two mirror annuli are uploaded, the corners of the frame and the hole in the middle are not.
  fe->bands.clear();
  fe->omni_bytes_per_frame = (int64_t)c.src_h * c.src_w * c.channels;
  const char* full = getenv("SOS_FULL_UPLOAD");   // A/B switch (tests run both): upload whole images
  if (full == nullptr || full[0] == '0') {
    const int pitch = c.src_w * c.channels;
    const int n_blocks = (pitch + 63) / 64, words = (n_blocks + 63) / 64;
    unsigned long long* dblk = nullptr;
    int rc = fe_alloc(fe, &dblk, (size_t)c.src_h * words);
    if (rc) return rc;
    const size_t n = (size_t)2 * c.pano_rows * c.pano_cols;
    lut_row_blocks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, fe->ctx->stream>>>(fe->lut, n, c.src_h, c.src_w, c.channels, words, dblk);
    SOS_CUDA(cudaStreamSynchronize(fe->ctx->stream));
    std::vector<unsigned long long> blk((size_t)c.src_h * words);
    SOS_CUDA(cudaMemcpy(blk.data(), dblk, blk.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const int band_rows = 64, merge_gap = 16;   // C2 rig: 32 copies, 0.74 of the image (32-row bands: 62 copies, 0.73)
    int64_t total = 0;
    std::vector<sos_frontend::Band> bands;
    std::vector<unsigned long long> acc(words);
    for (int r0 = 0; r0 < c.src_h; r0 += band_rows) {
      const int nr = std::min(band_rows, c.src_h - r0);
      std::fill(acc.begin(), acc.end(), 0ull);
      for (int r = r0; r < r0 + nr; ++r)
        for (int w = 0; w < words; ++w) acc[w] |= blk[(size_t)r * words + w];
      int run_lo = -1, run_hi = -1;   // current run of used blocks [run_lo, run_hi]
      auto flush = [&]() {
        if (run_lo < 0) return;
        const int x0 = run_lo * 64, x1 = std::min(pitch, (run_hi + 1) * 64);
        bands.push_back({r0, nr, x0, x1 - x0});
        total += (int64_t)nr * (x1 - x0);
        run_lo = run_hi = -1;
      };
      for (int b = 0; b < n_blocks; ++b) {
        if (!((acc[b >> 6] >> (b & 63)) & 1ull)) continue;
        if (run_lo >= 0 && b - run_hi > merge_gap) flush();
        if (run_lo < 0) run_lo = b;
        run_hi = b;
      }
      flush();
    }
    if (total < fe->omni_bytes_per_frame * 97 / 100) {   // otherwise one plain copy is cheaper than the band copies
      fe->bands = bands;
      fe->omni_bytes_per_frame = total;
    }
  }
  return SOS_OK;
}

extern "C" int sos_frontend_host_bytes(sos_frontend* fe, int64_t* h2d_bytes_per_step, int64_t* d2h_bytes_per_step) {
  SOS_CHECK_ARG(fe, "fe is NULL");
  SOS_CUDA(cudaSetDevice(fe->ctx->device));
  const int rc = ensure_staging(fe);
  if (rc) return rc;
  const sos_frontend_config& c = fe->cfg;
  const int64_t B = c.batch, F = c.max_feat_per_view;
  if (h2d_bytes_per_step)
    *h2d_bytes_per_step = B * (fe->omni_bytes_per_frame + 2 * F * 2 * 4 + 2 * F * 32 + 2 * (c.n_buckets + 1) * 4);
  if (d2h_bytes_per_step) *d2h_bytes_per_step = B * (12 * 4 + 4 * 4);
  return SOS_OK;
}

extern "C" int sos_frontend_submit_host(sos_frontend* fe, const uint8_t* omni, const float* px_top,
                                        const uint32_t* desc_top, const int32_t* bucket_off_top, const float* px_bot,
                                        const uint32_t* desc_bot, const int32_t* bucket_off_bot, int* ticket) {
  SOS_CHECK_ARG(fe && ticket, "NULL argument");
  SOS_CHECK_ARG(omni && px_top && desc_top && bucket_off_top && px_bot && desc_bot && bucket_off_bot, "NULL input");
  SOS_CUDA(cudaSetDevice(fe->ctx->device));
  int rc = ensure_staging(fe);
  if (rc) return rc;
  const int k = fe->next_stage;
  sos_frontend::Stage& s = fe->stage[k];
  if (s.in_flight) {
    sos_set_error("sos_frontend_submit_host: both staging slots are in flight; call sos_frontend_wait_host first");
    return SOS_ERR_INVALID;
  }
  const sos_frontend_config& c = fe->cfg;
  const size_t B = c.batch, F = c.max_feat_per_view;
  cudaStream_t cs = fe->copy_stream;
  // H2D on the copy stream so that it overlaps the kernels of the previous step
  if (fe->bands.empty()) {
    SOS_CUDA(cudaMemcpyAsync(s.omni, omni, B * c.src_h * c.src_w * c.channels, cudaMemcpyHostToDevice, cs));
  } else {
    // one strided copy per band of rows, covering that band in every frame of the batch (depth = frames)
    const size_t pitch = (size_t)c.src_w * c.channels;
    for (const auto& bd : fe->bands) {
      cudaMemcpy3DParms p = {};
      p.srcPtr = make_cudaPitchedPtr((void*)omni, pitch, pitch, c.src_h);
      p.dstPtr = make_cudaPitchedPtr((void*)s.omni, pitch, pitch, c.src_h);
      p.srcPos = make_cudaPos(bd.x0_bytes, bd.row0, 0);
      p.dstPos = make_cudaPos(bd.x0_bytes, bd.row0, 0);
      p.extent = make_cudaExtent(bd.width_bytes, bd.nrows, B);
      p.kind = cudaMemcpyHostToDevice;
      SOS_CUDA(cudaMemcpy3DAsync(&p, cs));
    }
  }
  SOS_CUDA(cudaMemcpyAsync(s.px_top, px_top, B * F * 2 * sizeof(float), cudaMemcpyHostToDevice, cs));
  SOS_CUDA(cudaMemcpyAsync(s.px_bot, px_bot, B * F * 2 * sizeof(float), cudaMemcpyHostToDevice, cs));
  SOS_CUDA(cudaMemcpyAsync(s.desc_top, desc_top, B * F * 32, cudaMemcpyHostToDevice, cs));
  SOS_CUDA(cudaMemcpyAsync(s.desc_bot, desc_bot, B * F * 32, cudaMemcpyHostToDevice, cs));
  SOS_CUDA(cudaMemcpyAsync(s.boff_top, bucket_off_top, B * (c.n_buckets + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
  SOS_CUDA(cudaMemcpyAsync(s.boff_bot, bucket_off_bot, B * (c.n_buckets + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
  SOS_CUDA(cudaEventRecord(s.copied, cs));
  SOS_CUDA(cudaStreamWaitEvent(fe->ctx->stream, s.copied, 0));
  const int64_t before = fe->ctx->launches;
  rc = run_step(fe, s.omni, s.px_top, s.desc_top, s.boff_top, s.px_bot, s.desc_bot, s.boff_bot);
  fe->parent->launches += fe->ctx->launches - before;
  if (rc) return rc;
  SOS_CUDA(cudaMemcpyAsync(s.h_poses, fe->d.pose, B * 12 * sizeof(float), cudaMemcpyDeviceToHost, fe->ctx->stream));
  SOS_CUDA(cudaMemcpyAsync(s.h_stats, fe->d.stats, B * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, fe->ctx->stream));
  SOS_CUDA(cudaEventRecord(s.done, fe->ctx->stream));
  // slot reuse is safe: sos_frontend_wait_host(ticket) must return before this slot is submitted to again
  s.in_flight = true;
  *ticket = k;
  fe->next_stage = (k + 1) % sos_frontend::DEPTH;
  return SOS_OK;
}

extern "C" int sos_frontend_wait_host(sos_frontend* fe, int ticket, float* poses, int32_t* stats) {
  SOS_CHECK_ARG(fe, "fe is NULL");
  SOS_CHECK_ARG(ticket >= 0 && ticket < sos_frontend::DEPTH && fe->stage[ticket].in_flight, "unknown ticket");
  sos_frontend::Stage& s = fe->stage[ticket];
  SOS_CUDA(cudaEventSynchronize(s.done));
  const size_t B = fe->cfg.batch;
  if (poses) memcpy(poses, s.h_poses, B * 12 * sizeof(float));
  if (stats) memcpy(stats, s.h_stats, B * 4 * sizeof(int32_t));
  s.in_flight = false;
  return SOS_OK;
}

extern "C" int sos_frontend_step_host(sos_frontend* fe, const uint8_t* omni, const float* px_top, const uint32_t* desc_top,
                                      const int32_t* bucket_off_top, const float* px_bot, const uint32_t* desc_bot,
                                      const int32_t* bucket_off_bot, float* poses, int32_t* stats) {
  int ticket = -1;
  int rc = sos_frontend_submit_host(fe, omni, px_top, desc_top, bucket_off_top, px_bot, desc_bot, bucket_off_bot, &ticket);
  if (rc) return rc;
  return sos_frontend_wait_host(fe, ticket, poses, stats);
}
