// Non-linear pose refinement on the RANSAC inliers.
//
// replaces: pyopengv.absolute_pose_noncentral_optimize_nonlinear / absolute_pose_optimize_nonlinear
//           (pose_est_tools.py:830, :937).  OpenGV is not vendored; its published algorithm is Levenberg-Marquardt on
//           x = (t, rotation) with one residual per correspondence, r_i = 1 - f_i . normalize(Rc^T (R^T (p_i - t) - tc)),
//           i.e. the same bearing residual the RANSAC scores (pose_est_tools.py:150-203).  Minimised here: sum r_i^2.
//
// One thread-block CLUSTER per problem (S CTAs, S <= 8): every CTA accumulates cost, J^T J (21) and J^T r (6) over its
// slice of the correspondences in float64, the partials meet in rank 0 through distributed shared memory, rank 0 solves
// the damped 6x6 system and pushes the trial pose + control word into every CTA's shared memory.  Two cluster barriers
// per iteration, no global-memory round trips, no host involvement: the whole refinement is one launch in the CUDA graph.
//
// Rotation updates are local, R <- R exp([d]x); the minimiser does not depend on the parametrisation, so the result can be
// compared with any LM on OpenGV's Cayley parameters (tests do, with scipy's MINPACK driver).
#include <cooperative_groups.h>
#include <math_constants.h>

#include "sos_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int RF_THREADS = 256;
constexpr int RF_MAX_CAMS = 2;
constexpr int RF_NACC = 28;          // 21 (upper triangle of J^T J) + 6 (J^T r) + 1 (sum r^2)
constexpr int RF_MAX_CLUSTER = 8;
constexpr double RF_STEP_TOL = 1e-10;
constexpr int RF_CACHE_ROWS = 8;               // rows per thread staged in shared memory (56 KB); the rest re-read from HBM
constexpr double RF_SECOND_ORDER_MAX = 0.05;  // 1 - cos(18 deg)

struct RefineRig {
  double Rt[RF_MAX_CAMS][12];
};

struct RefineState {                 // lives in every CTA's shared memory; written by rank 0 of the cluster
  double R[9], t[3];                 // trial pose
  int stop;                          // 1 = leave the loop
  int pad;
};

__device__ inline void rodrigues(const double* d, double* E) {  // E = exp([d]x) = I + a [d]x + b [d]x^2
  const double th2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  double a, b;
  if (th2 < 0.25) {   // LM steps are small: Taylor series of sin(t)/t and (1 - cos t)/t^2 in t^2, < 1e-17 at t = 0.5
    a = 1.0 + th2 * (-1.0 / 6 + th2 * (1.0 / 120 + th2 * (-1.0 / 5040 + th2 * (1.0 / 362880 + th2 * (-1.0 / 39916800 +
        th2 * (1.0 / 6227020800.0 + th2 * (-1.0 / 1307674368000.0)))))));
    b = 0.5 + th2 * (-1.0 / 24 + th2 * (1.0 / 720 + th2 * (-1.0 / 40320 + th2 * (1.0 / 3628800 + th2 * (-1.0 / 479001600 +
        th2 * (1.0 / 87178291200.0 + th2 * (-1.0 / 20922789888000.0)))))));
  } else {
    const double th = sqrt(th2);
    a = sin(th) / th;
    b = (1.0 - cos(th)) / th2;
  }
  const double x = d[0], y = d[1], z = d[2];
  E[0] = 1.0 - b * (y * y + z * z); E[1] = -a * z + b * x * y;        E[2] = a * y + b * x * z;
  E[3] = a * z + b * x * y;         E[4] = 1.0 - b * (x * x + z * z); E[5] = -a * x + b * y * z;
  E[6] = -a * y + b * x * z;        E[7] = a * x + b * y * z;         E[8] = 1.0 - b * (x * x + y * y);
}

// Solve (A + lambda diag(A)) x = -g for the symmetric 6x6 A given as its upper triangle; false if not positive definite.
__device__ __forceinline__ bool solve_damped(const double* Au, const double* g, double lambda, double* x) {
  double L[6][6];   // every loop below is fully unrolled: L stays in registers
  int k = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j) { L[i][j] = Au[k]; L[j][i] = Au[k]; ++k; }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double dgl = L[i][i];
    L[i][i] = dgl + lambda * (dgl > 0.0 ? dgl : 1.0);
  }
  bool pd = true;
  double dinv[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {       // Cholesky, lower triangle in place; one rsqrt per column, no divisions
    double s = L[j][j];
#pragma unroll
    for (int m = 0; m < j; ++m) s -= L[j][m] * L[j][m];
    pd = pd && (s > 0.0);
    const double di = rsqrt(s);
    dinv[j] = di;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = L[i][j];
#pragma unroll
      for (int m = 0; m < j; ++m) v -= L[i][m] * L[j][m];
      L[i][j] = v * di;
    }
  }
  if (!pd) return false;
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double v = -g[i];
#pragma unroll
    for (int m = 0; m < i; ++m) v -= L[i][m] * y[m];
    y[i] = v * dinv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double v = y[i];
#pragma unroll
    for (int m = i + 1; m < 6; ++m) v -= L[m][i] * x[m];
    x[i] = v * dinv[i];
  }
  return true;
}

__global__ void __launch_bounds__(RF_THREADS)
refine_kernel(const float* __restrict__ p_ref, const float* __restrict__ f_cur, const uint8_t* __restrict__ cam,
              const uint8_t* __restrict__ mask, const int32_t* __restrict__ n_arr, int cap,
              const __grid_constant__ RefineRig rig, const float* __restrict__ pose_in, int max_iters,
              float* __restrict__ pose_out, double* __restrict__ pose_out64, double* __restrict__ stats) {
  cg::cluster_group cluster = cg::this_cluster();
  const int S = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  extern __shared__ float cache[];                       // [RF_CACHE_ROWS][7][RF_THREADS]
  __shared__ RefineState st;
  __shared__ double partial[RF_NACC];                    // this CTA's sums; read by rank 0 through DSMEM
  __shared__ double wred[RF_THREADS / 32][RF_NACC];
  __shared__ double Msh[RF_MAX_CAMS][9];
  __shared__ float Mf[RF_MAX_CAMS][9], Rcf[RF_MAX_CAMS][9];   // float copies for the float32 Hessian
  // rank 0 only
  __shared__ double curR[9], curT[3], curAG[27], tot[RF_NACC];
  __shared__ double cur_cost, lambda, first_cost, nu, pred, step_max;
  __shared__ int have_cur, iters_done, n_used_sh;

  const int n = min(n_arr[b], cap);
  const size_t base = (size_t)b * cap;

  if (tid == 0) {
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) st.R[r * 3 + c] = (double)pose_in[(size_t)b * 12 + r * 4 + c];
      st.t[r] = (double)pose_in[(size_t)b * 12 + r * 4 + 3];
    }
    {  // pose_in is float32: re-orthonormalise (Gram-Schmidt on the rows) so that R stays a rotation to 1e-16
      double* R = st.R;
      double nrm = 1.0 / sqrt(R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
      for (int i = 0; i < 3; ++i) R[i] *= nrm;
      const double d = R[3] * R[0] + R[4] * R[1] + R[5] * R[2];
      for (int i = 0; i < 3; ++i) R[3 + i] -= d * R[i];
      nrm = 1.0 / sqrt(R[3] * R[3] + R[4] * R[4] + R[5] * R[5]);
      for (int i = 0; i < 3; ++i) R[3 + i] *= nrm;
      R[6] = R[1] * R[5] - R[2] * R[4];
      R[7] = R[2] * R[3] - R[0] * R[5];
      R[8] = R[0] * R[4] - R[1] * R[3];
    }
    st.stop = 0;
    if (rank == 0) { have_cur = 0; iters_done = 0; lambda = 1e-4; nu = 2.0; pred = 1.0; step_max = 1.0; n_used_sh = 0; }
  }
  // stage this thread's rows (the same in every evaluation): SoA in shared memory, slot 6 = camera index or -1
  const int rows_per_thread = (n + S * RF_THREADS - 1) / (S * RF_THREADS);
  for (int k = 0; k < min(rows_per_thread, RF_CACHE_ROWS); ++k) {
    const int j = (k * S + rank) * RF_THREADS + tid;
    const bool live = j < n && (!mask || mask[base + j]);
    const size_t o = (base + (live ? j : 0)) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      cache[(k * 7 + c) * RF_THREADS + tid] = live ? p_ref[o + c] : 0.f;
      cache[(k * 7 + 3 + c) * RF_THREADS + tid] = live ? f_cur[o + c] : 0.f;
    }
    const int ci = live ? (cam ? min((int)cam[base + j], RF_MAX_CAMS - 1) : 0) : -1;
    cache[(k * 7 + 6) * RF_THREADS + tid] = __int_as_float(ci);
  }
  cluster.sync();  // every CTA of the cluster is resident and rank 0's counters are initialised before any DSMEM access

  for (int it = 0;; ++it) {
    // ---- accumulate over this CTA's slice at the trial pose st ----
    if (tid < RF_MAX_CAMS * 9) {  // M_c = R Rc, so that Rc^T R^T = M_c^T
      const int c = tid / 9, r = (tid % 9) / 3, k = tid % 3;
      const double* Rt = rig.Rt[c];
      const double v = st.R[r * 3] * Rt[k] + st.R[r * 3 + 1] * Rt[4 + k] + st.R[r * 3 + 2] * Rt[8 + k];
      Msh[c][r * 3 + k] = v;
      Mf[c][r * 3 + k] = (float)v;
      Rcf[c][r * 3 + k] = (float)Rt[r * 4 + k];
    }
    __syncthreads();
    double R[9], t[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = st.R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = st.t[i];
    // float64: cost (1) and gradient J^T r (6) — they define the minimiser.  float32: the model Hessian (21) — it only
    // shapes the step, so single precision costs convergence nothing and takes the 84 FMAs per row off the FP64 pipe.
    double accg[7];
    float acch[21];
#pragma unroll
    for (int i = 0; i < 7; ++i) accg[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 21; ++i) acch[i] = 0.f;
    int used = 0;
    for (int k = 0; k < rows_per_thread; ++k) {
      float pf[6];
      int ci;
      if (k < RF_CACHE_ROWS) {          // rows of this thread are the same in every evaluation: staged once in smem
        ci = __float_as_int(cache[(k * 7 + 6) * RF_THREADS + tid]);
        if (ci < 0) continue;
#pragma unroll
        for (int c = 0; c < 6; ++c) pf[c] = cache[(k * 7 + c) * RF_THREADS + tid];
      } else {
        const int j = (k * S + rank) * RF_THREADS + tid;
        if (j >= n || (mask && !mask[base + j])) continue;
        const size_t o = (base + j) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) { pf[c] = p_ref[o + c]; pf[3 + c] = f_cur[o + c]; }
        ci = cam ? min((int)cam[base + j], RF_MAX_CAMS - 1) : 0;
      }
      const double* Rt = rig.Rt[ci];
      const double d0 = (double)pf[0] - t[0], d1 = (double)pf[1] - t[1], d2 = (double)pf[2] - t[2];
      const double q[3] = {R[0] * d0 + R[3] * d1 + R[6] * d2, R[1] * d0 + R[4] * d1 + R[7] * d2,
                           R[2] * d0 + R[5] * d1 + R[8] * d2};                      // R^T (p - t)
      const double e0 = q[0] - Rt[3], e1 = q[1] - Rt[7], e2 = q[2] - Rt[11];
      const double y[3] = {Rt[0] * e0 + Rt[4] * e1 + Rt[8] * e2, Rt[1] * e0 + Rt[5] * e1 + Rt[9] * e2,
                           Rt[2] * e0 + Rt[6] * e1 + Rt[10] * e2};                  // Rc^T (q - tc)
      const double ny2 = y[0] * y[0] + y[1] * y[1] + y[2] * y[2];
      if (!(ny2 > 0.0)) continue;
      const double inv = rsqrt(ny2);
      const double nv[3] = {y[0] * inv, y[1] * inv, y[2] * inv};
      const double f[3] = {(double)pf[3], (double)pf[4], (double)pf[5]};
      const double cosv = f[0] * nv[0] + f[1] * nv[1] + f[2] * nv[2];
      const double r = 1.0 - cosv;
      // float64 Jacobian row: dr/dy = g = -(f - cos n)/|y|;  h = Rc g;  dq = -R^T dt + [q]x d  ->  J = (-R h, h x q)
      const double g0 = -(f[0] - cosv * nv[0]) * inv, g1 = -(f[1] - cosv * nv[1]) * inv, g2 = -(f[2] - cosv * nv[2]) * inv;
      const double h[3] = {Rt[0] * g0 + Rt[1] * g1 + Rt[2] * g2, Rt[4] * g0 + Rt[5] * g1 + Rt[6] * g2,
                           Rt[8] * g0 + Rt[9] * g1 + Rt[10] * g2};
      double J[6];
      J[0] = -(R[0] * h[0] + R[1] * h[1] + R[2] * h[2]);
      J[1] = -(R[3] * h[0] + R[4] * h[1] + R[5] * h[2]);
      J[2] = -(R[6] * h[0] + R[7] * h[1] + R[8] * h[2]);
      J[3] = h[1] * q[2] - h[2] * q[1];
      J[4] = h[2] * q[0] - h[0] * q[2];
      J[5] = h[0] * q[1] - h[1] * q[0];
#pragma unroll
      for (int a = 0; a < 6; ++a) accg[a] += J[a] * r;
      accg[6] += r * r;
      // float32 model Hessian J J^T + r N^T N with N = dn/dx = (I - n n^T) Dy / |y|, Dy = (-M^T | Rc^T [q]x).
      // The second term is the part of r * d2r/dx2 that Gauss-Newton drops; because r = |f - n|^2 / 2 has a vanishing
      // gradient at a perfect fit it is as large as J J^T and without it LM only converges linearly.  It is valid
      // for small residuals, so rows with r >= RF_SECOND_ORDER_MAX get plain Gauss-Newton.
      const float qf[3] = {(float)q[0], (float)q[1], (float)q[2]};
      const float nf[3] = {(float)nv[0], (float)nv[1], (float)nv[2]};
      const float invf = (float)inv;
      float N[3][6];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        N[i][0] = -Mf[ci][i];
        N[i][1] = -Mf[ci][3 + i];
        N[i][2] = -Mf[ci][6 + i];
        const float a0 = Rcf[ci][i], a1 = Rcf[ci][3 + i], a2 = Rcf[ci][6 + i];   // column i of Rc
        N[i][3] = a1 * qf[2] - a2 * qf[1];
        N[i][4] = a2 * qf[0] - a0 * qf[2];
        N[i][5] = a0 * qf[1] - a1 * qf[0];
      }
      float Jf[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const float nd = nf[0] * N[0][c] + nf[1] * N[1][c] + nf[2] * N[2][c];
#pragma unroll
        for (int i = 0; i < 3; ++i) N[i][c] = (N[i][c] - nf[i] * nd) * invf;
        Jf[c] = (float)J[c];
      }
      const float w = r < RF_SECOND_ORDER_MAX ? (float)r : 0.f;
      int kk = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = a; c < 6; ++c)
          acch[kk++] += Jf[a] * Jf[c] + w * (N[0][a] * N[0][c] + N[1][a] * N[1][c] + N[2][a] * N[2][c]);
      ++used;
    }
    double acc[RF_NACC];
#pragma unroll
    for (int i = 0; i < 21; ++i) acc[i] = (double)acch[i];
#pragma unroll
    for (int i = 0; i < 7; ++i) acc[21 + i] = accg[i];
#pragma unroll
    for (int i = 0; i < RF_NACC; ++i) {
      double v = acc[i];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
      if (lane == 0) wred[warp][i] = v;
    }
    if (it == 0) {
      for (int off = 16; off > 0; off >>= 1) used += __shfl_xor_sync(0xFFFFFFFFu, used, off);
      if (lane == 0) atomicAdd(cluster.map_shared_rank(&n_used_sh, 0), used);
    }
    __syncthreads();
    if (tid < RF_NACC) {
      double v = 0.0;
      for (int w = 0; w < RF_THREADS / 32; ++w) v += wred[w][tid];
      partial[tid] = v;
    }
    cluster.sync();

    // ---- rank 0, warp 0: gather, decide, solve, scatter ----
    // A lone thread issues one dependent instruction every ~8 cycles, so the control step is kept short: every lane of
    // the warp runs the scalar logic redundantly (no divergence, no broadcasts), copies and DSMEM traffic are spread
    // over the lanes, and the rotation update uses series instead of sin/cos.
    if (rank == 0 && warp == 0) {
      if (lane < RF_NACC) {               // one lane per accumulator, fixed rank order -> deterministic sums
        double v = 0.0;
        for (int s = 0; s < S; ++s) v += cluster.map_shared_rank(partial, s)[lane];
        tot[lane] = v;
      }
      __syncwarp();
      const double cost = tot[27];
      const double prev_cost = cur_cost;
      double lam = lambda, nuv = nu;
      int stop = 0;
      bool accepted = false;
      if (!have_cur) {
        accepted = true;
      } else if (cost < prev_cost) {
        accepted = true;
        if (prev_cost - cost <= 1e-14 * prev_cost) stop = 1;   // converged: no measurable decrease left
        const double rho = (prev_cost - cost) / pred;           // Nielsen's gain-ratio damping update
        const double g3 = (2.0 * rho - 1.0) * (2.0 * rho - 1.0) * (2.0 * rho - 1.0);
        lam = fmax(lam * fmax(1.0 / 3.0, 1.0 - g3), 1e-12);
        nuv = 2.0;
      } else {
        if (step_max < 1e-9) stop = 1;                          // rejected although the step is at rounding level
        lam *= nuv;
        nuv *= 2.0;
        if (lam > 1e12) stop = 1;
      }
      if (it + 1 >= max_iters || n_used_sh < 6 || !(cost == cost)) stop = 1;
      const bool first = !have_cur;
      __syncwarp();                        // everyone has read the scalars that lane 0 rewrites below
      if (accepted) {
        if (lane < 9) curR[lane] = st.R[lane];
        else if (lane < 12) curT[lane - 9] = st.t[lane - 9];
        if (lane < 27) curAG[lane] = tot[lane];
      }
      __syncwarp();
      double A[21], g[6], dx[6];
#pragma unroll
      for (int i = 0; i < 21; ++i) A[i] = curAG[i];
#pragma unroll
      for (int i = 0; i < 6; ++i) g[i] = curAG[21 + i];
      double new_pred = pred, new_step = step_max;
      double nR[9], nT[3];
      if (!stop) {
        bool ok = false;
        for (int tries = 0; tries < 40 && !ok; ++tries) {
          ok = solve_damped(A, g, lam, dx);
          if (!ok) lam *= 10.0;
        }
        double m = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) m = fmax(m, fabs(dx[i]));
        if (!ok || !(m == m) || m < RF_STEP_TOL) {   // the proposed step is below what the float32 pose output resolves
          stop = 1;
        } else {
          new_step = m;
          double quad = 0.0, lin = 0.0;   // predicted decrease of sum r^2 under the model: -(2 g.dx + dx^T A dx)
          int k = 0;
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            lin += g[a] * dx[a];
#pragma unroll
            for (int c = a; c < 6; ++c) { quad += (a == c ? 1.0 : 2.0) * A[k] * dx[a] * dx[c]; ++k; }
          }
          new_pred = -(2.0 * lin + quad);
          if (!(new_pred > 0.0)) new_pred = 1e-300;
          double E[9];
          rodrigues(dx + 3, E);
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
              nR[r * 3 + c] = curR[r * 3] * E[c] + curR[r * 3 + 1] * E[3 + c] + curR[r * 3 + 2] * E[6 + c];
#pragma unroll
          for (int i = 0; i < 3; ++i) nT[i] = curT[i] + dx[i];
        }
      }
      if (stop) {
#pragma unroll
        for (int i = 0; i < 9; ++i) nR[i] = curR[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) nT[i] = curT[i];
      }
      if (lane == 0) {
        if (first) first_cost = cost;
        if (accepted) { cur_cost = cost; have_cur = 1; }
        lambda = lam; nu = nuv; pred = new_pred; step_max = new_step;
        iters_done = it + 1;
      }
      // lane l < 13 owns element l of the state (R 0..8, t 9..11, stop 12) and pushes it into every CTA of the cluster
      double mine = 0.0;
#pragma unroll
      for (int i = 0; i < 9; ++i) if (lane == i) mine = nR[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) if (lane == 9 + i) mine = nT[i];
      if (lane < 13)
        for (int s = 0; s < S; ++s) {
          RefineState* q = cluster.map_shared_rank(&st, s);
          if (lane < 9) q->R[lane] = mine;
          else if (lane < 12) q->t[lane - 9] = mine;
          else q->stop = stop;
        }
    }
    cluster.sync();
    if (st.stop) break;
  }

  if (rank == 0 && tid == 0) {
    const bool good = n_used_sh >= 6 && cur_cost == cur_cost;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) {
        const double v = good ? st.R[r * 3 + c] : (double)pose_in[(size_t)b * 12 + r * 4 + c];
        if (pose_out) pose_out[(size_t)b * 12 + r * 4 + c] = (float)v;
        if (pose_out64) pose_out64[(size_t)b * 12 + r * 4 + c] = v;
      }
      const double v = good ? st.t[r] : (double)pose_in[(size_t)b * 12 + r * 4 + 3];
      if (pose_out) pose_out[(size_t)b * 12 + r * 4 + 3] = (float)v;
      if (pose_out64) pose_out64[(size_t)b * 12 + r * 4 + 3] = v;
    }
    if (stats) {
      stats[(size_t)b * 4 + 0] = first_cost;
      stats[(size_t)b * 4 + 1] = cur_cost;
      stats[(size_t)b * 4 + 2] = (double)iters_done;
      stats[(size_t)b * 4 + 3] = (double)n_used_sh;
    }
  }
}

}  // namespace

extern "C" int sos_refine_pose(sos_ctx* ctx, const float* p_ref, const float* f_cur, const uint8_t* cam,
                               const uint8_t* inlier_mask, const int32_t* n, int n_problems, int cap, const double* rig,
                               int n_cams, const float* pose_in, int max_iters, int cluster_size, float* pose_out,
                               double* pose_out64, double* stats) {
  SOS_CHECK_ARG(ctx, "ctx is NULL");
  SOS_CHECK_ARG(n_problems >= 0 && cap >= 0, "negative size");
  SOS_CHECK_ARG(n_cams >= 0 && n_cams <= RF_MAX_CAMS, "at most 2 cameras in the rig");
  SOS_CHECK_ARG(max_iters >= 1, "max_iters must be >= 1");
  SOS_CHECK_ARG(cluster_size >= 0 && cluster_size <= RF_MAX_CLUSTER, "cluster_size must be 0 (auto) .. 8");
  SOS_CHECK_ARG(n_problems <= 65535, "too many problems");
  if (n_problems == 0) return SOS_OK;
  SOS_CHECK_ARG(p_ref && f_cur && n && pose_in && (pose_out || pose_out64), "NULL array");
  SOS_CHECK_ARG(n_cams == 0 || rig, "rig is NULL");
  SOS_CUDA(cudaSetDevice(ctx->device));
  RefineRig rr;
  for (int c = 0; c < RF_MAX_CAMS; ++c)
    for (int i = 0; i < 12; ++i) rr.Rt[c][i] = (c < n_cams) ? rig[c * 12 + i] : ((i == 0 || i == 5 || i == 10) ? 1.0 : 0.0);
  int S = cluster_size;
  if (S == 0) S = cap >= 4096 ? 8 : (cap >= 1024 ? 4 : 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S, n_problems, 1);
  cfg.blockDim = dim3(RF_THREADS, 1, 1);
  cfg.dynamicSmemBytes = RF_CACHE_ROWS * 7 * RF_THREADS * sizeof(float);
  SOS_CUDA(cudaFuncSetAttribute(refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SOS_CUDA(cudaLaunchKernelEx(&cfg, refine_kernel, p_ref, f_cur, cam, inlier_mask, n, cap, rr, pose_in, max_iters, pose_out,
                              pose_out64, stats));
  SOS_LAUNCHED(ctx);
  return SOS_OK;
}
