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
constexpr double RF_SECOND_ORDER_MAX = 0.05;  // 1 - cos(18 deg)

struct RefineRig {
  double Rt[RF_MAX_CAMS][12];
};

struct RefineState {                 // lives in every CTA's shared memory; written by rank 0 of the cluster
  double R[9], t[3];                 // trial pose
  int stop;                          // 1 = leave the loop
  int pad;
};

__device__ inline void rodrigues(const double* d, double* E) {  // E = exp([d]x)
  const double th2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  double a, b;                        // E = I + a [d]x + b [d]x^2
  if (th2 < 1e-16) {
    a = 1.0 - th2 / 6.0;
    b = 0.5 - th2 / 24.0;
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
__device__ bool solve_damped(const double* Au, const double* g, double lambda, double* x) {
  double L[6][6];
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) { L[i][j] = Au[k]; L[j][i] = Au[k]; ++k; }
  for (int i = 0; i < 6; ++i) {
    const double dgl = L[i][i];
    L[i][i] = dgl + lambda * (dgl > 0.0 ? dgl : 1.0);
  }
  for (int j = 0; j < 6; ++j) {       // Cholesky, lower triangle in place
    double s = L[j][j];
    for (int m = 0; m < j; ++m) s -= L[j][m] * L[j][m];
    if (!(s > 0.0)) return false;
    const double d = sqrt(s);
    L[j][j] = d;
    for (int i = j + 1; i < 6; ++i) {
      double v = L[i][j];
      for (int m = 0; m < j; ++m) v -= L[i][m] * L[j][m];
      L[i][j] = v / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double v = -g[i];
    for (int m = 0; m < i; ++m) v -= L[i][m] * y[m];
    y[i] = v / L[i][i];
  }
  for (int i = 5; i >= 0; --i) {
    double v = y[i];
    for (int m = i + 1; m < 6; ++m) v -= L[m][i] * x[m];
    x[i] = v / L[i][i];
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

  __shared__ RefineState st;
  __shared__ double partial[RF_NACC];                    // this CTA's sums; read by rank 0 through DSMEM
  __shared__ double wred[RF_THREADS / 32][RF_NACC];
  __shared__ double Msh[RF_MAX_CAMS][9];
  // rank 0 only
  __shared__ double curR[9], curT[3], curA[21], curG[6];
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
  cluster.sync();  // every CTA of the cluster is resident and rank 0's counters are initialised before any DSMEM access

  for (int it = 0;; ++it) {
    // ---- accumulate over this CTA's slice at the trial pose st ----
    if (tid < RF_MAX_CAMS * 9) {  // M_c = R Rc, so that Rc^T R^T = M_c^T
      const int c = tid / 9, r = (tid % 9) / 3, k = tid % 3;
      const double* Rt = rig.Rt[c];
      Msh[c][r * 3 + k] = st.R[r * 3] * Rt[k] + st.R[r * 3 + 1] * Rt[4 + k] + st.R[r * 3 + 2] * Rt[8 + k];
    }
    __syncthreads();
    double R[9], t[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = st.R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = st.t[i];
    double acc[RF_NACC];
#pragma unroll
    for (int i = 0; i < RF_NACC; ++i) acc[i] = 0.0;
    int used = 0;
    for (int j = rank * RF_THREADS + tid; j < n; j += S * RF_THREADS) {
      if (mask && !mask[base + j]) continue;
      const size_t o = (base + j) * 3;
      const int ci = cam ? min((int)cam[base + j], RF_MAX_CAMS - 1) : 0;
      const double* Rt = rig.Rt[ci];
      const double* M = Msh[ci];
      const double d0 = (double)p_ref[o] - t[0], d1 = (double)p_ref[o + 1] - t[1], d2 = (double)p_ref[o + 2] - t[2];
      const double q[3] = {R[0] * d0 + R[3] * d1 + R[6] * d2, R[1] * d0 + R[4] * d1 + R[7] * d2,
                           R[2] * d0 + R[5] * d1 + R[8] * d2};                      // R^T (p - t)
      const double e0 = q[0] - Rt[3], e1 = q[1] - Rt[7], e2 = q[2] - Rt[11];
      const double y[3] = {Rt[0] * e0 + Rt[4] * e1 + Rt[8] * e2, Rt[1] * e0 + Rt[5] * e1 + Rt[9] * e2,
                           Rt[2] * e0 + Rt[6] * e1 + Rt[10] * e2};                  // Rc^T (q - tc)
      const double ny2 = y[0] * y[0] + y[1] * y[1] + y[2] * y[2];
      if (!(ny2 > 0.0)) continue;
      const double inv = 1.0 / sqrt(ny2);
      const double nv[3] = {y[0] * inv, y[1] * inv, y[2] * inv};
      const double f[3] = {(double)f_cur[o], (double)f_cur[o + 1], (double)f_cur[o + 2]};
      const double r = 1.0 - (f[0] * nv[0] + f[1] * nv[1] + f[2] * nv[2]);
      // Dy = dy/dx (3x6): translation columns -M^T (dq = -R^T dt), rotation columns Rc^T [q]x (dq = [q]x d)
      double N[3][6];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        N[i][0] = -M[i];          // -(M^T)[i][0] = -M[0][i]
        N[i][1] = -M[3 + i];
        N[i][2] = -M[6 + i];
        // Rc^T [q]x: row i = (column i of Rc) x ... : ([q]x)^T Rc[:,i] = -q x Rc[:,i] -> row_i = Rc[:,i] x q
        const double a0 = Rt[i], a1 = Rt[4 + i], a2 = Rt[8 + i];
        N[i][3] = a1 * q[2] - a2 * q[1];
        N[i][4] = a2 * q[0] - a0 * q[2];
        N[i][5] = a0 * q[1] - a1 * q[0];
      }
      // N = dn/dx = (I - n n^T) Dy / |y|;  J = dr/dx = -f^T N
      double J[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double nd = nv[0] * N[0][k] + nv[1] * N[1][k] + nv[2] * N[2][k];
#pragma unroll
        for (int i = 0; i < 3; ++i) N[i][k] = (N[i][k] - nv[i] * nd) * inv;
        J[k] = -(f[0] * N[0][k] + f[1] * N[1][k] + f[2] * N[2][k]);
      }
      // model Hessian: J J^T + r N^T N.  The second term is the part of r * d2r/dx2 that Gauss-Newton drops; because
      // r = |f - n|^2 / 2 has a vanishing gradient at a perfect fit it is as large as J J^T and without it LM only
      // converges linearly.  It is valid for small residuals, so rows with r >= RF_SECOND_ORDER_MAX get plain GN.
      const double w = r < RF_SECOND_ORDER_MAX ? r : 0.0;
      int k = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = a; c < 6; ++c)
          acc[k++] += J[a] * J[c] + w * (N[0][a] * N[0][c] + N[1][a] * N[1][c] + N[2][a] * N[2][c]);
#pragma unroll
      for (int a = 0; a < 6; ++a) acc[21 + a] += J[a] * r;
      acc[27] += r * r;
      ++used;
    }
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

    // ---- rank 0: gather, decide, solve, scatter ----
    if (rank == 0 && tid == 0) {
      double tot[RF_NACC];
      for (int i = 0; i < RF_NACC; ++i) tot[i] = 0.0;
      for (int s = 0; s < S; ++s) {
        const double* pp = cluster.map_shared_rank(partial, s);
        for (int i = 0; i < RF_NACC; ++i) tot[i] += pp[i];
      }
      const double cost = tot[27];
      int stop = 0;
      bool accepted = false;
      if (!have_cur) {
        first_cost = cost;
        accepted = true;
      } else if (cost < cur_cost) {
        accepted = true;
        if (cur_cost - cost <= 1e-14 * cur_cost) stop = 1;   // converged: no measurable decrease left
        const double rho = (cur_cost - cost) / pred;          // Nielsen's gain-ratio damping update
        const double g3 = (2.0 * rho - 1.0) * (2.0 * rho - 1.0) * (2.0 * rho - 1.0);
        lambda = fmax(lambda * fmax(1.0 / 3.0, 1.0 - g3), 1e-12);
        nu = 2.0;
      } else {
        if (step_max < 1e-9) stop = 1;                        // rejected although the step is at rounding level
        lambda *= nu;
        nu *= 2.0;
        if (lambda > 1e12) stop = 1;
      }
      if (accepted) {
        for (int i = 0; i < 9; ++i) curR[i] = st.R[i];
        for (int i = 0; i < 3; ++i) curT[i] = st.t[i];
        for (int i = 0; i < 21; ++i) curA[i] = tot[i];
        for (int i = 0; i < 6; ++i) curG[i] = tot[21 + i];
        cur_cost = cost;
        have_cur = 1;
      }
      iters_done = it + 1;
      if (it + 1 >= max_iters || n_used_sh < 6 || !(cost == cost)) stop = 1;
      double nR[9], nT[3];
      if (!stop) {
        double dx[6];
        bool ok = false;
        for (int tries = 0; tries < 40 && !ok; ++tries) {
          ok = solve_damped(curA, curG, lambda, dx);
          if (!ok) lambda *= 10.0;
        }
        double m = 0.0;
        for (int i = 0; i < 6; ++i) m = fmax(m, fabs(dx[i]));
        if (!ok || !(m == m) || m < 1e-14) {
          stop = 1;
        } else {
          step_max = m;
          double quad = 0.0, lin = 0.0;   // predicted decrease of sum r^2 under the model: -(2 g.dx + dx^T A dx)
          int k = 0;
          for (int a = 0; a < 6; ++a) {
            lin += curG[a] * dx[a];
            for (int c = a; c < 6; ++c) { quad += (a == c ? 1.0 : 2.0) * curA[k] * dx[a] * dx[c]; ++k; }
          }
          pred = -(2.0 * lin + quad);
          if (!(pred > 0.0)) pred = 1e-300;
          double E[9];
          rodrigues(dx + 3, E);
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
              nR[r * 3 + c] = curR[r * 3] * E[c] + curR[r * 3 + 1] * E[3 + c] + curR[r * 3 + 2] * E[6 + c];
          for (int i = 0; i < 3; ++i) nT[i] = curT[i] + dx[i];
        }
      }
      if (stop) {
        for (int i = 0; i < 9; ++i) nR[i] = curR[i];
        for (int i = 0; i < 3; ++i) nT[i] = curT[i];
      }
      for (int s = 0; s < S; ++s) {
        RefineState* q = cluster.map_shared_rank(&st, s);
        for (int i = 0; i < 9; ++i) q->R[i] = nR[i];
        for (int i = 0; i < 3; ++i) q->t[i] = nT[i];
        q->stop = stop;
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
  cfg.dynamicSmemBytes = 0;
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
