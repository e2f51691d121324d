// icp.cu — K8: point-to-plane ICP (north_star only; Open3D registration_icp
// with TransformationEstimationPointToPlane; restated in
// oracle/t3d_oracle.c: o_icp_point_to_plane).
//
// One kernel per linearisation: every source point is transformed (f64),
// matched to its nearest target point within max_corr_dist through the grid
// hash, and contributes r = (s - t).n, J = [s x n ; n] to the 6x6 normal
// equations.  29 doubles (21 upper JtJ + 6 Jtr + sum d^2 + count) are reduced
// with warp shuffles, then per-CTA partials are summed by one CTA in a fixed
// order (deterministic).  No tensor cores: this is a gather + reduction.
#include <math.h>

#include "knn.cuh"

namespace {

constexpr int NACC = 29;
constexpr int ICP_THREADS = 128;

__global__ void __launch_bounds__(ICP_THREADS)
    icp_linearize_kernel(const __grid_constant__ GridDev g, const float* __restrict__ src,
                         long long n_src, const float* __restrict__ tgt_nrm, double r2,
                         const double* __restrict__ Tdev /* 12: row-major 3x4 */,
                         double* partial /* gridDim.x * NACC */) {
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  double T[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) T[k] = Tdev[k];
  const float* tpts = reinterpret_cast<const float*>(g.sorted_xyz);

  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_src;
       i += (long long)gridDim.x * blockDim.x) {
    const double px = src[3 * i], py = src[3 * i + 1], pz = src[3 * i + 2];
    const double sx = T[0] * px + T[1] * py + T[2] * pz + T[3];
    const double sy = T[4] * px + T[5] * py + T[6] * pz + T[7];
    const double sz = T[8] * px + T[9] * py + T[10] * pz + T[11];
    double d2;
    const int j = t3d_nn_within(g, sx, sy, sz, r2, &d2);
    if (j < 0) continue;
    const long long o = g.sorted_idx[j];
    const double tx = tpts[3ll * j], ty = tpts[3ll * j + 1], tz = tpts[3ll * j + 2];
    const double nx = tgt_nrm[3 * o], ny = tgt_nrm[3 * o + 1], nz = tgt_nrm[3 * o + 2];
    const double r = (sx - tx) * nx + (sy - ty) * ny + (sz - tz) * nz;
    double J[6];
    J[0] = sy * nz - sz * ny;
    J[1] = sz * nx - sx * nz;
    J[2] = sx * ny - sy * nx;
    J[3] = nx; J[4] = ny; J[5] = nz;
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = a; b < 6; ++b) acc[q++] += J[a] * J[b];
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] += J[a] * r;
    acc[27] += d2;
    acc[28] += 1.0;
  }
  // warp tree, then cross-warp through shared memory
#pragma unroll
  for (int k = 0; k < NACC; ++k)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], d);
  __shared__ double s[ICP_THREADS / 32][NACC];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0)
    for (int k = 0; k < NACC; ++k) s[w][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < NACC) {
    double t = 0.0;
    for (int ww = 0; ww < ICP_THREADS / 32; ++ww) t += s[ww][threadIdx.x];
    partial[(long long)blockIdx.x * NACC + threadIdx.x] = t;
  }
}

__global__ void icp_final_kernel(const double* partial, int nblocks, double* out29) {
  const int k = threadIdx.x;
  if (k >= NACC) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += partial[(long long)b * NACC + k];
  out29[k] = t;
}

// ---- host 6x6 helpers -----------------------------------------------------
double det6(const double* A) {
  double M[36];
  memcpy(M, A, sizeof(M));
  double det = 1.0;
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r)
      if (fabs(M[r * 6 + c]) > fabs(M[p * 6 + c])) p = r;
    if (M[p * 6 + c] == 0.0) return 0.0;
    if (p != c) {
      for (int k = 0; k < 6; ++k) { double t = M[c * 6 + k]; M[c * 6 + k] = M[p * 6 + k]; M[p * 6 + k] = t; }
      det = -det;
    }
    det *= M[c * 6 + c];
    for (int r = c + 1; r < 6; ++r) {
      const double f = M[r * 6 + c] / M[c * 6 + c];
      for (int k = c; k < 6; ++k) M[r * 6 + k] -= f * M[c * 6 + k];
    }
  }
  return det;
}

// LDL^T solve of the symmetric system A x = b (no pivoting; A is PSD here)
bool ldlt_solve6(const double* A, const double* b, double* x) {
  double L[36] = {0}, D[6];
  for (int j = 0; j < 6; ++j) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
    D[j] = d;
    if (d == 0.0 || !isfinite(d)) return false;
    L[j * 6 + j] = 1.0;
    for (int i = j + 1; i < 6; ++i) {
      double v = A[i * 6 + j];
      for (int k = 0; k < j; ++k) v -= L[i * 6 + k] * L[j * 6 + k] * D[k];
      L[i * 6 + j] = v / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double v = b[i];
    for (int k = 0; k < i; ++k) v -= L[i * 6 + k] * y[k];
    y[i] = v;
  }
  for (int i = 0; i < 6; ++i) y[i] /= D[i];
  for (int i = 5; i >= 0; --i) {
    double v = y[i];
    for (int k = i + 1; k < 6; ++k) v -= L[k * 6 + i] * x[k];
    x[i] = v;
  }
  return true;
}

void mat4_mul(const double* A, const double* B, double* C) {
  double R[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
      R[i * 4 + j] = s;
    }
  memcpy(C, R, sizeof(R));
}

// x = (alpha, beta, gamma, tx, ty, tz) -> [Rz(gamma) Ry(beta) Rx(alpha) | t]
void vec6_to_mat4(const double* x, double* M) {
  const double ca = cos(x[0]), sa = sin(x[0]), cb = cos(x[1]), sb = sin(x[1]), cg = cos(x[2]),
               sg = sin(x[2]);
  M[0] = cg * cb; M[1] = cg * sb * sa - sg * ca; M[2] = cg * sb * ca + sg * sa; M[3] = x[3];
  M[4] = sg * cb; M[5] = sg * sb * sa + cg * ca; M[6] = sg * sb * ca - cg * sa; M[7] = x[4];
  M[8] = -sb;     M[9] = cb * sa;                M[10] = cb * ca;               M[11] = x[5];
  M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
}

// solve the normal equations -> update matrix (identity when ill-posed, R8)
void solve_update(const double* acc, double* U) {
  double A[36], b[6], x[6];
  int q = 0;
  for (int a = 0; a < 6; ++a)
    for (int c = a; c < 6; ++c) { A[a * 6 + c] = acc[q]; A[c * 6 + a] = acc[q]; ++q; }
  for (int a = 0; a < 6; ++a) b[a] = -acc[21 + a];
  const double det = det6(A);
  bool ok = isfinite(det) && fabs(det) >= 1e-6;
  if (ok) ok = ldlt_solve6(A, b, x);
  if (ok)
    for (int a = 0; a < 6; ++a) ok = ok && isfinite(x[a]);
  if (!ok) {
    for (int i = 0; i < 16; ++i) U[i] = (i % 5 == 0) ? 1.0 : 0.0;
    return;
  }
  vec6_to_mat4(x, U);
}

int linearize(t3d_ctx* ctx, const GridDev& g, const float* src, long long n_src,
              const float* tgt_nrm, double max_corr, const double* T, double* acc_h,
              cudaStream_t st) {
  const long long want = (n_src + ICP_THREADS - 1) / ICP_THREADS;
  const int grid = (int)(want < ctx->num_sms * 8 ? (want > 0 ? want : 1) : ctx->num_sms * 8);
  int rc = ctx->scratch[6].reserve(sizeof(double) * ((size_t)grid * NACC + NACC + 16));
  if (rc != T3D_OK) return rc;
  double* partial = ctx->scratch[6].as<double>();
  double* out29 = partial + (size_t)grid * NACC;
  double* Tdev = out29 + NACC;
  double* hp = reinterpret_cast<double*>(ctx->pinned);
  for (int k = 0; k < 12; ++k) hp[64 + k] = T[k];  // rows 0..2 of the 4x4
  T3D_CUDA(cudaMemcpyAsync(Tdev, hp + 64, 12 * sizeof(double), cudaMemcpyHostToDevice, st));
  icp_linearize_kernel<<<grid, ICP_THREADS, 0, st>>>(g, src, n_src, tgt_nrm, max_corr * max_corr,
                                                      Tdev, partial);
  T3D_LAUNCH_CHECK();
  icp_final_kernel<<<1, 32, 0, st>>>(partial, grid, out29);
  T3D_LAUNCH_CHECK();
  ctx->launches += 2;
  T3D_CUDA(cudaMemcpyAsync(hp, out29, NACC * sizeof(double), cudaMemcpyDeviceToHost, st));
  T3D_CUDA(cudaStreamSynchronize(st));
  memcpy(acc_h, hp, NACC * sizeof(double));
  return T3D_OK;
}

}  // namespace

extern "C" int t3d_icp_linearize(t3d_ctx* ctx, const float* src, int64_t n_src, const float* tgt,
                                 const float* tgt_nrm, int64_t n_tgt, double max_corr_dist,
                                 const double* T_h, double* out27_h, double* out_stats_h,
                                 t3d_stream stream) {
  T3D_REQUIRE(ctx && T_h && out27_h && out_stats_h, "t3d_icp_linearize: null argument");
  T3D_REQUIRE(max_corr_dist > 0.0, "t3d_icp_linearize: max_corr_dist must be > 0");
  for (int k = 0; k < 27; ++k) out27_h[k] = 0.0;
  out_stats_h[0] = out_stats_h[1] = 0.0;
  if (n_src == 0 || n_tgt == 0) return T3D_OK;
  T3D_REQUIRE(src && tgt && tgt_nrm, "t3d_icp_linearize: null clouds");
  cudaStream_t st = as_stream(stream);
  GridDev g;
  int rc = t3d_grid_build(ctx, tgt, 0, n_tgt, max_corr_dist, &g, st);
  if (rc != T3D_OK) return rc;
  double acc[NACC];
  rc = linearize(ctx, g, src, n_src, tgt_nrm, max_corr_dist, T_h, acc, st);
  if (rc != T3D_OK) return rc;
  memcpy(out27_h, acc, 27 * sizeof(double));
  out_stats_h[0] = acc[27];
  out_stats_h[1] = acc[28];
  return T3D_OK;
}

extern "C" int t3d_icp_point_to_plane(t3d_ctx* ctx, const float* src, int64_t n_src,
                                      const float* tgt, const float* tgt_nrm, int64_t n_tgt,
                                      double max_corr_dist, const double* T0_h, int max_iter,
                                      double rel_fitness, double rel_rmse,
                                      t3d_icp_result* res, t3d_stream stream) {
  T3D_REQUIRE(ctx && res, "t3d_icp_point_to_plane: null argument");
  T3D_REQUIRE(max_corr_dist > 0.0 && max_iter >= 0, "t3d_icp_point_to_plane: bad parameters");
  memset(res, 0, sizeof(*res));
  for (int i = 0; i < 16; ++i) res->T[i] = T0_h ? T0_h[i] : ((i % 5 == 0) ? 1.0 : 0.0);
  if (n_src == 0 || n_tgt == 0) return T3D_OK;
  T3D_REQUIRE(src && tgt && tgt_nrm, "t3d_icp_point_to_plane: null clouds");
  cudaStream_t st = as_stream(stream);
  GridDev g;
  int rc = t3d_grid_build(ctx, tgt, 0, n_tgt, max_corr_dist, &g, st);
  if (rc != T3D_OK) return rc;
  double acc[NACC];
  rc = linearize(ctx, g, src, n_src, tgt_nrm, max_corr_dist, res->T, acc, st);
  if (rc != T3D_OK) return rc;
  double fitness = acc[28] / (double)n_src;
  double rmse = acc[28] > 0.0 ? sqrt(acc[27] / acc[28]) : 0.0;
  int it = 0;
  int converged = 0;
  for (; it < max_iter; ++it) {
    double U[16];
    solve_update(acc, U);
    mat4_mul(U, res->T, res->T);
    rc = linearize(ctx, g, src, n_src, tgt_nrm, max_corr_dist, res->T, acc, st);
    if (rc != T3D_OK) return rc;
    const double f2 = acc[28] / (double)n_src;
    const double r2 = acc[28] > 0.0 ? sqrt(acc[27] / acc[28]) : 0.0;
    const bool stop = fabs(fitness - f2) < rel_fitness && fabs(rmse - r2) < rel_rmse;
    fitness = f2;
    rmse = r2;
    if (stop) { converged = 1; ++it; break; }
  }
  res->fitness = fitness;
  res->inlier_rmse = rmse;
  res->iterations = it;
  res->converged = converged;
  res->correspondences = (int64_t)acc[28];
  return T3D_OK;
}
