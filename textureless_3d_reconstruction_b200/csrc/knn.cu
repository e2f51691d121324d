// knn.cu — grid build + K3 statistical outlier removal + K7 normals + k=1 NN.
//
// K3 replaces Open3D remove_statistical_outlier(nb_neighbors=20, std_ratio=2.0)
//    as called by merge_pointclouds (depth_to_reconstruction.py:413-415); R3.
// K7 is north_star-only (Open3D estimate_normals, KNN search); R7.
// Both restated in oracle/t3d_oracle.c (o_sor_mean_dist, o_estimate_normals).
#include "knn.cuh"

#include <math.h>

namespace {

template <typename T>
__global__ void cell_key_kernel(const T* __restrict__ xyz, long long n, const __grid_constant__ GridDev g,
                                unsigned long long* keys, unsigned* vals) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int cx, cy, cz;
    grid_cell_of(g, (double)xyz[i * 3], (double)xyz[i * 3 + 1], (double)xyz[i * 3 + 2], cx, cy, cz);
    cx = min(max(cx, 0), g.dims[0] - 1);
    cy = min(max(cy, 0), g.dims[1] - 1);
    cz = min(max(cz, 0), g.dims[2] - 1);
    keys[i] = grid_pack(g, cx, cy, cz);
    vals[i] = (unsigned)i;
  }
}

template <typename T>
__global__ void gather_sorted_kernel(const T* __restrict__ xyz, const unsigned* __restrict__ idx,
                                     long long n, T* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long s = idx[i];
    out[i * 3 + 0] = xyz[s * 3 + 0];
    out[i * 3 + 1] = xyz[s * 3 + 1];
    out[i * 3 + 2] = xyz[s * 3 + 2];
  }
}

__device__ __forceinline__ unsigned long long cell_slot_insert(unsigned long long* hkeys,
                                                               unsigned long long hmask,
                                                               unsigned long long key) {
  unsigned long long slot = mix64(key) & hmask;
  while (true) {
    const unsigned long long old = atomicCAS(hkeys + slot, T3D_KEY_EMPTY, key);
    if (old == T3D_KEY_EMPTY || old == key) return slot;
    slot = (slot + 1) & hmask;
  }
}

// boundaries of the sorted key array -> hash of [start, end)
__global__ void cell_ranges_kernel(const unsigned long long* __restrict__ keys, long long n,
                                   unsigned long long* hkeys, unsigned* hstart, unsigned* hend,
                                   unsigned long long hmask, unsigned long long* n_cells) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    if (i == 0 || keys[i - 1] != k) {
      hstart[cell_slot_insert(hkeys, hmask, k)] = (unsigned)i;
      if (n_cells) atomicAdd(n_cells, 1ull);
    }
    if (i == n - 1 || keys[i + 1] != k) hend[cell_slot_insert(hkeys, hmask, k)] = (unsigned)(i + 1);
  }
}

int ceil_log2(long long v) {
  int b = 0;
  while ((1ll << b) < v) ++b;
  return b;
}

}  // namespace

// One build with a given cell size (h <= 0: density heuristic).  mn/mx: bounds of the cloud.
// n_cells_h (nullable): receives the number of occupied cells (costs one host sync).
static int grid_build_once(t3d_ctx* ctx, const void* xyz, int is_f64, long long n, double h, const double* mn,
                           const double* mx, GridDev* out, long long* n_cells_h, cudaStream_t st) {
  int rc;
  const double ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
  if (h <= 0.0) {
    // surface-like clouds: spacing ~ diag / sqrt(n); aim for a few points per cell
    const double diag = sqrt(ex * ex + ey * ey + ez * ez);
    h = 2.0 * diag / sqrt((double)n);
    if (!(h > 0.0)) h = 1.0;
  }
  GridDev g;
  memset(&g, 0, sizeof(g));
  // keep the packed key within 63 bits: grow h until it fits
  for (int iter = 0; iter < 64; ++iter) {
    const double e[3] = {ex, ey, ez};
    int total = 0;
    for (int c = 0; c < 3; ++c) {
      const double d = floor(e[c] / h) + 1.0;
      g.dims[c] = d > 2.0e9 ? 2000000000 : (int)d;
      g.bits[c] = ceil_log2(g.dims[c]);
      if (g.bits[c] == 0) g.bits[c] = 1;
      total += g.bits[c];
    }
    if (total <= 60 && g.dims[0] < 2000000000 && g.dims[1] < 2000000000 && g.dims[2] < 2000000000)
      break;
    h *= 2.0;
  }
  g.h = h;
  for (int c = 0; c < 3; ++c) g.minb[c] = mn[c];
  g.n = n;
  const int key_bits = g.bits[0] + g.bits[1] + g.bits[2];

  const size_t esz = is_f64 ? 8 : 4;
  if ((rc = ctx->scratch[0].reserve((size_t)n * 8)) != T3D_OK) return rc;
  if ((rc = ctx->scratch[1].reserve((size_t)n * 4)) != T3D_OK) return rc;
  if ((rc = ctx->scratch[2].reserve((size_t)n * 8)) != T3D_OK) return rc;
  if ((rc = ctx->scratch[3].reserve((size_t)n * 4)) != T3D_OK) return rc;
  if ((rc = ctx->scratch[4].reserve((size_t)n * 3 * esz)) != T3D_OK) return rc;
  unsigned long long hc = 1024;
  while (hc < 2ull * (unsigned long long)n) hc <<= 1;
  if ((rc = ctx->scratch[5].reserve(hc * 16)) != T3D_OK) return rc;
  unsigned long long* keys = ctx->scratch[0].as<unsigned long long>();
  unsigned* vals = ctx->scratch[1].as<unsigned>();
  unsigned long long* hkeys = ctx->scratch[5].as<unsigned long long>();
  unsigned* hstart = reinterpret_cast<unsigned*>(hkeys + hc);
  unsigned* hend = hstart + hc;

  const int grid = ctx->num_sms * 8;
  if (is_f64)
    cell_key_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(xyz), n, g, keys, vals);
  else
    cell_key_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(xyz), n, g, keys, vals);
  T3D_LAUNCH_CHECK();
  rc = t3d_radix_sort_u64(ctx, keys, vals, ctx->scratch[2].as<unsigned long long>(),
                          ctx->scratch[3].as<unsigned>(), n, key_bits, st);
  if (rc != T3D_OK) return rc;
  if (is_f64)
    gather_sorted_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(xyz), vals, n,
                                                        ctx->scratch[4].as<double>());
  else
    gather_sorted_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(xyz), vals, n,
                                                       ctx->scratch[4].as<float>());
  T3D_LAUNCH_CHECK();
  T3D_CUDA(cudaMemsetAsync(hkeys, 0xFF, hc * 8, st));
  unsigned long long* d_cells = nullptr;
  if (n_cells_h) {
    if ((rc = ctx->scratch[11].reserve(64)) != T3D_OK) return rc;
    d_cells = ctx->scratch[11].as<unsigned long long>();
    T3D_CUDA(cudaMemsetAsync(d_cells, 0, 8, st));
  }
  cell_ranges_kernel<<<grid, 256, 0, st>>>(keys, n, hkeys, hstart, hend, hc - 1, d_cells);
  T3D_LAUNCH_CHECK();
  ctx->launches += 3;
  if (n_cells_h) {
    unsigned long long* hp = reinterpret_cast<unsigned long long*>(ctx->pinned) + 512;
    T3D_CUDA(cudaMemcpyAsync(hp, d_cells, 8, cudaMemcpyDeviceToHost, st));
    T3D_CUDA(cudaStreamSynchronize(st));
    *n_cells_h = (long long)*hp;
  }
  g.sorted_idx = vals;
  g.sorted_xyz = ctx->scratch[4].p;
  g.hkeys = hkeys;
  g.hstart = hstart;
  g.hend = hend;
  g.hmask = hc - 1;
  *out = g;
  return T3D_OK;
}

// Builds the index.  h > 0: that cell size.  h <= 0: chosen for a k-nearest search — a density
// heuristic first, then (target_k > 0) one rebuild if the measured occupancy of the non-empty
// cells is far from ~target_k/3 points per cell, the point where the 27 cells around a query
// hold 2-3x k candidates (fewer hash probes per query than with sparse cells, fewer wasted
// distance evaluations than with crowded ones).  Surface-like clouds: occupancy ~ h^2.
int t3d_grid_build(t3d_ctx* ctx, const void* xyz, int is_f64, long long n, double h,
                   GridDev* out, cudaStream_t st, int target_k) {
  T3D_REQUIRE(n > 0 && n < (1ll << 31), "grid: n=%lld out of range", n);
  double mn[3], mx[3];
  int rc = t3d_bounds(ctx, xyz, is_f64, n, mn, mx, reinterpret_cast<t3d_stream>(st));
  if (rc != T3D_OK) return rc;
  for (int c = 0; c < 3; ++c) {
    if (!isfinite(mn[c]) || !isfinite(mx[c])) {
      t3d_set_error("grid: non-finite coordinates");
      return T3D_E_NUMERIC;
    }
  }
  if (h > 0.0 || target_k <= 0) return grid_build_once(ctx, xyz, is_f64, n, h, mn, mx, out, nullptr, st);
  long long cells = 0;
  rc = grid_build_once(ctx, xyz, is_f64, n, h, mn, mx, out, &cells, st);
  if (rc != T3D_OK || cells <= 0) return rc;
  const double occ = (double)n / (double)cells;
  const double want = target_k / 3.0 > 2.0 ? target_k / 3.0 : 2.0;
  const double f = sqrt(want / occ);
  if (f > 0.8 && f < 1.25) return T3D_OK;
  return grid_build_once(ctx, xyz, is_f64, n, out->h * f, mn, mx, out, nullptr, st);
}

namespace {

constexpr int KMAX = 32;

// Exact k-nearest search of query q (excluding nothing: the query itself is a
// neighbour at distance 0, as in Open3D's SearchKNN on its own cloud).
// Keeps (d2, sorted position) ascending, ties broken by ORIGINAL index.  The k-entry
// lists live in shared memory (entry j of thread t at [j * blockDim.x + t]: conflict-free,
// dynamically indexable); the current k-th distance is mirrored in a register so that the
// common case — candidate farther than the k-th best — costs no memory access.  Cells whose
// box is farther than the k-th best are skipped without a hash probe.
template <typename T>
__device__ __forceinline__ int knn_search(const GridDev& g, double qx, double qy, double qz,
                                          int k, double* bd, unsigned* bi) {
  const T* pts = reinterpret_cast<const T*>(g.sorted_xyz);
  const int S = blockDim.x;
  int cx, cy, cz;
  grid_cell_of(g, qx, qy, qz, cx, cy, cz);
  cx = min(max(cx, 0), g.dims[0] - 1);
  cy = min(max(cy, 0), g.dims[1] - 1);
  cz = min(max(cz, 0), g.dims[2] - 1);
  int found = 0;
  double worst = 1e300;  // bd[k-1] once the list is full
  const int max_ring = max(g.dims[0], max(g.dims[1], g.dims[2]));
  // position of q inside its own cell / distance to its faces
  const double lx = qx - (g.minb[0] + cx * g.h), ly = qy - (g.minb[1] + cy * g.h),
               lz = qz - (g.minb[2] + cz * g.h);
  const double near_face = fmax(0.0, fmin(fmin(fmin(lx, g.h - lx), fmin(ly, g.h - ly)),
                                          fmin(lz, g.h - lz)));
  for (int r = 0; r <= max_ring; ++r) {
    for (int dz = -r; dz <= r; ++dz) {
      const double az = dz == 0 ? 0.0 : (dz < 0 ? lz + (double)(-dz - 1) * g.h : (g.h - lz) + (double)(dz - 1) * g.h);
      for (int dy = -r; dy <= r; ++dy) {
        const double ay = dy == 0 ? 0.0 : (dy < 0 ? ly + (double)(-dy - 1) * g.h : (g.h - ly) + (double)(dy - 1) * g.h);
        const bool shell_yz = (abs(dz) == r) || (abs(dy) == r);
        const int step = shell_yz ? 1 : 2 * r;  // interior rows: only dx = -r and +r
        for (int dx = -r; dx <= r; dx += (step > 0 ? step : 1)) {
          const double ax = dx == 0 ? 0.0 : (dx < 0 ? lx + (double)(-dx - 1) * g.h : (g.h - lx) + (double)(dx - 1) * g.h);
          if (found == k && ax * ax + ay * ay + az * az > worst) continue;  // box farther than the k-th best
          unsigned s, e;
          if (!grid_lookup(g, cx + dx, cy + dy, cz + dz, s, e)) continue;
          for (unsigned j = s; j < e; ++j) {
            const double ddx = (double)pts[3ll * j] - qx, ddy = (double)pts[3ll * j + 1] - qy,
                         ddz = (double)pts[3ll * j + 2] - qz;
            const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
            if (found == k && !(d2 < worst ||
                                (d2 == worst && g.sorted_idx[j] < g.sorted_idx[bi[(k - 1) * S]])))
              continue;
            // insertion (ascending by d2, then by original index)
            int pos = found < k ? found : k - 1;
            const unsigned oj = g.sorted_idx[j];
            while (pos > 0) {
              const double pd = bd[(pos - 1) * S];
              if (!(pd > d2 || (pd == d2 && g.sorted_idx[bi[(pos - 1) * S]] > oj))) break;
              bd[pos * S] = pd;
              bi[pos * S] = bi[(pos - 1) * S];
              --pos;
            }
            bd[pos * S] = d2;
            bi[pos * S] = j;
            if (found < k) ++found;
            if (found == k) worst = bd[(k - 1) * S];
          }
        }
      }
    }
    if (found == k) {
      const double reach = near_face + r * g.h;  // everything closer than this has been seen
      if (worst <= reach * reach) break;
    }
    if (r > 0 && cx - r < 0 && cy - r < 0 && cz - r < 0 && cx + r >= g.dims[0] &&
        cy + r >= g.dims[1] && cz + r >= g.dims[2])
      break;  // whole grid covered
  }
  return found;
}

constexpr int KNN_THREADS = 128;
// dynamic shared memory: k doubles + k unsigneds per thread
__device__ __forceinline__ void knn_lists(int k, double** bd, unsigned** bi) {
  extern __shared__ double knn_smem[];
  *bd = knn_smem + threadIdx.x;
  *bi = reinterpret_cast<unsigned*>(knn_smem + (size_t)k * blockDim.x) + threadIdx.x;
}
inline size_t knn_smem_bytes(int k) { return (size_t)k * KNN_THREADS * (sizeof(double) + sizeof(unsigned)); }

// K3: mean distance to the nb nearest neighbours (self included), R3.
__global__ void __launch_bounds__(128)
    sor_mean_kernel(const __grid_constant__ GridDev g, int nb, double* mean_dist, long long lo, long long hi) {
  // [lo, hi): range of GRID-SORTED positions handled by this launch (multi-GPU: a rank's share;
  // contiguous in cell order, so every warp stays fully busy)
  const double* pts = reinterpret_cast<const double*>(g.sorted_xyz);
  for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi;
       i += (long long)gridDim.x * blockDim.x) {
    double* bd;
    unsigned* bi;
    knn_lists(nb, &bd, &bi);
    const int found = knn_search<double>(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], nb, bd, bi);
    double sum = 0.0;
    for (int k = 0; k < found; ++k) sum += sqrt(bd[k * blockDim.x]);
    mean_dist[g.sorted_idx[i]] = found > 0 ? sum / (double)found : -1.0;
  }
}

// deterministic two-level sums for mu / sigma
__global__ void __launch_bounds__(256)
    sor_stats_kernel(const double* __restrict__ mean_dist, long long n, double mu, int pass,
                     double* partial /* gridDim.x * 2 */) {
  double a = 0.0, b = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double m = mean_dist[i];
    if (m >= 0.0) {
      if (pass == 0) { if (m > 0.0) a += m; b += 1.0; }
      else if (m > 0.0) { const double d = m - mu; a += d * d; }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, d);
    b += __shfl_xor_sync(0xffffffffu, b, d);
  }
  __shared__ double sa[8], sb[8];
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tb2 = 0;
    for (int k = 0; k < 8; ++k) { ta += sa[k]; tb2 += sb[k]; }
    partial[blockIdx.x * 2] = ta;
    partial[blockIdx.x * 2 + 1] = tb2;
  }
}

__global__ void sor_mask_kernel(const double* __restrict__ mean_dist, long long n, double thr,
                                uint8_t* keep, unsigned long long* kept) {
  unsigned long long c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double m = mean_dist[i];
    const bool k = m > 0.0 && m < thr;
    keep[i] = k ? 1 : 0;
    c += k;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(kept, c);
}

// ---- 3x3 symmetric eigen-solver (smallest eigenvector), shared with oracle ----
__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// eigenvector of the eigenvalue `ev` of symmetric A (a00,a01,a02,a11,a12,a22):
// the largest cross product of two rows of (A - ev I).
__device__ __forceinline__ void eigvec_for(const double* A, double ev, double* out) {
  const double r0[3] = {A[0] - ev, A[1], A[2]};
  const double r1[3] = {A[1], A[3] - ev, A[4]};
  const double r2[3] = {A[2], A[4], A[5] - ev};
  double c0[3], c1[3], c2[3];
  cross3(r0, r1, c0);
  cross3(r0, r2, c1);
  cross3(r1, r2, c2);
  const double d0 = c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2];
  const double d1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
  const double d2 = c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2];
  const double* best = c0;
  double dm = d0;
  if (d1 > dm) { dm = d1; best = c1; }
  if (d2 > dm) { dm = d2; best = c2; }
  if (dm > 0.0) {
    const double inv = 1.0 / sqrt(dm);
    out[0] = best[0] * inv; out[1] = best[1] * inv; out[2] = best[2] * inv;
  } else {
    out[0] = 0.0; out[1] = 0.0; out[2] = 1.0;
  }
}

__device__ __forceinline__ void smallest_eigvec(const double* C, double* nrm) {
  // scale by the largest |coefficient|
  double mc = 0.0;
  for (int i = 0; i < 6; ++i) mc = fmax(mc, fabs(C[i]));
  if (!(mc > 0.0)) { nrm[0] = 0; nrm[1] = 0; nrm[2] = 1; return; }
  double A[6];
  for (int i = 0; i < 6; ++i) A[i] = C[i] / mc;
  const double off = A[1] * A[1] + A[2] * A[2] + A[4] * A[4];
  double ev_min;
  if (off > 0.0) {
    const double q = (A[0] + A[3] + A[5]) / 3.0;
    const double b00 = A[0] - q, b11 = A[3] - q, b22 = A[5] - q;
    const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * off) / 6.0);
    const double c00 = b11 * b22 - A[4] * A[4];
    const double c01 = A[1] * b22 - A[4] * A[2];
    const double c02 = A[1] * A[4] - b11 * A[2];
    const double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
    const double half_det = fmin(fmax(det * 0.5, -1.0), 1.0);
    const double angle = acos(half_det) / 3.0;
    const double beta0 = 2.0 * cos(angle + 2.0943951023931954923);  // smallest root
    ev_min = q + p * beta0;
  } else {
    ev_min = fmin(A[0], fmin(A[3], A[5]));
    if (A[0] == ev_min) { nrm[0] = 1; nrm[1] = 0; nrm[2] = 0; }
    else if (A[3] == ev_min) { nrm[0] = 0; nrm[1] = 1; nrm[2] = 0; }
    else { nrm[0] = 0; nrm[1] = 0; nrm[2] = 1; }
    return;
  }
  eigvec_for(A, ev_min, nrm);
}

// K7: covariance of the knn neighbourhood from cumulants (R7), f64.
__global__ void __launch_bounds__(128)
    normals_kernel(const __grid_constant__ GridDev g, int knn, int orient, double ox, double oy,
                   double oz, float* nrm_out) {
  const float* pts = reinterpret_cast<const float*>(g.sorted_xyz);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < g.n;
       i += (long long)gridDim.x * blockDim.x) {
    double* bd;
    unsigned* bi;
    knn_lists(knn, &bd, &bi);
    const double qx = pts[3 * i], qy = pts[3 * i + 1], qz = pts[3 * i + 2];
    const int found = knn_search<float>(g, qx, qy, qz, knn, bd, bi);
    double n[3] = {0.0, 0.0, 1.0};
    if (found >= 3) {
      double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = 0; k < found; ++k) {
        const unsigned jj = bi[k * blockDim.x];
        const double x = pts[3ll * jj], y = pts[3ll * jj + 1], z = pts[3ll * jj + 2];
        c[0] += x; c[1] += y; c[2] += z;
        c[3] += x * x; c[4] += x * y; c[5] += x * z;
        c[6] += y * y; c[7] += y * z; c[8] += z * z;
      }
      const double inv = 1.0 / (double)found;
      for (int k = 0; k < 9; ++k) c[k] *= inv;
      double C[6];
      C[0] = c[3] - c[0] * c[0];
      C[1] = c[4] - c[0] * c[1];
      C[2] = c[5] - c[0] * c[2];
      C[3] = c[6] - c[1] * c[1];
      C[4] = c[7] - c[1] * c[2];
      C[5] = c[8] - c[2] * c[2];
      smallest_eigvec(C, n);
    }
    if (orient) {
      const double d = n[0] * (ox - qx) + n[1] * (oy - qy) + n[2] * (oz - qz);
      if (d < 0.0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    }
    const long long o = g.sorted_idx[i];
    nrm_out[o * 3 + 0] = (float)n[0];
    nrm_out[o * 3 + 1] = (float)n[1];
    nrm_out[o * 3 + 2] = (float)n[2];
  }
}

}  // namespace

namespace {
__global__ void __launch_bounds__(128)
    nn_kernel(const __grid_constant__ GridDev g, const float* __restrict__ q, long long nq,
              double r2, int* out_idx, float* out_d2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nq;
       i += (long long)gridDim.x * blockDim.x) {
    double d2;
    const int j = t3d_nn_within(g, q[3 * i], q[3 * i + 1], q[3 * i + 2], r2, &d2);
    out_idx[i] = j >= 0 ? (int)g.sorted_idx[j] : -1;
    if (out_d2) out_d2[i] = j >= 0 ? (float)d2 : -1.0f;
  }
}
}  // namespace

static int sor_from_means(t3d_ctx* ctx, const double* mean, int64_t n, double std_ratio, uint8_t* keep_mask,
                          int64_t* out_kept, double* stats_h, cudaStream_t st);

extern "C" int t3d_statistical_outlier(t3d_ctx* ctx, const double* xyz, int64_t n, int nb,
                                       double std_ratio, double* out_mean_dist,
                                       uint8_t* keep_mask, int64_t* out_kept, double* stats_h,
                                       t3d_stream stream) {
  T3D_REQUIRE(ctx && keep_mask && out_kept, "t3d_statistical_outlier: null argument");
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(nb >= 1 && nb <= KMAX && std_ratio > 0.0,
              "t3d_statistical_outlier: nb_neighbors must be in [1,%d], std_ratio > 0", KMAX);
  cudaStream_t st = as_stream(stream);
  T3D_CUDA(cudaMemsetAsync(out_kept, 0, sizeof(int64_t), st));
  if (n == 0) return T3D_OK;
  T3D_REQUIRE(xyz, "t3d_statistical_outlier: null xyz");
  GridDev g;
  int rc = t3d_grid_build(ctx, xyz, 1, n, -1.0, &g, st, nb);
  if (rc != T3D_OK) return rc;
  const int grid = ctx->num_sms * 8;
  double* mean = out_mean_dist;
  if (!mean) {
    if ((rc = ctx->scratch[6].reserve((size_t)n * 8 + 4096 * 16)) != T3D_OK) return rc;
    mean = ctx->scratch[6].as<double>();
  }
  const long long want = (n + 127) / 128;
  const int mgrid = (int)(want < 2ll * grid ? want : 2ll * grid);
  sor_mean_kernel<<<mgrid, KNN_THREADS, knn_smem_bytes(nb), st>>>(g, nb, mean, 0, n);
  T3D_LAUNCH_CHECK();
  return sor_from_means(ctx, mean, n, std_ratio, keep_mask, out_kept, stats_h, st);
}

// R3 steps 2-4 from the per-point mean distances: mu, sigma (Bessel), threshold, keep mask.
static int sor_from_means(t3d_ctx* ctx, const double* mean, int64_t n, double std_ratio, uint8_t* keep_mask,
                          int64_t* out_kept, double* stats_h, cudaStream_t st) {
  int rc;
  const int grid = ctx->num_sms * 8;
  // mu, sigma: per-CTA partials summed on the host in a fixed order
  const int sgrid = 256;
  if ((rc = ctx->scratch[7].reserve(sizeof(double) * 2 * sgrid)) != T3D_OK) return rc;
  double* partial = ctx->scratch[7].as<double>();
  double* h = reinterpret_cast<double*>(ctx->pinned);
  sor_stats_kernel<<<sgrid, 256, 0, st>>>(mean, n, 0.0, 0, partial);
  T3D_LAUNCH_CHECK();
  T3D_CUDA(cudaMemcpyAsync(h, partial, sizeof(double) * 2 * sgrid, cudaMemcpyDeviceToHost, st));
  T3D_CUDA(cudaStreamSynchronize(st));
  double sum = 0.0, valid = 0.0;
  for (int b = 0; b < sgrid; ++b) { sum += h[2 * b]; valid += h[2 * b + 1]; }
  ctx->launches += 2;
  if (valid == 0.0) {
    T3D_CUDA(cudaMemsetAsync(keep_mask, 0, (size_t)n, st));
    if (stats_h) { stats_h[0] = stats_h[1] = stats_h[2] = 0.0; }
    return T3D_OK;
  }
  const double mu = sum / valid;
  sor_stats_kernel<<<sgrid, 256, 0, st>>>(mean, n, mu, 1, partial);
  T3D_LAUNCH_CHECK();
  T3D_CUDA(cudaMemcpyAsync(h, partial, sizeof(double) * 2 * sgrid, cudaMemcpyDeviceToHost, st));
  T3D_CUDA(cudaStreamSynchronize(st));
  double sq = 0.0;
  for (int b = 0; b < sgrid; ++b) sq += h[2 * b];
  const double sigma = sqrt(sq / (valid - 1.0));
  const double thr = mu + std_ratio * sigma;
  if (stats_h) { stats_h[0] = mu; stats_h[1] = sigma; stats_h[2] = thr; }
  sor_mask_kernel<<<grid, 256, 0, st>>>(mean, n, thr, keep_mask,
                                        reinterpret_cast<unsigned long long*>(out_kept));
  T3D_LAUNCH_CHECK();
  ctx->launches += 2;
  T3D_CUDA(cudaStreamSynchronize(st));
  return T3D_OK;
}

// Sharded K3 (replicate the cloud, shard the queries).  Mean neighbour distance of the
// points at grid-sorted positions [part * n / parts, (part + 1) * n / parts) only; the other
// entries of out_mean_dist are left untouched (callers initialise them to -inf and combine the
// ranks' vectors with an all_reduce(MAX)).  The grid is built from the full cloud with a stable
// sort, so every rank sees the same order and the parts tile the cloud exactly.
extern "C" int t3d_sor_mean_distances_part(t3d_ctx* ctx, const double* xyz, int64_t n, int nb, int part,
                                           int parts, double* out_mean_dist, t3d_stream stream) {
  T3D_REQUIRE(ctx && out_mean_dist && parts >= 1 && part >= 0 && part < parts, "t3d_sor_mean_distances_part: bad argument");
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(nb >= 1 && nb <= KMAX, "t3d_sor_mean_distances_part: nb_neighbors must be in [1,%d]", KMAX);
  if (n == 0) return T3D_OK;
  T3D_REQUIRE(xyz, "t3d_sor_mean_distances_part: null xyz");
  cudaStream_t st = as_stream(stream);
  GridDev g;
  int rc = t3d_grid_build(ctx, xyz, 1, n, -1.0, &g, st, nb);
  if (rc != T3D_OK) return rc;
  const long long lo = (long long)n * part / parts, hi = (long long)n * (part + 1) / parts;
  if (hi > lo) {
    const long long want = (hi - lo + 127) / 128;
    const long long cap = 2ll * ctx->num_sms * 8;
    sor_mean_kernel<<<(int)(want < cap ? want : cap), KNN_THREADS, knn_smem_bytes(nb), st>>>(g, nb, out_mean_dist, lo, hi);
    T3D_LAUNCH_CHECK();
    ctx->launches++;
  }
  T3D_CUDA(cudaStreamSynchronize(st));
  return T3D_OK;
}

// The rest of R3 from a complete vector of mean distances (all ranks hold the same one).
extern "C" int t3d_sor_from_mean_distances(t3d_ctx* ctx, const double* mean_dist, int64_t n, double std_ratio,
                                           uint8_t* keep_mask, int64_t* out_kept, double* stats_h,
                                           t3d_stream stream) {
  T3D_REQUIRE(ctx && keep_mask && out_kept && std_ratio > 0.0, "t3d_sor_from_mean_distances: bad argument");
  T3D_ON_DEVICE(ctx->device);
  cudaStream_t st = as_stream(stream);
  T3D_CUDA(cudaMemsetAsync(out_kept, 0, sizeof(int64_t), st));
  if (n == 0) return T3D_OK;
  T3D_REQUIRE(mean_dist, "t3d_sor_from_mean_distances: null mean_dist");
  return sor_from_means(ctx, mean_dist, n, std_ratio, keep_mask, out_kept, stats_h, st);
}

extern "C" int t3d_estimate_normals(t3d_ctx* ctx, const float* xyz, int64_t n, int knn,
                                    const double* orient_to_h, float* nrm, t3d_stream stream) {
  T3D_REQUIRE(ctx && (n == 0 || (xyz && nrm)), "t3d_estimate_normals: null argument");
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(knn >= 1 && knn <= KMAX, "t3d_estimate_normals: knn must be in [1,%d]", KMAX);
  if (n == 0) return T3D_OK;
  cudaStream_t st = as_stream(stream);
  GridDev g;
  int rc = t3d_grid_build(ctx, xyz, 0, n, -1.0, &g, st, knn);
  if (rc != T3D_OK) return rc;
  const long long want = (n + 127) / 128;
  const int grid = (int)(want < ctx->num_sms * 16 ? want : ctx->num_sms * 16);
  normals_kernel<<<grid, KNN_THREADS, knn_smem_bytes(knn), st>>>(g, knn, orient_to_h != nullptr,
                                       orient_to_h ? orient_to_h[0] : 0.0,
                                       orient_to_h ? orient_to_h[1] : 0.0,
                                       orient_to_h ? orient_to_h[2] : 0.0, nrm);
  T3D_LAUNCH_CHECK();
  ctx->launches++;
  T3D_CUDA(cudaStreamSynchronize(st));
  return T3D_OK;
}

extern "C" int t3d_nearest_neighbor(t3d_ctx* ctx, const float* query, int64_t n_q,
                                    const float* ref, int64_t n_ref, double radius,
                                    int32_t* out_idx, float* out_d2, t3d_stream stream) {
  T3D_REQUIRE(ctx && (n_q == 0 || (query && out_idx)), "t3d_nearest_neighbor: null argument");
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(radius > 0.0, "t3d_nearest_neighbor: radius must be > 0");
  if (n_q == 0) return T3D_OK;
  cudaStream_t st = as_stream(stream);
  if (n_ref == 0) {
    T3D_CUDA(cudaMemsetAsync(out_idx, 0xFF, (size_t)n_q * 4, st));
    return T3D_OK;
  }
  GridDev g;
  int rc = t3d_grid_build(ctx, ref, 0, n_ref, radius, &g, st, 0);
  if (rc != T3D_OK) return rc;
  const long long want = (n_q + 127) / 128;
  const int grid = (int)(want < ctx->num_sms * 16 ? want : ctx->num_sms * 16);
  nn_kernel<<<grid, 128, 0, st>>>(g, query, n_q, radius * radius, out_idx, out_d2);
  T3D_LAUNCH_CHECK();
  ctx->launches++;
  T3D_CUDA(cudaStreamSynchronize(st));
  return T3D_OK;
}
