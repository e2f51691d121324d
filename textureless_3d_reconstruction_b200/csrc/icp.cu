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
// t3d_icp_point_to_plane keeps the whole registration on the device
// (icp_nn_kernel below); t3d_icp_linearize keeps the
// one-linearisation-per-call form for the multi-GPU all-reduce path.
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

// ---- small matrix helpers (host + device) ----------------------------------
__host__ __device__ void mat4_mul(const double* A, const double* B, double* C) {
  double R[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
      R[i * 4 + j] = s;
    }
  for (int i = 0; i < 16; ++i) C[i] = R[i];
}

// x = (alpha, beta, gamma, tx, ty, tz) -> [Rz(gamma) Ry(beta) Rx(alpha) | t]
__host__ __device__ void vec6_to_mat4(const double* x, double* M) {
  const double ca = cos(x[0]), sa = sin(x[0]), cb = cos(x[1]), sb = sin(x[1]), cg = cos(x[2]),
               sg = sin(x[2]);
  M[0] = cg * cb; M[1] = cg * sb * sa - sg * ca; M[2] = cg * sb * ca + sg * sa; M[3] = x[3];
  M[4] = sg * cb; M[5] = sg * sb * sa + cg * ca; M[6] = sg * sb * ca - cg * sa; M[7] = x[4];
  M[8] = -sb;     M[9] = cb * sa;                M[10] = cb * ca;               M[11] = x[5];
  M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
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

// ===========================================================================
// Device-resident registration: hashed target grid (no bounds pass, no sort) and
// a registration state that lives in HBM.  Every iteration is ONE launch
// (correspondences with the accumulation fused in; reduce + 6x6 solve by the
// last CTA) that turns into a no-op once the state says "done", so the host
// enqueues a handful of iterations at a time and reads one small result record
// back — no host round trip, H2D pose upload or D2H reduction per iteration.
// ===========================================================================

namespace {

struct __align__(16) HSlot {  // one 16-byte load answers a probe: the cell's key and its point range
  unsigned long long key1;    // pack_key + 1; 0 = empty (the table is cleared with one memset)
  unsigned start;             // first sorted position of the cell
  unsigned count;             // points in the cell
};

struct HGrid {            // cells of size h keyed by floor(p / h) (biased 21-bit pack_key)
  HSlot* slots;
  unsigned* fill;         // scatter cursor
  unsigned long long* coarse_keys;  // set of occupied coarse cells (input of the dilation)
  unsigned long long* near_keys;  // set of coarse cells (2h = max_corr) with a target within one coarse
                                  // cell in every direction: a query whose coarse cell is absent has no
                                  // correspondence — decided by ONE probe instead of 125
  unsigned long long mask;
  float4* xyzi;           // n: {x, y, z, original index as bits}, cell-contiguous
  float* nrm;             // n*3, same order
  unsigned* cursor;       // [0] running start allocator, [1] out-of-range flag
  double inv_h, h;        // cell size h = max_corr / 2: the search covers the 5^3 cells around the query
  long long n;            // number of target points (upper bound when n_dev is set)
  const long long* n_dev; // nullable: actual count lives in device memory (no host sync)
};

__device__ __forceinline__ long long hg_n(const HGrid& g) {
  if (!g.n_dev) return g.n;
  const long long v = *g.n_dev;
  return v < g.n ? v : g.n;
}

// 5x5x5 cell offsets ordered by ring (max-norm) and, inside a ring, by distance:
// ring 0 = 1 cell, ring 1 = 26, ring 2 = 98.
__constant__ signed char c_ofs[125][4];


__device__ __forceinline__ bool hg_cell(const HGrid& g, double x, double y, double z, int& cx, int& cy, int& cz) {
  const double fx = floor(x * g.inv_h), fy = floor(y * g.inv_h), fz = floor(z * g.inv_h);
  if (!(fabs(fx) < 1048575.0 && fabs(fy) < 1048575.0 && fabs(fz) < 1048575.0)) return false;
  cx = (int)fx; cy = (int)fy; cz = (int)fz;
  return true;
}

__global__ void hg_count_kernel(const float* __restrict__ tgt, const __grid_constant__ HGrid g) {
  const long long n = hg_n(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int cx, cy, cz;
    if (!hg_cell(g, tgt[3 * i], tgt[3 * i + 1], tgt[3 * i + 2], cx, cy, cz)) { g.cursor[1] = 1; continue; }
    const unsigned long long key = pack_key(cx, cy, cz), key1 = key + 1ull;
    unsigned long long slot = mix64(key) & g.mask;
    while (true) {
      unsigned long long k = ld_volatile_u64(reinterpret_cast<const uint64_t*>(&g.slots[slot].key1));
      if (k == 0ull) k = atomicCAS(&g.slots[slot].key1, 0ull, key1);
      if (k == 0ull || k == key1) break;
      slot = (slot + 1) & g.mask;
    }
    atomicAdd(&g.slots[slot].count, 1u);
  }
}

// cells get their [start, start+count) range from a running cursor: cell order is
// irrelevant, only contiguity inside a cell matters
__global__ void hg_alloc_kernel(const __grid_constant__ HGrid g, int with_coarse) {
  // the loop bound is warp-uniform (the table size is a multiple of 32 and so is the stride), so the
  // cursor can be advanced once per warp: tens of thousands of returning atomics on one address
  // would otherwise serialise in L2
  const unsigned lane = lane_id();
  for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s <= g.mask;
       s += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned c = g.slots[s].count;
    unsigned incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= (unsigned)d) incl += t;
    }
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    unsigned base = 0;
    if (lane == 31) base = atomicAdd(g.cursor, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (!c) continue;
    g.slots[s].start = base + (incl - c);
    if (!with_coarse) continue;
    // note this fine cell's coarse cell (2h); hg_dilate_kernel then marks the 27 coarse cells around
    // every occupied coarse cell — one dilation per coarse cell instead of one per fine cell
    int fx, fy, fz;
    unpack_key(g.slots[s].key1 - 1ull, fx, fy, fz);
    const unsigned long long key = pack_key(fx >> 1, fy >> 1, fz >> 1);
    unsigned long long slot = mix64(key) & g.mask;
    while (true) {
      unsigned long long k = ld_volatile_u64(reinterpret_cast<const uint64_t*>(g.coarse_keys + slot));
      if (k == key) break;
      if (k == T3D_KEY_EMPTY) {
        k = atomicCAS(g.coarse_keys + slot, T3D_KEY_EMPTY, key);
        if (k == T3D_KEY_EMPTY || k == key) break;
      }
      slot = (slot + 1) & g.mask;
    }
  }
}

__global__ void hg_dilate_kernel(const __grid_constant__ HGrid g) {
  for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s <= g.mask;
       s += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long ck = g.coarse_keys[s];
    if (ck == T3D_KEY_EMPTY) continue;
    int qx, qy, qz;
    unpack_key(ck, qx, qy, qz);
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const unsigned long long key = pack_key(qx + dx, qy + dy, qz + dz);
          unsigned long long slot = mix64(key) & g.mask;
          while (true) {
            unsigned long long k = ld_volatile_u64(reinterpret_cast<const uint64_t*>(g.near_keys + slot));
            if (k == key) break;
            if (k == T3D_KEY_EMPTY) {
              k = atomicCAS(g.near_keys + slot, T3D_KEY_EMPTY, key);
              if (k == T3D_KEY_EMPTY || k == key) break;
            }
            slot = (slot + 1) & g.mask;
          }
        }
  }
}

__device__ __forceinline__ bool hg_near_target(const HGrid& g, int cx, int cy, int cz) {
  const unsigned long long key = pack_key(cx >> 1, cy >> 1, cz >> 1);
  unsigned long long slot = mix64(key) & g.mask;
  while (true) {
    const unsigned long long k = __ldg(g.near_keys + slot);
    if (k == key) return true;
    if (k == T3D_KEY_EMPTY) return false;
    slot = (slot + 1) & g.mask;
  }
}

__global__ void hg_scatter_kernel(const float* __restrict__ tgt, const float* __restrict__ tgt_nrm,
                                  const __grid_constant__ HGrid g) {
  const long long n = hg_n(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int cx, cy, cz;
    const float x = tgt[3 * i], y = tgt[3 * i + 1], z = tgt[3 * i + 2];
    if (!hg_cell(g, x, y, z, cx, cy, cz)) continue;
    const unsigned long long key = pack_key(cx, cy, cz);
    unsigned long long slot = mix64(key) & g.mask;
    while (g.slots[slot].key1 != key + 1ull) slot = (slot + 1) & g.mask;
    const unsigned pos = g.slots[slot].start + atomicAdd(g.fill + slot, 1u);
    g.xyzi[pos] = make_float4(x, y, z, __uint_as_float((unsigned)i));
    g.nrm[3ll * pos] = tgt_nrm[3 * i]; g.nrm[3ll * pos + 1] = tgt_nrm[3 * i + 1]; g.nrm[3ll * pos + 2] = tgt_nrm[3 * i + 2];
  }
}

// Nearest target within sqrt(r2); ties -> lowest original index (as t3d_nn_within).
// Exact: cells are visited ring by ring around the query's cell; a cell is skipped
// only if its box is farther than the best distance so far, and the search stops
// after ring k once best <= (k*h + distance of q to its own cell's faces)^2.
//
// The result is decided in f64 (the oracle's arithmetic), but nearly every candidate and cell is
// REJECTED in f32: a candidate's f32 squared distance to the f32-rounded query differs from the exact
// one by at most M (nn_margin), so `d2f > best + M` proves it cannot beat or tie the incumbent; only the
// survivors (the eventual answer and its near-ties) pay the f32->f64 conversions and the f64 chain.
// Cell boxes are tested the same way with their own error bound.  Rejecting less is always safe.
struct NNState {
  double best;
  int bj;
  unsigned bo;
  float thr;      // candidates with an f32 squared distance above this cannot win
  float thr_box;  // cells whose f32 box distance exceeds this cannot hold a winner
};

// |d2_f32 - d2_exact| <= M for every candidate within r of the query: with e = 2^-23 (max|q| + 2r) bounding the
// error of one f32 coordinate difference (rounding q to f32 + the subtraction; |difference| <= 3h < 2r),
// sum(af^2) - sum(a^2) <= e (2 sqrt(3) r + 3 e), and the three roundings of the f32 sum add < 2^-22 r^2.
__device__ __forceinline__ double nn_margin(double qx, double qy, double qz, double r, double r2) {
  const double qmax = fmax(fmax(fabs(qx), fabs(qy)), fabs(qz));
  const double e = (qmax + 2.0 * r) * (1.0 / 8388608.0);
  return 4.0 * e * r + 4.0 * e * e + r2 * (1.0 / 2097152.0);
}

__device__ __forceinline__ void nn_refresh(NNState& st, double M, double box_abs) {
  st.thr = __double2float_ru(st.best + M);
  st.thr_box = __double2float_ru(st.best + (st.best * 1e-6 + box_abs));
}

// f32 squared distance from the query (offset (lx,ly,lz) inside its own cell) to the box of the cell at
// integer offset (dx,dy,dz); absolute error < 4e-6 h^2 (box_abs) + 1e-6 of the value
__device__ __forceinline__ float hg_box_d2f(float h, float lx, float ly, float lz, int dx, int dy, int dz) {
  const float ax = dx == 0 ? 0.0f : (dx < 0 ? lx + (float)(-dx - 1) * h : (h - lx) + (float)(dx - 1) * h);
  const float ay = dy == 0 ? 0.0f : (dy < 0 ? ly + (float)(-dy - 1) * h : (h - ly) + (float)(dy - 1) * h);
  const float az = dz == 0 ? 0.0f : (dz < 0 ? lz + (float)(-dz - 1) * h : (h - lz) + (float)(dz - 1) * h);
  return fmaf(az, az, fmaf(ay, ay, ax * ax));
}
// the same for the 3^3 neighbourhood (|d| <= 1)
__device__ __forceinline__ float hg_box_d2f_ring1(float h, float lx, float ly, float lz, int dx, int dy, int dz) {
  const float ax = dx == 0 ? 0.0f : (dx < 0 ? lx : h - lx);
  const float ay = dy == 0 ? 0.0f : (dy < 0 ? ly : h - ly);
  const float az = dz == 0 ? 0.0f : (dz < 0 ? lz : h - lz);
  return fmaf(az, az, fmaf(ay, ay, ax * ax));
}

struct NNQuery {
  double x, y, z;     // exact (f64) query
  float fx, fy, fz;   // rounded to f32
  double M, box_abs;  // error bounds of the f32 candidate / box tests
};

// one candidate (already loaded): f32 rejection, f64 decision
__device__ __forceinline__ void nn_candidate_v(const float4 t, unsigned j, const NNQuery& q, double r2, NNState& st) {
  const float ax = t.x - q.fx, ay = t.y - q.fy, az = t.z - q.fz;
  if (fmaf(az, az, fmaf(ay, ay, ax * ax)) > st.thr) return;
  const double ddx = (double)t.x - q.x, ddy = (double)t.y - q.y, ddz = (double)t.z - q.z;
  const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
  if (d2 < st.best || (d2 == st.best && d2 <= r2 && __float_as_uint(t.w) < st.bo)) {
    st.best = d2;
    st.bj = (int)j;
    st.bo = __float_as_uint(t.w);
    nn_refresh(st, q.M, q.box_abs);
  }
}
__device__ __forceinline__ void nn_candidate(const HGrid& g, unsigned j, const NNQuery& q, double r2, NNState& st) {
  nn_candidate_v(__ldg(g.xyzi + j), j, q, r2, st);
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint4 hg_load_slot(const HGrid& g, unsigned long long slot) {
  return __ldg(reinterpret_cast<const uint4*>(g.slots + slot));
}

// Point range [s, e) of the cell with key `key`, starting from its home slot whose contents `sl` were loaded
// earlier (the caller issues that load as soon as the cell is known, long before it needs the answer).
__device__ __forceinline__ void hg_cell_range(const HGrid& g, unsigned long long key, unsigned long long slot, uint4 sl,
                                              unsigned& s, unsigned& e) {
  const unsigned k1lo = (unsigned)(key + 1ull), k1hi = (unsigned)((key + 1ull) >> 32);
  s = 0; e = 0;
  while (true) {
    if (sl.x == k1lo && sl.y == k1hi) { s = sl.z; e = s + sl.w; break; }
    if ((sl.x | sl.y) == 0u) break;
    slot = (slot + 1) & g.mask;
    sl = hg_load_slot(g, slot);
  }
}

__device__ __forceinline__ void hg_scan_cell_from(const HGrid& g, unsigned long long key, unsigned long long slot,
                                                  uint4 sl, const NNQuery& q, double r2, NNState& st) {
  unsigned s, e;
  hg_cell_range(g, key, slot, sl, s, e);
  for (unsigned j = s; j < e; ++j) nn_candidate(g, j, q, r2, st);
}

__device__ __forceinline__ void hg_scan_cell(const HGrid& g, int cx, int cy, int cz, const NNQuery& q, double r2,
                                             NNState& st) {
  const unsigned long long key = pack_key(cx, cy, cz);
  const unsigned long long slot = mix64(key) & g.mask;
  hg_scan_cell_from(g, key, slot, hg_load_slot(g, slot), q, r2, st);
}

// (d2, original index) arg-min across the G lanes of a query group
template <int G>
__device__ __forceinline__ void nn_group_min(NNState& st, unsigned gmask) {
#pragma unroll
  for (int d = G / 2; d > 0; d >>= 1) {
    const double ob = __shfl_xor_sync(gmask, st.best, d);
    const int oj = __shfl_xor_sync(gmask, st.bj, d);
    const unsigned oo = __shfl_xor_sync(gmask, st.bo, d);
    if (oj >= 0 && (st.bj < 0 || ob < st.best || (ob == st.best && oo < st.bo))) {
      st.best = ob; st.bj = oj; st.bo = oo;
    }
  }
}

// 6x6 SPD solve by one warp: lane j < 6 holds column j of A, lane 6 the right-hand side.
// Gaussian elimination without pivoting (== LDL^T pivots, det = product of pivots), then
// back substitution; every lane ends with the same x.  Returns false when ill-posed (R8:
// |det| < 1e-6 or non-finite -> identity update).
__device__ __forceinline__ bool warp_solve6(const double* acc, double* x, unsigned lane) {
  double a[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double v = 0.0;
    if (lane < 6) {  // A[i][lane] from the packed upper triangle
      const int r = i < (int)lane ? i : (int)lane, c = i < (int)lane ? (int)lane : i;
      v = acc[r * 6 - r * (r - 1) / 2 + (c - r)];
    } else if (lane == 6) {
      v = -acc[21 + i];
    }
    a[i] = v;
  }
  double det = 1.0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double pk = __shfl_sync(0xffffffffu, a[k], k);
    det *= pk;
    if (pk == 0.0 || !isfinite(pk)) ok = false;
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      const double f = __shfl_sync(0xffffffffu, a[i], k) / pk;
      a[i] -= f * a[k];
    }
  }
  if (!ok || !isfinite(det) || fabs(det) < 1e-6) return false;
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double sacc = __shfl_sync(0xffffffffu, a[i], 6);
#pragma unroll
    for (int j = i + 1; j < 6; ++j) sacc -= __shfl_sync(0xffffffffu, a[i], j) * x[j];
    x[i] = sacc / __shfl_sync(0xffffffffu, a[i], i);
  }
  bool fin = true;
#pragma unroll
  for (int i = 0; i < 6; ++i) fin = fin && isfinite(x[i]);
  return fin;
}

struct IcpState {       // device-resident registration state (also the D2H result record)
  double T[16];
  double fitness, rmse;
  double acc[NACC];     // last linearisation
  int iterations, converged, done, round;
  unsigned ticket;
  int skipped;          // device-count mode: a cloud was smaller than min_points
  unsigned long long dbg[8];  // T3D_ICP_TIMING: globaltimer stamps of the last linearisation (ns)
};

// fixed pairwise tree over the per-warp rows of s (NW a power of two)
template <int NW>
__device__ __forceinline__ double tree_sum(const double (*s)[NACC], int k) {
  double t[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) t[i] = s[i][k];
#pragma unroll
  for (int h = 1; h < NW; h <<= 1)
#pragma unroll
    for (int i = 0; i + h < NW; i += 2 * h) t[i] += t[i + h];
  return t[0];
}

// The last CTA of a linearisation finishes the round: fixed-order sum of the per-CTA partials (warp w takes
// every NW-th row with 8 independent accumulators — deterministic), convergence bookkeeping as R8, the 6x6
// solve by one warp and the pose update.  Called by all NW warps of that CTA.
template <int NW>
__device__ __forceinline__ void icp_finish_round(IcpState* st, const double* partial, unsigned nb, long long n_src,
                                                 const HGrid& g, int min_points, int max_iter, double rel_fitness,
                                                 double rel_rmse, double (*s)[NACC]) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  {
    const int k = l;
    double a8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (k < NACC) {
      unsigned b = w;
      for (; b + 7 * NW < nb; b += 8 * NW) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          a8[u] += *reinterpret_cast<const volatile double*>(&partial[(long long)(b + NW * u) * NACC + k]);
      }
      for (int u = 0; b < nb; b += NW, ++u)
        a8[u & 7] += *reinterpret_cast<const volatile double*>(&partial[(long long)b * NACC + k]);
      s[w][k] = ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
    }
  }
  __syncthreads();
  if (threadIdx.x < NACC) st->acc[threadIdx.x] = tree_sum<NW>(s, threadIdx.x);
  __syncthreads();
  if (threadIdx.x == 0) st->dbg[3] = gtime();
  if (w == 0) {  // convergence bookkeeping (uniform across the warp) + warp-parallel 6x6 solve
    const int round = st->round;
    const double cnt = st->acc[28];
    const double f2 = cnt / (double)n_src;
    const double e2 = cnt > 0.0 ? sqrt(st->acc[27] / cnt) : 0.0;
    bool done = false;
    int converged = 0;
    if (round > 0 && fabs(st->fitness - f2) < rel_fitness && fabs(st->rmse - e2) < rel_rmse) {
      converged = 1;
      done = true;
    }
    if (!done && round >= max_iter) done = true;
    // device-count mode: clouds below min_points are not registered (pose = initial guess)
    const bool skipped = min_points > 0 && (n_src < min_points || hg_n(g) < min_points);
    if (skipped) done = true;
    double x[6] = {0, 0, 0, 0, 0, 0};
    bool solved = false;
    if (!done) solved = warp_solve6(st->acc, x, l);
    __syncwarp();
    if (l == 0) {
      st->dbg[4] = gtime();
      if (skipped) st->skipped = 1;
      if (round > 0) st->iterations = round;  // this linearisation closes iteration `round`
      if (converged) st->converged = 1;
      st->fitness = f2;
      st->rmse = e2;
      if (!done && solved) {  // ill-posed -> identity update (R8)
        double U[16], Tn[16], Tc[16];
        for (int i = 0; i < 16; ++i) Tc[i] = st->T[i];
        vec6_to_mat4(x, U);
        mat4_mul(U, Tc, Tn);
        for (int i = 0; i < 16; ++i) st->T[i] = Tn[i];
      }
      st->round = round + 1;
      st->ticket = 0;
      st->dbg[5] = gtime();
      __threadfence();
      st->done = done ? 1 : 0;
    }
  }
}

// The linearisation as its own launch (after a search with fewer than 8 lanes per query): one thread per
// correspondence accumulates the 29 sums, warp shuffles + a fixed tree per CTA, the last CTA finishes the round.
__global__ void __launch_bounds__(ICP_THREADS)
    icp_acc_kernel(const __grid_constant__ HGrid g, const float* __restrict__ src, long long n_src,
                   const long long* n_src_dev, int min_points, int max_iter, double rel_fitness, double rel_rmse,
                   IcpState* st, const int* __restrict__ corr, const double* __restrict__ corr_d2,
                   double* partial /* gridDim.x * NACC */) {
  if (*reinterpret_cast<volatile int*>(&st->done)) return;
  if (n_src_dev) { const long long v = *n_src_dev; n_src = v < n_src ? v : n_src; }
  __shared__ double s[ICP_THREADS / 32][NACC];
  __shared__ double s_T[12];
  __shared__ unsigned s_last;
  if (threadIdx.x < 12) s_T[threadIdx.x] = *reinterpret_cast<volatile double*>(&st->T[threadIdx.x]);
  __syncthreads();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_src;
       i += (long long)gridDim.x * blockDim.x) {
    const int j = corr[i];
    if (j < 0) continue;
    const double px = src[3 * i], py = src[3 * i + 1], pz = src[3 * i + 2];
    const double sx = s_T[0] * px + s_T[1] * py + s_T[2] * pz + s_T[3];
    const double sy = s_T[4] * px + s_T[5] * py + s_T[6] * pz + s_T[7];
    const double sz = s_T[8] * px + s_T[9] * py + s_T[10] * pz + s_T[11];
    const float4 tp = g.xyzi[j];
    const double tx = tp.x, ty = tp.y, tz = tp.z;
    const double nx = g.nrm[3ll * j], ny = g.nrm[3ll * j + 1], nz = g.nrm[3ll * j + 2];
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
    acc[27] += corr_d2[i];
    acc[28] += 1.0;
  }
#pragma unroll
  for (int k = 0; k < NACC; ++k)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], d);
  if (l == 0)
    for (int k = 0; k < NACC; ++k) s[w][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < NACC) partial[(long long)blockIdx.x * NACC + threadIdx.x] = tree_sum<ICP_THREADS / 32>(s, threadIdx.x);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  icp_finish_round<ICP_THREADS / 32>(st, partial, gridDim.x, n_src, g, min_points, max_iter, rel_fitness, rel_rmse, s);
}

// One linearisation = ONE launch that returns at once when the registration has already
// finished (st->done), so the host can enqueue several rounds without a round trip per
// iteration: icp_nn_kernel finds the correspondences and, fused into the same pass, the 8
// lanes of each query group share out the 29 sums of J^T J / J^T r; the last CTA to finish
// adds the per-CTA partials in blockIdx order (deterministic), checks convergence and
// solves the 6x6 system (icp_finish_round).
constexpr int NN_THREADS = 256;
constexpr int NN_GROUP = 8;  // lanes per query

// Exact nearest neighbour, EIGHT lanes per query (the clouds of one frame hold ~1e5 queries:
// one thread per query leaves the GPU short of warps and gives every warp a long dependent
// chain of hash probes).  Step 1: the 8 lanes probe the 8 cells of the query's octant (its own
// cell and the neighbours on the side the query sits in) at once — with h ~ 2-3 surface
// samples these almost always contain the answer.  Step 2: the other 19 cells of ring 1 are
// dealt round-robin; a cell is probed only if its box is closer than the best so far.
// Step 3 (rare): ring 2 the same way, only if something closer could still hide there.
// Between steps the group takes an arg-min over (d2, original index) with 3 shuffles.
__global__ void __launch_bounds__(NN_THREADS, 4)
    icp_nn_kernel(const __grid_constant__ HGrid g, const float* __restrict__ src, long long n_src,
                  const long long* n_src_dev, double r2, IcpState* st, int* __restrict__ corr,
                  double* __restrict__ corr_d2, int near_filter, int fuse, int min_points, int max_iter,
                  double rel_fitness, double rel_rmse, double* partial /* gridDim.x * NACC */) {
  if (*reinterpret_cast<const volatile int*>(&st->done)) return;
  // from the second linearisation on, corr[] holds the previous round's answer for this registration: it
  // seeds the search as the incumbent (exact: any valid candidate may), and because the pose moves little
  // between rounds the seed is almost always the answer already, so nearly every cell fails the box test
  const bool warm = *reinterpret_cast<const volatile int*>(&st->round) > 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) st->dbg[0] = gtime();
  if (n_src_dev) { const long long v = *n_src_dev; n_src = v < n_src ? v : n_src; }
  __shared__ double s_T[12];
  if (threadIdx.x < 12) s_T[threadIdx.x] = *reinterpret_cast<const volatile double*>(&st->T[threadIdx.x]);
  __syncthreads();
  const double r = sqrt(r2);
  const unsigned l = lane_id();
  const unsigned grp = l / NN_GROUP, sub = l % NN_GROUP;
  const unsigned gmask = 0xFFu << (grp * NN_GROUP);
  // Linearisation fused into the search (fuse != 0): every lane of a query group holds J0..J5 (computed from the
  // same broadcast loads), lane `sub` owns ONE of the group's 8 values {J0..J5, r, d2} and accumulates its products
  // with the six J's — lanes 0..5 the rows of J^T J, lane 6 J^T r, lane 7 sum d2 and the count: no shuffles.
  double accl[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long total_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  constexpr int QPW = 32 / NN_GROUP;  // queries per warp
  for (long long qb = warp_global * QPW; qb < n_src; qb += total_warps * QPW) {
    const long long i = qb + grp;
    if (i >= n_src) continue;  // group-uniform
    // every load that does not depend on a decision is issued up front: the point, last round's answer and the
    // target it names, and the home slots of this lane's octant cells — two dependent round trips to memory
    // before the candidates instead of five
    const int pj = warm ? corr[i] : -1;
    const double px = src[3 * i], py = src[3 * i + 1], pz = src[3 * i + 2];
    float4 seed = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pj >= 0) seed = __ldg(g.xyzi + pj);
    const double sx = s_T[0] * px + s_T[1] * py + s_T[2] * pz + s_T[3];
    const double sy = s_T[4] * px + s_T[5] * py + s_T[6] * pz + s_T[7];
    const double sz = s_T[8] * px + s_T[9] * py + s_T[10] * pz + s_T[11];
    int cx, cy, cz;
    const bool in_grid = hg_cell(g, sx, sy, sz, cx, cy, cz) && (!(near_filter & 1) || hg_near_target(g, cx, cy, cz));  // group-uniform
    float lx = 0.f, ly = 0.f, lz = 0.f;
    int ox = 1, oy = 1, oz = 1;
    const float hf = (float)g.h;
    unsigned long long okey = 0, oslot = 0;  // this lane's cell of the query's octant: cell `sub`
    uint4 osl = make_uint4(0u, 0u, 0u, 0u);
    if (in_grid) {
      lx = (float)(sx - (double)cx * g.h); ly = (float)(sy - (double)cy * g.h); lz = (float)(sz - (double)cz * g.h);
      const float hh = 0.5f * hf;
      ox = lx < hh ? -1 : 1; oy = ly < hh ? -1 : 1; oz = lz < hh ? -1 : 1;
      okey = pack_key(cx + ((sub & 1) ? ox : 0), cy + ((sub & 2) ? oy : 0), cz + ((sub & 4) ? oz : 0));
      oslot = mix64(okey) & g.mask;
      osl = hg_load_slot(g, oslot);
    }
    NNState nn;
    nn.best = r2; nn.bj = -1; nn.bo = 0xFFFFFFFFu;
    if (pj >= 0) {
      const double ddx = (double)seed.x - sx, ddy = (double)seed.y - sy, ddz = (double)seed.z - sz;
      const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
      if (d2 <= r2) { nn.best = d2; nn.bj = pj; nn.bo = __float_as_uint(seed.w); }
    }
    if (in_grid) {
      NNQuery q;
      q.x = sx; q.y = sy; q.z = sz;
      q.fx = (float)sx; q.fy = (float)sy; q.fz = (float)sz;
      q.M = nn_margin(sx, sy, sz, r, r2);
      q.box_abs = 4e-6 * g.h * g.h;
      nn_refresh(nn, q.M, q.box_abs);
      {  // step 1: the 8 cells of the octant.  Lane t resolved cell t's range and box distance; the GROUP then
         // visits the cells that can still hold a winner, one candidate per lane (a cell holds about as many
         // points as the group has lanes).
        unsigned cs, ce;
        hg_cell_range(g, okey, oslot, osl, cs, ce);
        const float bd = hg_box_d2f_ring1(hf, lx, ly, lz, (sub & 1) ? ox : 0, (sub & 2) ? oy : 0, (sub & 4) ? oz : 0);
        unsigned pend = (__ballot_sync(gmask, ce > cs && bd <= nn.thr_box) >> (grp * NN_GROUP)) & 0xFFu;
        // The cells that pass the box test against the incumbent at this point (the seed, from the second round on)
        // are visited two per trip — both candidate loads in flight together — with private incumbents per lane
        // and ONE arg-min at the end: a visit costs a memory round trip, an arg-min three dependent shuffle rounds.
        while (pend) {  // group-uniform
          const int t0 = __ffs(pend) - 1;
          pend &= pend - 1u;
          const int t1 = pend ? __ffs(pend) - 1 : t0;
          const bool two = pend != 0u;
          pend &= pend - 1u;
          const int f0 = (int)(grp * NN_GROUP) + t0, f1 = (int)(grp * NN_GROUP) + t1;
          const unsigned s0 = __shfl_sync(gmask, cs, f0), e0 = __shfl_sync(gmask, ce, f0);
          const unsigned s1 = __shfl_sync(gmask, cs, f1), e1 = two ? __shfl_sync(gmask, ce, f1) : s1;
          const bool va = s0 + sub < e0, vb = s1 + sub < e1;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          if (va) a = __ldg(g.xyzi + s0 + sub);
          if (vb) b = __ldg(g.xyzi + s1 + sub);
          if (va) nn_candidate_v(a, s0 + sub, q, r2, nn);
          if (vb) nn_candidate_v(b, s1 + sub, q, r2, nn);
          for (unsigned j0 = s0 + NN_GROUP; j0 < e0; j0 += NN_GROUP)  // cells with more points than lanes (rare)
            if (j0 + sub < e0) nn_candidate(g, j0 + sub, q, r2, nn);
          for (unsigned j0 = s1 + NN_GROUP; j0 < e1; j0 += NN_GROUP)
            if (j0 + sub < e1) nn_candidate(g, j0 + sub, q, r2, nn);
        }
        nn_group_min<NN_GROUP>(nn, gmask);
        nn_refresh(nn, q.M, q.box_abs);
      }
      // step 2: the rest of ring 1 (cells in index order; the octant's 8 were done in step 1).  Every one of those
      // 19 cells lies beyond the FAR face of the query's cell on at least one axis, i.e. at least
      // min_axis(max(l, h - l)) >= h / 2 away: with h ~ 2-3 surface samples the incumbent is almost always
      // closer than that, and the whole step (and its arg-min) is skipped — group-uniform.
      const float far_face = fminf(fminf(fmaxf(lx, hf - lx), fmaxf(ly, hf - ly)), fmaxf(lz, hf - lz));
      if (far_face * far_face <= nn.thr_box) {
#pragma unroll 1
        for (int c = (int)sub; c < 27; c += NN_GROUP) {
          const int dz = c / 9 - 1, dy = (c - (dz + 1) * 9) / 3 - 1, dx = c - (dz + 1) * 9 - (dy + 1) * 3 - 1;
          if ((dx == 0 || dx == ox) && (dy == 0 || dy == oy) && (dz == 0 || dz == oz)) continue;  // octant: done
          if (hg_box_d2f_ring1(hf, lx, ly, lz, dx, dy, dz) > nn.thr_box) continue;
          hg_scan_cell(g, cx + dx, cy + dy, cz + dz, q, r2, nn);
        }
        nn_group_min<NN_GROUP>(nn, gmask);
      }
      // after rings 0+1 everything closer than h + (distance of q to its cell's faces) has been seen; the f32
      // form of that radius is shrunk by more than its rounding error, so ring 2 is never skipped wrongly
      const float face = fmaxf(0.0f, fminf(fminf(fminf(lx, hf - lx), fminf(ly, hf - ly)), fminf(lz, hf - lz)));
      const float reach1 = (hf + face) * 0.99999f;
      if (!(near_filter & 2) && !(__double2float_ru(nn.best) <= __fmul_rd(reach1, reach1))) {  // step 3: ring 2 (group-uniform branch)
        nn_refresh(nn, q.M, q.box_abs);
#pragma unroll 1
        for (int c = 27 + (int)sub; c < 125; c += NN_GROUP) {
          const int dx = c_ofs[c][0], dy = c_ofs[c][1], dz = c_ofs[c][2];
          if (hg_box_d2f(hf, lx, ly, lz, dx, dy, dz) > nn.thr_box) continue;
          hg_scan_cell(g, cx + dx, cy + dy, cz + dz, q, r2, nn);
        }
        nn_group_min<NN_GROUP>(nn, gmask);
      }
    }
    if (sub == 0) {
      corr[i] = nn.bj;
      corr_d2[i] = nn.best;
    }
    if (fuse && nn.bj >= 0) {  // group-uniform
      const float4 tp = __ldg(g.xyzi + nn.bj);
      const double tx = tp.x, ty = tp.y, tz = tp.z;
      const double nx = __ldg(g.nrm + 3ll * nn.bj), ny = __ldg(g.nrm + 3ll * nn.bj + 1), nz = __ldg(g.nrm + 3ll * nn.bj + 2);
      const double res = (sx - tx) * nx + (sy - ty) * ny + (sz - tz) * nz;
      const double j0 = sy * nz - sz * ny, j1 = sz * nx - sx * nz, j2 = sx * ny - sy * nx;
      if (sub < 7) {
        const double v = sub == 0 ? j0 : sub == 1 ? j1 : sub == 2 ? j2 : sub == 3 ? nx : sub == 4 ? ny : sub == 5 ? nz : res;
        accl[0] += j0 * v; accl[1] += j1 * v; accl[2] += j2 * v;
        accl[3] += nx * v; accl[4] += ny * v; accl[5] += nz * v;
      } else {
        accl[0] += nn.best;
        accl[1] += 1.0;
      }
    }
  }
  if (!fuse) return;
  const unsigned long long t_loop = gtime();
  // warp: the four groups' sums; CTA: the eight warps' in a fixed tree; grid: the last CTA
  __syncwarp();
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    accl[a] += __shfl_xor_sync(0xffffffffu, accl[a], 8);
    accl[a] += __shfl_xor_sync(0xffffffffu, accl[a], 16);
  }
  __shared__ double s_acc[NN_THREADS / 32][NACC];
  __shared__ unsigned s_last;
  const int w = threadIdx.x >> 5;
  if (grp == 0) {  // lane `sub` holds row `sub` of the 7x6 product table [J;r] J^T: keep the packed upper triangle
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      if (sub < 6 && a >= (int)sub) s_acc[w][(int)sub * 6 - (int)sub * ((int)sub - 1) / 2 + (a - (int)sub)] = accl[a];
      if (sub == 6) s_acc[w][21 + a] = accl[a];
    }
    if (sub == 7) { s_acc[w][27] = accl[0]; s_acc[w][28] = accl[1]; }
  }
  __syncthreads();
  if (threadIdx.x < NACC) partial[(long long)blockIdx.x * NACC + threadIdx.x] = tree_sum<NN_THREADS / 32>(s_acc, threadIdx.x);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) { st->dbg[1] = t_loop; st->dbg[2] = gtime(); }
  icp_finish_round<NN_THREADS / 32>(st, partial, gridDim.x, n_src, g, min_points, max_iter, rel_fitness, rel_rmse, s_acc);
}

// correspondences as ORIGINAL target indices (t3d_icp_correspondences)
__global__ void icp_corr_export_kernel(const __grid_constant__ HGrid g, const int* __restrict__ corr,
                                       const double* __restrict__ corr_d2, long long n, int* __restrict__ out_idx,
                                       double* __restrict__ out_d2) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = corr[i];
  out_idx[i] = j < 0 ? -1 : (int)__float_as_uint(g.xyzi[j].w);
  out_d2[i] = corr_d2[i];
}

}  // namespace

// n_src / n_tgt are exact counts, or capacities when n_src_dev / n_tgt_dev point at the actual
// counts in device memory (frame-to-model tracking: the clouds were just produced on the device and
// their sizes never visit the host).
static int icp_fused(t3d_ctx* ctx, const float* src, int64_t n_src, const long long* n_src_dev, const float* tgt,
                     const float* tgt_nrm, int64_t n_tgt, const long long* n_tgt_dev, int min_points,
                     double max_corr, const double* T0, int max_iter, double rel_fitness,
                     double rel_rmse, t3d_icp_result* res, int* skipped_h, cudaStream_t st,
                     int* out_idx = nullptr, double* out_d2 = nullptr, const double* T1 = nullptr) {
  T3D_REQUIRE(n_tgt < (1ll << 31) && n_src < (1ll << 40), "icp: cloud too large");
  unsigned long long hc = 1024;
  while (hc < 2ull * (unsigned long long)n_tgt) hc <<= 1;
  int rc;
  if ((rc = ctx->scratch[0].reserve(hc * 36)) != T3D_OK) return rc;
  if ((rc = ctx->scratch[1].reserve((size_t)n_tgt * 28 + 64)) != T3D_OK) return rc;
  if (!ctx->icp_offsets_ready) {  // 5^3 offsets sorted by (ring, squared length)
    signed char h_ofs[125][4];
    int order[125], key[125];
    for (int i = 0; i < 125; ++i) {
      const int dx = i % 5 - 2, dy = (i / 5) % 5 - 2, dz = i / 25 - 2;
      const int ring = abs(dx) > abs(dy) ? (abs(dx) > abs(dz) ? abs(dx) : abs(dz)) : (abs(dy) > abs(dz) ? abs(dy) : abs(dz));
      key[i] = ring * 1000 + dx * dx + dy * dy + dz * dz;
      order[i] = i;
    }
    for (int a = 1; a < 125; ++a)  // insertion sort, stable
      for (int b = a; b > 0 && key[order[b - 1]] > key[order[b]]; --b) { int t = order[b]; order[b] = order[b - 1]; order[b - 1] = t; }
    for (int i = 0; i < 125; ++i) {
      const int o = order[i];
      h_ofs[i][0] = (signed char)(o % 5 - 2); h_ofs[i][1] = (signed char)((o / 5) % 5 - 2);
      h_ofs[i][2] = (signed char)(o / 25 - 2); h_ofs[i][3] = 0;
    }
    T3D_CUDA(cudaMemcpyToSymbol(c_ofs, h_ofs, sizeof(h_ofs)));
    ctx->icp_offsets_ready = true;
  }
  HGrid g;
  g.slots = ctx->scratch[0].as<HSlot>();
  g.fill = reinterpret_cast<unsigned*>(g.slots + hc);
  g.near_keys = reinterpret_cast<unsigned long long*>(g.fill + hc);
  g.coarse_keys = g.near_keys + hc;
  g.mask = hc - 1;
  g.xyzi = ctx->scratch[1].as<float4>();
  g.nrm = reinterpret_cast<float*>(g.xyzi + n_tgt);
  g.h = max_corr * 0.5;
  g.inv_h = 1.0 / g.h;
  g.n = n_tgt;
  g.n_dev = n_tgt_dev;
  // T3D_ICP_FUSE=0 (tuning): search and linearisation as two launches.  One search CTA per resident slot (4 per SM
  // at 64 registers): the persistent loop covers the queries, and the last CTA's pass over the per-CTA partials
  // shrinks with the grid.
  static int fuse = -1;
  if (fuse < 0) { const char* e = getenv("T3D_ICP_FUSE"); fuse = (e && atoi(e) == 0) ? 0 : 1; }
  const long long want_nn = (n_src * NN_GROUP + NN_THREADS - 1) / NN_THREADS;
  const int grid_nn = (int)(want_nn < (long long)ctx->num_sms * 4 ? (want_nn > 0 ? want_nn : 1) : (long long)ctx->num_sms * 4);
  const long long want_acc = (n_src + ICP_THREADS - 1) / ICP_THREADS;
  const int grid_acc = (int)(want_acc < (long long)ctx->num_sms ? (want_acc > 0 ? want_acc : 1) : (long long)ctx->num_sms);
  const size_t part_rows = (size_t)(grid_nn > grid_acc ? grid_nn : grid_acc);
  if ((rc = ctx->scratch[6].reserve(sizeof(double) * (part_rows * NACC) + sizeof(IcpState) + 64)) != T3D_OK) return rc;
  double* partial = ctx->scratch[6].as<double>();
  IcpState* dst = reinterpret_cast<IcpState*>(partial + part_rows * NACC);
  if ((rc = ctx->scratch[2].reserve(64)) != T3D_OK) return rc;
  g.cursor = ctx->scratch[2].as<unsigned>();
  if ((rc = ctx->scratch[3].reserve((size_t)n_src * 4)) != T3D_OK) return rc;
  if ((rc = ctx->scratch[4].reserve((size_t)n_src * 8)) != T3D_OK) return rc;
  int* corr = ctx->scratch[3].as<int>();
  double* corr_d2 = ctx->scratch[4].as<double>();

  IcpState* hst = reinterpret_cast<IcpState*>(reinterpret_cast<char*>(ctx->pinned) + 2048);
  unsigned* hflag = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ctx->pinned) + 1024);
  memset(hst, 0, sizeof(IcpState));
  for (int i = 0; i < 16; ++i) hst->T[i] = T0[i];
  const double r2 = max_corr * max_corr;
  const int bgrid = (int)((n_tgt + 255) / 256 < (long long)ctx->num_sms * 8 ? (n_tgt + 255) / 256 : (long long)ctx->num_sms * 8);
  const int hgrid = (int)((hc + 255) / 256 < 2368 ? (hc + 255) / 256 : 2368);

  // grid build + initial state + `rounds` linearisations + result read-back, all on stream st
  // The dilated coarse occupancy set answers "no target within reach" with one probe; it pays when many
  // queries are far from the target (a bad initial guess, partial overlap).  Frame-to-model tracking
  // (device-count mode) registers a frame against the surface it was predicted from: nearly every query has
  // a target nearby, so the set's construction (a table scan + 27 inserts per coarse cell) is skipped.
  const int near_filter = (n_src_dev == nullptr ? 1 : 0) | (getenv("T3D_ICP_DEBUG_NO_RING2") ? 2 : 0);
  auto enqueue_all = [&](int rounds, bool with_build) -> int {
    if (with_build) {
      T3D_CUDA(cudaMemsetAsync(g.slots, 0, hc * 20, st));  // slots + fill
      if (near_filter & 1) T3D_CUDA(cudaMemsetAsync(g.near_keys, 0xFF, hc * 16, st));  // near_keys + coarse_keys
      T3D_CUDA(cudaMemsetAsync(g.cursor, 0, 64, st));
      hg_count_kernel<<<bgrid, 256, 0, st>>>(tgt, g);
      hg_alloc_kernel<<<hgrid, 256, 0, st>>>(g, near_filter & 1);
      if (near_filter & 1) hg_dilate_kernel<<<hgrid, 256, 0, st>>>(g);
      hg_scatter_kernel<<<bgrid, 256, 0, st>>>(tgt, tgt_nrm, g);
      T3D_LAUNCH_CHECK();
      T3D_CUDA(cudaMemcpyAsync(dst, hst, sizeof(IcpState), cudaMemcpyHostToDevice, st));
    }
    for (int r = 0; r < rounds; ++r) {
      icp_nn_kernel<<<grid_nn, NN_THREADS, 0, st>>>(g, src, (long long)n_src, n_src_dev, r2, dst, corr, corr_d2, near_filter,
                                                    fuse, min_points, max_iter, rel_fitness, rel_rmse, partial);
      if (!fuse)
        icp_acc_kernel<<<grid_acc, ICP_THREADS, 0, st>>>(g, src, (long long)n_src, n_src_dev, min_points, max_iter,
                                                          rel_fitness, rel_rmse, dst, corr, corr_d2, partial);
    }
    T3D_LAUNCH_CHECK();
    T3D_CUDA(cudaMemcpyAsync(hst, dst, sizeof(IcpState), cudaMemcpyDeviceToHost, st));
    T3D_CUDA(cudaMemcpyAsync(hflag, g.cursor, 8, cudaMemcpyDeviceToHost, st));
    return T3D_OK;
  };

  if (out_idx) {  // t3d_icp_correspondences: one cold search at T0, optionally one warm search at T1
    if ((rc = enqueue_all(0, true)) != T3D_OK) return rc;
    icp_nn_kernel<<<grid_nn, NN_THREADS, 0, st>>>(g, src, (long long)n_src, n_src_dev, r2, dst, corr, corr_d2, near_filter,
                                                     0, 0, 0, 0.0, 0.0, partial);
    T3D_LAUNCH_CHECK();
    T3D_CUDA(cudaStreamSynchronize(st));
    if (T1) {
      for (int i = 0; i < 16; ++i) hst->T[i] = T1[i];
      hst->round = 1;
      T3D_CUDA(cudaMemcpyAsync(dst, hst, sizeof(IcpState), cudaMemcpyHostToDevice, st));
      icp_nn_kernel<<<grid_nn, NN_THREADS, 0, st>>>(g, src, (long long)n_src, n_src_dev, r2, dst, corr, corr_d2, near_filter,
                                                     0, 0, 0, 0.0, 0.0, partial);
    }
    icp_corr_export_kernel<<<(unsigned)((n_src + 255) / 256), 256, 0, st>>>(g, corr, corr_d2, (long long)n_src, out_idx, out_d2);
    T3D_LAUNCH_CHECK();
    T3D_CUDA(cudaStreamSynchronize(st));
    if (hflag[1]) {
      t3d_set_error("icp: target coordinates exceed +-2^19 * max_corr_dist");
      return T3D_E_NUMERIC;
    }
    return T3D_OK;
  }

  // rounds needed: at most max_iter + 1.  Enqueue speculatively (finished rounds are no-op
  // launches), read the state back, continue only if the registration is still running.
  // On a capturable stream the first chunk (build + 6 rounds + read-back, ~24 operations that
  // are each a few microseconds long) is one cached CUDA graph launch.
  int enqueued = 0;
  const int first = max_iter + 1 < 6 ? max_iter + 1 : 6;
  static int use_graph = -1;
  if (use_graph < 0) { const char* e = getenv("T3D_ICP_GRAPH"); use_graph = (e && atoi(e) == 0) ? 0 : 1; }
  bool launched = false;
  if (use_graph && st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread) {
    char key[512];
    snprintf(key, sizeof(key), "icp|%p|%p|%p|%p|%p|%lld|%lld|%.17g|%d|%d|%.17g|%.17g|%p|%p|%p|%p|%p|%p|%d", (const void*)src,
             (const void*)n_src_dev, (const void*)tgt, (const void*)tgt_nrm, (const void*)n_tgt_dev, (long long)n_src,
             (long long)n_tgt, max_corr, max_iter, min_points, rel_fitness, rel_rmse, ctx->scratch[0].p,
             ctx->scratch[1].p, ctx->scratch[2].p, ctx->scratch[3].p, ctx->scratch[4].p, ctx->scratch[6].p, first + 100 * fuse);
    cudaGraphExec_t exec = nullptr;
    for (auto& cg : ctx->graphs)
      if (cg.key == key) exec = cg.exec;
    if (!exec) {
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int erc = enqueue_all(first, true);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (erc == T3D_OK && ce == cudaSuccess && graph &&
            cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
          if (ctx->graphs.size() >= 8) {  // bounded cache
            for (auto& cg : ctx->graphs) cudaGraphExecDestroy(cg.exec);
            ctx->graphs.clear();
          }
          ctx->graphs.push_back({key, exec});
        } else {
          exec = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();  // a failed capture must not poison later calls
      }
    }
    if (exec) {
      T3D_CUDA(cudaGraphLaunch(exec, st));
      launched = true;
    }
  }
  if (!launched) {
    if ((rc = enqueue_all(first, true)) != T3D_OK) return rc;
  }
  ctx->launches += 4 + first * (fuse ? 1 : 2);
  enqueued = first;
  T3D_CUDA(cudaStreamSynchronize(st));
  while (!hst->done && enqueued < max_iter + 1) {
    int n = max_iter + 1 - enqueued;
    if (n > 4) n = 4;  // a finished round costs ~4.5 us as a no-op launch, a host round trip ~20 us
    if ((rc = enqueue_all(n, false)) != T3D_OK) return rc;
    ctx->launches += n * (fuse ? 1 : 2);
    enqueued += n;
    T3D_CUDA(cudaStreamSynchronize(st));
  }
  if (hflag[1]) {
    t3d_set_error("icp: target coordinates exceed +-2^19 * max_corr_dist");
    return T3D_E_NUMERIC;
  }
  if (getenv("T3D_ICP_TIMING"))
    fprintf(stderr, "icp timing (last round, us): start->lastCTA loop end %.1f | ->ticket %.1f | ->partials summed %.1f | ->solved %.1f | ->state written %.1f  (rounds %d)\n",
            (hst->dbg[1] - hst->dbg[0]) * 1e-3, (hst->dbg[2] - hst->dbg[1]) * 1e-3, (hst->dbg[3] - hst->dbg[2]) * 1e-3,
            (hst->dbg[4] - hst->dbg[3]) * 1e-3, (hst->dbg[5] - hst->dbg[4]) * 1e-3, hst->round);
  for (int i = 0; i < 16; ++i) res->T[i] = hst->T[i];
  res->fitness = hst->fitness;
  res->inlier_rmse = hst->rmse;
  res->iterations = hst->iterations;
  res->converged = hst->converged;
  res->correspondences = (int64_t)hst->acc[28];
  if (skipped_h) *skipped_h = hst->skipped;
  return T3D_OK;
}

extern "C" int t3d_icp_linearize(t3d_ctx* ctx, const float* src, int64_t n_src, const float* tgt,
                                 const float* tgt_nrm, int64_t n_tgt, double max_corr_dist,
                                 const double* T_h, double* out27_h, double* out_stats_h,
                                 t3d_stream stream) {
  T3D_REQUIRE(ctx && T_h && out27_h && out_stats_h, "t3d_icp_linearize: null argument");
  T3D_ON_DEVICE(ctx->device);
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
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(max_corr_dist > 0.0 && max_iter >= 0, "t3d_icp_point_to_plane: bad parameters");
  memset(res, 0, sizeof(*res));
  for (int i = 0; i < 16; ++i) res->T[i] = T0_h ? T0_h[i] : ((i % 5 == 0) ? 1.0 : 0.0);
  if (n_src == 0 || n_tgt == 0) return T3D_OK;
  T3D_REQUIRE(src && tgt && tgt_nrm, "t3d_icp_point_to_plane: null clouds");
  cudaStream_t st = as_stream(stream);
  return icp_fused(ctx, src, n_src, nullptr, tgt, tgt_nrm, n_tgt, nullptr, 0, max_corr_dist, res->T, max_iter,
                   rel_fitness, rel_rmse, res, nullptr, st);
}

extern "C" int t3d_icp_correspondences(t3d_ctx* ctx, const float* src, int64_t n_src, const float* tgt,
                                       const float* tgt_nrm, int64_t n_tgt, double max_corr_dist,
                                       const double* T0_h, const double* T1_h, int32_t* out_idx, double* out_d2,
                                       t3d_stream stream) {
  T3D_REQUIRE(ctx && src && tgt && tgt_nrm && T0_h && out_idx && out_d2 && n_src > 0 && n_tgt > 0,
              "t3d_icp_correspondences: null argument");
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(max_corr_dist > 0.0, "t3d_icp_correspondences: bad parameters");
  t3d_icp_result res;
  memset(&res, 0, sizeof(res));
  return icp_fused(ctx, src, n_src, nullptr, tgt, tgt_nrm, n_tgt, nullptr, 0, max_corr_dist, T0_h, 0, 0.0, 0.0, &res,
                   nullptr, as_stream(stream), out_idx, out_d2, T1_h);
}

extern "C" int t3d_icp_point_to_plane_dev(t3d_ctx* ctx, const float* src, int64_t src_capacity,
                                          const int64_t* n_src_dev, const float* tgt, const float* tgt_nrm,
                                          int64_t tgt_capacity, const int64_t* n_tgt_dev, int min_points,
                                          double max_corr_dist, const double* T0_h, int max_iter,
                                          double rel_fitness, double rel_rmse, t3d_icp_result* res,
                                          int* skipped_h, t3d_stream stream) {
  T3D_REQUIRE(ctx && res && src && tgt && tgt_nrm && n_src_dev && n_tgt_dev && src_capacity > 0 && tgt_capacity > 0,
              "t3d_icp_point_to_plane_dev: null argument");
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(max_corr_dist > 0.0 && max_iter >= 0, "t3d_icp_point_to_plane_dev: bad parameters");
  memset(res, 0, sizeof(*res));
  for (int i = 0; i < 16; ++i) res->T[i] = T0_h ? T0_h[i] : ((i % 5 == 0) ? 1.0 : 0.0);
  return icp_fused(ctx, src, src_capacity, reinterpret_cast<const long long*>(n_src_dev), tgt, tgt_nrm,
                   tgt_capacity, reinterpret_cast<const long long*>(n_tgt_dev), min_points, max_corr_dist, res->T,
                   max_iter, rel_fitness, rel_rmse, res, skipped_h, as_stream(stream));
}
