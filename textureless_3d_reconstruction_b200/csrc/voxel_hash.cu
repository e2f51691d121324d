// voxel_hash.cu — K2: voxel-grid downsample / dedup on a GPU open-addressing
// hash, plus the small device-wide utilities it shares with K3 (bounds,
// ordered row compaction, radix sort of packed voxel keys).
//
// Replaces Open3D PointCloud::VoxelDownSample as called by
//   DenseReconstructor.merge_pointclouds        depth_to_reconstruction.py:386-420
//   DensePointCloudGenerator.merge_pointclouds  depth_enhanced_reconstruction.py:615-645
// Semantics: SURVEY.md §8c R2 (restated in oracle/t3d_oracle.c: o_voxel_downsample).
//   minb = min(p) - 0.5*v ; idx = floor((p - minb) / v) in f64, true division;
//   per voxel: mean xyz (f64), mean colour.  Colours are accumulated as exact
//   integer sums and divided once.
//
// Table layout (wait-free): 64-bit packed key per slot; accumulators are
// indexed by slot (SoA: sum x/y/z f64, sum r/g/b u32, count u32) so a thread
// that loses the CAS race never waits for the winner.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------
// bounds
// ---------------------------------------------------------------------------
template <typename T>
__global__ void bounds_kernel(const T* __restrict__ xyz, long long n, double* partial) {
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double v = (double)xyz[i * 3 + c];
      mn[c] = fmin(mn[c], v);
      mx[c] = fmax(mx[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn[c] = fmin(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], d));
      mx[c] = fmax(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
    }
  }
  __shared__ double s[8][6];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    for (int c = 0; c < 3; ++c) { s[w][c] = mn[c]; s[w][3 + c] = mx[c]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double r = s[0][threadIdx.x];
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
      r = threadIdx.x < 3 ? fmin(r, s[k][threadIdx.x]) : fmax(r, s[k][threadIdx.x]);
    partial[blockIdx.x * 6 + threadIdx.x] = r;
  }
}

__global__ void bounds_final_kernel(const double* partial, int nblocks, double* out6) {
  const int c = threadIdx.x;
  if (c >= 6) return;
  double r = partial[c];
  for (int b = 1; b < nblocks; ++b)
    r = c < 3 ? fmin(r, partial[b * 6 + c]) : fmax(r, partial[b * 6 + c]);
  out6[c] = r;
}

// ---------------------------------------------------------------------------
// voxel hash
// ---------------------------------------------------------------------------
struct VoxTable {
  unsigned long long* keys;
  double* sx; double* sy; double* sz;
  unsigned* sr; unsigned* sg; unsigned* sb;
  unsigned* cnt;
  unsigned long long mask;
  int* flags;  // [0] table full, [1] index out of packable range
};

template <typename T>
__global__ void __launch_bounds__(256)
    voxel_insert_kernel(const T* __restrict__ xyz, const uint8_t* __restrict__ rgb,
                        long long n, double minx, double miny, double minz, double voxel,
                        const __grid_constant__ VoxTable tb) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double px = (double)xyz[i * 3 + 0], py = (double)xyz[i * 3 + 1],
                 pz = (double)xyz[i * 3 + 2];
    // floor((p - minb) / v): IEEE subtract + true division, no contraction
    const double fx = floor(__ddiv_rn(__dsub_rn(px, minx), voxel));
    const double fy = floor(__ddiv_rn(__dsub_rn(py, miny), voxel));
    const double fz = floor(__ddiv_rn(__dsub_rn(pz, minz), voxel));
    if (!(fx >= 0.0 && fx < 2097152.0 && fy >= 0.0 && fy < 2097152.0 && fz >= 0.0 &&
          fz < 2097152.0)) {
      tb.flags[1] = 1;
      continue;
    }
    const unsigned long long key = ((unsigned long long)(unsigned)(int)fx << 42) |
                                   ((unsigned long long)(unsigned)(int)fy << 21) |
                                   (unsigned long long)(unsigned)(int)fz;
    unsigned long long slot = mix64(key) & tb.mask;
    bool placed = false;
    for (unsigned long long probe = 0; probe <= tb.mask; ++probe) {
      unsigned long long k = ld_volatile_u64(reinterpret_cast<const uint64_t*>(tb.keys + slot));
      if (k == T3D_KEY_EMPTY) k = atomicCAS(tb.keys + slot, T3D_KEY_EMPTY, key);
      if (k == T3D_KEY_EMPTY || k == key) { placed = true; break; }
      slot = (slot + 1) & tb.mask;
    }
    if (!placed) { tb.flags[0] = 1; continue; }
    atomicAdd(tb.sx + slot, px);
    atomicAdd(tb.sy + slot, py);
    atomicAdd(tb.sz + slot, pz);
    if (rgb) {
      atomicAdd(tb.sr + slot, (unsigned)rgb[i * 3 + 0]);
      atomicAdd(tb.sg + slot, (unsigned)rgb[i * 3 + 1]);
      atomicAdd(tb.sb + slot, (unsigned)rgb[i * 3 + 2]);
    }
    atomicAdd(tb.cnt + slot, 1u);
  }
}

// occupied slots -> dense list of (key, slot); warp-aggregated append
__global__ void voxel_collect_kernel(const __grid_constant__ VoxTable tb,
                                     unsigned long long* out_keys, unsigned* out_slots,
                                     unsigned long long* out_m) {
  const unsigned long long cap = tb.mask + 1;
  for (unsigned long long s0 = (unsigned long long)blockIdx.x * blockDim.x; s0 < cap;
       s0 += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long s = s0 + threadIdx.x;
    const bool occ = s < cap && tb.keys[s] != T3D_KEY_EMPTY;
    const unsigned b = __ballot_sync(0xffffffffu, occ);
    if (b == 0) continue;
    unsigned long long base = 0;
    if (lane_id() == 0) base = atomicAdd(out_m, (unsigned long long)__popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (occ) {
      const unsigned long long o = base + __popc(b & lanemask_lt());
      out_keys[o] = tb.keys[s];
      out_slots[o] = (unsigned)s;
    }
  }
}

__global__ void voxel_finalize_kernel(const __grid_constant__ VoxTable tb,
                                      const unsigned* __restrict__ slots, long long m,
                                      int has_rgb, double* out_xyz, uint8_t* out_rgb,
                                      unsigned* out_rgb_sum, unsigned* out_count,
                                      int* out_idx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned s = slots[i];
    const unsigned c = tb.cnt[s];
    const double dn = (double)c;
    out_xyz[i * 3 + 0] = __ddiv_rn(tb.sx[s], dn);
    out_xyz[i * 3 + 1] = __ddiv_rn(tb.sy[s], dn);
    out_xyz[i * 3 + 2] = __ddiv_rn(tb.sz[s], dn);
    const unsigned r = has_rgb ? tb.sr[s] : 0u, g = has_rgb ? tb.sg[s] : 0u,
                   b = has_rgb ? tb.sb[s] : 0u;
    if (out_rgb && has_rgb) {
      // (mean(c/255) * 255).astype(uint8) — truncation, d2r:417-418
      out_rgb[i * 3 + 0] = (uint8_t)(int)__dmul_rn(__ddiv_rn(__ddiv_rn((double)r, 255.0), dn), 255.0);
      out_rgb[i * 3 + 1] = (uint8_t)(int)__dmul_rn(__ddiv_rn(__ddiv_rn((double)g, 255.0), dn), 255.0);
      out_rgb[i * 3 + 2] = (uint8_t)(int)__dmul_rn(__ddiv_rn(__ddiv_rn((double)b, 255.0), dn), 255.0);
    }
    if (out_rgb_sum) { out_rgb_sum[i * 3] = r; out_rgb_sum[i * 3 + 1] = g; out_rgb_sum[i * 3 + 2] = b; }
    if (out_count) out_count[i] = c;
    if (out_idx) {
      const unsigned long long k = tb.keys[s];
      out_idx[i * 3 + 0] = (int)((k >> 42) & 0x1FFFFF);
      out_idx[i * 3 + 1] = (int)((k >> 21) & 0x1FFFFF);
      out_idx[i * 3 + 2] = (int)(k & 0x1FFFFF);
    }
  }
}

// ---------------------------------------------------------------------------
// LSD radix sort of (64-bit key, 32-bit payload), 8 bits per pass.  Stable.
// Used to give K2's output a canonical (ix,iy,iz)-ascending order.
// ---------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;                    // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per CTA

__global__ void __launch_bounds__(RS_THREADS)
    rs_hist_kernel(const unsigned long long* __restrict__ keys, long long n, int shift,
                   unsigned* __restrict__ hist /* [256][ntiles] */, int ntiles) {
  __shared__ unsigned s_h[256];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * RS_TILE;
  for (int j = 0; j < RS_ITEMS; ++j) {
    const long long i = base + j * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 255], 1u);
  }
  __syncthreads();
  hist[(long long)threadIdx.x * ntiles + blockIdx.x] = s_h[threadIdx.x];
}

// exclusive scan over hist[256*ntiles] in digit-major order (single CTA)
__global__ void __launch_bounds__(1024) rs_scan_kernel(unsigned* hist, long long total) {
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (long long base = 0; base < total; base += 1024) {
    const long long i = base + threadIdx.x;
    const unsigned v = i < total ? hist[i] : 0u;
    unsigned inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
      if ((threadIdx.x & 31) >= d) inc += t;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned w = s_w[threadIdx.x];
      unsigned winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, winc, d);
        if (threadIdx.x >= d) winc += t;
      }
      s_w[threadIdx.x] = winc - w;
    }
    __syncthreads();
    const unsigned carry = s_carry;
    if (i < total) hist[i] = carry + s_w[threadIdx.x >> 5] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_w[31] + inc;
    __syncthreads();
  }
}

// stable scatter: ranks inside the tile come from per-warp match + ordered
// warp-by-warp / item-by-item offsets.
__global__ void __launch_bounds__(RS_THREADS)
    rs_scatter_kernel(const unsigned long long* __restrict__ keys_in,
                      const unsigned* __restrict__ vals_in, unsigned long long* keys_out,
                      unsigned* vals_out, long long n, int shift,
                      const unsigned* __restrict__ hist, int ntiles) {
  __shared__ unsigned s_off[256];                 // running global offset per digit
  __shared__ unsigned short s_wc[RS_THREADS / 32][256];  // per-warp digit counts for one item row
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  s_off[threadIdx.x] = hist[(long long)threadIdx.x * ntiles + blockIdx.x];
  const long long base = (long long)blockIdx.x * RS_TILE;
  for (int j = 0; j < RS_ITEMS; ++j) {
    for (int d = lane; d < 256; d += 32) s_wc[warp][d] = 0;
    __syncthreads();
    const long long i = base + j * RS_THREADS + threadIdx.x;
    const bool ok = i < n;
    unsigned long long k = 0;
    unsigned digit = 0, peers = 0;
    if (ok) {
      k = keys_in[i];
      digit = (unsigned)(k >> shift) & 255u;
    }
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      peers = __match_any_sync(act, digit);
      if (lane == __ffs(peers) - 1) s_wc[warp][digit] = (unsigned short)__popc(peers);
    }
    __syncthreads();
    if (ok) {
      unsigned before = 0;
      for (int w = 0; w < warp; ++w) before += s_wc[w][digit];
      const unsigned o = s_off[digit] + before + __popc(peers & lanemask_lt());
      keys_out[o] = k;
      vals_out[o] = vals_in[i];
    }
    __syncthreads();
    {  // advance the running offsets by this row's totals
      unsigned tot = 0;
      for (int w = 0; w < RS_THREADS / 32; ++w) tot += s_wc[w][threadIdx.x];
      s_off[threadIdx.x] += tot;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// ordered row compaction
// ---------------------------------------------------------------------------
constexpr int CP_TILE = 2048;

__global__ void __launch_bounds__(256)
    compact_count_kernel(const uint8_t* __restrict__ mask, long long n, unsigned* counts) {
  const long long base = (long long)blockIdx.x * CP_TILE;
  unsigned c = 0;
  for (int j = 0; j < CP_TILE / 256; ++j) {
    const long long i = base + j * 256 + threadIdx.x;
    c += (i < n && mask[i] != 0) ? 1u : 0u;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  __shared__ unsigned s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int k = 0; k < 8; ++k) t += s[k];
    counts[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256)
    compact_scatter_kernel(const uint8_t* __restrict__ rows, const uint8_t* __restrict__ mask,
                           long long n, int row_bytes, const unsigned* __restrict__ offs,
                           long long ntiles, uint8_t* out, long long* out_n) {
  __shared__ unsigned s_w[8];
  __shared__ unsigned s_run;
  const long long base = (long long)blockIdx.x * CP_TILE;
  if (threadIdx.x == 0) s_run = offs[blockIdx.x];
  __syncthreads();
  for (int j = 0; j < CP_TILE / 256; ++j) {
    const long long i = base + j * 256 + threadIdx.x;
    const bool keep = i < n && mask[i] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    unsigned before = s_run;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += s_w[w];
    if (keep) {
      const long long o = (long long)before + __popc(b & lanemask_lt());
      const uint8_t* src = rows + i * row_bytes;
      uint8_t* dst = out + o * row_bytes;
      if ((row_bytes & 7) == 0) {
        for (int q = 0; q < row_bytes / 8; ++q)
          reinterpret_cast<uint64_t*>(dst)[q] = reinterpret_cast<const uint64_t*>(src)[q];
      } else if ((row_bytes & 3) == 0) {
        for (int q = 0; q < row_bytes / 4; ++q)
          reinterpret_cast<uint32_t*>(dst)[q] = reinterpret_cast<const uint32_t*>(src)[q];
      } else {
        for (int q = 0; q < row_bytes; ++q) dst[q] = src[q];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned t = 0;
      for (int w = 0; w < 8; ++w) t += s_w[w];
      s_run += t;
    }
    __syncthreads();
  }
  if (blockIdx.x == ntiles - 1 && threadIdx.x == 0) *out_n = (long long)s_run;
}

}  // namespace

// exclusive scan helper shared with other files
int t3d_exclusive_scan_u32(unsigned* data, long long n, cudaStream_t st) {
  rs_scan_kernel<<<1, 1024, 0, st>>>(data, n);
  T3D_LAUNCH_CHECK();
  return T3D_OK;
}

// sort (keys, vals) ascending by key; uses bits [0, key_bits).  Result ends in
// keys_a/vals_a (ping-pong with keys_b/vals_b, even number of passes enforced).
int t3d_radix_sort_u64(t3d_ctx* ctx, unsigned long long* keys_a, unsigned* vals_a,
                       unsigned long long* keys_b, unsigned* vals_b, long long n,
                       int key_bits, cudaStream_t st) {
  if (n <= 1) return T3D_OK;
  int passes = (key_bits + 7) / 8;
  if (passes & 1) ++passes;
  const int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
  int rc = ctx->scratch[7].reserve(sizeof(unsigned) * 256ull * ntiles);
  if (rc != T3D_OK) return rc;
  unsigned* hist = ctx->scratch[7].as<unsigned>();
  unsigned long long* ki = keys_a; unsigned* vi = vals_a;
  unsigned long long* ko = keys_b; unsigned* vo = vals_b;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * 8;
    rs_hist_kernel<<<ntiles, RS_THREADS, 0, st>>>(ki, n, shift, hist, ntiles);
    T3D_LAUNCH_CHECK();
    rs_scan_kernel<<<1, 1024, 0, st>>>(hist, 256ll * ntiles);
    T3D_LAUNCH_CHECK();
    rs_scatter_kernel<<<ntiles, RS_THREADS, 0, st>>>(ki, vi, ko, vo, n, shift, hist, ntiles);
    T3D_LAUNCH_CHECK();
    ctx->launches += 3;
    unsigned long long* tk = ki; ki = ko; ko = tk;
    unsigned* tv = vi; vi = vo; vo = tv;
  }
  return T3D_OK;
}

extern "C" int t3d_bounds(t3d_ctx* ctx, const void* xyz, int xyz_is_f64, int64_t n,
                          double* min_h, double* max_h, t3d_stream stream) {
  T3D_REQUIRE(ctx && min_h && max_h, "t3d_bounds: null argument");
  T3D_REQUIRE(n > 0 && xyz, "t3d_bounds: empty cloud");
  cudaStream_t st = as_stream(stream);
  const int grid = ctx->num_sms * 4;
  int rc = ctx->scratch[6].reserve(sizeof(double) * 6 * (grid + 1));
  if (rc != T3D_OK) return rc;
  double* partial = ctx->scratch[6].as<double>();
  if (xyz_is_f64)
    bounds_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(xyz), n, partial);
  else
    bounds_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(xyz), n, partial);
  T3D_LAUNCH_CHECK();
  bounds_final_kernel<<<1, 32, 0, st>>>(partial, grid, partial + 6 * grid);
  T3D_LAUNCH_CHECK();
  ctx->launches += 2;
  double* h = reinterpret_cast<double*>(ctx->pinned);
  T3D_CUDA(cudaMemcpyAsync(h, partial + 6 * grid, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
  T3D_CUDA(cudaStreamSynchronize(st));
  for (int c = 0; c < 3; ++c) { min_h[c] = h[c]; max_h[c] = h[3 + c]; }
  return T3D_OK;
}

extern "C" int t3d_voxel_downsample(t3d_ctx* ctx, const void* xyz, int xyz_is_f64,
                                    const uint8_t* rgb, int64_t n, double voxel,
                                    const double* min_bound_h, int sorted, double* out_xyz,
                                    uint8_t* out_rgb, uint32_t* out_rgb_sum, uint32_t* out_count,
                                    int32_t* out_vox_idx, int64_t capacity, int64_t* out_m,
                                    double* out_min_bound_h, t3d_stream stream) {
  T3D_REQUIRE(ctx && out_m, "t3d_voxel_downsample: null argument");
  T3D_REQUIRE(voxel > 0.0, "t3d_voxel_downsample: voxel_size <= 0");  // Open3D raises
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    T3D_CUDA(cudaMemsetAsync(out_m, 0, sizeof(int64_t), st));
    return T3D_OK;
  }
  T3D_REQUIRE(xyz && out_xyz && n > 0, "t3d_voxel_downsample: null xyz/out_xyz");
  double mn[3], mx[3];
  int rc = t3d_bounds(ctx, xyz, xyz_is_f64, n, mn, mx, stream);
  if (rc != T3D_OK) return rc;
  double minb[3];
  for (int c = 0; c < 3; ++c) {
    minb[c] = min_bound_h ? min_bound_h[c] : mn[c] - voxel * 0.5;
    if (out_min_bound_h) out_min_bound_h[c] = minb[c];
    const double ext = (mx[c] + voxel * 0.5) - minb[c];
    if (voxel * 2147483647.0 < ext) {
      t3d_set_error("t3d_voxel_downsample: voxel_size is too small");  // Open3D's message
      return T3D_E_NUMERIC;
    }
    if (ext / voxel >= 2097151.0 || mn[c] < minb[c]) {
      t3d_set_error("t3d_voxel_downsample: grid exceeds 2^21 voxels per axis (or min_bound > data)");
      return T3D_E_NUMERIC;
    }
  }
  // table capacity: load factor <= 0.5 w.r.t. the worst case (every point its own voxel)
  unsigned long long cap = 1024;
  while (cap < 2ull * (unsigned long long)n) cap <<= 1;
  const size_t per_slot = 8 + 3 * 8 + 3 * 4 + 4;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  while (cap > 1024 && cap * per_slot > free_b / 2 + ctx->scratch[0].cap && cap > (unsigned long long)n)
    cap >>= 1;  // degrade to load factor <= 1 before failing
  rc = ctx->scratch[0].reserve(cap * per_slot);
  if (rc != T3D_OK) return rc;
  uint8_t* base = ctx->scratch[0].as<uint8_t>();
  VoxTable tb;
  tb.keys = reinterpret_cast<unsigned long long*>(base);
  tb.sx = reinterpret_cast<double*>(base + cap * 8);
  tb.sy = tb.sx + cap;
  tb.sz = tb.sy + cap;
  tb.sr = reinterpret_cast<unsigned*>(base + cap * 32);
  tb.sg = tb.sr + cap;
  tb.sb = tb.sg + cap;
  tb.cnt = tb.sb + cap;
  tb.mask = cap - 1;
  rc = ctx->scratch[1].reserve(64);
  if (rc != T3D_OK) return rc;
  tb.flags = ctx->scratch[1].as<int>();
  unsigned long long* d_m = reinterpret_cast<unsigned long long*>(tb.flags + 4);
  T3D_CUDA(cudaMemsetAsync(tb.keys, 0xFF, cap * 8, st));
  T3D_CUDA(cudaMemsetAsync(base + cap * 8, 0, cap * (per_slot - 8), st));
  T3D_CUDA(cudaMemsetAsync(tb.flags, 0, 64, st));

  const int grid = ctx->num_sms * 8;
  if (xyz_is_f64)
    voxel_insert_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(xyz), rgb, n,
                                                      minb[0], minb[1], minb[2], voxel, tb);
  else
    voxel_insert_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(xyz), rgb, n,
                                                     minb[0], minb[1], minb[2], voxel, tb);
  T3D_LAUNCH_CHECK();
  ctx->launches++;

  // collect occupied slots
  rc = ctx->scratch[2].reserve((size_t)n * 8);   // keys   (M <= n)
  if (rc != T3D_OK) return rc;
  rc = ctx->scratch[3].reserve((size_t)n * 4);   // slots
  if (rc != T3D_OK) return rc;
  unsigned long long* ckeys = ctx->scratch[2].as<unsigned long long>();
  unsigned* cslots = ctx->scratch[3].as<unsigned>();
  voxel_collect_kernel<<<grid, 256, 0, st>>>(tb, ckeys, cslots, d_m);
  T3D_LAUNCH_CHECK();
  ctx->launches++;
  int* h = reinterpret_cast<int*>(ctx->pinned);
  T3D_CUDA(cudaMemcpyAsync(h, tb.flags, 64, cudaMemcpyDeviceToHost, st));
  T3D_CUDA(cudaStreamSynchronize(st));
  if (h[0]) { t3d_set_error("t3d_voxel_downsample: hash table full"); return T3D_E_CAPACITY; }
  if (h[1]) { t3d_set_error("t3d_voxel_downsample: voxel index out of range"); return T3D_E_NUMERIC; }
  const long long m = (long long)*reinterpret_cast<unsigned long long*>(h + 4);
  if (m > capacity) {
    t3d_set_error("t3d_voxel_downsample: capacity %lld < voxels %lld", (long long)capacity, m);
    return T3D_E_CAPACITY;
  }
  if (sorted && m > 1) {
    rc = ctx->scratch[4].reserve((size_t)m * 8);
    if (rc != T3D_OK) return rc;
    rc = ctx->scratch[5].reserve((size_t)m * 4);
    if (rc != T3D_OK) return rc;
    rc = t3d_radix_sort_u64(ctx, ckeys, cslots, ctx->scratch[4].as<unsigned long long>(),
                            ctx->scratch[5].as<unsigned>(), m, 63, st);
    if (rc != T3D_OK) return rc;
  }
  if (m > 0) {
    voxel_finalize_kernel<<<grid, 256, 0, st>>>(tb, cslots, m, rgb != nullptr, out_xyz, out_rgb,
                                                out_rgb_sum, out_count, out_vox_idx);
    T3D_LAUNCH_CHECK();
    ctx->launches++;
  }
  const int64_t m64 = m;
  T3D_CUDA(cudaMemcpyAsync(out_m, &m64, sizeof(int64_t), cudaMemcpyHostToDevice, st));
  T3D_CUDA(cudaStreamSynchronize(st));
  return T3D_OK;
}

extern "C" int t3d_compact_rows(t3d_ctx* ctx, const void* rows, int64_t n, int32_t row_bytes,
                                const uint8_t* keep_mask, void* out_rows, int64_t* out_n,
                                t3d_stream stream) {
  T3D_REQUIRE(ctx && out_n && row_bytes > 0, "t3d_compact_rows: bad argument");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    T3D_CUDA(cudaMemsetAsync(out_n, 0, sizeof(int64_t), st));
    return T3D_OK;
  }
  T3D_REQUIRE(rows && keep_mask && out_rows, "t3d_compact_rows: null pointer");
  const long long ntiles = (n + CP_TILE - 1) / CP_TILE;
  int rc = ctx->scratch[6].reserve(sizeof(unsigned) * (size_t)(ntiles + 1));
  if (rc != T3D_OK) return rc;
  unsigned* counts = ctx->scratch[6].as<unsigned>();
  compact_count_kernel<<<(unsigned)ntiles, 256, 0, st>>>(keep_mask, n, counts);
  T3D_LAUNCH_CHECK();
  rc = t3d_exclusive_scan_u32(counts, ntiles, st);
  if (rc != T3D_OK) return rc;
  compact_scatter_kernel<<<(unsigned)ntiles, 256, 0, st>>>(
      reinterpret_cast<const uint8_t*>(rows), keep_mask, n, row_bytes, counts, ntiles,
      reinterpret_cast<uint8_t*>(out_rows), reinterpret_cast<long long*>(out_n));
  T3D_LAUNCH_CHECK();
  ctx->launches += 3;
  return T3D_OK;
}
