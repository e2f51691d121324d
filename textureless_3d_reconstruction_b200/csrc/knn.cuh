// knn.cuh — uniform-grid spatial index shared by K3 (statistical outlier
// removal), K7 (normal estimation) and K8 (ICP correspondences).
//
// Build: cell key = floor((p - min)/h) packed into as few bits as the extent
// needs; (key, point index) pairs are radix-sorted (stable, so points inside a
// cell are in ascending index order => deterministic tie-breaks); cell
// [start,end) ranges go into an open-addressing hash keyed by the cell key.
// Point coordinates are copied in sorted order so a cell is one contiguous,
// coalesced run in HBM/L2.
#pragma once
#include "common.cuh"

struct GridDev {
  double minb[3];
  double h, inv_h_unused;
  int bits[3];
  int dims[3];
  long long n;
  const unsigned* sorted_idx;   // sorted position -> original index
  const void* sorted_xyz;       // n*3 (T) in sorted order
  const unsigned long long* hkeys;
  const unsigned* hstart;
  const unsigned* hend;
  unsigned long long hmask;
};

__device__ __forceinline__ unsigned long long grid_pack(const GridDev& g, int cx, int cy,
                                                        int cz) {
  return ((unsigned long long)cx << (g.bits[1] + g.bits[2])) |
         ((unsigned long long)cy << g.bits[2]) | (unsigned long long)cz;
}

__device__ __forceinline__ void grid_cell_of(const GridDev& g, double x, double y, double z,
                                             int& cx, int& cy, int& cz) {
  cx = (int)floor((x - g.minb[0]) / g.h);
  cy = (int)floor((y - g.minb[1]) / g.h);
  cz = (int)floor((z - g.minb[2]) / g.h);
}

// returns false if the cell is empty / outside
__device__ __forceinline__ bool grid_lookup(const GridDev& g, int cx, int cy, int cz,
                                            unsigned& start, unsigned& end) {
  if (cx < 0 || cy < 0 || cz < 0 || cx >= g.dims[0] || cy >= g.dims[1] || cz >= g.dims[2])
    return false;
  const unsigned long long key = grid_pack(g, cx, cy, cz);
  unsigned long long slot = mix64(key) & g.hmask;
  while (true) {
    const unsigned long long k = __ldg(g.hkeys + slot);
    if (k == key) {
      start = __ldg(g.hstart + slot);
      end = __ldg(g.hend + slot);
      return true;
    }
    if (k == T3D_KEY_EMPTY) return false;
    slot = (slot + 1) & g.hmask;
  }
}

// k=1 search within radius (cell size >= radius => 27 cells).  Ties -> lowest
// original index.  Returns sorted position or -1.
__device__ __forceinline__ int t3d_nn_within(const GridDev& g, double qx, double qy, double qz, double r2,
                             double* out_d2) {
  const float* pts = reinterpret_cast<const float*>(g.sorted_xyz);
  int cx, cy, cz;
  grid_cell_of(g, qx, qy, qz, cx, cy, cz);
  double best = r2;
  int bj = -1;
  unsigned bo = 0xFFFFFFFFu;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        unsigned s, e;
        if (!grid_lookup(g, cx + dx, cy + dy, cz + dz, s, e)) continue;
        for (unsigned j = s; j < e; ++j) {
          const double ddx = (double)pts[3ll * j] - qx, ddy = (double)pts[3ll * j + 1] - qy,
                       ddz = (double)pts[3ll * j + 2] - qz;
          const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
          if (d2 < best || (d2 == best && d2 <= r2 && g.sorted_idx[j] < bo)) {
            best = d2;
            bj = (int)j;
            bo = g.sorted_idx[j];
          }
        }
      }
  *out_d2 = best;
  return bj;
}


// Host side: builds the index over `xyz` (n*3, f32 or f64) with cell size h
// (h <= 0: pick from the point density).  Buffers live in ctx->scratch[0..5];
// the GridDev stays valid until the next build on the same ctx.
int t3d_grid_build(t3d_ctx* ctx, const void* xyz, int is_f64, long long n, double h,
                   GridDev* out, cudaStream_t st, int target_k = 0);

int t3d_radix_sort_u64(t3d_ctx* ctx, unsigned long long* keys_a, unsigned* vals_a,
                       unsigned long long* keys_b, unsigned* vals_b, long long n,
                       int key_bits, cudaStream_t st);
int t3d_exclusive_scan_u32(unsigned* data, long long n, cudaStream_t st);
