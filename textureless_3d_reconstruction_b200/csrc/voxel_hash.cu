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

// one warp per component (6 warps): lanes stride over the per-CTA partials, shuffle tree at the end
__global__ void bounds_final_kernel(const double* partial, int nblocks, double* out6) {
  const int c = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (c >= 6) return;
  const bool is_min = c < 3;
  double r = is_min ? 1e300 : -1e300;
  for (int b = l; b < nblocks; b += 32)
    r = is_min ? fmin(r, partial[b * 6 + c]) : fmax(r, partial[b * 6 + c]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, r, d);
    r = is_min ? fmin(r, o) : fmax(r, o);
  }
  if (l == 0) out6[c] = r;
}

// ---------------------------------------------------------------------------
// voxel hash — two passes over the points, both aggregated over RUNS of equal keys in
// consecutive lanes (fused clouds are scan-ordered: a voxel's points sit next to each other).
//
//   pass 1 (voxel_keys_kernel)   key table only: the first lane of every run inserts its key
//       (a voxel that re-appears later is found by the probe).  A warp takes four rows of 32
//       points per iteration and issues the first probe of all of them before resolving any;
//       the CAS winners of a warp iteration take their dense voxel ids from ONE atomicAdd.
//       The table starts at n/2 slots (retry with the worst-case 2n if it fills beyond 0.7).
//   (host reads M — the only size the accumulators need)
//   pass 2 (voxel_accum_kernel)  a 5-step segmented shuffle-down reduces every run of the
//       warp at once; run heads look the voxel id up and issue ONE red.add per accumulator
//       field.  Accumulators are M 64-byte records (one per voxel), not per-slot, so nothing
//       of size O(table) is ever scanned.
//
// Memory: 12 B per table slot (>= 2 N slots) + 8 B per point of key scratch + 64 B
// per output voxel, instead of 48 B per slot.  f64 sums of f32 inputs that lie
// within one voxel are exact, so the means do not depend on the reduction order.
// ---------------------------------------------------------------------------
// one 64-byte, 64-byte-aligned accumulator per output voxel: a group's seven
// reductions land in two adjacent 32 B sectors instead of seven scattered ones
struct __align__(64) VoxAcc {
  double sx, sy, sz;
  unsigned r, g, b, cnt;
  unsigned pad[6];
};

struct VoxTable {
  unsigned long long* keys;   // cap slots, T3D_KEY_EMPTY = free
  unsigned* ids;              // cap slots: dense voxel id of the slot's key
  unsigned long long mask;    // cap - 1
  unsigned long long max_probe;   // insert gives up (flags[0]) after this many slots
  unsigned long long* key_of_id;  // n entries (M used)
  unsigned long long* counter;    // [0] M
  int* flags;                 // [0] table full, [1] index out of packable range
  struct VoxAcc* acc;         // M records
  int sh_x, sh_y;             // key = ix << sh_x | iy << sh_y | iz (ascending key == (x,y,z) order)
};

struct VoxGrid {
  double minx, miny, minz, voxel;
  double lim_x, lim_y, lim_z;  // number of cells per axis (exclusive upper bound of the index)
};

template <typename T>
__device__ __forceinline__ bool voxel_key_of(const T* __restrict__ xyz, long long i,
                                             const VoxGrid& g, const VoxTable& tb, double& px,
                                             double& py, double& pz, unsigned long long& key) {
  px = (double)xyz[i * 3 + 0];
  py = (double)xyz[i * 3 + 1];
  pz = (double)xyz[i * 3 + 2];
  // floor((p - minb) / v): IEEE subtract + true division, no contraction
  const double fx = floor(__ddiv_rn(__dsub_rn(px, g.minx), g.voxel));
  const double fy = floor(__ddiv_rn(__dsub_rn(py, g.miny), g.voxel));
  const double fz = floor(__ddiv_rn(__dsub_rn(pz, g.minz), g.voxel));
  if (!(fx >= 0.0 && fx < g.lim_x && fy >= 0.0 && fy < g.lim_y && fz >= 0.0 && fz < g.lim_z)) {
    tb.flags[1] = 1;
    return false;
  }
  key = ((unsigned long long)(unsigned)(int)fx << tb.sh_x) |
        ((unsigned long long)(unsigned)(int)fy << tb.sh_y) | (unsigned long long)(unsigned)(int)fz;
  return true;
}

template <typename T>
__global__ void __launch_bounds__(256)
    voxel_keys_kernel(const T* __restrict__ xyz, long long n, const __grid_constant__ VoxGrid g,
                      const __grid_constant__ VoxTable tb) {
  // A warp takes KR rows of 32 consecutive points per iteration.  Fused clouds are scan-ordered:
  // lanes that share a voxel are almost always neighbours, so only the first lane of every run of
  // equal keys inserts (a voxel that re-appears later is found by the probe, like one that another
  // warp inserted).  The first probe of all KR rows is issued before any of them is resolved: the
  // kernel is bound by the latency of these random table reads, not by their number.
  constexpr int KR = 4;
  const unsigned lane = lane_id();
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long base = warp * (32 * KR); base < n; base += nwarps * (32 * KR)) {
    unsigned long long key[KR], slot[KR], first[KR];
    bool ins[KR];
#pragma unroll
    for (int r = 0; r < KR; ++r) {
      const long long i = base + r * 32 + lane;
      double px, py, pz;
      key[r] = 0;
      const bool valid = i < n && voxel_key_of(xyz, i, g, tb, px, py, pz, key[r]);
      const unsigned long long prev = __shfl_up_sync(0xffffffffu, key[r], 1);
      const unsigned vmask = __ballot_sync(0xffffffffu, valid);
      ins[r] = valid && !(lane != 0 && ((vmask >> (lane - 1)) & 1u) && prev == key[r]);
      slot[r] = mix64(key[r]) & tb.mask;
      first[r] = T3D_KEY_EMPTY;
    }
#pragma unroll
    for (int r = 0; r < KR; ++r)
      if (ins[r]) first[r] = ld_volatile_u64(reinterpret_cast<const uint64_t*>(tb.keys + slot[r]));
    unsigned created = 0;  // bit r: this lane's CAS created the voxel of row r (at slot[r])
#pragma unroll
    for (int r = 0; r < KR; ++r) {
      if (!ins[r]) continue;
      unsigned long long s = slot[r], k = first[r];
      bool placed = false;
      for (unsigned long long probe = 0; probe < tb.max_probe; ++probe) {
        if (k == T3D_KEY_EMPTY) {
          k = atomicCAS(tb.keys + s, T3D_KEY_EMPTY, key[r]);
          if (k == T3D_KEY_EMPTY) {  // this thread created the voxel
            created |= 1u << r;
            slot[r] = s;
            placed = true;
            break;
          }
        }
        if (k == key[r]) { placed = true; break; }
        s = (s + 1) & tb.mask;
        k = ld_volatile_u64(reinterpret_cast<const uint64_t*>(tb.keys + s));
      }
      if (!placed) tb.flags[0] = 1;
    }
    // dense voxel ids: ONE atomicAdd per warp iteration (a million returning atomics on a single
    // counter would serialise in L2 and cost more than the whole table traffic)
    const unsigned mine = __popc(created);
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= (unsigned)d) incl += t;
    }
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;  // warp-uniform
    unsigned long long id0 = 0;
    if (lane == 31) id0 = atomicAdd(tb.counter, (unsigned long long)total);
    id0 = __shfl_sync(0xffffffffu, id0, 31) + (incl - mine);
#pragma unroll
    for (int r = 0; r < KR; ++r) {
      if (!(created & (1u << r))) continue;
      tb.ids[slot[r]] = (unsigned)id0;
      tb.key_of_id[id0] = key[r];
      ++id0;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    voxel_accum_kernel(const T* __restrict__ xyz, const uint8_t* __restrict__ rgb, long long n,
                       const __grid_constant__ VoxGrid g, const __grid_constant__ VoxTable tb) {
  // Segmented reduction over RUNS of equal keys in consecutive lanes (fused clouds are scan-ordered,
  // so a voxel's points sit next to each other): five shuffle-down steps reduce every run of the
  // warp at once, whatever the number of distinct voxels; the head lane of a run looks the voxel
  // id up and issues one red.add per field.  A voxel that shows up in two separate runs simply gets
  // two updates.  f64 sums of f32 coordinates inside a voxel are exact, so grouping is free.
  constexpr int AR = 2;  // rows of 32 points per warp iteration: two table look-ups in flight per head lane
  const unsigned lane = lane_id();
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long base = warp * (32 * AR); base < n; base += nwarps * (32 * AR)) {
    double px[AR], py[AR], pz[AR];
    unsigned cr[AR], cg[AR], cb[AR], cnt[AR];
    unsigned long long key[AR], slot[AR], first[AR];
    bool emit[AR];
#pragma unroll
    for (int r = 0; r < AR; ++r) {
      const long long i = base + r * 32 + lane;
      px[r] = py[r] = pz[r] = 0.0;
      key[r] = 0;
      const bool valid = i < n && voxel_key_of(xyz, i, g, tb, px[r], py[r], pz[r], key[r]);
      cr[r] = cg[r] = cb[r] = 0;
      if (valid && rgb) { cr[r] = rgb[i * 3 + 0]; cg[r] = rgb[i * 3 + 1]; cb[r] = rgb[i * 3 + 2]; }
      if (!valid) { px[r] = py[r] = pz[r] = 0.0; key[r] = T3D_KEY_EMPTY; }  // its own run, contributes nothing
      const unsigned long long prev = __shfl_up_sync(0xffffffffu, key[r], 1);
      const bool head = lane == 0 || prev != key[r];
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      // last lane of this lane's run: one below the next head above it
      const unsigned above = lane == 31 ? 0u : (heads >> (lane + 1));
      const int tail = above ? (int)lane + __ffs(above) - 1 : 31;
      cnt[r] = 1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double ox = __shfl_down_sync(0xffffffffu, px[r], d);
        const double oy = __shfl_down_sync(0xffffffffu, py[r], d);
        const double oz = __shfl_down_sync(0xffffffffu, pz[r], d);
        const unsigned orr = __shfl_down_sync(0xffffffffu, cr[r], d);
        const unsigned og = __shfl_down_sync(0xffffffffu, cg[r], d);
        const unsigned ob = __shfl_down_sync(0xffffffffu, cb[r], d);
        const unsigned oc = __shfl_down_sync(0xffffffffu, cnt[r], d);
        if ((int)lane + d <= tail) {
          px[r] += ox; py[r] += oy; pz[r] += oz; cr[r] += orr; cg[r] += og; cb[r] += ob; cnt[r] += oc;
        }
      }
      emit[r] = head && valid;
      slot[r] = mix64(key[r]) & tb.mask;
      first[r] = 0;
    }
#pragma unroll
    for (int r = 0; r < AR; ++r)
      if (emit[r]) first[r] = tb.keys[slot[r]];
#pragma unroll
    for (int r = 0; r < AR; ++r) {
      if (!emit[r]) continue;  // pass 1 inserted every key
      unsigned long long sl = slot[r], k = first[r];
      while (k != key[r]) { sl = (sl + 1) & tb.mask; k = tb.keys[sl]; }
      VoxAcc* a = tb.acc + tb.ids[sl];
      atomicAdd(&a->sx, px[r]); atomicAdd(&a->sy, py[r]); atomicAdd(&a->sz, pz[r]);
      if (rgb) { atomicAdd(&a->r, cr[r]); atomicAdd(&a->g, cg[r]); atomicAdd(&a->b, cb[r]); }
      atomicAdd(&a->cnt, cnt[r]);
    }
  }
}

// ---------------------------------------------------------------------------
// Sharded K2: a rank's per-voxel partial sums travel as 56-byte records; the owner
// (mix64(voxel index) mod world) adds the partials of every rank.  f64 sums of f32 coordinates
// inside one voxel are exact, so the merged means equal the single-GPU means bit for bit.
// ---------------------------------------------------------------------------
struct __align__(8) VoxRec {  // wire format, 7 x 8 bytes
  int ix, iy, iz;
  unsigned cnt;
  double sx, sy, sz;
  unsigned r, g, b, pad;
};
static_assert(sizeof(VoxRec) == 56, "VoxRec is the 56-byte wire record");

__device__ __forceinline__ bool rec_key(const VoxRec& q, const VoxGrid& g, const VoxTable& tb,
                                        unsigned long long& key) {
  if (!(q.ix >= 0 && (double)q.ix < g.lim_x && q.iy >= 0 && (double)q.iy < g.lim_y && q.iz >= 0 &&
        (double)q.iz < g.lim_z)) {
    tb.flags[1] = 1;
    return false;
  }
  key = ((unsigned long long)(unsigned)q.ix << tb.sh_x) | ((unsigned long long)(unsigned)q.iy << tb.sh_y) |
        (unsigned long long)(unsigned)q.iz;
  return true;
}

__global__ void __launch_bounds__(256)
    rec_keys_kernel(const VoxRec* __restrict__ recs, long long n, const __grid_constant__ VoxGrid g,
                    const __grid_constant__ VoxTable tb) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    unsigned long long key = 0;
    if (!rec_key(recs[i], g, tb, key)) continue;
    unsigned long long slot = mix64(key) & tb.mask;
    bool placed = false;
    for (unsigned long long probe = 0; probe < tb.max_probe; ++probe) {
      unsigned long long k = ld_volatile_u64(reinterpret_cast<const uint64_t*>(tb.keys + slot));
      if (k == T3D_KEY_EMPTY) {
        k = atomicCAS(tb.keys + slot, T3D_KEY_EMPTY, key);
        if (k == T3D_KEY_EMPTY) {
          const unsigned long long id = atomicAdd(tb.counter, 1ull);
          tb.ids[slot] = (unsigned)id;
          tb.key_of_id[id] = key;
          placed = true;
          break;
        }
      }
      if (k == key) { placed = true; break; }
      slot = (slot + 1) & tb.mask;
    }
    if (!placed) tb.flags[0] = 1;
  }
}

__global__ void __launch_bounds__(256)
    rec_accum_kernel(const VoxRec* __restrict__ recs, long long n, int has_rgb,
                     const __grid_constant__ VoxGrid g, const __grid_constant__ VoxTable tb) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const VoxRec q = recs[i];
    unsigned long long key = 0;
    if (!rec_key(q, g, tb, key)) continue;
    unsigned long long slot = mix64(key) & tb.mask;
    while (tb.keys[slot] != key) slot = (slot + 1) & tb.mask;
    VoxAcc* a = tb.acc + tb.ids[slot];
    atomicAdd(&a->sx, q.sx); atomicAdd(&a->sy, q.sy); atomicAdd(&a->sz, q.sz);
    if (has_rgb) { atomicAdd(&a->r, q.r); atomicAdd(&a->g, q.g); atomicAdd(&a->b, q.b); }
    atomicAdd(&a->cnt, q.cnt);
  }
}

__device__ __forceinline__ int rec_owner(int ix, int iy, int iz, int world) {
  return (int)(mix64(pack_key(ix, iy, iz)) % (unsigned long long)world);
}

// pass 0: per-owner counts; pass 1: scatter (offsets = exclusive scan of the counts, done by
// thread 0 of the first CTA of pass 1's predecessor kernel below)
__global__ void __launch_bounds__(256)
    partials_count_kernel(const __grid_constant__ VoxTable tb, long long m, int world, int* counts) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long k = tb.key_of_id[i];
    const int ix = (int)(k >> tb.sh_x);
    const int iy = (int)((k >> tb.sh_y) & ((1ull << (tb.sh_x - tb.sh_y)) - 1ull));
    const int iz = (int)(k & ((1ull << tb.sh_y) - 1ull));
    atomicAdd(counts + rec_owner(ix, iy, iz, world), 1);
  }
}

__global__ void partials_offsets_kernel(const int* counts, int world, int* offs, int* fill) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int run = 0;
    for (int d = 0; d < world; ++d) { offs[d] = run; run += counts[d]; fill[d] = 0; }
  }
}

__global__ void __launch_bounds__(256)
    partials_scatter_kernel(const __grid_constant__ VoxTable tb, long long m, int world, int has_rgb,
                            const int* offs, int* fill, VoxRec* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long k = tb.key_of_id[i];
    VoxRec q;
    q.ix = (int)(k >> tb.sh_x);
    q.iy = (int)((k >> tb.sh_y) & ((1ull << (tb.sh_x - tb.sh_y)) - 1ull));
    q.iz = (int)(k & ((1ull << tb.sh_y) - 1ull));
    const VoxAcc a = tb.acc[i];
    q.cnt = a.cnt;
    q.sx = a.sx; q.sy = a.sy; q.sz = a.sz;
    q.r = has_rgb ? a.r : 0u; q.g = has_rgb ? a.g : 0u; q.b = has_rgb ? a.b : 0u;
    q.pad = 0u;
    const int o = rec_owner(q.ix, q.iy, q.iz, world);
    out[offs[o] + atomicAdd(fill + o, 1)] = q;
  }
}

// order[i] = voxel id written at output row i (identity when unsorted)
__global__ void voxel_finalize_kernel(const __grid_constant__ VoxTable tb,
                                      const unsigned* __restrict__ order, long long m,
                                      int has_rgb, double* out_xyz, uint8_t* out_rgb,
                                      unsigned* out_rgb_sum, unsigned* out_count,
                                      int* out_idx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned s = order ? order[i] : (unsigned)i;
    const VoxAcc a = tb.acc[s];
    const unsigned c = a.cnt;
    const double dn = (double)c;
    out_xyz[i * 3 + 0] = __ddiv_rn(a.sx, dn);
    out_xyz[i * 3 + 1] = __ddiv_rn(a.sy, dn);
    out_xyz[i * 3 + 2] = __ddiv_rn(a.sz, dn);
    const unsigned r = has_rgb ? a.r : 0u, g = has_rgb ? a.g : 0u, b = has_rgb ? a.b : 0u;
    if (out_rgb && has_rgb) {
      // (mean(c/255) * 255).astype(uint8) — truncation, d2r:417-418
      out_rgb[i * 3 + 0] = (uint8_t)(int)__dmul_rn(__ddiv_rn(__ddiv_rn((double)r, 255.0), dn), 255.0);
      out_rgb[i * 3 + 1] = (uint8_t)(int)__dmul_rn(__ddiv_rn(__ddiv_rn((double)g, 255.0), dn), 255.0);
      out_rgb[i * 3 + 2] = (uint8_t)(int)__dmul_rn(__ddiv_rn(__ddiv_rn((double)b, 255.0), dn), 255.0);
    }
    if (out_rgb_sum) { out_rgb_sum[i * 3] = r; out_rgb_sum[i * 3 + 1] = g; out_rgb_sum[i * 3 + 2] = b; }
    if (out_count) out_count[i] = c;
    if (out_idx) {
      const unsigned long long k = tb.key_of_id[s];
      out_idx[i * 3 + 0] = (int)(k >> tb.sh_x);
      out_idx[i * 3 + 1] = (int)((k >> tb.sh_y) & ((1ull << (tb.sh_x - tb.sh_y)) - 1ull));
      out_idx[i * 3 + 2] = (int)(k & ((1ull << tb.sh_y) - 1ull));
    }
  }
}

__global__ void iota_kernel(unsigned* v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    v[i] = (unsigned)i;
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

// exclusive scan over hist[256*ntiles] in digit-major order (single CTA).  Every thread owns 16 consecutive
// entries per round (four 128-bit loads; a warp covers 2 KB contiguously), scans them in registers, and one
// block scan of the 1024 thread totals per round links them: 16 Ki entries per round instead of 1 Ki —
// a sort pass of a million keys needs 4 rounds instead of 60 (49 -> ~6 us per pass on B200).
__global__ void __launch_bounds__(1024) rs_scan_kernel(unsigned* hist, long long total) {
  constexpr int IPT = 16;
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool aligned = (reinterpret_cast<uintptr_t>(hist) & 15) == 0;
  for (long long base = 0; base < total; base += 1024 * IPT) {
    const long long i0 = base + (long long)threadIdx.x * IPT;
    unsigned v[IPT];
    if (aligned && i0 + IPT <= total) {
#pragma unroll
      for (int q = 0; q < IPT / 4; ++q) {
        const uint4 t = reinterpret_cast<const uint4*>(hist + i0)[q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < IPT; ++k) v[k] = i0 + k < total ? hist[i0 + k] : 0u;
    }
    unsigned sum = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {  // exclusive scan inside the thread
      const unsigned t = v[k];
      v[k] = sum;
      sum += t;
    }
    unsigned inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= (unsigned)d) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const unsigned w = s_w[lane];
      unsigned winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= (unsigned)d) winc += t;
      }
      s_w[lane] = winc - w;
    }
    __syncthreads();
    const unsigned carry = s_carry;
    const unsigned off = carry + s_w[warp] + inc - sum;
    if (aligned && i0 + IPT <= total) {
#pragma unroll
      for (int q = 0; q < IPT / 4; ++q)
        reinterpret_cast<uint4*>(hist + i0)[q] = make_uint4(off + v[4 * q], off + v[4 * q + 1], off + v[4 * q + 2], off + v[4 * q + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < IPT; ++k)
        if (i0 + k < total) hist[i0 + k] = off + v[k];
    }
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
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(n > 0 && xyz, "t3d_bounds: empty cloud");
  cudaStream_t st = as_stream(stream);
  const int grid = ctx->num_sms * 16;
  int rc = ctx->scratch[6].reserve(sizeof(double) * 6 * (grid + 1));
  if (rc != T3D_OK) return rc;
  double* partial = ctx->scratch[6].as<double>();
  if (xyz_is_f64)
    bounds_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(xyz), n, partial);
  else
    bounds_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(xyz), n, partial);
  T3D_LAUNCH_CHECK();
  bounds_final_kernel<<<1, 192, 0, st>>>(partial, grid, partial + 6 * grid);
  T3D_LAUNCH_CHECK();
  ctx->launches += 2;
  double* h = reinterpret_cast<double*>(ctx->pinned);
  T3D_CUDA(cudaMemcpyAsync(h, partial + 6 * grid, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
  T3D_CUDA(cudaStreamSynchronize(st));
  for (int c = 0; c < 3; ++c) { min_h[c] = h[c]; max_h[c] = h[3 + c]; }
  return T3D_OK;
}

// One K2 run.  Source = points (xyz [+ rgb]) or, when `recs` is given, partial records of other
// ranks (then min/max bounds must be supplied).  Sink = the usual per-voxel outputs or, when
// `part_out` is given, partial records grouped by owner (part_world owners, counts in
// part_counts, device int32[part_world]).  max_bound_h (nullable): data maximum to size the grid
// with (multi-GPU: the global maximum, so that every rank packs keys identically).
static int voxel_run(t3d_ctx* ctx, const void* xyz, int xyz_is_f64, const uint8_t* rgb,
                     const VoxRec* recs, int recs_have_rgb, int64_t n, double voxel,
                     const double* min_bound_h, const double* max_bound_h, int sorted, double* out_xyz,
                     uint8_t* out_rgb, uint32_t* out_rgb_sum, uint32_t* out_count,
                     int32_t* out_vox_idx, VoxRec* part_out, int part_world, int32_t* part_counts,
                     int64_t capacity, int64_t* out_m, double* out_min_bound_h, t3d_stream stream) {
  T3D_REQUIRE(ctx && out_m, "t3d_voxel_downsample: null argument");
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(voxel > 0.0, "t3d_voxel_downsample: voxel_size <= 0");  // Open3D raises
  cudaStream_t st = as_stream(stream);
  if (part_out) T3D_CUDA(cudaMemsetAsync(part_counts, 0, sizeof(int32_t) * (size_t)part_world, st));
  if (n == 0) {
    T3D_CUDA(cudaMemsetAsync(out_m, 0, sizeof(int64_t), st));
    return T3D_OK;
  }
  T3D_REQUIRE((xyz || recs) && (out_xyz || part_out) && n > 0, "t3d_voxel_downsample: null xyz/out_xyz");
  T3D_REQUIRE(n < (1ll << 32) - 64, "t3d_voxel_downsample: more than 2^32 points in one call");
  T3D_REQUIRE(!recs || (min_bound_h && max_bound_h), "t3d_voxel_merge_partials: bounds are required");
  T3DTrace tr("voxel_downsample");
  double mn[3], mx[3];
  int rc = T3D_OK;
  if (min_bound_h && max_bound_h) {
    for (int c = 0; c < 3; ++c) { mn[c] = min_bound_h[c]; mx[c] = max_bound_h[c]; }
  } else {
    rc = t3d_bounds(ctx, xyz, xyz_is_f64, n, mn, mx, stream);
    if (rc != T3D_OK) return rc;
    if (max_bound_h)
      for (int c = 0; c < 3; ++c) mx[c] = max_bound_h[c] > mx[c] ? max_bound_h[c] : mx[c];
  }
  tr.mark("bounds+sync");
  double minb[3];
  int bits[3];
  VoxGrid g;
  for (int c = 0; c < 3; ++c) {
    minb[c] = min_bound_h ? min_bound_h[c] : mn[c] - voxel * 0.5;
    if (out_min_bound_h) out_min_bound_h[c] = minb[c];
    const double ext = (mx[c] + voxel * 0.5) - minb[c];
    if (voxel * 2147483647.0 < ext) {
      t3d_set_error("t3d_voxel_downsample: voxel_size is too small");  // Open3D's message
      return T3D_E_NUMERIC;
    }
    if (ext / voxel >= 2097151.0 || (!(min_bound_h && max_bound_h) && mn[c] < minb[c])) {
      t3d_set_error("t3d_voxel_downsample: grid exceeds 2^21 voxels per axis (or min_bound > data)");
      return T3D_E_NUMERIC;
    }
    // cells per axis (+2 guards the rounding of the device-side division)
    const double cells = floor(ext / voxel) + 2.0;
    bits[c] = 1;
    while ((double)(1ull << bits[c]) < cells) ++bits[c];
  }
  g.minx = minb[0]; g.miny = minb[1]; g.minz = minb[2]; g.voxel = voxel;
  g.lim_x = (double)(1ull << bits[0]); g.lim_y = (double)(1ull << bits[1]); g.lim_z = (double)(1ull << bits[2]);
  const int key_bits = bits[0] + bits[1] + bits[2];

  // key table.  Worst case (every point its own voxel) needs >= 2 n slots for a load factor <= 0.5,
  // but depth-map clouds put ~10 points into a voxel: the first attempt uses n / 2 slots — 4x less
  // memory to clear and, for clouds of tens of millions of points, a table that stays in the
  // 126 MB L2 instead of costing one DRAM access per probe.  If that table fills beyond 0.7 or an
  // insert gives up after 64 slots, pass 1 is simply repeated with the worst-case table.
  unsigned long long cap_full = 1024;
  while (cap_full < 2ull * (unsigned long long)n) cap_full <<= 1;
  if (cap_full * 12 > ctx->scratch[0].cap && cap_full * 12 > (8ull << 30)) {  // only probe the driver for big tables
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    while (cap_full > 1024 && cap_full * 12 > free_b / 2 + ctx->scratch[0].cap &&
           cap_full / 2 > (unsigned long long)n + n / 4)
      cap_full >>= 1;  // degrade to load factor <= 0.8 before failing
  }
  unsigned long long cap_try = 1024;
  while (cap_try < (unsigned long long)n / 2) cap_try <<= 1;
  if (recs || cap_try > cap_full) cap_try = cap_full;  // partial records: one per voxel and rank already
  if ((rc = ctx->scratch[1].reserve(64)) != T3D_OK) return rc;
  if ((rc = ctx->scratch[2].reserve((size_t)n * 8)) != T3D_OK) return rc;
  VoxTable tb;
  memset(&tb, 0, sizeof(tb));
  const long long warps = (n + 31) / 32;
  const long long want = (warps + 7) / 8;  // 256-thread CTAs
  const int grid = (int)(want < (long long)ctx->num_sms * 16 ? want : (long long)ctx->num_sms * 16);
  int* h = reinterpret_cast<int*>(ctx->pinned);
  long long m = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const unsigned long long cap = attempt == 0 ? cap_try : cap_full;
    if ((rc = ctx->scratch[0].reserve(cap * 12)) != T3D_OK) return rc;
    tb.keys = ctx->scratch[0].as<unsigned long long>();
    tb.ids = reinterpret_cast<unsigned*>(tb.keys + cap);
    tb.mask = cap - 1;
    tb.max_probe = cap < cap_full ? 64 : cap;
    tb.key_of_id = ctx->scratch[2].as<unsigned long long>();
    tb.flags = ctx->scratch[1].as<int>();
    tb.counter = reinterpret_cast<unsigned long long*>(tb.flags + 4);
    tb.sh_y = bits[2];
    tb.sh_x = bits[1] + bits[2];
    tr.mark("reserve");
    T3D_CUDA(cudaMemsetAsync(tb.keys, 0xFF, cap * 8, st));
    T3D_CUDA(cudaMemsetAsync(tb.flags, 0, 64, st));
    if (recs)
      rec_keys_kernel<<<grid, 256, 0, st>>>(recs, n, g, tb);
    else if (xyz_is_f64)
      voxel_keys_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(xyz), n, g, tb);
    else
      voxel_keys_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(xyz), n, g, tb);
    T3D_LAUNCH_CHECK();
    ctx->launches++;
    T3D_CUDA(cudaMemcpyAsync(h, tb.flags, 64, cudaMemcpyDeviceToHost, st));
    T3D_CUDA(cudaStreamSynchronize(st));
    tr.mark("pass1+sync");
    if (h[1]) { t3d_set_error("t3d_voxel_downsample: voxel index out of range"); return T3D_E_NUMERIC; }
    m = (long long)*reinterpret_cast<unsigned long long*>(h + 4);
    if (cap < cap_full && (h[0] || (double)m > 0.7 * (double)cap)) continue;  // too optimistic: worst-case table
    if (h[0]) { t3d_set_error("t3d_voxel_downsample: hash table full"); return T3D_E_CAPACITY; }
    break;
  }
  if (m > capacity) {
    t3d_set_error("t3d_voxel_downsample: capacity %lld < voxels %lld", (long long)capacity, m);
    return T3D_E_CAPACITY;
  }

  // accumulators: m 64-byte records (cudaMalloc returns >= 256 B aligned memory)
  if ((rc = ctx->scratch[8].reserve((size_t)m * sizeof(VoxAcc) + 64)) != T3D_OK) return rc;
  tb.acc = ctx->scratch[8].as<VoxAcc>();
  T3D_CUDA(cudaMemsetAsync(tb.acc, 0, (size_t)m * sizeof(VoxAcc), st));
  const int has_rgb = recs ? recs_have_rgb : (rgb != nullptr);
  if (recs)
    rec_accum_kernel<<<grid, 256, 0, st>>>(recs, n, has_rgb, g, tb);
  else if (xyz_is_f64)
    voxel_accum_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(xyz), rgb, n, g, tb);
  else
    voxel_accum_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(xyz), rgb, n, g, tb);
  T3D_LAUNCH_CHECK();
  ctx->launches++;
  if (part_out) {  // sink = partial records grouped by owner
    if ((rc = ctx->scratch[3].reserve(sizeof(int) * 2 * 64)) != T3D_OK) return rc;
    int* offs = ctx->scratch[3].as<int>();
    int* fill = offs + 64;
    const int pgrid = ctx->num_sms * 8;
    if (m > 0) {
      partials_count_kernel<<<pgrid, 256, 0, st>>>(tb, m, part_world, part_counts);
      T3D_LAUNCH_CHECK();
      partials_offsets_kernel<<<1, 32, 0, st>>>(part_counts, part_world, offs, fill);
      T3D_LAUNCH_CHECK();
      partials_scatter_kernel<<<pgrid, 256, 0, st>>>(tb, m, part_world, has_rgb, offs, fill, part_out);
      T3D_LAUNCH_CHECK();
      ctx->launches += 3;
    }
    T3D_CUDA(cudaMemcpyAsync(out_m, tb.counter, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    return T3D_OK;
  }

  const unsigned* order = nullptr;
  const int fgrid = ctx->num_sms * 8;
  if (sorted && m > 1) {
    if ((rc = ctx->scratch[3].reserve((size_t)m * 4)) != T3D_OK) return rc;
    if ((rc = ctx->scratch[4].reserve((size_t)m * 8)) != T3D_OK) return rc;
    if ((rc = ctx->scratch[5].reserve((size_t)m * 4)) != T3D_OK) return rc;
    if ((rc = ctx->scratch[9].reserve((size_t)m * 8)) != T3D_OK) return rc;
    unsigned* ids = ctx->scratch[3].as<unsigned>();
    iota_kernel<<<fgrid, 256, 0, st>>>(ids, m);
    T3D_LAUNCH_CHECK();
    ctx->launches++;
    // sort a copy of the keys (key_of_id must stay addressable by id for out_idx)
    unsigned long long* skeys = ctx->scratch[9].as<unsigned long long>();
    T3D_CUDA(cudaMemcpyAsync(skeys, tb.key_of_id, (size_t)m * 8, cudaMemcpyDeviceToDevice, st));
    rc = t3d_radix_sort_u64(ctx, skeys, ids, ctx->scratch[4].as<unsigned long long>(),
                            ctx->scratch[5].as<unsigned>(), m, key_bits, st);
    if (rc != T3D_OK) return rc;
    order = ids;
  }
  if (m > 0) {
    voxel_finalize_kernel<<<fgrid, 256, 0, st>>>(tb, order, m, has_rgb, out_xyz, out_rgb,
                                                 out_rgb_sum, out_count, out_vox_idx);
    T3D_LAUNCH_CHECK();
    ctx->launches++;
  }
  T3D_CUDA(cudaMemcpyAsync(out_m, tb.counter, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  if (tr.on) { cudaStreamSynchronize(st); tr.mark("pass2+finalize"); }
  return T3D_OK;
}

extern "C" int t3d_voxel_downsample(t3d_ctx* ctx, const void* xyz, int xyz_is_f64,
                                    const uint8_t* rgb, int64_t n, double voxel,
                                    const double* min_bound_h, int sorted, double* out_xyz,
                                    uint8_t* out_rgb, uint32_t* out_rgb_sum, uint32_t* out_count,
                                    int32_t* out_vox_idx, int64_t capacity, int64_t* out_m,
                                    double* out_min_bound_h, t3d_stream stream) {
  return voxel_run(ctx, xyz, xyz_is_f64, rgb, nullptr, 0, n, voxel, min_bound_h, nullptr, sorted, out_xyz, out_rgb,
                   out_rgb_sum, out_count, out_vox_idx, nullptr, 0, nullptr, capacity, out_m, out_min_bound_h,
                   stream);
}

extern "C" int t3d_voxel_partials(t3d_ctx* ctx, const void* xyz, int xyz_is_f64, const uint8_t* rgb,
                                  int64_t n, double voxel, const double* min_bound_h,
                                  const double* max_bound_h, int world, void* out_records,
                                  int64_t capacity, int32_t* out_counts, int64_t* out_m,
                                  t3d_stream stream) {
  T3D_REQUIRE(min_bound_h && max_bound_h && out_counts && world >= 1 && world <= 64 && (capacity == 0 || out_records),
              "t3d_voxel_partials: bad argument");
  return voxel_run(ctx, xyz, xyz_is_f64, rgb, nullptr, 0, n, voxel, min_bound_h, max_bound_h, 0, nullptr, nullptr,
                   nullptr, nullptr, nullptr, reinterpret_cast<VoxRec*>(out_records), world, out_counts, capacity,
                   out_m, nullptr, stream);
}

extern "C" int t3d_voxel_merge_partials(t3d_ctx* ctx, const void* records, int64_t n_records, int has_rgb,
                                        double voxel, const double* min_bound_h,
                                        const double* max_bound_h, int sorted, double* out_xyz,
                                        uint8_t* out_rgb, uint32_t* out_rgb_sum, uint32_t* out_count,
                                        int32_t* out_vox_idx, int64_t capacity, int64_t* out_m,
                                        t3d_stream stream) {
  T3D_REQUIRE(min_bound_h && max_bound_h && (n_records == 0 || records), "t3d_voxel_merge_partials: bad argument");
  return voxel_run(ctx, nullptr, 0, nullptr, reinterpret_cast<const VoxRec*>(records), has_rgb, n_records, voxel,
                   min_bound_h, max_bound_h, sorted, out_xyz, out_rgb, out_rgb_sum, out_count, out_vox_idx, nullptr, 0,
                   nullptr, capacity, out_m, nullptr, stream);
}

extern "C" int t3d_compact_rows(t3d_ctx* ctx, const void* rows, int64_t n, int32_t row_bytes,
                                const uint8_t* keep_mask, void* out_rows, int64_t* out_n,
                                t3d_stream stream) {
  T3D_REQUIRE(ctx && out_n && row_bytes > 0, "t3d_compact_rows: bad argument");
  T3D_ON_DEVICE(ctx->device);
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
