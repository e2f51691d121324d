// backproject.cu — K1: depth -> coloured world-frame points, one HBM pass.
//
// Replaces (reference file:line, relative to the upstream repo):
//   DenseReconstructor.depth_to_pointcloud        depth_to_reconstruction.py:328-384
//   DensePointCloudGenerator.depth_to_pointcloud  depth_enhanced_reconstruction.py:554-613
//   PointCloudGenerator.generate                  depth_processor.py:371-422
//
// Design (B200): one CTA owns a tile of 2048 sampled pixels (256 threads x 8,
// striped so every warp-level load is a coalesced 128 B line).  The boolean
// mask is compacted *in order* (NumPy boolean-mask order) with warp ballots,
// a 64-entry block scan and a single-pass *flat* look-back across tiles (every
// thread of the CTA loads predecessor aggregates in parallel, one memory round
// trip instead of a 32-tile-per-hop chain; checkpoint tiles every 1024 tiles
// publish inclusive prefixes so the work stays O(T*1024)), so depth/colour are
// read exactly once and nothing but the packed outputs is written.  Points and colours are staged in shared memory at an offset that
// is congruent (mod 16 B) to their final global address, then streamed out
// with 128-bit stores.  The arithmetic is f64 like NumPy's (x_factor tables
// are built with IEEE division, the 3x3 pose product uses the k-ordered FMA
// chain of a dgemm micro-kernel) and rounded to f32 once.
#include "common.cuh"

namespace {

constexpr int BP_THREADS = 256;
constexpr int BP_WARPS = BP_THREADS / 32;
constexpr int BP_CHK = 1024;  // checkpoint spacing of the flat look-back

constexpr int BP_MAX_BATCH = 32;  // frames per launch (per-frame data rides in the kernel params)

struct BPFrame {
  const void* depth;
  const uint8_t* bgr;
  const uint8_t* conf;
  double Rt[9];  // R^T
  double c[3];   // R^T t
};

struct BPParams {
  BPFrame f[BP_MAX_BATCH];
  int n_frames, tiles_per_frame;
  long long* out_offsets;       // batch: [n_frames + 1] running point offsets (device), else null
  const long long* base_ptr;    // batch chunk > 0: offset where this launch starts writing, else null
  const double* xf;
  const double* yf;
  int W, s, Ws;
  int P;  // sampled pixels = Hs*Ws
  int has_pose, has_color;
  int wide_bgr;  // colour frames are 4-byte aligned and a multiple of 4 bytes long: 32-bit colour loads are safe
  double scale, min_d, max_d;
  float scale32, min32, max32;
  float* out_xyz;
  void* out_rgb;
  long long* out_n;
  unsigned long long* tile_state;  // per tile: FLAG_AGG | count
  unsigned long long* inc_state;   // per checkpoint (tile % BP_CHK == 0): FLAG_INC | inclusive prefix
  unsigned* ticket;
  int num_tiles;
};

constexpr unsigned long long FLAG_AGG = 1ull << 62;
constexpr unsigned long long FLAG_INC = 2ull << 62;
constexpr unsigned long long VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ void st_state(unsigned long long* p,
                                         unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Copy `total` 4-byte words staged in shared memory to global memory.  The
// staging buffer was filled starting at word `shift = g0 % 4`, so that every
// 16-byte global line maps to a 16-byte aligned shared-memory chunk.
__device__ __forceinline__ void flush_words(const uint32_t* sbuf, uint32_t* gdst,
                                            long long g0, int total) {
  const int shift = (int)(g0 & 3);
  int head = (4 - shift) & 3;
  if (head > total) head = total;
  uint32_t* g = gdst + g0;
  const uint32_t* s = sbuf + shift;
  if ((int)threadIdx.x < head) g[threadIdx.x] = s[threadIdx.x];
  const int nv = (total - head) >> 2;
  uint4* g4 = reinterpret_cast<uint4*>(g + head);
  const uint4* s4 = reinterpret_cast<const uint4*>(s + head);
  for (int k = threadIdx.x; k < nv; k += BP_THREADS) st_stream_u4(g4 + k, s4[k]);
  const int done = head + (nv << 2);
  const int tail = total - done;
  if ((int)threadIdx.x < tail) g[done + threadIdx.x] = s[done + threadIdx.x];
}

// Same for bytes (uint8 colours): shift = g0 % 16.
__device__ __forceinline__ void flush_bytes(const uint8_t* sbuf, uint8_t* gdst,
                                            long long g0, int total) {
  const int shift = (int)(g0 & 15);
  int head = (16 - shift) & 15;
  if (head > total) head = total;
  uint8_t* g = gdst + g0;
  const uint8_t* s = sbuf + shift;
  if ((int)threadIdx.x < head) g[threadIdx.x] = s[threadIdx.x];
  const int nv = (total - head) >> 4;
  uint4* g4 = reinterpret_cast<uint4*>(g + head);
  const uint4* s4 = reinterpret_cast<const uint4*>(s + head);
  for (int k = threadIdx.x; k < nv; k += BP_THREADS) st_stream_u4(g4 + k, s4[k]);
  const int done = head + (nv << 4);
  const int tail = total - done;
  if ((int)threadIdx.x < tail) g[done + threadIdx.x] = s[done + threadIdx.x];
}

// DEPTH_MODE: 0 = f32 depth, f32 mask arithmetic (python-float scale)
//             1 = f32 depth, f64 mask arithmetic (np.float64 scale)
//             2 = f64 depth
template <int DEPTH_MODE, bool RGB_F32, int BP_PPT, int MINB>
__global__ void __launch_bounds__(BP_THREADS, MINB)
    backproject_kernel(const __grid_constant__ BPParams p) {
  constexpr int BP_TILE = BP_THREADS * BP_PPT;  // sampled pixels per CTA
  __shared__ __align__(16) uint32_t s_xyz[BP_TILE * 3 + 4];
  __shared__ __align__(16) uint8_t s_rgb[BP_TILE * 3 + 16];
  __shared__ int s_wcount[BP_PPT * BP_WARPS];
  __shared__ int s_woff[BP_PPT * BP_WARPS];
  __shared__ int s_tile, s_total;
  __shared__ long long s_base;
  __shared__ long long s_part[BP_WARPS];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid == 0) s_tile = (int)atomicAdd(p.ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int fi = tile / p.tiles_per_frame;            // frame of this tile (uniform per CTA)
  const int ltile = tile - fi * p.tiles_per_frame;    // tile inside the frame
  const BPFrame& fr = p.f[fi];
  const int p0 = ltile * BP_TILE + tid;

  // ---- pass A: issue ALL of this thread's loads first (depth + colour of
  // BP_PPT pixels in flight at once), then evaluate the masks and count -------
  double z[BP_PPT];
  unsigned bal[BP_PPT];
  uint32_t crgb[BP_PPT];  // packed r | g<<8 | b<<16
  float d32[BP_PPT];
  double d64[DEPTH_MODE == 2 ? BP_PPT : 1];
  uint8_t cf[BP_PPT];
  int u = 0, v = 0;
  if (p0 < p.P) {
    v = p0 / p.Ws;
    u = p0 - v * p.Ws;
  }
  int uu = u, vv = v;
#pragma unroll
  for (int j = 0; j < BP_PPT; ++j) {
    const int pj = p0 + j * BP_THREADS;
    crgb[j] = 0;
    d32[j] = 0.0f;
    cf[j] = 1;
    if (DEPTH_MODE == 2) d64[j] = 0.0;
    if (pj < p.P) {
      const long long sp = (long long)(vv * p.s) * p.W + (long long)uu * p.s;
      if (DEPTH_MODE == 2) d64[j] = __ldg(reinterpret_cast<const double*>(fr.depth) + sp);
      else d32[j] = __ldg(reinterpret_cast<const float*>(fr.depth) + sp);
      if (fr.conf != nullptr) cf[j] = __ldg(fr.conf + sp);
      if (p.has_color) {  // colour rides with the depth: pass B then has no DRAM loads
        if (p.wide_bgr) {
          // the aligned 32-bit word that holds the pixel's first byte (+ the next word for the half of the pixels
          // that reach into it) instead of three byte loads: a third of the L1 wavefronts per pixel
          const unsigned long long o = (unsigned long long)sp * 3ull;
          const unsigned* wp = reinterpret_cast<const unsigned*>(fr.bgr + (o & ~3ull));
          const unsigned w0 = __ldg(wp);
          unsigned w1 = 0u;
          if ((o & 3ull) >= 2ull) w1 = __ldg(wp + 1);
          crgb[j] = __byte_perm(__funnelshift_r(w0, w1, (unsigned)(o & 3ull) * 8u), 0u, 0x4012);  // b g r -> r | g<<8 | b<<16
        } else {
          const uint8_t* c = fr.bgr + sp * 3;
          crgb[j] = (uint32_t)__ldg(c + 2) | ((uint32_t)__ldg(c + 1) << 8) | ((uint32_t)__ldg(c) << 16);
        }
      }
      // advance (u,v) by 256 sampled pixels for the next j
      uu += BP_THREADS;
      while (uu >= p.Ws) {
        uu -= p.Ws;
        ++vv;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < BP_PPT; ++j) {
    const int pj = p0 + j * BP_THREADS;
    bool valid = false;
    double zj = 0.0;
    if (pj < p.P) {
      if (DEPTH_MODE == 0) {
        const float ds = __fmul_rn(d32[j], p.scale32);
        valid = (ds > p.min32) && (ds < p.max32) && isfinite(ds);
        zj = (double)ds;
      } else {
        const double ds = __dmul_rn(DEPTH_MODE == 2 ? d64[j] : (double)d32[j], p.scale);
        valid = (ds > p.min_d) && (ds < p.max_d) && isfinite(ds);
        zj = ds;
      }
      valid = valid && cf[j] != 0;
    }
    z[j] = zj;
    const unsigned b = __ballot_sync(0xffffffffu, valid);
    bal[j] = valid ? b : 0u;  // 0 marks "this thread's pixel is invalid"
    if (lane == 0) s_wcount[j * BP_WARPS + warp] = __popc(b);
  }
  __syncthreads();

  // ---- block scan of the BP_PPT*8 (j, warp) counts; publish this tile's aggregate --
  if (warp == 0) {
    constexpr int PER = BP_PPT * BP_WARPS / 32;  // entries per lane (1 or 2)
    int cnt[PER];
    int sum = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) { cnt[q] = s_wcount[PER * lane + q]; sum += cnt[q]; }
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    int run = inc - sum;
#pragma unroll
    for (int q = 0; q < PER; ++q) { s_woff[PER * lane + q] = run; run += cnt[q]; }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 0) {
      st_state(p.tile_state + tile, FLAG_AGG | (unsigned long long)total);
      s_total = total;
    }
  }
  // ---- flat look-back: base = INC[c] + sum AGG(c, tile), c = last checkpoint < tile.
  // Tickets are handed out in order, so every lower tile is already running and
  // publishes its aggregate before it waits on anything: no deadlock.
  long long part = 0;
  if (tile > 0) {
    const int c = ((tile - 1) / BP_CHK) * BP_CHK;
    for (int j = c + tid; j < tile; j += BP_THREADS) {
      const unsigned long long* src = (j == c) ? (p.inc_state + c / BP_CHK) : (p.tile_state + j);
      unsigned long long st;
      do {
        st = ld_volatile_u64(reinterpret_cast<const uint64_t*>(src));
      } while ((st >> 62) == 0);
      part += (long long)(st & VAL_MASK);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  if (tid == 0) {
    long long base = 0;
#pragma unroll
    for (int w = 0; w < BP_WARPS; ++w) base += s_part[w];
    const int total = s_total;
    if (tile % BP_CHK == 0)
      st_state(p.inc_state + tile / BP_CHK, FLAG_INC | (unsigned long long)(base + total));
    if (p.base_ptr) base += *p.base_ptr;  // later chunk of a batch: continue after the previous launch
    s_base = base;
    if (p.out_offsets) {
      if (ltile == 0) p.out_offsets[fi] = base;
      if (tile == p.num_tiles - 1) p.out_offsets[p.n_frames] = base + total;
    } else if (tile == p.num_tiles - 1) {
      *p.out_n = base + total;
    }
  }
  __syncthreads();
  const long long base = s_base;
  const int total = s_total;
  if (total == 0) return;

  // ---- pass B: back-project valid pixels into the staging buffers ---------
  const int shift_w = (int)((base * 3) & 3);
  const int shift_b = (int)((base * 3) & 15);
  uu = u;
  vv = v;
#pragma unroll
  for (int j = 0; j < BP_PPT; ++j) {
    const int pj = p0 + j * BP_THREADS;
    if (pj < p.P) {
      if (bal[j] != 0u) {
        const int rank =
            s_woff[j * BP_WARPS + warp] + __popc(bal[j] & lanemask_lt());
        const int su = uu * p.s, sv = vv * p.s;
        const double zz = z[j];
        const double x = __dmul_rn(__ldg(p.xf + su), zz);
        const double y = __dmul_rn(__ldg(p.yf + sv), zz);
        double wx = x, wy = y, wz = zz;
        if (p.has_pose) {
          wx = __dsub_rn(
              __fma_rn(fr.Rt[2], zz, __fma_rn(fr.Rt[1], y, __dmul_rn(fr.Rt[0], x))),
              fr.c[0]);
          wy = __dsub_rn(
              __fma_rn(fr.Rt[5], zz, __fma_rn(fr.Rt[4], y, __dmul_rn(fr.Rt[3], x))),
              fr.c[1]);
          wz = __dsub_rn(
              __fma_rn(fr.Rt[8], zz, __fma_rn(fr.Rt[7], y, __dmul_rn(fr.Rt[6], x))),
              fr.c[2]);
        }
        uint32_t* dst = s_xyz + shift_w + rank * 3;
        dst[0] = __float_as_uint(__double2float_rn(wx));
        dst[1] = __float_as_uint(__double2float_rn(wy));
        dst[2] = __float_as_uint(__double2float_rn(wz));
        if (p.has_color && !RGB_F32) {
          uint8_t* cd = s_rgb + shift_b + rank * 3;
          cd[0] = (uint8_t)(crgb[j] & 255u);  // BGR -> RGB (d2r:381-382)
          cd[1] = (uint8_t)((crgb[j] >> 8) & 255u);
          cd[2] = (uint8_t)(crgb[j] >> 16);
        }
      }
      uu += BP_THREADS;
      while (uu >= p.Ws) {
        uu -= p.Ws;
        ++vv;
      }
    }
  }
  __syncthreads();
  flush_words(s_xyz, reinterpret_cast<uint32_t*>(p.out_xyz), base * 3, total * 3);
  if (!p.has_color) return;
  if (!RGB_F32) {
    flush_bytes(s_rgb, reinterpret_cast<uint8_t*>(p.out_rgb), base * 3, total * 3);
  } else {
    // second phase through the same word buffer: colours as f32 in [0,1]
    // (u8.astype(f32) / 255, depth_processor.py:417)
    __syncthreads();
#pragma unroll
    for (int j = 0; j < BP_PPT; ++j) {
      if (bal[j] != 0u) {
        const int rank =
            s_woff[j * BP_WARPS + warp] + __popc(bal[j] & lanemask_lt());
        uint32_t* dst = s_xyz + shift_w + rank * 3;
        dst[0] = __float_as_uint(__fdiv_rn((float)(crgb[j] & 255u), 255.0f));
        dst[1] = __float_as_uint(__fdiv_rn((float)((crgb[j] >> 8) & 255u), 255.0f));
        dst[2] = __float_as_uint(__fdiv_rn((float)((crgb[j] >> 16) & 255u), 255.0f));
      }
    }
    __syncthreads();
    flush_words(s_xyz, reinterpret_cast<uint32_t*>(p.out_rgb), base * 3, total * 3);
  }
}

// ---------------------------------------------------------------------------
// K1s — the subsample = 1 streaming variant (the 1080p / 4K roofline case), f32 depth.
//
// At s = 1 a tile of 2048 sampled pixels is one contiguous run of the frame: 8 KiB of depth and
// 6 KiB of BGR.  Persistent CTAs (3 per SM) take tiles from the ordered ticket and keep two
// shared-memory rings filled by the TMA bulk-copy engine (cp.async.bulk + mbarrier complete_tx,
// one elected thread): a 3-slot depth ring and a 2-slot colour ring.  In iteration k a CTA
//   - tickets tile T(k+2) and starts its depth copy,
//   - fires the look-back polls of T(k) (its predecessors published long ago: no spinning),
//   - runs pass A of T(k+1): masks of 2 x 4 CONSECUTIVE pixels per thread (128-bit shared
//     loads), one packed warp scan instead of 8 ballots, and PUBLISHES T(k+1)'s aggregate — a
//     tile's count is visible as soon as its depth has landed, so a ticket held for prefetching
//     never delays another CTA's look-back,
//   - consumes the polls -> base of T(k), runs pass B of T(k) (f64 back-projection, straight-line
//     4-pixel fast path), starts the colour copy of T(k+2) into the slot just consumed, and
//     flushes T(k) with 128-bit stores.
// Every byte is requested a full iteration before it is needed, so HBM latency is hidden by the
// rings, not by occupancy.  Ordered compaction, staging, flush and every floating-point operation
// are those of the generic kernel (bit-identical outputs).  The last (partial) tile of a frame
// and the optional confidence mask are read with plain loads.
// ---------------------------------------------------------------------------
constexpr int ST_TILE = 2048;
constexpr int ST_THREADS = 256;
constexpr int ST_DSLOTS = 3;
constexpr int ST_CSLOTS = 2;
constexpr int ST_CTAS_PER_SM = 3;
constexpr int ST_DEPTH_BYTES = ST_TILE * 4;
constexpr int ST_BGR_BYTES = ST_TILE * 3;
constexpr int ST_XYZ_BYTES = (ST_TILE * 3 + 4) * 4;
constexpr int ST_RGB_BYTES = ST_TILE * 3 + 16;
constexpr int ST_SMEM_BYTES = ST_DSLOTS * ST_DEPTH_BYTES + ST_CSLOTS * ST_BGR_BYTES + ST_XYZ_BYTES + ST_RGB_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t a = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}

// F64_MASK: mask arithmetic in f64 (np.float64 scale), else f32 (python-float scale)
template <bool F64_MASK, bool RGB_F32>
__global__ void __launch_bounds__(ST_THREADS, ST_CTAS_PER_SM)
    backproject_stream_kernel(const __grid_constant__ BPParams p) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint8_t* const s_depth = st_smem;
  uint8_t* const s_bgr = st_smem + ST_DSLOTS * ST_DEPTH_BYTES;
  uint32_t* const s_xyz = reinterpret_cast<uint32_t*>(s_bgr + ST_CSLOTS * ST_BGR_BYTES);
  uint8_t* const s_rgb = reinterpret_cast<uint8_t*>(s_xyz) + ST_XYZ_BYTES;
  __shared__ __align__(8) uint64_t s_dfull[ST_DSLOTS];
  __shared__ __align__(8) uint64_t s_cfull[ST_CSLOTS];
  __shared__ int s_tiles[ST_DSLOTS];
  __shared__ int s_fi[ST_DSLOTS];
  __shared__ int s_ltile[ST_DSLOTS];
  __shared__ int s_wcount[2 * BP_WARPS];
  __shared__ int s_woff[ST_DSLOTS][2 * BP_WARPS];
  __shared__ int s_total[ST_DSLOTS];
  __shared__ long long s_part[BP_WARPS];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  // elected thread: ticket for depth slot `ds` + bulk copy of the tile's depth (full tiles only;
  // a partial tile is loaded co-operatively when it is consumed)
  auto ticket_and_depth = [&](int ds) {
    const int tile = (int)atomicAdd(p.ticket, 1u);
    if (tile >= p.num_tiles) {
      s_tiles[ds] = -1;
      return;
    }
    s_tiles[ds] = tile;
    const int fi = tile / p.tiles_per_frame;
    const int ltile = tile - fi * p.tiles_per_frame;
    s_fi[ds] = fi;          // one division per tile instead of one per thread and use
    s_ltile[ds] = ltile;
    if (p.P - ltile * ST_TILE < ST_TILE) return;
    mbar_expect_tx(&s_dfull[ds], ST_DEPTH_BYTES);
    tma_bulk_g2s(s_depth + ds * ST_DEPTH_BYTES,
                 reinterpret_cast<const uint8_t*>(p.f[fi].depth) + (size_t)ltile * ST_DEPTH_BYTES, ST_DEPTH_BYTES,
                 &s_dfull[ds]);
  };
  // elected thread: bulk copy of the colours of `tile` into colour slot `cs`
  auto colour = [&](int tile, int cs) {
    if (tile < 0 || !p.has_color) return;
    const int fi = tile / p.tiles_per_frame;
    const int ltile = tile - fi * p.tiles_per_frame;
    if (p.P - ltile * ST_TILE < ST_TILE) return;
    mbar_expect_tx(&s_cfull[cs], ST_BGR_BYTES);
    tma_bulk_g2s(s_bgr + cs * ST_BGR_BYTES, p.f[fi].bgr + (size_t)ltile * ST_BGR_BYTES, ST_BGR_BYTES, &s_cfull[cs]);
  };

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < ST_DSLOTS; ++i) mbar_init(&s_dfull[i], 1);
#pragma unroll
    for (int i = 0; i < ST_CSLOTS; ++i) mbar_init(&s_cfull[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ticket_and_depth(0);
    ticket_and_depth(1);
    s_tiles[2] = -1;
    colour(s_tiles[0], 0);
    colour(s_tiles[1], 1);
  }
  __syncthreads();

  uint32_t dphase = 0, cphase = 0;  // bit i = parity the next TMA-filled use of slot i completes

  // Pass A of the tile in depth slot `ds` (see the header comment).
  auto pass_a = [&](int ds, unsigned& m_out, unsigned& excl_out) {
    m_out = 0;
    excl_out = 0;
    const int tile = s_tiles[ds];
    if (tile < 0) return;  // uniform
    const int fi = s_fi[ds];
    const int ltile = s_ltile[ds];
    const BPFrame& fr = p.f[fi];
    const int npx = min(ST_TILE, p.P - ltile * ST_TILE);
    const float* sd = reinterpret_cast<const float*>(s_depth + ds * ST_DEPTH_BYTES);
    if (npx == ST_TILE) {
      mbar_wait(&s_dfull[ds], (dphase >> ds) & 1u);
      dphase ^= 1u << ds;
    } else {
      const float* src = reinterpret_cast<const float*>(fr.depth) + (size_t)ltile * ST_TILE;
      float* dst = reinterpret_cast<float*>(s_depth + ds * ST_DEPTH_BYTES);
      for (int i = tid; i < npx; i += ST_THREADS) dst[i] = __ldg(src + i);
      __syncthreads();
    }
    unsigned m = 0;
    if (npx == ST_TILE && fr.conf == nullptr) {  // the common case: no per-pixel bounds / mask tests
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float4 a = *reinterpret_cast<const float4*>(sd + g * (ST_TILE / 2) + 4 * tid);
        const float d32[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          bool valid;
          if (!F64_MASK) {
            const float ds_ = __fmul_rn(d32[k], p.scale32);
            valid = (ds_ > p.min32) && (ds_ < p.max32);
          } else {
            const double ds_ = __dmul_rn((double)d32[k], p.scale);
            valid = (ds_ > p.min_d) && (ds_ < p.max_d);
          }
          if (valid) m |= 1u << (4 * g + k);
        }
      }
    } else
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int idx = g * (ST_TILE / 2) + 4 * tid;
      const float4 a = *reinterpret_cast<const float4*>(sd + idx);
      const float d32[4] = {a.x, a.y, a.z, a.w};
      unsigned cf = 0x01010101u;
      if (fr.conf != nullptr) {  // optional extension: plain loads (tile offsets are multiples of 4)
        const uint8_t* c = fr.conf + (size_t)ltile * ST_TILE + idx;
        cf = idx + 3 < npx ? __ldg(reinterpret_cast<const unsigned*>(c))
                           : (idx < npx ? (unsigned)__ldg(c) : 0u) | (idx + 1 < npx ? (unsigned)__ldg(c + 1) << 8 : 0u) |
                                 (idx + 2 < npx ? (unsigned)__ldg(c + 2) << 16 : 0u);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        bool valid;
        if (!F64_MASK) {
          const float ds_ = __fmul_rn(d32[k], p.scale32);
          valid = (ds_ > p.min32) && (ds_ < p.max32);  // implies isfinite (d2r:359-361)
        } else {
          const double ds_ = __dmul_rn((double)d32[k], p.scale);
          valid = (ds_ > p.min_d) && (ds_ < p.max_d);
        }
        valid = valid && ((cf >> (8 * k)) & 255u) != 0u && (idx + k < npx);
        m |= (valid ? 1u : 0u) << (4 * g + k);
      }
    }
    const unsigned packed = (unsigned)__popc(m & 15u) | ((unsigned)__popc(m >> 4) << 16);
    unsigned inc = packed;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    m_out = m;
    excl_out = inc - packed;
    if (lane == 31) {
      s_wcount[warp] = (int)(inc & 0xffffu);
      s_wcount[BP_WARPS + warp] = (int)(inc >> 16);
    }
  };
  // second half of pass A, after a barrier: warp 0 scans the 16 (group, warp) counts and
  // publishes the tile's aggregate
  auto pass_a_publish = [&](int ds) {
    const int tile = s_tiles[ds];
    if (tile < 0) return;
    if (warp == 0) {
      const int cnt = lane < 2 * BP_WARPS ? s_wcount[lane] : 0;
      int sc = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, sc, d);
        if (lane >= d) sc += t;
      }
      if (lane < 2 * BP_WARPS) s_woff[ds][lane] = sc - cnt;
      const int total = __shfl_sync(0xffffffffu, sc, 31);
      if (lane == 0) {
        st_state(p.tile_state + tile, FLAG_AGG | (unsigned long long)total);
        s_total[ds] = total;
      }
    }
  };

  // Flat look-back in two halves so that its memory latency hides under pass A of the next
  // tile: lb_issue() fires this thread's (<= 4) polls, lb_finish() consumes them.
  constexpr int LB_POLLS = BP_CHK / ST_THREADS;
  auto lb_src = [&](int tile, int k) -> const uint64_t* {
    const int c = ((tile - 1) / BP_CHK) * BP_CHK;
    const int j = c + tid + k * ST_THREADS;
    if (tile <= 0 || j >= tile) return nullptr;
    return reinterpret_cast<const uint64_t*>(j == c ? p.inc_state + c / BP_CHK : p.tile_state + j);
  };
  auto lb_issue = [&](int tile, unsigned long long (&pv)[LB_POLLS]) {
#pragma unroll
    for (int k = 0; k < LB_POLLS; ++k) {
      const uint64_t* src = lb_src(tile, k);
      pv[k] = src ? ld_volatile_u64(src) : FLAG_AGG;
    }
  };
  auto lb_finish = [&](int tile, unsigned long long (&pv)[LB_POLLS]) {
    long long part = 0;
#pragma unroll
    for (int k = 0; k < LB_POLLS; ++k) {
      unsigned long long stt = pv[k];
      if ((stt >> 62) == 0) {
        const uint64_t* src = lb_src(tile, k);
        do {
          stt = ld_volatile_u64(src);
        } while ((stt >> 62) == 0);
      }
      part += (long long)(stt & VAL_MASK);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) s_part[warp] = part;
  };

  unsigned m = 0, excl = 0, m_nxt = 0, excl_nxt = 0;
  unsigned long long pv[LB_POLLS];
  pass_a(0, m, excl);
  __syncthreads();
  pass_a_publish(0);
  __syncthreads();  // s_wcount is rewritten by the next pass A
  for (int it = 0;; ++it) {
    const int ds = it % ST_DSLOTS;          // depth slot of the current tile
    const int ds_n = (it + 1) % ST_DSLOTS;  // ... of the next tile (pass A now)
    const int ds_nn = (it + 2) % ST_DSLOTS; // ... of the tile ticketed now
    const int cs = it % ST_CSLOTS;
    const int tile = s_tiles[ds];
    if (tile < 0) break;  // tickets are monotonic: the other slots hold nothing either
    const int fi = s_fi[ds];
    const int ltile = s_ltile[ds];
    const BPFrame& fr = p.f[fi];
    const int npx = min(ST_TILE, p.P - ltile * ST_TILE);
    const uint8_t* sd = s_depth + ds * ST_DEPTH_BYTES;
    uint8_t* sc = s_bgr + cs * ST_BGR_BYTES;
    const int total = s_total[ds];
    if (tid == 0) ticket_and_depth(ds_nn);  // slot of tile it-1: free since the last barrier
    lb_issue(tile, pv);
    pass_a(ds_n, m_nxt, excl_nxt);
    lb_finish(tile, pv);
    if (npx < ST_TILE && p.has_color)  // partial tile: colours with plain loads
      for (int i = tid; i < npx * 3; i += ST_THREADS) sc[i] = __ldg(fr.bgr + (size_t)ltile * ST_BGR_BYTES + i);
    __syncthreads();  // s_wcount of the next tile, s_part of this one, partial-tile colours
    pass_a_publish(ds_n);
    long long base = 0;
#pragma unroll
    for (int w = 0; w < BP_WARPS; ++w) base += s_part[w];
    if (tid == 0 && tile % BP_CHK == 0)
      st_state(p.inc_state + tile / BP_CHK, FLAG_INC | (unsigned long long)(base + total));
    if (p.base_ptr) base += *p.base_ptr;  // later chunk of a batch
    if (tid == 0) {
      if (p.out_offsets) {
        if (ltile == 0) p.out_offsets[fi] = base;
        if (tile == p.num_tiles - 1) p.out_offsets[p.n_frames] = base + total;
      } else if (tile == p.num_tiles - 1) {
        *p.out_n = base + total;
      }
    }
    if (npx == ST_TILE && p.has_color) {
      mbar_wait(&s_cfull[cs], (cphase >> cs) & 1u);
      cphase ^= 1u << cs;
    }

    // ---- pass B: back-project the valid pixels into the staging buffers ----------------
    const int shift_w = (int)((base * 3) & 3);
    const int shift_b = (int)((base * 3) & 15);
    if (total > 0 && m != 0u) {
      double Rt[9], cc[3];
      if (p.has_pose) {
#pragma unroll
        for (int i = 0; i < 9; ++i) Rt[i] = fr.Rt[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) cc[i] = fr.c[i];
      }
      // one pixel: depth at tile index i (pixel column uk, row vk) -> staged point `rank`
      auto point = [&](int i, int uk, int vk, int rank) {
        const float d = *reinterpret_cast<const float*>(sd + (size_t)i * 4);
        const double zz = F64_MASK ? __dmul_rn((double)d, p.scale) : (double)__fmul_rn(d, p.scale32);
        const double x = __dmul_rn(__ldg(p.xf + uk), zz);
        const double y = __dmul_rn(__ldg(p.yf + vk), zz);
        double wx = x, wy = y, wz = zz;
        if (p.has_pose) {
          wx = __dsub_rn(__fma_rn(Rt[2], zz, __fma_rn(Rt[1], y, __dmul_rn(Rt[0], x))), cc[0]);
          wy = __dsub_rn(__fma_rn(Rt[5], zz, __fma_rn(Rt[4], y, __dmul_rn(Rt[3], x))), cc[1]);
          wz = __dsub_rn(__fma_rn(Rt[8], zz, __fma_rn(Rt[7], y, __dmul_rn(Rt[6], x))), cc[2]);
        }
        uint32_t* dst = s_xyz + shift_w + rank * 3;
        dst[0] = __float_as_uint(__double2float_rn(wx));
        dst[1] = __float_as_uint(__double2float_rn(wy));
        dst[2] = __float_as_uint(__double2float_rn(wz));
      };
      // pixel coordinates of the tile's first pixel: one division per thread and tile
      const int lin0 = ltile * ST_TILE;
      const int vt = lin0 / p.W;
      const int ut = lin0 - vt * p.W;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const unsigned mg = (m >> (4 * g)) & 15u;
        if (mg == 0u) continue;
        const int idx = g * (ST_TILE / 2) + 4 * tid;
        const int rank0 = s_woff[ds][g * BP_WARPS + warp] + (int)((excl >> (16 * g)) & 0xffffu);
        int u0 = ut + idx, v0 = vt;
        if (p.W >= 256) {  // a tile spans at most 2048 / W + 1 rows
          while (u0 >= p.W) { u0 -= p.W; ++v0; }
        } else {
          const int q = u0 / p.W;
          u0 -= q * p.W;
          v0 += q;
        }
        uint32_t cw[3] = {0u, 0u, 0u};
        if (p.has_color) {
          const uint32_t* c32 = reinterpret_cast<const uint32_t*>(sc + idx * 3);
          cw[0] = c32[0]; cw[1] = c32[1]; cw[2] = c32[2];
        }
        if (mg == 15u) {
          // all four pixels valid: straight-line code, four independent f64 chains in flight
          int uk[4], vk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uk[k] = u0 + k;
            vk[k] = v0;
            if (uk[k] >= p.W) { uk[k] -= p.W; ++vk[k]; }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) point(idx + k, uk[k], vk[k], rank0 + k);
          if (p.has_color && !RGB_F32) {
            // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3  ->  R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
            const uint32_t o0 = __byte_perm(cw[0], cw[1], 0x5012);
            const uint32_t o1 = __byte_perm(__byte_perm(cw[1], cw[2], 0x3400), cw[0], 0x3270);
            const uint32_t o2 = __byte_perm(cw[1], cw[2], 0x5672);
            uint8_t* cd = s_rgb + shift_b + rank0 * 3;
            if ((((uintptr_t)cd) & 3u) == 0u) {
              uint32_t* c4 = reinterpret_cast<uint32_t*>(cd);
              c4[0] = o0; c4[1] = o1; c4[2] = o2;
            } else {
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                cd[b] = (uint8_t)(o0 >> (8 * b));
                cd[4 + b] = (uint8_t)(o1 >> (8 * b));
                cd[8 + b] = (uint8_t)(o2 >> (8 * b));
              }
            }
          }
        } else {
          int rank = rank0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (!(mg & (1u << k))) continue;
            int uk = u0 + k, vk = v0;
            if (uk >= p.W) { uk -= p.W; ++vk; }
            point(idx + k, uk, vk, rank);
            if (p.has_color && !RGB_F32) {
              // bytes 3k..3k+2 of the 12-byte group are B, G, R
              const uint64_t lo = (uint64_t)cw[0] | ((uint64_t)cw[1] << 32);
              const uint32_t bgr = k < 2 ? (uint32_t)(lo >> (24 * k)) & 0xffffffu
                                         : (k == 2 ? (uint32_t)((lo >> 48) | ((uint64_t)cw[2] << 16)) & 0xffffffu
                                                   : (cw[2] >> 8));
              uint8_t* cd = s_rgb + shift_b + rank * 3;
              cd[0] = (uint8_t)(bgr >> 16);  // BGR -> RGB (d2r:381-382)
              cd[1] = (uint8_t)(bgr >> 8);
              cd[2] = (uint8_t)bgr;
            }
            ++rank;
          }
        }
      }
    }
    __syncthreads();  // staging complete (and, unless RGB_F32, the colour slot is no longer read)
    if (!RGB_F32 && tid == 0) colour(s_tiles[ds_nn], cs);
    if (total > 0) {
      flush_words(s_xyz, reinterpret_cast<uint32_t*>(p.out_xyz), base * 3, total * 3);
      if (p.has_color) {
        if (!RGB_F32) {
          flush_bytes(s_rgb, reinterpret_cast<uint8_t*>(p.out_rgb), base * 3, total * 3);
        } else {
          // second phase through the word buffer: colours as f32 in [0,1] (dp:417)
          __syncthreads();
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const unsigned mg = (m >> (4 * g)) & 15u;
            const int idx = g * (ST_TILE / 2) + 4 * tid;
            int rank = s_woff[ds][g * BP_WARPS + warp] + (int)((excl >> (16 * g)) & 0xffffu);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (!(mg & (1u << k))) continue;
              const uint8_t* c = sc + (idx + k) * 3;
              uint32_t* dst = s_xyz + shift_w + rank * 3;
              dst[0] = __float_as_uint(__fdiv_rn((float)c[2], 255.0f));
              dst[1] = __float_as_uint(__fdiv_rn((float)c[1], 255.0f));
              dst[2] = __float_as_uint(__fdiv_rn((float)c[0], 255.0f));
              ++rank;
            }
          }
          __syncthreads();
          flush_words(s_xyz, reinterpret_cast<uint32_t*>(p.out_rgb), base * 3, total * 3);
        }
      }
    }
    if (RGB_F32) {
      __syncthreads();
      if (tid == 0) colour(s_tiles[ds_nn], cs);
    }
    __syncthreads();  // staging buffers reusable; s_tiles[ds_nn] visible
    m = m_nxt;
    excl = excl_nxt;
  }
}

__global__ void proj_table_kernel(double* xf, double* yf, int W, int H, double fx,
                                  double fy, double cx, double cy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // (u - cx) / fx with IEEE subtraction and division, as NumPy (d2r:292-293)
  if (i < W) xf[i] = __ddiv_rn(__dsub_rn((double)i, cx), fx);
  if (i < H) yf[i] = __ddiv_rn(__dsub_rn((double)i, cy), fy);
}

int get_proj_table(t3d_ctx* ctx, int H, int W, double fx, double fy, double cx,
                   double cy, cudaStream_t st, const ProjTable** out) {
  for (const ProjTable& t : ctx->proj) {
    if (t.H == H && t.W == W && t.fx == fx && t.fy == fy && t.cx == cx &&
        t.cy == cy) {
      *out = &t;
      return T3D_OK;
    }
  }
  if (ctx->proj.size() >= 16) {  // bounded cache
    for (ProjTable& t : ctx->proj) {
      cudaFree(t.xf);
      cudaFree(t.yf);
    }
    ctx->proj.clear();
  }
  ProjTable t;
  t.H = H; t.W = W; t.fx = fx; t.fy = fy; t.cx = cx; t.cy = cy;
  T3D_CUDA(cudaMalloc(&t.xf, sizeof(double) * (size_t)(W > 0 ? W : 1)));
  T3D_CUDA(cudaMalloc(&t.yf, sizeof(double) * (size_t)(H > 0 ? H : 1)));
  const int n = H > W ? H : W;
  if (n > 0) {
    proj_table_kernel<<<(n + 255) / 256, 256, 0, st>>>(t.xf, t.yf, W, H, fx, fy,
                                                        cx, cy);
    T3D_LAUNCH_CHECK();
    ctx->launches++;
  }
  ctx->proj.push_back(t);
  *out = &ctx->proj.back();
  return T3D_OK;
}

}  // namespace

// Shared launcher: frames[0..n) (n <= BP_MAX_BATCH) of identical geometry in one launch.
static int bp_launch(t3d_ctx* ctx, const t3d_backproject_params* q, const BPFrame* frames, int n,
                     float* out_xyz, void* out_rgb, long long* out_n, long long* out_offsets,
                     const long long* base_ptr, cudaStream_t st) {
  const int s = q->subsample;
  const int Hs = (q->H + s - 1) / s, Ws = (q->W + s - 1) / s;
  const int64_t P64 = (int64_t)Hs * Ws;
  const ProjTable* tab = nullptr;
  int rc = get_proj_table(ctx, q->H, q->W, q->fx, q->fy, q->cx, q->cy, st, &tab);
  if (rc != T3D_OK) return rc;

  BPParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < n; ++i) p.f[i] = frames[i];
  p.n_frames = n;
  p.out_offsets = out_offsets;
  p.base_ptr = base_ptr;
  p.xf = tab->xf;
  p.yf = tab->yf;
  p.W = q->W;
  p.s = s;
  p.Ws = Ws;
  p.P = (int)P64;
  p.has_pose = q->has_pose;
  p.has_color = q->has_color;
  p.scale = q->scale;
  p.min_d = q->min_depth;
  p.max_d = q->max_depth;
  p.scale32 = (float)q->scale;      // NumPy weak-scalar cast (NEP 50)
  p.min32 = (float)q->min_depth;
  p.max32 = (float)q->max_depth;
  p.out_xyz = out_xyz;
  p.out_rgb = out_rgb;
  p.out_n = out_n;
  static int ppt = -1, minb = -1;  // tuning knobs: T3D_K1_PPT = 4 | 8 pixels per thread, T3D_K1_MINB = CTAs/SM
  if (ppt < 0) {
    const char* e = getenv("T3D_K1_PPT");
    ppt = e ? atoi(e) : 8;
    if (ppt != 4 && ppt != 8) ppt = 8;
    const char* m = getenv("T3D_K1_MINB");
    minb = m ? atoi(m) : 5;
  }
  // s = 1 with 16-byte aligned frames: the TMA-staged streaming kernel (K1s)
  static int use_stream = -1;  // T3D_K1_STREAM=0 forces the generic kernel (A/B measurements)
  if (use_stream < 0) {
    const char* e = getenv("T3D_K1_STREAM");
    use_stream = e ? atoi(e) : 1;
  }
  p.wide_bgr = (q->has_color && ((long long)q->H * q->W * 3) % 4 == 0) ? 1 : 0;
  for (int i = 0; i < n && p.wide_bgr; ++i) p.wide_bgr = ((uintptr_t)frames[i].bgr % 4 == 0) ? 1 : 0;
  bool stream = use_stream != 0 && s == 1 && q->W >= 4 && !q->depth_is_f64;
  for (int i = 0; i < n && stream; ++i) {
    const BPFrame& f = frames[i];
    stream = ((uintptr_t)f.depth % 16 == 0) && (!q->has_color || (uintptr_t)f.bgr % 16 == 0) &&
             ((uintptr_t)f.conf % 4 == 0);
  }
  const int tile_px = stream ? ST_TILE : BP_THREADS * ppt;
  p.tiles_per_frame = (p.P + tile_px - 1) / tile_px;
  p.num_tiles = p.tiles_per_frame * n;

  const size_t n_chk = (size_t)p.num_tiles / BP_CHK + 1;
  const size_t state_bytes = sizeof(uint64_t) * ((size_t)p.num_tiles + n_chk + 2);
  rc = ctx->scan_state.reserve(state_bytes);
  if (rc != T3D_OK) return rc;
  T3D_CUDA(cudaMemsetAsync(ctx->scan_state.p, 0, state_bytes, st));
  p.ticket = ctx->scan_state.as<unsigned>();
  p.tile_state = ctx->scan_state.as<unsigned long long>() + 1;
  p.inc_state = p.tile_state + p.num_tiles;

  const int mode = q->depth_is_f64 ? 2 : (q->scale_is_f64 ? 1 : 0);
  if (stream) {
    const int want = ctx->num_sms * ST_CTAS_PER_SM;
    const dim3 sgrid(p.num_tiles < want ? p.num_tiles : want), sblock(ST_THREADS);
#define ST_LAUNCH(M, F)                                                                              \
  do {                                                                                               \
    if (ctx->func_attr_needed(reinterpret_cast<const void*>(&backproject_stream_kernel<M, F>)))      \
      T3D_CUDA(cudaFuncSetAttribute(backproject_stream_kernel<M, F>,                                 \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM_BYTES));    \
    backproject_stream_kernel<M, F><<<sgrid, sblock, ST_SMEM_BYTES, st>>>(p);                        \
  } while (0)
    if (q->rgb_out_f32) {
      if (mode == 0) ST_LAUNCH(false, true);
      else ST_LAUNCH(true, true);
    } else {
      if (mode == 0) ST_LAUNCH(false, false);
      else ST_LAUNCH(true, false);
    }
#undef ST_LAUNCH
    T3D_LAUNCH_CHECK();
    ctx->launches++;
    return T3D_OK;
  }
  const dim3 grid(p.num_tiles), block(BP_THREADS);
#define BP_LAUNCH(M, F)                                                        \
  do {                                                                         \
    if (ppt == 8 && minb >= 5) backproject_kernel<M, F, 8, 5><<<grid, block, 0, st>>>(p); \
    else if (ppt == 8) backproject_kernel<M, F, 8, 4><<<grid, block, 0, st>>>(p);         \
    else if (minb >= 6) backproject_kernel<M, F, 4, 6><<<grid, block, 0, st>>>(p);        \
    else backproject_kernel<M, F, 4, 5><<<grid, block, 0, st>>>(p);                       \
  } while (0)
  if (q->rgb_out_f32) {
    if (mode == 0) BP_LAUNCH(0, true);
    else if (mode == 1) BP_LAUNCH(1, true);
    else BP_LAUNCH(2, true);
  } else {
    if (mode == 0) BP_LAUNCH(0, false);
    else if (mode == 1) BP_LAUNCH(1, false);
    else BP_LAUNCH(2, false);
  }
#undef BP_LAUNCH
  T3D_LAUNCH_CHECK();
  ctx->launches++;
  return T3D_OK;
}

// Rt = R^T ; c = R^T t with the dgemm k-ordered FMA chain (d2r:376)
static void bp_pose(const double* R, const double* t, BPFrame* f) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) f->Rt[i * 3 + j] = R[j * 3 + i];
  for (int i = 0; i < 3; ++i)
    f->c[i] = fma(f->Rt[i * 3 + 2], t[2], fma(f->Rt[i * 3 + 1], t[1], f->Rt[i * 3 + 0] * t[0]));
}

static int bp_check(const t3d_backproject_params* q, int64_t* P64) {
  T3D_REQUIRE(q->H >= 0 && q->W >= 0 && q->subsample >= 1,
              "t3d_backproject: bad H=%d W=%d subsample=%d", q->H, q->W, q->subsample);
  const int s = q->subsample;
  const int Hs = (q->H + s - 1) / s, Ws = (q->W + s - 1) / s;
  *P64 = (int64_t)Hs * Ws;
  T3D_REQUIRE(*P64 * BP_MAX_BATCH < (1ll << 31) - 4096 * BP_MAX_BATCH, "t3d_backproject: frame too large");
  return T3D_OK;
}

extern "C" int t3d_backproject(t3d_ctx* ctx, const void* depth, const uint8_t* bgr,
                               const uint8_t* conf_mask,
                               const t3d_backproject_params* q, float* out_xyz,
                               void* out_rgb, int64_t capacity, int64_t* out_n,
                               t3d_stream stream) {
  T3D_REQUIRE(ctx && q && out_n, "t3d_backproject: null ctx/params/out_n");
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(!q->has_color || bgr, "t3d_backproject: has_color but bgr is NULL");
  cudaStream_t st = as_stream(stream);
  int64_t P64 = 0;
  int rc = bp_check(q, &P64);
  if (rc != T3D_OK) return rc;
  if (P64 == 0) {
    T3D_CUDA(cudaMemsetAsync(out_n, 0, sizeof(int64_t), st));
    return T3D_OK;
  }
  T3D_REQUIRE(depth && out_xyz, "t3d_backproject: null depth/out_xyz");
  if (capacity < P64) {
    t3d_set_error("t3d_backproject: capacity %lld < sampled pixels %lld",
                  (long long)capacity, (long long)P64);
    return T3D_E_CAPACITY;
  }
  T3D_REQUIRE(!q->has_color || out_rgb, "t3d_backproject: null out_rgb");
  BPFrame f;
  memset(&f, 0, sizeof(f));
  f.depth = depth;
  f.bgr = bgr;
  f.conf = conf_mask;
  if (q->has_pose) bp_pose(q->R, q->t, &f);
  return bp_launch(ctx, q, &f, 1, out_xyz, out_rgb, reinterpret_cast<long long*>(out_n), nullptr,
                   nullptr, st);
}

extern "C" int t3d_backproject_batch(t3d_ctx* ctx, const t3d_backproject_frame* frames_h,
                                     int n_frames, const t3d_backproject_params* q,
                                     float* out_xyz, void* out_rgb, int64_t capacity,
                                     int64_t* out_offsets, t3d_stream stream) {
  T3D_REQUIRE(ctx && q && out_offsets && (n_frames == 0 || frames_h),
              "t3d_backproject_batch: null argument");
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(n_frames >= 0, "t3d_backproject_batch: n_frames < 0");
  cudaStream_t st = as_stream(stream);
  int64_t P64 = 0;
  int rc = bp_check(q, &P64);
  if (rc != T3D_OK) return rc;
  if (P64 == 0 || n_frames == 0) {
    T3D_CUDA(cudaMemsetAsync(out_offsets, 0, sizeof(int64_t) * (size_t)(n_frames + 1), st));
    return T3D_OK;
  }
  T3D_REQUIRE(out_xyz && (!q->has_color || out_rgb), "t3d_backproject_batch: null outputs");
  if (capacity < P64 * n_frames) {
    t3d_set_error("t3d_backproject_batch: capacity %lld < sampled pixels %lld",
                  (long long)capacity, (long long)(P64 * n_frames));
    return T3D_E_CAPACITY;
  }
  long long* offs = reinterpret_cast<long long*>(out_offsets);
  for (int c0 = 0; c0 < n_frames; c0 += BP_MAX_BATCH) {
    const int n = n_frames - c0 < BP_MAX_BATCH ? n_frames - c0 : BP_MAX_BATCH;
    BPFrame f[BP_MAX_BATCH];
    memset(f, 0, sizeof(f));
    for (int i = 0; i < n; ++i) {
      const t3d_backproject_frame& src = frames_h[c0 + i];
      T3D_REQUIRE(src.depth && (!q->has_color || src.bgr), "t3d_backproject_batch: frame %d has null depth/bgr",
                  c0 + i);
      f[i].depth = src.depth;
      f[i].bgr = src.bgr;
      f[i].conf = src.conf_mask;
      if (q->has_pose) bp_pose(src.R, src.t, &f[i]);
    }
    // chunk c writes offsets[c0 .. c0+n]; its starting offset is offsets[c0] of the previous launch
    rc = bp_launch(ctx, q, f, n, out_xyz, out_rgb, nullptr, offs + c0, c0 > 0 ? offs + c0 : nullptr, st);
    if (rc != T3D_OK) return rc;
  }
  return T3D_OK;
}
