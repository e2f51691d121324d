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
        const uint8_t* c = fr.bgr + sp * 3;
        crgb[j] = (uint32_t)__ldg(c + 2) | ((uint32_t)__ldg(c + 1) << 8) | ((uint32_t)__ldg(c) << 16);
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
  const int tile_px = BP_THREADS * ppt;
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
