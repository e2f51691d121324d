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
// a 64-entry block scan and a single-pass decoupled look-back across tiles, so
// depth/colour are read exactly once and nothing but the packed outputs is
// written.  Points and colours are staged in shared memory at an offset that
// is congruent (mod 16 B) to their final global address, then streamed out
// with 128-bit stores.  The arithmetic is f64 like NumPy's (x_factor tables
// are built with IEEE division, the 3x3 pose product uses the k-ordered FMA
// chain of a dgemm micro-kernel) and rounded to f32 once.
#include "common.cuh"

namespace {

constexpr int BP_THREADS = 256;
constexpr int BP_PPT = 8;
constexpr int BP_TILE = BP_THREADS * BP_PPT;  // 2048 sampled pixels
constexpr int BP_WARPS = BP_THREADS / 32;

struct BPParams {
  const void* depth;
  const uint8_t* bgr;
  const uint8_t* conf;
  const double* xf;
  const double* yf;
  int W, s, Ws;
  int P;  // sampled pixels = Hs*Ws
  int has_pose, has_color;
  double scale, min_d, max_d;
  float scale32, min32, max32;
  double Rt[9];
  double c[3];
  float* out_xyz;
  void* out_rgb;
  long long* out_n;
  unsigned long long* tile_state;
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
template <int DEPTH_MODE, bool RGB_F32>
__global__ void __launch_bounds__(BP_THREADS)
    backproject_kernel(const __grid_constant__ BPParams p) {
  __shared__ __align__(16) uint32_t s_xyz[BP_TILE * 3 + 4];
  __shared__ __align__(16) uint8_t s_rgb[BP_TILE * 3 + 16];
  __shared__ int s_wcount[BP_PPT * BP_WARPS];
  __shared__ int s_woff[BP_PPT * BP_WARPS];
  __shared__ int s_tile, s_total;
  __shared__ long long s_base;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid == 0) s_tile = (int)atomicAdd(p.ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int p0 = tile * BP_TILE + tid;

  // ---- pass A: load depth, evaluate the mask, count -----------------------
  double z[BP_PPT];
  unsigned bal[BP_PPT];
  int u = 0, v = 0;
  if (p0 < p.P) {
    v = p0 / p.Ws;
    u = p0 - v * p.Ws;
  }
  int uu = u, vv = v;
#pragma unroll
  for (int j = 0; j < BP_PPT; ++j) {
    const int pj = p0 + j * BP_THREADS;
    bool valid = false;
    double zj = 0.0;
    if (pj < p.P) {
      const long long sp = (long long)(vv * p.s) * p.W + (long long)uu * p.s;
      if (DEPTH_MODE == 0) {
        const float d = __ldg(reinterpret_cast<const float*>(p.depth) + sp);
        const float ds = __fmul_rn(d, p.scale32);
        valid = (ds > p.min32) && (ds < p.max32) && isfinite(ds);
        zj = (double)ds;
      } else if (DEPTH_MODE == 1) {
        const float d = __ldg(reinterpret_cast<const float*>(p.depth) + sp);
        const double ds = __dmul_rn((double)d, p.scale);
        valid = (ds > p.min_d) && (ds < p.max_d) && isfinite(ds);
        zj = ds;
      } else {
        const double d = __ldg(reinterpret_cast<const double*>(p.depth) + sp);
        const double ds = __dmul_rn(d, p.scale);
        valid = (ds > p.min_d) && (ds < p.max_d) && isfinite(ds);
        zj = ds;
      }
      if (valid && p.conf != nullptr) valid = __ldg(p.conf + sp) != 0;
      // advance (u,v) by 256 sampled pixels for the next j
      uu += BP_THREADS;
      while (uu >= p.Ws) {
        uu -= p.Ws;
        ++vv;
      }
    }
    z[j] = zj;
    const unsigned b = __ballot_sync(0xffffffffu, valid);
    bal[j] = valid ? b : 0u;  // 0 marks "this thread's pixel is invalid"
    if (lane == 0) s_wcount[j * BP_WARPS + warp] = __popc(b);
  }
  __syncthreads();

  // ---- block scan of the 64 (j, warp) counts + decoupled look-back --------
  if (warp == 0) {
    const int a = s_wcount[2 * lane];
    const int b = s_wcount[2 * lane + 1];
    int sum = a + b;
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    const int exc = inc - sum;
    s_woff[2 * lane] = exc;
    s_woff[2 * lane + 1] = exc + a;
    const int total = __shfl_sync(0xffffffffu, inc, 31);

    long long base = 0;
    if (tile == 0) {
      if (lane == 0) st_state(p.tile_state, FLAG_INC | (unsigned long long)total);
    } else {
      if (lane == 0)
        st_state(p.tile_state + tile, FLAG_AGG | (unsigned long long)total);
      int look = tile - 1;  // lanes inspect tiles look, look-1, ...
      while (true) {
        const int t = look - lane;
        unsigned long long st = FLAG_INC;  // out-of-range lanes: "prefix 0"
        if (t >= 0) {
          do {
            st = ld_volatile_u64(
                reinterpret_cast<const uint64_t*>(p.tile_state + t));
          } while ((st >> 62) == 0);
        }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (st >> 62) == 2);
        long long val = (long long)(st & VAL_MASK);
        if (inc_mask != 0) {
          const int first = __ffs(inc_mask) - 1;
          if (lane > first) val = 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
          val += __shfl_xor_sync(0xffffffffu, val, d);
        base += val;
        if (inc_mask != 0) break;
        look -= 32;
      }
      if (lane == 0)
        st_state(p.tile_state + tile,
                 FLAG_INC | (unsigned long long)(base + total));
    }
    if (lane == 0) {
      s_base = base;
      s_total = total;
      if (tile == p.num_tiles - 1) *p.out_n = base + total;
    }
  }
  __syncthreads();
  const long long base = s_base;
  const int total = s_total;
  if (total == 0) return;

  // ---- pass B: back-project valid pixels into the staging buffers ---------
  const int shift_w = (int)((base * 3) & 3);
  const int shift_b = (int)((base * 3) & 15);
  uint32_t crgb[BP_PPT];  // packed colours, kept for the RGB_F32 second phase
  uu = u;
  vv = v;
#pragma unroll
  for (int j = 0; j < BP_PPT; ++j) {
    const int pj = p0 + j * BP_THREADS;
    crgb[j] = 0;
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
              __fma_rn(p.Rt[2], zz, __fma_rn(p.Rt[1], y, __dmul_rn(p.Rt[0], x))),
              p.c[0]);
          wy = __dsub_rn(
              __fma_rn(p.Rt[5], zz, __fma_rn(p.Rt[4], y, __dmul_rn(p.Rt[3], x))),
              p.c[1]);
          wz = __dsub_rn(
              __fma_rn(p.Rt[8], zz, __fma_rn(p.Rt[7], y, __dmul_rn(p.Rt[6], x))),
              p.c[2]);
        }
        uint32_t* dst = s_xyz + shift_w + rank * 3;
        dst[0] = __float_as_uint(__double2float_rn(wx));
        dst[1] = __float_as_uint(__double2float_rn(wy));
        dst[2] = __float_as_uint(__double2float_rn(wz));
        if (p.has_color) {
          const long long sp = (long long)sv * p.W + su;
          const uint8_t* c = p.bgr + sp * 3;
          const uint32_t b = __ldg(c), g = __ldg(c + 1), r = __ldg(c + 2);
          if (RGB_F32) {
            crgb[j] = r | (g << 8) | (b << 16);
          } else {
            uint8_t* cd = s_rgb + shift_b + rank * 3;
            cd[0] = (uint8_t)r;  // BGR -> RGB (d2r:381-382)
            cd[1] = (uint8_t)g;
            cd[2] = (uint8_t)b;
          }
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

extern "C" int t3d_backproject(t3d_ctx* ctx, const void* depth, const uint8_t* bgr,
                               const uint8_t* conf_mask,
                               const t3d_backproject_params* q, float* out_xyz,
                               void* out_rgb, int64_t capacity, int64_t* out_n,
                               t3d_stream stream) {
  T3D_REQUIRE(ctx && q && out_n, "t3d_backproject: null ctx/params/out_n");
  T3D_REQUIRE(q->H >= 0 && q->W >= 0 && q->subsample >= 1,
              "t3d_backproject: bad H=%d W=%d subsample=%d", q->H, q->W,
              q->subsample);
  T3D_REQUIRE(!q->has_color || bgr, "t3d_backproject: has_color but bgr is NULL");
  cudaStream_t st = as_stream(stream);
  const int s = q->subsample;
  const int Hs = (q->H + s - 1) / s, Ws = (q->W + s - 1) / s;
  const int64_t P64 = (int64_t)Hs * Ws;
  T3D_REQUIRE(P64 < (1ll << 31) - BP_TILE, "t3d_backproject: frame too large");
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

  const ProjTable* tab = nullptr;
  int rc = get_proj_table(ctx, q->H, q->W, q->fx, q->fy, q->cx, q->cy, st, &tab);
  if (rc != T3D_OK) return rc;

  BPParams p;
  memset(&p, 0, sizeof(p));
  p.depth = depth;
  p.bgr = bgr;
  p.conf = conf_mask;
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
  if (q->has_pose) {
    // Rt = R^T ; c = R^T t with the dgemm k-ordered FMA chain (d2r:376)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) p.Rt[i * 3 + j] = q->R[j * 3 + i];
    for (int i = 0; i < 3; ++i)
      p.c[i] = fma(p.Rt[i * 3 + 2], q->t[2],
                   fma(p.Rt[i * 3 + 1], q->t[1], p.Rt[i * 3 + 0] * q->t[0]));
  }
  p.out_xyz = out_xyz;
  p.out_rgb = out_rgb;
  p.out_n = reinterpret_cast<long long*>(out_n);
  p.num_tiles = (p.P + BP_TILE - 1) / BP_TILE;

  const size_t state_bytes = sizeof(uint64_t) * ((size_t)p.num_tiles + 2);
  rc = ctx->scan_state.reserve(state_bytes);
  if (rc != T3D_OK) return rc;
  T3D_CUDA(cudaMemsetAsync(ctx->scan_state.p, 0, state_bytes, st));
  p.ticket = ctx->scan_state.as<unsigned>();
  p.tile_state = ctx->scan_state.as<unsigned long long>() + 1;

  const int mode = q->depth_is_f64 ? 2 : (q->scale_is_f64 ? 1 : 0);
  const dim3 grid(p.num_tiles), block(BP_THREADS);
#define BP_LAUNCH(M, F) backproject_kernel<M, F><<<grid, block, 0, st>>>(p)
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
