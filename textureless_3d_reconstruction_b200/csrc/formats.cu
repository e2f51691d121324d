// formats.cu — the data formats either side of the hot path:
//   * 16-bit millimetre depth <-> float32 metres (writer depth_processor.py:919-921,
//     reader depth_to_reconstruction.py:85-90): decoded / encoded on the GPU so a depth
//     PNG costs 2 B/pixel over PCIe instead of 4;
//   * depth -> RGB-size bilinear resize (cv2.resize INTER_LINEAR,
//     depth_to_reconstruction.py:465-467);
//   * depth-scale estimation: gather depth at sparse feature points, ratio Z/d, sanity
//     gate, median (depth_to_reconstruction.py:297-326, depth_enhanced_reconstruction.py:652-697);
//   * PointCloud2 record packing x,y,z,rgb(b,g,r,0 as float bits)
//     (depth_processor.py:744-758, a per-point Python loop in the reference).
// All of them are streaming or tiny gathers: HBM-bound, no tensor cores.
#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace {

__global__ void u16_to_f32_kernel(const unsigned short* __restrict__ raw, long long n, float divisor,
                                  float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __fdiv_rn((float)raw[i], divisor);  // raw.astype(np.float32) / 1000.0
}

// (depth * 1000).astype(np.uint16): f32 product, C conversion as NumPy does it on x86-64
// (cvttss2si to int32 — INT_MIN for NaN / overflow — then truncation to 16 bits).
__global__ void f32_to_u16_kernel(const float* __restrict__ depth, long long n, float factor,
                                  unsigned short* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = __fmul_rn(depth[i], factor);
    int t;
    if (!(v > -2147483904.0f && v < 2147483648.0f)) t = (int)0x80000000;  // NaN, +-inf, overflow
    else t = (int)v;                                                        // truncation toward zero
    out[i] = (unsigned short)((unsigned)t & 0xFFFFu);
  }
}

// cv2.resize(src, (dw, dh), interpolation=INTER_LINEAR) for CV_32FC1: source coordinate
// (dx + 0.5) * scale - 0.5 in double, sx = floor, fractional weight rounded to float once
// (this is what OpenCV 4.13 produces: its outputs sit within 1e-6 of the exact bilinear
// value, which float-precision coordinates do not), clamped at both ends; horizontal pass
// S[sx]*(1-fx) + S[sx+1]*fx, vertical pass fma(h0, 1-fy, h1*fy).  <= 2 ulp from cv2.
__global__ void resize_linear_kernel(const float* __restrict__ src, int sh, int sw, float* __restrict__ dst,
                                     int dh, int dw, double scale_x, double scale_y) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  const int dy = blockIdx.y * blockDim.y + threadIdx.y;
  if (dx >= dw || dy >= dh) return;
  const double fxd = ((double)dx + 0.5) * scale_x - 0.5;
  int sx = (int)floor(fxd);
  float fx = (float)(fxd - (double)sx);
  if (sx < 0) { fx = 0.f; sx = 0; }
  int sx1 = sx + 1;
  if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; sx1 = sw - 1; }
  const double fyd = ((double)dy + 0.5) * scale_y - 0.5;
  const int sy = (int)floor(fyd);
  const float fy = (float)(fyd - (double)sy);
  int sy0 = min(max(sy, 0), sh - 1), sy1 = min(max(sy + 1, 0), sh - 1);
  const float a0 = __fsub_rn(1.f, fx), a1 = fx, b0 = __fsub_rn(1.f, fy), b1 = fy;
  const float* r0 = src + (long long)sy0 * sw;
  const float* r1 = src + (long long)sy1 * sw;
  const float h0 = __fadd_rn(__fmul_rn(__ldg(r0 + sx), a0), __fmul_rn(__ldg(r0 + sx1), a1));
  const float h1 = __fadd_rn(__fmul_rn(__ldg(r1 + sx), a0), __fmul_rn(__ldg(r1 + sx1), a1));
  dst[(long long)dy * dw + dx] = __fmaf_rn(h0, b0, __fmul_rn(h1, b1));
}

// ratio[i] = Z_i / depth[int(v_i), int(u_i)] (f64) or NaN when the sample is rejected
__global__ void scale_ratio_kernel(const float* __restrict__ depth, int H, int W,
                                   const double* __restrict__ pts3d, const double* __restrict__ pts2d,
                                   int n, int gate, double* __restrict__ ratio) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r = nan("");
  const double u = pts2d[2 * i], v = pts2d[2 * i + 1];
  if (isfinite(u) && isfinite(v) && fabs(u) < 2.0e9 && fabs(v) < 2.0e9) {
    const int x = (int)u, y = (int)v;  // Python int(): truncation toward zero
    if (x >= 0 && x < W && y >= 0 && y < H) {
      const double dn = (double)depth[(long long)y * W + x];
      const double ds = pts3d[3 * i + 2];
      if (dn > 0.0 && ds > 0.0) {
        const double s = __ddiv_rn(ds, dn);
        if (!gate || (0.001 < s && s < 1000.0)) r = s;
      }
    }
  }
  ratio[i] = r;
}

// record i = {x, y, z, float-bits(b | g<<8 | r<<16)} with r,g,b = (c*255).astype(uint8)
template <bool COLOR_F32>
__global__ void pack_pc2_kernel(const float* __restrict__ xyz, const void* __restrict__ colors,
                                long long n, float4* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    unsigned r, g, b;
    if (COLOR_F32) {
      const float* c = reinterpret_cast<const float*>(colors) + 3 * i;
      r = (unsigned)(int)__fmul_rn(c[0], 255.0f) & 255u;
      g = (unsigned)(int)__fmul_rn(c[1], 255.0f) & 255u;
      b = (unsigned)(int)__fmul_rn(c[2], 255.0f) & 255u;
    } else {
      const uint8_t* c = reinterpret_cast<const uint8_t*>(colors) + 3 * i;
      r = c[0]; g = c[1]; b = c[2];
    }
    out[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], __uint_as_float(b | (g << 8) | (r << 16)));
  }
}

int stream_grid(const t3d_ctx* ctx, long long n) {
  const long long want = (n + 255) / 256;
  const long long cap = (long long)ctx->num_sms * 16;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

extern "C" int t3d_depth_u16_to_f32(t3d_ctx* ctx, const uint16_t* raw, int64_t n, float divisor,
                                    float* out, t3d_stream stream) {
  T3D_REQUIRE(ctx && n >= 0 && (n == 0 || (raw && out)) && divisor != 0.f, "t3d_depth_u16_to_f32: bad argument");
  T3D_ON_DEVICE(ctx->device);
  if (n == 0) return T3D_OK;
  u16_to_f32_kernel<<<stream_grid(ctx, n), 256, 0, as_stream(stream)>>>(raw, n, divisor, out);
  T3D_LAUNCH_CHECK();
  ctx->launches++;
  return T3D_OK;
}

extern "C" int t3d_depth_f32_to_u16(t3d_ctx* ctx, const float* depth, int64_t n, float factor,
                                    uint16_t* out, t3d_stream stream) {
  T3D_REQUIRE(ctx && n >= 0 && (n == 0 || (depth && out)), "t3d_depth_f32_to_u16: bad argument");
  T3D_ON_DEVICE(ctx->device);
  if (n == 0) return T3D_OK;
  f32_to_u16_kernel<<<stream_grid(ctx, n), 256, 0, as_stream(stream)>>>(depth, n, factor, out);
  T3D_LAUNCH_CHECK();
  ctx->launches++;
  return T3D_OK;
}

extern "C" int t3d_resize_bilinear_f32(t3d_ctx* ctx, const float* src, int src_h, int src_w, float* dst,
                                       int dst_h, int dst_w, t3d_stream stream) {
  T3D_REQUIRE(ctx && src && dst && src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0,
              "t3d_resize_bilinear_f32: bad argument");
  T3D_ON_DEVICE(ctx->device);
  const dim3 block(32, 8), grid((dst_w + 31) / 32, (dst_h + 7) / 8);
  resize_linear_kernel<<<grid, block, 0, as_stream(stream)>>>(src, src_h, src_w, dst, dst_h, dst_w,
                                                               (double)src_w / dst_w, (double)src_h / dst_h);
  T3D_LAUNCH_CHECK();
  ctx->launches++;
  return T3D_OK;
}

extern "C" int t3d_estimate_scale(t3d_ctx* ctx, const float* depth, int H, int W, const double* pts3d_h,
                                  const double* pts2d_h, int64_t n, int gate, int min_input_points,
                                  double* out_scale_h, int64_t* out_samples_h, t3d_stream stream) {
  T3D_REQUIRE(ctx && depth && out_scale_h && H > 0 && W > 0 && n >= 0 && n < (1 << 24) &&
                  (n == 0 || (pts3d_h && pts2d_h)), "t3d_estimate_scale: bad argument");
  T3D_ON_DEVICE(ctx->device);
  *out_scale_h = 1.0;
  if (out_samples_h) *out_samples_h = 0;
  if (n < min_input_points || n == 0) return T3D_OK;  // der:673-674
  cudaStream_t st = as_stream(stream);
  int rc = ctx->scratch[0].reserve((size_t)n * 6 * sizeof(double));
  if (rc != T3D_OK) return rc;
  double* d3 = ctx->scratch[0].as<double>();
  double* d2 = d3 + 3 * n;
  double* dr = d2 + 2 * n;
  T3D_CUDA(cudaMemcpyAsync(d3, pts3d_h, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
  T3D_CUDA(cudaMemcpyAsync(d2, pts2d_h, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  scale_ratio_kernel<<<(int)((n + 127) / 128), 128, 0, st>>>(depth, H, W, d3, d2, (int)n, gate, dr);
  T3D_LAUNCH_CHECK();
  ctx->launches++;
  std::vector<double> r((size_t)n);
  T3D_CUDA(cudaMemcpyAsync(r.data(), dr, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  T3D_CUDA(cudaStreamSynchronize(st));
  std::vector<double> s;
  s.reserve((size_t)n);
  for (double v : r)
    if (v == v) s.push_back(v);
  if (out_samples_h) *out_samples_h = (int64_t)s.size();
  if (s.size() < 3) return T3D_OK;  // "Too few scale samples, using default scale=1.0" (d2r:317-319)
  std::sort(s.begin(), s.end());
  const size_t m = s.size();
  *out_scale_h = (m & 1) ? s[m / 2] : 0.5 * (s[m / 2 - 1] + s[m / 2]);  // np.median
  return T3D_OK;
}

extern "C" int t3d_pack_pointcloud2(t3d_ctx* ctx, const float* xyz, const void* colors, int colors_are_f32,
                                    int64_t n, float* out_records, t3d_stream stream) {
  T3D_REQUIRE(ctx && n >= 0 && (n == 0 || (xyz && colors && out_records)), "t3d_pack_pointcloud2: bad argument");
  T3D_ON_DEVICE(ctx->device);
  if (n == 0) return T3D_OK;
  cudaStream_t st = as_stream(stream);
  if (colors_are_f32)
    pack_pc2_kernel<true><<<stream_grid(ctx, n), 256, 0, st>>>(xyz, colors, n, reinterpret_cast<float4*>(out_records));
  else
    pack_pc2_kernel<false><<<stream_grid(ctx, n), 256, 0, st>>>(xyz, colors, n, reinterpret_cast<float4*>(out_records));
  T3D_LAUNCH_CHECK();
  ctx->launches++;
  return T3D_OK;
}
