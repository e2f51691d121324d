// synth.cu — device-side synthetic RGB-D scenes.  The
// reference ships no depth maps, poses or ground truth, and its depth network
// is out of scope, so every benchmark/property test runs on these analytic
// scenes.  textureless_3d_reconstruction_b200/synthetic.py is the NumPy twin
// used by the CPU-side tests (same formulas, same integer hash).
//
//  scene 0  "T1": tunnel along +z, radius r(theta,z) = 1.5 + 0.08 sin(2 pi z/1.7)
//           + 0.05 cos(3 theta + 0.9 z); camera i at (0.1 sin(0.05 i), 0, 0.25 i),
//           yaw 2deg*sin(0.03 i); grey albedo 128 +-2 LSB hash noise; optional
//           Gaussian depth noise.
//  scene 1  "S1": plane z = 3 + 0.25 sin(2 pi x/0.9) cos(2 pi y/1.3); camera
//           moves +x 0.05 m/frame, yaw 0.5deg/frame; 1% zero, 0.1% NaN, 0.1% +inf
//           pixels and a 64-px border at 60 m; RGB pattern from (u,v,i).
#include <math.h>

#include "common.cuh"

namespace {

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU;
  x ^= x >> 15; x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

struct SynthParams {
  int scene, frame, H, W;
  double fx, fy, cx, cy;
  uint32_t seed;
  float noise_sigma;
  double Rwc[9];  // camera -> world
  double o[3];    // camera position
};

__device__ __forceinline__ double tunnel_radius(double x, double y, double z) {
  const double th = atan2(y, x);
  return 1.5 + 0.08 * sin(6.283185307179586 * z / 1.7) + 0.05 * cos(3.0 * th + 0.9 * z);
}

__global__ void __launch_bounds__(256)
    synth_kernel(const __grid_constant__ SynthParams p, float* depth, uint8_t* bgr) {
  const long long P = (long long)p.H * p.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i / p.W), u = (int)(i - (long long)v * p.W);
    const double xc = ((double)u - p.cx) / p.fx, yc = ((double)v - p.cy) / p.fy;
    const double dx = p.Rwc[0] * xc + p.Rwc[1] * yc + p.Rwc[2];
    const double dy = p.Rwc[3] * xc + p.Rwc[4] * yc + p.Rwc[5];
    const double dz = p.Rwc[6] * xc + p.Rwc[7] * yc + p.Rwc[8];
    const uint32_t h0 = lowbias32((uint32_t)i ^ lowbias32((uint32_t)p.frame * 0x9E3779B9u ^ p.seed));
    float d = 0.0f;
    uint8_t B, G, R;
    if (p.scene == 0) {
      // bracket the hit between the cylinders rho = 1.37 and rho = 1.63
      const double a = dx * dx + dy * dy;
      const double b = p.o[0] * dx + p.o[1] * dy;
      const double c0 = p.o[0] * p.o[0] + p.o[1] * p.o[1];
      double t_hit = 0.0;
      if (a > 1e-12) {
        const double disc_in = b * b - a * (c0 - 1.37 * 1.37);
        const double disc_out = b * b - a * (c0 - 1.63 * 1.63);
        double t_lo = disc_in > 0.0 ? (-b + sqrt(disc_in)) / a : 0.0;
        const double t_hi = (-b + sqrt(fmax(disc_out, 0.0))) / a;
        if (t_lo < 0.0) t_lo = 0.0;
        if (t_hi > t_lo && t_lo < 12.0) {
          // first sign change over 8 uniform sub-intervals, then bisection
          double lo = t_lo, hi = t_hi;
          const double st = (t_hi - t_lo) / 8.0;
          for (int k = 1; k <= 8; ++k) {
            const double t = t_lo + st * k;
            const double x = p.o[0] + t * dx, y = p.o[1] + t * dy, z = p.o[2] + t * dz;
            if (sqrt(x * x + y * y) - tunnel_radius(x, y, z) >= 0.0) { hi = t; break; }
            lo = t;
          }
          for (int k = 0; k < 30; ++k) {
            const double t = 0.5 * (lo + hi);
            const double x = p.o[0] + t * dx, y = p.o[1] + t * dy, z = p.o[2] + t * dz;
            if (sqrt(x * x + y * y) - tunnel_radius(x, y, z) >= 0.0) hi = t; else lo = t;
          }
          t_hit = 0.5 * (lo + hi);
        }
      }
      if (t_hit > 0.0 && t_hit < 12.0) {
        double dn = t_hit;
        if (p.noise_sigma > 0.f) {
          const uint32_t h1 = lowbias32(h0 ^ 0x68bc21ebU), h2 = lowbias32(h0 ^ 0x02e5be93U);
          const double u1 = ((double)h1 + 1.0) / 4294967297.0, u2 = (double)h2 / 4294967296.0;
          dn += (double)p.noise_sigma * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        }
        d = (float)dn;
      }
      const int n0 = (int)(h0 % 5u) - 2, n1 = (int)((h0 >> 8) % 5u) - 2, n2 = (int)((h0 >> 16) % 5u) - 2;
      R = (uint8_t)(128 + n0); G = (uint8_t)(128 + n1); B = (uint8_t)(128 + n2);
    } else {
      double t = (3.0 - p.o[2]) / dz;
      for (int k = 0; k < 8; ++k) {
        const double x = p.o[0] + t * dx, y = p.o[1] + t * dy;
        const double s = 3.0 + 0.25 * sin(6.283185307179586 * x / 0.9) * cos(6.283185307179586 * y / 1.3);
        t = (s - p.o[2]) / dz;
      }
      d = (float)t;
      const uint32_t sel = h0 % 1000u;
      if (sel < 10u) d = 0.0f;
      else if (sel == 10u) d = __int_as_float(0x7fc00000);
      else if (sel == 11u) d = __int_as_float(0x7f800000);
      if (u < 64 || v < 64 || u >= p.W - 64 || v >= p.H - 64) d = 60.0f;
      R = (uint8_t)((u * 7 + v * 13 + p.frame * 29) & 255);
      G = (uint8_t)((u ^ v) & 255);
      B = (uint8_t)((u + v + p.frame) & 255);
    }
    depth[i] = d;
    if (bgr) { bgr[i * 3 + 0] = B; bgr[i * 3 + 1] = G; bgr[i * 3 + 2] = R; }
  }
}

}  // namespace

// pose of frame i (camera->world rotation + position), shared with Python via T_cw_h
static void scene_pose(int scene, int i, double* Rwc, double* o) {
  double yaw;
  if (scene == 0) {
    yaw = (2.0 * M_PI / 180.0) * sin(0.03 * i);
    o[0] = 0.1 * sin(0.05 * i); o[1] = 0.0; o[2] = 0.25 * i;
  } else {
    yaw = (0.5 * M_PI / 180.0) * i;
    o[0] = 0.05 * i; o[1] = 0.0; o[2] = 0.0;
  }
  const double c = cos(yaw), s = sin(yaw);
  Rwc[0] = c;  Rwc[1] = 0; Rwc[2] = s;
  Rwc[3] = 0;  Rwc[4] = 1; Rwc[5] = 0;
  Rwc[6] = -s; Rwc[7] = 0; Rwc[8] = c;
}

extern "C" int t3d_synth_frame(t3d_ctx* ctx, int scene, int frame_index, int H, int W, double fx,
                               double fy, double cx, double cy, uint64_t seed, float noise_sigma,
                               float* depth, uint8_t* bgr, double* T_cw_h, t3d_stream stream) {
  T3D_REQUIRE(ctx, "t3d_synth_frame: null ctx");
  T3D_ON_DEVICE(ctx->device);
  T3D_REQUIRE(scene == 0 || scene == 1, "t3d_synth_frame: unknown scene %d", scene);
  T3D_REQUIRE(H > 0 && W > 0, "t3d_synth_frame: bad size");
  SynthParams p;
  p.scene = scene; p.frame = frame_index; p.H = H; p.W = W;
  p.fx = fx; p.fy = fy; p.cx = cx; p.cy = cy;
  p.seed = (uint32_t)(seed ^ (seed >> 32));
  p.noise_sigma = noise_sigma;
  scene_pose(scene, frame_index, p.Rwc, p.o);
  if (T_cw_h) {  // world->camera: [R^T | -R^T o]
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) T_cw_h[r * 4 + c] = p.Rwc[c * 3 + r];
      T_cw_h[r * 4 + 3] = -(p.Rwc[0 * 3 + r] * p.o[0] + p.Rwc[1 * 3 + r] * p.o[1] + p.Rwc[2 * 3 + r] * p.o[2]);
    }
  }
  if (depth) {
    synth_kernel<<<ctx->num_sms * 8, 256, 0, as_stream(stream)>>>(p, depth, bgr);
    T3D_LAUNCH_CHECK();
    ctx->launches++;
  }
  return T3D_OK;
}
