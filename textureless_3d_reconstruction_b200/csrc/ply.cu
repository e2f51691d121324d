// ply.cu — K9: PLY serialisation (host side of the boundary).
//
// Replaces
//   DepthToReconstructionPipeline.save_reconstruction  depth_to_reconstruction.py:673-703
//   DepthEnhancedReconstruction._save_pointcloud        depth_enhanced_reconstruction.py:1283-1311
//   PointCloudGenerator.save_ply                        depth_processor.py:424-440
// Two layouts:
//   T3D_PLY_O3D_BINARY  what Open3D's write_point_cloud emits:
//                       binary_little_endian, double xyz [double normals], uchar rgb
//   T3D_PLY_REF_ASCII   the reference's own fallback (d2r:690-701): ascii,
//                       `property float`, one "x y z r g b" line per point where
//                       x,y,z are Python's repr of the value promoted to double.
#include <charconv>
#include <functional>
#include <thread>

#include "common.cuh"

namespace {

// Python float.__repr__: shortest round-trip digits; fixed notation when
// -4 <= exp10 < 16, otherwise d.ddde[+-]XX.
int py_repr(double v, char* out) {
  if (v != v) { memcpy(out, "nan", 3); return 3; }
  if (v == INFINITY) { memcpy(out, "inf", 3); return 3; }
  if (v == -INFINITY) { memcpy(out, "-inf", 4); return 4; }
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
  const int len = (int)(r.ptr - buf);
  // parse [-]d[.ddd]e[+-]XX
  int p = 0, o = 0;
  if (buf[p] == '-') { out[o++] = '-'; ++p; }
  char digits[32];
  int nd = 0;
  while (p < len && buf[p] != 'e') {
    if (buf[p] != '.') digits[nd++] = buf[p];
    ++p;
  }
  ++p;  // 'e'
  int esign = 1;
  if (buf[p] == '-') { esign = -1; ++p; } else if (buf[p] == '+') { ++p; }
  int e10 = 0;
  while (p < len) e10 = e10 * 10 + (buf[p++] - '0');
  e10 *= esign;
  if (nd == 1 && digits[0] == '0') e10 = 0;
  if (e10 >= -4 && e10 < 16) {
    if (e10 >= 0) {
      for (int i = 0; i <= e10; ++i) out[o++] = i < nd ? digits[i] : '0';
      out[o++] = '.';
      if (nd > e10 + 1) for (int i = e10 + 1; i < nd; ++i) out[o++] = digits[i];
      else out[o++] = '0';
    } else {
      out[o++] = '0';
      out[o++] = '.';
      for (int i = 0; i < -e10 - 1; ++i) out[o++] = '0';
      for (int i = 0; i < nd; ++i) out[o++] = digits[i];
    }
  } else {
    out[o++] = digits[0];
    if (nd > 1) {
      out[o++] = '.';
      for (int i = 1; i < nd; ++i) out[o++] = digits[i];
    }
    out[o++] = 'e';
    out[o++] = e10 < 0 ? '-' : '+';
    int a = e10 < 0 ? -e10 : e10;
    char eb[8];
    int ne = 0;
    while (a > 0) { eb[ne++] = (char)('0' + a % 10); a /= 10; }
    while (ne < 2) eb[ne++] = '0';
    while (ne > 0) out[o++] = eb[--ne];
  }
  return o;
}

int put_uint(unsigned v, char* out) {
  char b[4];
  int n = 0;
  do { b[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  for (int i = 0; i < n; ++i) out[i] = b[n - 1 - i];
  return n;
}

}  // namespace

extern "C" int t3d_write_ply_h(const char* path, const void* xyz_h, int xyz_is_f64,
                               const uint8_t* rgb_h, const void* nrm_h, int64_t n, int layout) {
  T3D_REQUIRE(path && (n == 0 || xyz_h) && n >= 0, "t3d_write_ply_h: null argument");
  T3D_REQUIRE(layout == T3D_PLY_O3D_BINARY || layout == T3D_PLY_REF_ASCII,
              "t3d_write_ply_h: unknown layout %d", layout);
  FILE* f = fopen(path, "wb");
  if (!f) {
    t3d_set_error("t3d_write_ply_h: cannot open %s", path);
    return T3D_E_IO;
  }
  const float* xf = reinterpret_cast<const float*>(xyz_h);
  const double* xd = reinterpret_cast<const double*>(xyz_h);
  std::vector<char> buf;
  buf.reserve(1 << 22);
  bool ok = true;
  auto flush = [&]() {
    if (!buf.empty()) ok = ok && fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    buf.clear();
  };
  if (layout == T3D_PLY_REF_ASCII) {
    ok = ok && fprintf(f,
                 "ply\nformat ascii 1.0\nelement vertex %lld\nproperty float x\nproperty float y\n"
                 "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n"
                 "end_header\n", (long long)n) > 0;
    // shortest-repr formatting is the whole cost (~0.2 us per number): super-chunks of 2 M points are
    // formatted by all host threads into per-thread buffers and written in order
    auto format_range = [&](int64_t a, int64_t b, std::vector<char>& out) {
      out.clear();
      out.reserve((size_t)(b - a) * 64);
      char line[160];
      for (int64_t i = a; i < b; ++i) {
        int o = 0;
        for (int c = 0; c < 3; ++c) {
          const double v = xyz_is_f64 ? xd[i * 3 + c] : (double)xf[i * 3 + c];
          o += py_repr(v, line + o);
          line[o++] = ' ';
        }
        for (int c = 0; c < 3; ++c) {
          o += put_uint(rgb_h ? rgb_h[i * 3 + c] : 0u, line + o);
          line[o++] = c == 2 ? '\n' : ' ';
        }
        out.insert(out.end(), line, line + o);
      }
    };
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 1 : (nt > 32 ? 32 : nt);
    if (n < 200000) nt = 1;
    const int64_t super = 2000000;
    std::vector<std::vector<char>> parts(nt);
    for (int64_t s0 = 0; s0 < n && ok; s0 += super) {
      const int64_t s1 = s0 + super < n ? s0 + super : n;
      const int64_t per = (s1 - s0 + nt - 1) / nt;
      std::vector<std::thread> th;
      for (unsigned t = 1; t < nt; ++t) {
        const int64_t a = s0 + (int64_t)t * per, b = a + per < s1 ? a + per : s1;
        if (a < b) th.emplace_back(format_range, a, b, std::ref(parts[t]));
        else parts[t].clear();
      }
      format_range(s0, s0 + per < s1 ? s0 + per : s1, parts[0]);
      for (auto& t : th) t.join();
      for (unsigned t = 0; t < nt && ok; ++t)
        if (!parts[t].empty()) ok = fwrite(parts[t].data(), 1, parts[t].size(), f) == parts[t].size();
    }
  } else {
    ok = ok && fprintf(f, "ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\n"
                          "element vertex %lld\nproperty double x\nproperty double y\n"
                          "property double z\n", (long long)n) > 0;
    if (nrm_h)
      ok = ok && fprintf(f, "property double nx\nproperty double ny\nproperty double nz\n") > 0;
    if (rgb_h)
      ok = ok && fprintf(f, "property uchar red\nproperty uchar green\nproperty uchar blue\n") > 0;
    ok = ok && fprintf(f, "end_header\n") > 0;
    const float* nf = reinterpret_cast<const float*>(nrm_h);
    const double* nd = reinterpret_cast<const double*>(nrm_h);
    for (int64_t i = 0; i < n; ++i) {
      double rec[6];
      int k = 0;
      for (int c = 0; c < 3; ++c) rec[k++] = xyz_is_f64 ? xd[i * 3 + c] : (double)xf[i * 3 + c];
      if (nrm_h)
        for (int c = 0; c < 3; ++c) rec[k++] = xyz_is_f64 ? nd[i * 3 + c] : (double)nf[i * 3 + c];
      const char* rb = reinterpret_cast<const char*>(rec);
      buf.insert(buf.end(), rb, rb + k * 8);
      if (rgb_h) buf.insert(buf.end(), rgb_h + i * 3, rgb_h + i * 3 + 3);
      if (buf.size() > (1 << 22) - 256) flush();
    }
    flush();
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) {
    t3d_set_error("t3d_write_ply_h: write to %s failed", path);
    return T3D_E_IO;
  }
  return T3D_OK;
}

extern "C" int t3d_write_ply_mesh_h(const char* path, const float* xyz_h, const float* nrm_h,
                                    const uint8_t* rgb_h, int64_t nv, const int32_t* tri_h, int64_t nt) {
  T3D_REQUIRE(path && nv >= 0 && nt >= 0 && (nv == 0 || xyz_h) && (nt == 0 || tri_h),
              "t3d_write_ply_mesh_h: null argument");
  FILE* f = fopen(path, "wb");
  if (!f) {
    t3d_set_error("t3d_write_ply_mesh_h: cannot open %s", path);
    return T3D_E_IO;
  }
  bool ok = fprintf(f, "ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\n"
                       "element vertex %lld\nproperty double x\nproperty double y\nproperty double z\n",
                    (long long)nv) > 0;
  if (nrm_h) ok = ok && fprintf(f, "property double nx\nproperty double ny\nproperty double nz\n") > 0;
  if (rgb_h) ok = ok && fprintf(f, "property uchar red\nproperty uchar green\nproperty uchar blue\n") > 0;
  ok = ok && fprintf(f, "element face %lld\nproperty list uchar uint vertex_indices\nend_header\n",
                     (long long)nt) > 0;
  std::vector<char> buf;
  buf.reserve(1 << 22);
  auto flush = [&]() {
    if (!buf.empty()) ok = ok && fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    buf.clear();
  };
  for (int64_t i = 0; i < nv; ++i) {
    double rec[6];
    int k = 0;
    for (int c = 0; c < 3; ++c) rec[k++] = (double)xyz_h[i * 3 + c];
    if (nrm_h) for (int c = 0; c < 3; ++c) rec[k++] = (double)nrm_h[i * 3 + c];
    const char* rb = reinterpret_cast<const char*>(rec);
    buf.insert(buf.end(), rb, rb + k * 8);
    if (rgb_h) buf.insert(buf.end(), rgb_h + i * 3, rgb_h + i * 3 + 3);
    if (buf.size() > (1 << 22) - 256) flush();
  }
  for (int64_t i = 0; i < nt; ++i) {
    char rec[13];
    rec[0] = 3;
    memcpy(rec + 1, tri_h + i * 3, 12);
    buf.insert(buf.end(), rec, rec + 13);
    if (buf.size() > (1 << 22) - 256) flush();
  }
  flush();
  ok = (fclose(f) == 0) && ok;
  if (!ok) {
    t3d_set_error("t3d_write_ply_mesh_h: write to %s failed", path);
    return T3D_E_IO;
  }
  return T3D_OK;
}
