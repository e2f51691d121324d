// common.cuh — shared host/device plumbing of libt3d.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/t3d.h"

#define T3D_NUM_SMS_DEFAULT 148

void t3d_set_error(const char* fmt, ...);

#define T3D_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      t3d_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,              \
                    cudaGetErrorString(e__));                                 \
      return T3D_E_CUDA;                                                      \
    }                                                                         \
  } while (0)

#define T3D_REQUIRE(cond, ...)                                                \
  do {                                                                        \
    if (!(cond)) {                                                            \
      t3d_set_error(__VA_ARGS__);                                             \
      return T3D_E_INVALID;                                                   \
    }                                                                         \
  } while (0)

#define T3D_LAUNCH_CHECK()                                                    \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      t3d_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__,          \
                    cudaGetErrorString(e__));                                 \
      return T3D_E_CUDA;                                                      \
    }                                                                         \
  } while (0)

// Every entry point runs on ITS context's device, whatever the caller's current device is (one process may
// hold a context per GPU), and leaves the caller's current device as it found it.
struct T3DDeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit T3DDeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) {
      err = cudaSetDevice(dev);
      switched = err == cudaSuccess;
    }
  }
  ~T3DDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define T3D_ON_DEVICE(dev)            \
  T3DDeviceGuard t3d_guard__(dev);    \
  T3D_CUDA(t3d_guard__.err)

// A grow-only device buffer owned by the ctx (scratch that survives calls so
// the steady state does no cudaMalloc).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return T3D_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      // retry exact
      want = bytes;
      e = cudaMalloc(&p, want);
      if (e != cudaSuccess) {
        t3d_set_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
        p = nullptr;
        return T3D_E_CUDA;
      }
    }
    cap = want;
    return T3D_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

struct ProjTable {  // cached (u-cx)/fx, (v-cy)/fy tables — d2r:287-295
  double fx, fy, cx, cy;
  int H, W;
  double* xf = nullptr;  // W doubles
  double* yf = nullptr;  // H doubles
};

struct t3d_ctx {
  int device = 0;
  int num_sms = T3D_NUM_SMS_DEFAULT;
  int64_t launches = 0;
  // K1
  std::vector<ProjTable> proj;
  DevBuf scan_state;  // decoupled look-back tile descriptors + ticket
  // generic scratch (hash tables, grid cells, reductions)
  DevBuf scratch[12];
  // pinned host staging for small synchronous results
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  // cached CUDA graphs of launch-bound inner loops (key = every baked-in pointer/scalar)
  struct CachedGraph {
    std::string key;
    cudaGraphExec_t exec = nullptr;
  };
  std::vector<CachedGraph> graphs;
  // one-time per-DEVICE initialisation (constant-memory tables, cudaFuncSetAttribute opt-ins): kept here,
  // not in function-local statics — a process may hold one ctx per GPU and both are per device
  bool icp_offsets_ready = false;
  std::vector<const void*> func_attrs_done;
  bool func_attr_needed(const void* fn) {
    for (const void* f : func_attrs_done)
      if (f == fn) return false;
    func_attrs_done.push_back(fn);
    return true;
  }
};

// host-side phase timing, printed to stderr when T3D_TRACE is set (debug aid)
#include <chrono>
struct T3DTrace {
  bool on;
  const char* fn;
  std::chrono::steady_clock::time_point t;
  explicit T3DTrace(const char* f) : on(getenv("T3D_TRACE") != nullptr), fn(f), t(std::chrono::steady_clock::now()) {}
  void mark(const char* label) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[t3d trace] %s: %-18s %9.3f ms\n", fn, label,
            std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

inline cudaStream_t as_stream(t3d_stream s) {
  return reinterpret_cast<cudaStream_t>(s);
}

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ unsigned lane_id() {  // hardware lane (valid for 2-D blocks too)
  unsigned l;
  asm("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// streaming 128-bit load that does not pollute L1 (inputs are read once)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_volatile_s32(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// 64-bit mix (murmur3 finaliser) used by every open-addressing table here
__device__ __host__ __forceinline__ uint64_t mix64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

// pack three signed 21-bit integers (biased) into one 63-bit key; axis x is
// the most significant so that ascending key order == (x,y,z) lexicographic
#define T3D_KEY_BIAS (1 << 20)
#define T3D_KEY_EMPTY 0xFFFFFFFFFFFFFFFFULL
__device__ __host__ __forceinline__ uint64_t pack_key(int x, int y, int z) {
  return ((uint64_t)(uint32_t)(x + T3D_KEY_BIAS) << 42) |
         ((uint64_t)(uint32_t)(y + T3D_KEY_BIAS) << 21) |
         ((uint64_t)(uint32_t)(z + T3D_KEY_BIAS));
}
__device__ __host__ __forceinline__ void unpack_key(uint64_t k, int& x, int& y,
                                                    int& z) {
  x = (int)((k >> 42) & 0x1FFFFF) - T3D_KEY_BIAS;
  y = (int)((k >> 21) & 0x1FFFFF) - T3D_KEY_BIAS;
  z = (int)(k & 0x1FFFFF) - T3D_KEY_BIAS;
}
__device__ __host__ __forceinline__ bool key_in_range(int x, int y, int z) {
  return x >= -T3D_KEY_BIAS && x < T3D_KEY_BIAS && y >= -T3D_KEY_BIAS &&
         y < T3D_KEY_BIAS && z >= -T3D_KEY_BIAS && z < T3D_KEY_BIAS;
}

#endif  // __CUDACC__
