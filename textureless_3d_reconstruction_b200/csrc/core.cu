// core.cu — ctx lifetime, error string, launch accounting.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void t3d_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* t3d_last_error(void) { return g_err; }
extern "C" int t3d_version(void) { return T3D_VERSION; }

extern "C" t3d_ctx* t3d_create(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    t3d_set_error("t3d_create: no CUDA device (%s); libt3d has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
    return nullptr;
  }
  if (device < 0 || device >= count) {
    t3d_set_error("t3d_create: device %d out of range [0,%d)", device, count);
    return nullptr;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    t3d_set_error("t3d_create: cudaSetDevice(%d) -> %s", device,
                  cudaGetErrorString(e));
    return nullptr;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    t3d_set_error("t3d_create: cudaGetDeviceProperties -> %s",
                  cudaGetErrorString(e));
    return nullptr;
  }
  if (prop.major != 10) {
    t3d_set_error(
        "t3d_create: device is sm_%d%d; libt3d is built for sm_100a (B200) only",
        prop.major, prop.minor);
    return nullptr;
  }
  t3d_ctx* ctx = new t3d_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->pinned_bytes = 1 << 16;
  e = cudaMallocHost(&ctx->pinned, ctx->pinned_bytes);
  if (e != cudaSuccess) {
    t3d_set_error("t3d_create: cudaMallocHost -> %s", cudaGetErrorString(e));
    delete ctx;
    return nullptr;
  }
  return ctx;
}

extern "C" void t3d_destroy(t3d_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (ProjTable& t : ctx->proj) {
    cudaFree(t.xf);
    cudaFree(t.yf);
  }
  ctx->scan_state.release();
  for (auto& g : ctx->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (DevBuf& b : ctx->scratch) b.release();
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  delete ctx;
}

extern "C" int64_t t3d_launch_count(const t3d_ctx* ctx) {
  return ctx ? ctx->launches : 0;
}

// ---------------------------------------------------------------------------
// Peer memory (multi-GPU block routing over NVLink/NVSwitch): buffers allocated here can be
// opened by the other ranks of the box through CUDA IPC handles, so an export kernel can
// store its records straight into the owner's memory.
// ---------------------------------------------------------------------------
extern "C" int t3d_ipc_alloc(t3d_ctx* ctx, size_t bytes, void** dev_ptr, uint8_t* handle_out64) {
  T3D_REQUIRE(ctx && dev_ptr && handle_out64 && bytes > 0, "t3d_ipc_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  T3D_ON_DEVICE(ctx->device);
  void* p = nullptr;
  T3D_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    t3d_set_error("t3d_ipc_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    return T3D_E_CUDA;
  }
  T3D_CUDA(cudaMemset(p, 0, bytes));
  memcpy(handle_out64, &h, 64);
  *dev_ptr = p;
  return T3D_OK;
}

extern "C" int t3d_ipc_open(t3d_ctx* ctx, const uint8_t* handle64, void** dev_ptr) {
  T3D_REQUIRE(ctx && handle64 && dev_ptr, "t3d_ipc_open: bad argument");
  T3D_ON_DEVICE(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  T3D_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return T3D_OK;
}

extern "C" int t3d_ipc_close(t3d_ctx* ctx, void* dev_ptr) {
  T3D_REQUIRE(ctx && dev_ptr, "t3d_ipc_close: bad argument");
  T3D_ON_DEVICE(ctx->device);
  T3D_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return T3D_OK;
}

extern "C" int t3d_ipc_free(t3d_ctx* ctx, void* dev_ptr) {
  T3D_REQUIRE(ctx && dev_ptr, "t3d_ipc_free: bad argument");
  T3D_ON_DEVICE(ctx->device);
  T3D_CUDA(cudaFree(dev_ptr));
  return T3D_OK;
}

// Device-to-device copy on the copy engine (also between peers mapped with t3d_ipc_open: NVLink, no SM work)
extern "C" int t3d_memcpy_async(void* dst, const void* src, size_t bytes, t3d_stream stream) {
  T3D_REQUIRE((dst && src) || bytes == 0, "t3d_memcpy_async: null pointer");
  if (bytes) T3D_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, as_stream(stream)));
  return T3D_OK;
}

extern "C" int t3d_memset_async(void* dev_ptr, int value, size_t bytes, t3d_stream stream) {
  T3D_REQUIRE(dev_ptr || bytes == 0, "t3d_memset_async: null pointer");
  if (bytes) T3D_CUDA(cudaMemsetAsync(dev_ptr, value, bytes, as_stream(stream)));
  return T3D_OK;
}
