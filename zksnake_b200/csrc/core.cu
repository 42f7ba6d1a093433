// core.cu -- context, error reporting, device scratch arena and raw memory entry points of libzkb200.so.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include "../../include/zkb200.h"
#include "zkb_internal.h"

namespace zkb {

struct Ctx {
  bool ready = false;
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  char* arena = nullptr;
  size_t arena_cap = 0, arena_off = 0;
  unsigned long long launches = 0;
};
static Ctx g_ctx;
static thread_local std::string g_err;

int set_error(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(int cuda_err, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", cuda_err, cudaGetErrorString((cudaError_t)cuda_err), file,
           line, what);
  return set_error(ZKB_ERR_CUDA, buf);
}
void* ctx_stream() { return (void*)g_ctx.stream; }
bool ctx_ready() { return g_ctx.ready; }
void count_launch(int n) { g_ctx.launches += n; }
unsigned long long launches() { return g_ctx.launches; }

int scratch_reserve(size_t bytes) {
  if (bytes <= g_ctx.arena_cap) return ZKB_OK;
  ZKB_CUDA(cudaStreamSynchronize(g_ctx.stream));
  if (g_ctx.arena) ZKB_CUDA(cudaFree(g_ctx.arena));
  g_ctx.arena = nullptr;
  g_ctx.arena_cap = 0;
  size_t cap = bytes + (bytes >> 3) + (1 << 20);
  ZKB_CUDA(cudaMalloc((void**)&g_ctx.arena, cap));
  g_ctx.arena_cap = cap;
  g_ctx.arena_off = 0;
  return ZKB_OK;
}
void scratch_reset() { g_ctx.arena_off = 0; }
void* scratch_take(size_t bytes) {
  size_t off = (g_ctx.arena_off + 255) & ~(size_t)255;
  if (off + bytes > g_ctx.arena_cap) return nullptr;
  g_ctx.arena_off = off + bytes;
  return g_ctx.arena + off;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int zkb_init(int device) {
  if (g_ctx.ready && g_ctx.device == device) return ZKB_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return set_error(ZKB_ERR_CUDA, "zkb_init: no CUDA device visible -- libzkb200 has no CPU fallback");
  if (device < 0 || device >= n) return set_error(ZKB_ERR_ARG, "zkb_init: bad device index");
  ZKB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ZKB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    char buf[256];
    snprintf(buf, sizeof(buf), "zkb_init: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
             prop.minor);
    return set_error(ZKB_ERR_CUDA, buf);
  }
  if (g_ctx.ready) zkb_shutdown();
  ZKB_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
  ZKB_CUDA(cudaEventCreate(&g_ctx.ev0));
  ZKB_CUDA(cudaEventCreate(&g_ctx.ev1));
  g_ctx.device = device;
  g_ctx.ready = true;
  return ZKB_OK;
}

void zkb_shutdown(void) {
  if (!g_ctx.ready) return;
  cudaStreamSynchronize(g_ctx.stream);
  if (g_ctx.arena) cudaFree(g_ctx.arena);
  g_ctx.arena = nullptr;
  g_ctx.arena_cap = g_ctx.arena_off = 0;
  cudaEventDestroy(g_ctx.ev0);
  cudaEventDestroy(g_ctx.ev1);
  cudaStreamDestroy(g_ctx.stream);
  g_ctx.stream = nullptr;
  g_ctx.ready = false;
}

const char* zkb_last_error(void) { return g_err.c_str(); }
void* zkb_stream(void) { return (void*)g_ctx.stream; }
unsigned long long zkb_launch_count(void) { return g_ctx.launches; }

#define NEED_INIT() \
  if (!g_ctx.ready) return set_error(ZKB_ERR_NOINIT, "zkb_init has not been called (no CUDA context; no CPU fallback)")

int zkb_sync(void) {
  NEED_INIT();
  ZKB_CUDA(cudaStreamSynchronize(g_ctx.stream));
  return ZKB_OK;
}
int zkb_dev_alloc(size_t bytes, void** out) {
  NEED_INIT();
  ZKB_CUDA(cudaMalloc(out, bytes ? bytes : 256));
  return ZKB_OK;
}
int zkb_dev_free(void* p) {
  NEED_INIT();
  ZKB_CUDA(cudaStreamSynchronize(g_ctx.stream));
  ZKB_CUDA(cudaFree(p));
  return ZKB_OK;
}
int zkb_host_alloc(size_t bytes, void** out) {
  NEED_INIT();
  ZKB_CUDA(cudaHostAlloc(out, bytes ? bytes : 256, cudaHostAllocDefault));
  return ZKB_OK;
}
int zkb_host_free(void* p) {
  NEED_INIT();
  ZKB_CUDA(cudaFreeHost(p));
  return ZKB_OK;
}
int zkb_h2d(void* dst, const void* src, size_t bytes) {
  NEED_INIT();
  ZKB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
  ZKB_CUDA(cudaStreamSynchronize(g_ctx.stream));
  return ZKB_OK;
}
int zkb_d2h(void* dst, const void* src, size_t bytes) {
  NEED_INIT();
  ZKB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
  ZKB_CUDA(cudaStreamSynchronize(g_ctx.stream));
  return ZKB_OK;
}
int zkb_d2d(void* dst, const void* src, size_t bytes) {
  NEED_INIT();
  ZKB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_ctx.stream));
  return ZKB_OK;
}
int zkb_memset(void* dst, int value, size_t bytes) {
  NEED_INIT();
  ZKB_CUDA(cudaMemsetAsync(dst, value, bytes, g_ctx.stream));
  return ZKB_OK;
}
int zkb_timer_start(void) {
  NEED_INIT();
  ZKB_CUDA(cudaEventRecord(g_ctx.ev0, g_ctx.stream));
  return ZKB_OK;
}
int zkb_timer_stop(float* ms) {
  NEED_INIT();
  ZKB_CUDA(cudaEventRecord(g_ctx.ev1, g_ctx.stream));
  ZKB_CUDA(cudaEventSynchronize(g_ctx.ev1));
  ZKB_CUDA(cudaEventElapsedTime(ms, g_ctx.ev0, g_ctx.ev1));
  return ZKB_OK;
}

}  // extern "C"
