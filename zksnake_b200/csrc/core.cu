// core.cu -- context, error reporting, device scratch arena and raw memory entry points of libzkb200.so.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <thread>
#include <vector>
#include "../../include/zkb200.h"
#include "zkb_internal.h"

namespace zkb {

struct Ctx {
  bool ready = false;
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t side[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
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
static std::mutex g_entry_mutex;
static std::thread::id g_entry_owner;
static int g_entry_depth = 0;
EntryGuard::EntryGuard() {
  std::lock_guard<std::mutex> lk(g_entry_mutex);
  ok = g_entry_depth == 0 || g_entry_owner == std::this_thread::get_id();
  if (ok) {
    g_entry_owner = std::this_thread::get_id();
    g_entry_depth++;
  }
}
EntryGuard::~EntryGuard() {
  if (!ok) return;
  std::lock_guard<std::mutex> lk(g_entry_mutex);
  g_entry_depth--;
}
void* ctx_stream() { return (void*)g_ctx.stream; }
// Every launch helper of the library takes its stream from ctx_stream(); a stage that has to run beside the library stream (the
// quotient's last step on the exchange stream of a multi-GPU proof, api.cu) swaps it for the duration of its enqueue calls and
// swaps it back.  Single host thread by contract (EntryGuard), so a plain global is enough.
void* ctx_stream_swap(void* s) {
  void* old = (void*)g_ctx.stream;
  g_ctx.stream = (cudaStream_t)s;
  return old;
}
void* ctx_side_stream(int i) {
  if (i < 0 || i >= 8) return nullptr;
  if (!g_ctx.side[i] && cudaStreamCreateWithFlags(&g_ctx.side[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  return (void*)g_ctx.side[i];
}
bool ctx_ready() { return g_ctx.ready; }
void count_launch(int n) { g_ctx.launches += n; }
static unsigned long long g_h2d_bytes = 0, g_d2h_bytes = 0;
void count_h2d(size_t bytes) { g_h2d_bytes += bytes; }
void count_d2h(size_t bytes) { g_d2h_bytes += bytes; }
unsigned long long launches() { return g_ctx.launches; }

// ---- per-kernel-family device timing (CUDA events on the library stream), off by default ----------------------------
struct ProfRec { int tag; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t g_prof_open[PROF_NTAGS];
static cudaEvent_t prof_event() {
  cudaEvent_t e = nullptr;
  if (!g_prof_pool.empty()) {
    e = g_prof_pool.back();
    g_prof_pool.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}
void prof_begin(int tag) {
  if (!g_prof_on) return;
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, g_ctx.stream);
  g_prof_open[tag] = e;
}
void prof_end(int tag) {
  if (!g_prof_on || !g_prof_open[tag]) return;
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, g_ctx.stream);
  g_prof.push_back({tag, g_prof_open[tag], e});
  g_prof_open[tag] = nullptr;
}
static void prof_clear() {
  for (auto& r : g_prof) {
    g_prof_pool.push_back(r.e0);
    g_prof_pool.push_back(r.e1);
  }
  g_prof.clear();
}

// A HELD arena (scratch_hold_begin .. scratch_hold_end) is one allocation epoch shared by several stages that would each start
// their own: their scratch_reserve only checks that the request still fits behind what has been taken, their scratch_reset does
// nothing.  Groth16 prove uses it so that digit sorts enqueued on a side stream keep their buffers while the transforms and the
// later MSM stages take theirs.
static bool g_arena_held = false;
int scratch_hold_begin(size_t total) {
  g_arena_held = false;
  int rc = scratch_reserve(total);
  if (rc) return rc;
  scratch_reset();
  g_arena_held = true;
  return ZKB_OK;
}
void scratch_hold_end() { g_arena_held = false; }
int scratch_reserve(size_t bytes) {
  if (g_arena_held) {
    if (((g_ctx.arena_off + 255) & ~(size_t)255) + bytes > g_ctx.arena_cap) return set_error(ZKB_ERR_CUDA, "held scratch arena too small");
    return ZKB_OK;
  }
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
void scratch_reset() {
  if (!g_arena_held) g_ctx.arena_off = 0;
}
void* scratch_take(size_t bytes) {
  size_t off = (g_ctx.arena_off + 255) & ~(size_t)255;
  if (off + bytes > g_ctx.arena_cap) return nullptr;
  g_ctx.arena_off = off + bytes;
  return g_ctx.arena + off;
}

}  // namespace zkb

using namespace zkb;

// ---- integer-pipe peak microbenchmark (MSM roofline denominator) ---------------------------------------------------------
// 8 independent multiply-add chains per thread, fully unrolled; the result is written so nothing is dead code.
template <int WIDE>
__global__ void __launch_bounds__(256) imad_peak_kernel(uint32_t* out, uint32_t a, uint32_t b, int iters) {
  uint32_t x0 = threadIdx.x, x1 = a ^ 1, x2 = a ^ 2, x3 = a ^ 3, x4 = a ^ 4, x5 = a ^ 5, x6 = a ^ 6, x7 = a ^ 7;
  unsigned long long y0 = threadIdx.x, y1 = 1, y2 = 2, y3 = 3, y4 = 4, y5 = 5, y6 = 6, y7 = 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (WIDE) {
        // the multiplicand is the running low word, so ptxas cannot hoist the product out of the loop (with loop-invariant
        // operands it does, and the "IMAD.WIDE rate" measured is that of a 64-bit add)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y0) : "r"((uint32_t)y0), "r"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y1) : "r"((uint32_t)y1), "r"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y2) : "r"((uint32_t)y2), "r"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y3) : "r"((uint32_t)y3), "r"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y4) : "r"((uint32_t)y4), "r"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y5) : "r"((uint32_t)y5), "r"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y6) : "r"((uint32_t)y6), "r"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y7) : "r"((uint32_t)y7), "r"(b));
      } else {
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x0) : "r"(a), "r"(b));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x1) : "r"(a), "r"(b));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x2) : "r"(a), "r"(b));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x3) : "r"(a), "r"(b));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x4) : "r"(a), "r"(b));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x5) : "r"(a), "r"(b));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x6) : "r"(a), "r"(b));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x7) : "r"(a), "r"(b));
      }
    }
  }
  uint32_t r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ (uint32_t)(y0 ^ y1 ^ y2 ^ y3 ^ y4 ^ y5 ^ y6 ^ y7) ^
               (uint32_t)((y0 ^ y1 ^ y2 ^ y3 ^ y4 ^ y5 ^ y6 ^ y7) >> 32);
  if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

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
  for (int i = 0; i < 8; i++) {
    if (g_ctx.side[i]) cudaStreamDestroy(g_ctx.side[i]);
    g_ctx.side[i] = nullptr;
  }
  cudaStreamDestroy(g_ctx.stream);
  g_ctx.stream = nullptr;
  g_ctx.ready = false;
}

const char* zkb_last_error(void) { return g_err.c_str(); }
void* zkb_stream(void) { return (void*)g_ctx.stream; }
unsigned long long zkb_launch_count(void) { return g_ctx.launches; }
void zkb_transfer_count(unsigned long long* h2d_bytes, unsigned long long* d2h_bytes) {
  if (h2d_bytes) *h2d_bytes = g_h2d_bytes;
  if (d2h_bytes) *d2h_bytes = g_d2h_bytes;
}

#define NEED_INIT() \
  if (!g_ctx.ready) return set_error(ZKB_ERR_NOINIT, "zkb_init has not been called (no CUDA context; no CPU fallback)"); \
  ZKB_ENTRY_GUARD()

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
  count_h2d(bytes);
  ZKB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
  ZKB_CUDA(cudaStreamSynchronize(g_ctx.stream));
  return ZKB_OK;
}
int zkb_h2d_async(void* dst, const void* src, size_t bytes) {
  NEED_INIT();
  count_h2d(bytes);
  ZKB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
  return ZKB_OK;
}
int zkb_d2h(void* dst, const void* src, size_t bytes) {
  NEED_INIT();
  count_d2h(bytes);
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
int zkb_prof_enable(int on) {
  NEED_INIT();
  ZKB_CUDA(cudaStreamSynchronize(g_ctx.stream));
  prof_clear();
  g_prof_on = on != 0;
  return ZKB_OK;
}
int zkb_prof_read(int tag, float* total_ms, unsigned long long* count) {
  NEED_INIT();
  if (tag < 0 || tag >= PROF_NTAGS) return set_error(ZKB_ERR_ARG, "unknown profile tag");
  ZKB_CUDA(cudaStreamSynchronize(g_ctx.stream));
  float sum = 0;
  unsigned long long n = 0;
  for (auto& r : g_prof) {
    if (r.tag != tag) continue;
    float ms = 0;
    ZKB_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    sum += ms;
    n++;
  }
  *total_ms = sum;
  *count = n;
  return ZKB_OK;
}
int zkb_imad_peak(int which, double* ops_per_s) {
  NEED_INIT();
  uint32_t* d = nullptr;
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  ZKB_CUDA(cudaMalloc((void**)&d, (size_t)blocks * threads * 4));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    ZKB_CUDA(cudaEventRecord(g_ctx.ev0, g_ctx.stream));
    if (which) imad_peak_kernel<1><<<blocks, threads, 0, g_ctx.stream>>>(d, 0x9e3779b9u, 0x7f4a7c15u, iters);
    else imad_peak_kernel<0><<<blocks, threads, 0, g_ctx.stream>>>(d, 0x9e3779b9u, 0x7f4a7c15u, iters);
    ZKB_CUDA(cudaEventRecord(g_ctx.ev1, g_ctx.stream));
    ZKB_CUDA(cudaEventSynchronize(g_ctx.ev1));
    float ms = 0;
    ZKB_CUDA(cudaEventElapsedTime(&ms, g_ctx.ev0, g_ctx.ev1));
    if (rep > 0 && ms < best) best = ms;
  }
  count_launch(5);
  ZKB_CUDA(cudaFree(d));
  *ops_per_s = (double)blocks * threads * iters * 64.0 / (best * 1e-3);
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
