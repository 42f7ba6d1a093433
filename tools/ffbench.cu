// ffbench.cu -- integer-pipe microbenchmarks that set the MSM / NTT compute rooflines on a B200:
//   1. IMAD.WIDE.U32 issue rate with loop-variant operands (nothing ptxas can hoist),
//   2. the carry-chained (mad.lo.cc / madc.hi.cc) rows the Montgomery multiplier is made of,
//   3. Montgomery multiplications per second (Fq BN254, Fq BLS12-381), inline vs out-of-line, 1/2/4 independent streams,
//   4. XYZZ mixed additions per second in isolation (no gather), inline vs out-of-line multiplier.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. -I../zksnake_b200/csrc -o build/ffbench ffbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define ZKB_EXPERIMENTAL_WIDE   // pulls tools/ff_wide.cuh into ff.cuh (build with -I tools -I zksnake_b200/csrc)
#include "ec.cuh"

using namespace zkb;

struct FqBN254Inl : FqBN254 { static constexpr bool NOINLINE_MUL = false; };
struct FqBLS381Inl : FqBLS381 { static constexpr bool NOINLINE_MUL = false; };
struct FqBN254Split : FqBN254 { static constexpr bool NOINLINE_MUL = false; static constexpr bool SPLIT_MUL = true; };
struct FqBN254SplitCall : FqBN254 { static constexpr bool SPLIT_MUL = true; };
struct FqBLS381Split : FqBLS381 { static constexpr bool NOINLINE_MUL = false; static constexpr bool SPLIT_MUL = true; };

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

// ---- 1. IMAD.WIDE with a loop-variant multiplicand ------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) k_imad_wide(uint32_t* out, uint32_t b, int iters) {
  unsigned long long y[CH];
#pragma unroll
  for (int k = 0; k < CH; k++) y[k] = threadIdx.x * 0x9e3779b97f4a7c15ull + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < CH; k++) {
        uint32_t lo = (uint32_t)y[k];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y[k]) : "r"(lo), "r"(b));
      }
    }
  }
  unsigned long long r = 0;
#pragma unroll
  for (int k = 0; k < CH; k++) r ^= y[k];
  if (r == 0x1234567812345678ull) out[0] = 1;
}

// plain 32-bit IMAD (lo) with a true dependency
template <int CH>
__global__ void __launch_bounds__(256) k_imad_lo(uint32_t* out, uint32_t b, int iters) {
  uint32_t y[CH];
#pragma unroll
  for (int k = 0; k < CH; k++) y[k] = threadIdx.x * 0x9e3779b9u + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < CH; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[k]) : "r"(b), "r"(b));
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < CH; k++) r ^= y[k];
  if (r == 0x12345678u) out[0] = 1;
}


// ---- 1b. FP64 pipe: DFMA.RZ chains (the Emmart-style 52-bit limb product uses 2 DFMA + 1 DADD per 52x52 product) -------
template <int CH>
__global__ void __launch_bounds__(256) k_dfma(double* out, double b, int iters) {
  double y[CH];
#pragma unroll
  for (int k = 0; k < CH; k++) y[k] = threadIdx.x * 1.25 + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < CH; k++) y[k] = __fma_rz(y[k], b, y[k]);
    }
  }
  double r = 0;
#pragma unroll
  for (int k = 0; k < CH; k++) r += y[k];
  if (r == 1234.5) out[0] = r;
}
// the real product step: hi = fma_rz(a,b,C1); lo = fma_rz(a,b,C2-hi); acc_hi += bits(hi); acc_lo += bits(lo)
template <int CH>
__global__ void __launch_bounds__(256) k_dprod(unsigned long long* out, double b0, int iters) {
  double a[CH];
  unsigned long long acc_hi[CH], acc_lo[CH];
#pragma unroll
  for (int k = 0; k < CH; k++) { a[k] = (double)(threadIdx.x * 977 + k * 13 + 1); acc_hi[k] = k; acc_lo[k] = k * 3; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int k = 0; k < CH; k++) {
        double b = b0 + (double)(u * 8);
        double hi = __fma_rz(a[k], b, 0x1p104);
        double sub = (0x1p104 + 0x1p52) - hi;
        double lo = __fma_rz(a[k], b, sub);
        acc_hi[k] += (unsigned long long)__double_as_longlong(hi);
        acc_lo[k] += (unsigned long long)__double_as_longlong(lo);
        a[k] = __longlong_as_double((__double_as_longlong(lo) & 0x000fffffffffffffll) | 0x4330000000000000ll) - 0x1p52 + 1.0;
      }
    }
  }
  unsigned long long r = 0;
#pragma unroll
  for (int k = 0; k < CH; k++) r ^= acc_hi[k] ^ acc_lo[k];
  if (r == 0x1234567812345678ull) out[0] = r;
}
// IMAD.HI.U32 with a true dependency
template <int CH>
__global__ void __launch_bounds__(256) k_imad_hi(uint32_t* out, uint32_t b, int iters) {
  uint32_t y[CH];
#pragma unroll
  for (int k = 0; k < CH; k++) y[k] = threadIdx.x * 0x9e3779b9u + k + 0x80000000u;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < CH; k++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(y[k]) : "r"(b), "r"(b));
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < CH; k++) r ^= y[k];
  if (r == 0x12345678u) out[0] = 1;
}
// IADD3 + IADD3.X pairs (64-bit three-input adds) -- the alu pipe
template <int CH>
__global__ void __launch_bounds__(256) k_iadd64(unsigned long long* out, unsigned long long b, int iters) {
  unsigned long long y[CH];
#pragma unroll
  for (int k = 0; k < CH; k++) y[k] = threadIdx.x * 0x9e3779b97f4a7c15ull + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < CH; k++) y[k] = y[k] + (y[(k + 1) % CH] ^ b) + b;
    }
  }
  unsigned long long r = 0;
#pragma unroll
  for (int k = 0; k < CH; k++) r ^= y[k];
  if (r == 0x1234567812345678ull) out[0] = r;
}

// ---- 2. carry-chained row: acc[0..8] += a[0..7] * b  (4 wide mads on even limbs, as in mont_mul) -------------------------
template <int CH>
__global__ void __launch_bounds__(256) k_row_cc(uint32_t* out, uint32_t b0, int iters) {
  uint32_t acc[CH][8], a[CH][8];
#pragma unroll
  for (int k = 0; k < CH; k++)
#pragma unroll
    for (int j = 0; j < 8; j++) { acc[k][j] = threadIdx.x + j + k; a[k][j] = threadIdx.x * 7 + j * 3 + k; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < CH; k++) {
      uint32_t b = acc[k][7] ^ b0;
      acc[k][0] = mad_lo_cc(a[k][0], b, acc[k][0]);
      acc[k][1] = madc_hi_cc(a[k][0], b, acc[k][1]);
      acc[k][2] = madc_lo_cc(a[k][2], b, acc[k][2]);
      acc[k][3] = madc_hi_cc(a[k][2], b, acc[k][3]);
      acc[k][4] = madc_lo_cc(a[k][4], b, acc[k][4]);
      acc[k][5] = madc_hi_cc(a[k][4], b, acc[k][5]);
      acc[k][6] = madc_lo_cc(a[k][6], b, acc[k][6]);
      acc[k][7] = madc_hi(a[k][6], b, acc[k][7]);
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < CH; k++)
#pragma unroll
    for (int j = 0; j < 8; j++) r ^= acc[k][j];
  if (r == 0x12345678u) out[0] = 1;
}

// ---- 3. Montgomery multiplications ------------------------------------------------------------------------------------
template <class F, int ST>
__global__ void __launch_bounds__(128) k_mul(F* io, int iters) {
  F x[ST], y[ST];
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int s = 0; s < ST; s++) {
    x[s] = io[(t * ST + s) & 1023];
    y[s] = io[(t * ST + s + 7) & 1023];
  }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int s = 0; s < ST; s++) x[s] = x[s] * y[s];
#pragma unroll
    for (int s = 0; s < ST; s++) y[s] = y[s] * x[s];
  }
  F r = x[0];
#pragma unroll
  for (int s = 0; s < ST; s++) r = r + x[s] + y[s];
  if (r.v[0] == 0x12345678u && r.v[1] == 0x9abcdef0u) io[t & 1023] = r;
}

// ---- 4. mixed additions -----------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) k_madd(const Affine<F>* pts, XYZZ<F>* out, int iters) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int i = 0; i < iters; i++) {
    Affine<F> p = pts[(t * 31 + i * 17) & 1023];
    madd(acc, p, (i & 1) != 0);
  }
  if (acc.ZZ == acc.X && acc.Y == acc.ZZZ) out[t & 1023] = acc;
}

#define P52_0 154029749239111LL
#define P52_1 2558044347618242LL
#define P52_2 423691504025962LL
#define P52_3 2817616741948264LL
#define P52_4 53207371014449LL
#define PINV52 571208714576777LL

// ---- 5. PROTOTYPE: Montgomery product on the FP64 pipe (52-bit limbs, R = 2^260; Fq BN254) --------------------------------
// Every 52x52-bit partial product is 2 DFMA.RZ + 1 DADD (hi = fma_rz(a,b,2^104), lo = fma_rz(a,b,2^104 + 2^52 - hi)) whose bit
// patterns are accumulated as int64 -- no IMAD at all.  B200 issues DFMA at twice the IMAD.WIDE rate.
struct D52 { double v[5]; };
__device__ __constant__ double D52_P[5] = {(double)P52_0, (double)P52_1, (double)P52_2, (double)P52_3, (double)P52_4};
__device__ __forceinline__ D52 d52_mul(const D52& a, const D52& b) {
  const double C1 = 0x1p104, C2 = 0x1p104 + 0x1p52, PINV = (double)PINV52;
  const long long LO = 0x4330000000000000LL, HI = 0x4670000000000000LL, MASK = 0x000fffffffffffffLL;
  const double P0 = (double)P52_0, P1 = (double)P52_1, P2 = (double)P52_2, P3 = (double)P52_3, P4 = (double)P52_4;
  const double P[5] = {P0, P1, P2, P3, P4};
  long long c[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 5; i++) {
#pragma unroll
    for (int j = 0; j < 5; j++) {
      double hi = __fma_rz(a.v[j], b.v[i], C1);
      double lo = __fma_rz(a.v[j], b.v[i], C2 - hi);
      c[j] += __double_as_longlong(lo) - LO;
      c[j + 1] += __double_as_longlong(hi) - HI;
    }
    double vd = __longlong_as_double((c[0] & MASK) | LO) - 0x1p52;
    double qh = __fma_rz(vd, PINV, C1);
    double qd = __fma_rz(vd, PINV, C2 - qh) - 0x1p52;   // (v * p') mod 2^52
#pragma unroll
    for (int j = 0; j < 5; j++) {
      double hi = __fma_rz(qd, P[j], C1);
      double lo = __fma_rz(qd, P[j], C2 - hi);
      c[j] += __double_as_longlong(lo) - LO;
      c[j + 1] += __double_as_longlong(hi) - HI;
    }
    long long carry = c[0] >> 52;     // c[0] is a multiple of 2^52 now
    c[0] = c[1] + carry; c[1] = c[2]; c[2] = c[3]; c[3] = c[4]; c[4] = c[5]; c[5] = 0;
  }
  // normalise, conditional subtraction of p, back to doubles
#pragma unroll
  for (int k = 0; k < 4; k++) { c[k + 1] += c[k] >> 52; c[k] &= MASK; }
  const long long PL[5] = {P52_0, P52_1, P52_2, P52_3, P52_4};
  long long d[5], borrow = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) { long long t = c[k] - PL[k] - borrow; borrow = (t >> 63) & 1; d[k] = t & MASK; }
  D52 r;
#pragma unroll
  for (int k = 0; k < 5; k++) r.v[k] = __longlong_as_double((borrow ? c[k] : d[k]) | LO) - 0x1p52;
  return r;
}
__device__ __forceinline__ D52 to52(const uint32_t* v) {
  unsigned long long w0 = v[0] | ((unsigned long long)v[1] << 32), w1 = v[2] | ((unsigned long long)v[3] << 32);
  unsigned long long w2 = v[4] | ((unsigned long long)v[5] << 32), w3 = v[6] | ((unsigned long long)v[7] << 32);
  const unsigned long long M = 0x000fffffffffffffull;
  D52 r;
  r.v[0] = (double)(long long)(w0 & M);
  r.v[1] = (double)(long long)(((w0 >> 52) | (w1 << 12)) & M);
  r.v[2] = (double)(long long)(((w1 >> 40) | (w2 << 24)) & M);
  r.v[3] = (double)(long long)(((w2 >> 28) | (w3 << 36)) & M);
  r.v[4] = (double)(long long)(w3 >> 16);
  return r;
}
__device__ __forceinline__ void from52(const D52& a, uint32_t* v) {
  unsigned long long l0 = (unsigned long long)a.v[0], l1 = (unsigned long long)a.v[1], l2 = (unsigned long long)a.v[2];
  unsigned long long l3 = (unsigned long long)a.v[3], l4 = (unsigned long long)a.v[4];
  unsigned long long w0 = l0 | (l1 << 52), w1 = (l1 >> 12) | (l2 << 40), w2 = (l2 >> 24) | (l3 << 28), w3 = (l3 >> 36) | (l4 << 16);
  v[0] = (uint32_t)w0; v[1] = (uint32_t)(w0 >> 32); v[2] = (uint32_t)w1; v[3] = (uint32_t)(w1 >> 32);
  v[4] = (uint32_t)w2; v[5] = (uint32_t)(w2 >> 32); v[6] = (uint32_t)w3; v[7] = (uint32_t)(w3 >> 32);
}
// check: d52_mul(a, b) = a b 2^-260 mod p  ==  mont_mul(mont_mul(a, b), 2^252)   (integer multiplier: x y 2^-256)
__global__ void k_d52_check(const Fp<FqBN254>* in, int n, int* bad) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  Fp<FqBN254> a = in[t], b = in[(t * 7 + 3) % n];
  Fp<FqBN254> k = Fp<FqBN254>::zero();
  k.v[7] = 0x10000000u;   // 2^252
  Fp<FqBN254> want = mont_mul(mont_mul(a, b), k);
  D52 r = d52_mul(to52(a.v), to52(b.v));
  Fp<FqBN254> got;
  from52(r, got.v);
  if (!(got == want)) atomicAdd(bad, 1);
}
template <int ST>
__global__ void __launch_bounds__(128) k_d52_mul(const Fp<FqBN254>* io, uint32_t* out, int iters) {
  D52 x[ST], y[ST];
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int s = 0; s < ST; s++) { x[s] = to52(io[(t * ST + s) & 1023].v); y[s] = to52(io[(t * ST + s + 7) & 1023].v); }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int s = 0; s < ST; s++) x[s] = d52_mul(x[s], y[s]);
#pragma unroll
    for (int s = 0; s < ST; s++) y[s] = d52_mul(y[s], x[s]);
  }
  double r = 0;
#pragma unroll
  for (int s = 0; s < ST; s++) r += x[s].v[0] + y[s].v[4];
  if (r == 1234.5) out[0] = 1;
}

template <class K, class... A>
static float run(const char* name, double ops_per_thread, int blocks, int threads, K kern, A... args) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    kern<<<blocks, threads>>>(args...);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double total = ops_per_thread * blocks * threads;
  printf("%-44s %9.3f ms  %10.2f Gop/s  (%6.2f lane-ops/clk/SM @1.965GHz,148SM)\n", name, best, total / best / 1e6,
         total / (best * 1e-3) / (148 * 1.965e9));
  fflush(stdout);
  return best;
}

template <class F>
static void fill_field(F* d, int n) {
  // arbitrary values below the modulus: small multiples in Montgomery form via host arithmetic
  F* h = (F*)malloc(n * sizeof(F));
  F x = F::one();
  F g = F::one() + F::one() + F::one();
  for (int i = 0; i < n; i++) { x = x * g + F::one(); h[i] = x; }
  CK(cudaMemcpy(d, h, n * sizeof(F), cudaMemcpyHostToDevice));
  free(h);
}

template <class F>
static Affine<F>* make_points(int n) {
  // multiples of a generator-ish point: take (x, y) on y^2 = x^3 + b by scalar-multiplying the known generator on the host
  Affine<F>* h = (Affine<F>*)malloc(n * sizeof(Affine<F>));
  Affine<F> g;
  g.x = F::one();
  g.y = F::one() + F::one();   // (1, 2) is the BN254 G1 generator; for BLS this is not on the curve but the adder's cost is the same
  XYZZ<F> acc = XYZZ<F>::from_affine(g);
  for (int i = 0; i < n; i++) {
    acc = dbl(acc);
    madd(acc, g);
    h[i] = to_affine(acc);
  }
  Affine<F>* d;
  CK(cudaMalloc(&d, n * sizeof(Affine<F>)));
  CK(cudaMemcpy(d, h, n * sizeof(Affine<F>), cudaMemcpyHostToDevice));
  free(h);
  return d;
}

template <class F>
static void bench_field(const char* tag, int blocks_per_sm) {
  F* d;
  CK(cudaMalloc(&d, 1024 * sizeof(F)));
  fill_field(d, 1024);
  char name[128];
  const int iters = 2000;
  const int blocks = 148 * blocks_per_sm;
  snprintf(name, sizeof(name), "mont_mul %s streams=1 (%d blk/SM x128)", tag, blocks_per_sm);
  run(name, 2.0 * iters * 1, blocks, 128, k_mul<F, 1>, d, iters);
  snprintf(name, sizeof(name), "mont_mul %s streams=2 (%d blk/SM x128)", tag, blocks_per_sm);
  run(name, 2.0 * iters * 2, blocks, 128, k_mul<F, 2>, d, iters);
  snprintf(name, sizeof(name), "mont_mul %s streams=4 (%d blk/SM x128)", tag, blocks_per_sm);
  run(name, 2.0 * iters * 4, blocks, 128, k_mul<F, 4>, d, iters);
  CK(cudaFree(d));
}

template <class F>
static void bench_madd(const char* tag, int blocks_per_sm) {
  Affine<F>* pts = make_points<F>(1024);
  XYZZ<F>* out;
  CK(cudaMalloc(&out, 1024 * sizeof(XYZZ<F>)));
  char name[128];
  const int iters = 512;
  snprintf(name, sizeof(name), "madd XYZZ %s (%d blk/SM x128)", tag, blocks_per_sm);
  run(name, (double)iters, 148 * blocks_per_sm, 128, k_madd<F>, (const Affine<F>*)pts, out, iters);
  CK(cudaFree(out));
  CK(cudaFree(pts));
}

int main(int argc, char** argv) {
  uint32_t* d;
  CK(cudaMalloc(&d, 4096));
  const int it = 2048;
  if (argc > 1 && atoi(argv[1]) == 4) {
    Fp<FqBN254>* dv;
    CK(cudaMalloc(&dv, 65536 * sizeof(Fp<FqBN254>)));
    fill_field(dv, 65536);
    int* bad;
    CK(cudaMalloc(&bad, 4));
    CK(cudaMemset(bad, 0, 4));
    k_d52_check<<<256, 256>>>(dv, 65536, bad);
    int hb = -1;
    CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
    printf("d52_mul correctness vs integer multiplier on 65536 pairs: %d mismatches\n", hb);
    const int iters = 2000;
    for (int bps : {2, 4, 8}) {
      char name[128];
      snprintf(name, sizeof(name), "d52 mont_mul FqBN254 streams=1 (%d blk/SM x128)", bps);
      run(name, 2.0 * iters, 148 * bps, 128, k_d52_mul<1>, (const Fp<FqBN254>*)dv, d, iters);
      snprintf(name, sizeof(name), "d52 mont_mul FqBN254 streams=2 (%d blk/SM x128)", bps);
      run(name, 4.0 * iters, 148 * bps, 128, k_d52_mul<2>, (const Fp<FqBN254>*)dv, d, iters);
    }
    bench_field<Fp<FqBN254Inl>>("FqBN254 integer fused inline", 4);
    return 0;
  }
#ifndef D52_ONLY
  if (argc > 1 && atoi(argv[1]) == 3) {
    for (int bps : {2, 4}) {
      bench_field<Fp<FqBN254Inl>>("FqBN254 fused inline", bps);
      bench_field<Fp<FqBN254Split>>("FqBN254 split inline", bps);
      bench_field<Fp<FqBN254SplitCall>>("FqBN254 split call", bps);
    }
    bench_field<Fp<FqBLS381Inl>>("FqBLS381 fused inline", 4);
    bench_field<Fp<FqBLS381Split>>("FqBLS381 split inline", 4);
    for (int bps : {3, 4}) {
      bench_madd<Fp<FqBN254Inl>>("G1 BN254 fused inline", bps);
      bench_madd<Fp<FqBN254Split>>("G1 BN254 split inline", bps);
      bench_madd<Fp<FqBN254SplitCall>>("G1 BN254 split call", bps);
    }
    bench_madd<Fp<FqBLS381Inl>>("G1 BLS381 fused inline", 3);
    bench_madd<Fp<FqBLS381Split>>("G1 BLS381 split inline", 3);
    bench_madd<Fp2<FqBN254>>("G2 BN254 fused call", 2);
    bench_madd<Fp2<FqBN254SplitCall>>("G2 BN254 split call", 2);
    return 0;
  }
  if (argc > 1 && atoi(argv[1]) == 2) {
    double* dd = (double*)d;
    unsigned long long* du = (unsigned long long*)d;
    run("DFMA.RZ dep chains=4 (8 blk/SM x256)", 8.0 * 4 * it, 148 * 8, 256, k_dfma<4>, dd, 1.0000001, it);
    run("DFMA.RZ dep chains=8 (8 blk/SM x256)", 8.0 * 8 * it, 148 * 8, 256, k_dfma<8>, dd, 1.0000001, it);
    run("dprod (2DFMA+DADD+2 i64 add+conv) ch=4 (8 blk/SM)", 4.0 * 4 * it, 148 * 8, 256, k_dprod<4>, du, 4503599627370001.0, it);
    run("dprod (2DFMA+DADD+2 i64 add+conv) ch=8 (4 blk/SM)", 4.0 * 8 * it, 148 * 4, 256, k_dprod<8>, du, 4503599627370001.0, it);
    run("IMAD.HI dep chains=8 (8 blk/SM x256)", 8.0 * 8 * it, 148 * 8, 256, k_imad_hi<8>, d, 0x7f4a7c15u, it);
    run("i64 add3 (IADD3+IADD3.X) chains=8 (8 blk/SM x256)", 8.0 * 8 * it, 148 * 8, 256, k_iadd64<8>, du, 0x7f4a7c15ull, it);
    return 0;
  }
  run("IMAD.lo  dep  chains=4  (8 blk/SM x256)", 8.0 * 4 * it, 148 * 8, 256, k_imad_lo<4>, d, 0x7f4a7c15u, it);
  run("IMAD.lo  dep  chains=8  (8 blk/SM x256)", 8.0 * 8 * it, 148 * 8, 256, k_imad_lo<8>, d, 0x7f4a7c15u, it);
  run("IMAD.WIDE dep chains=4  (8 blk/SM x256)", 8.0 * 4 * it, 148 * 8, 256, k_imad_wide<4>, d, 0x7f4a7c15u, it);
  run("IMAD.WIDE dep chains=8  (8 blk/SM x256)", 8.0 * 8 * it, 148 * 8, 256, k_imad_wide<8>, d, 0x7f4a7c15u, it);
  run("IMAD.WIDE dep chains=8  (2 blk/SM x256)", 8.0 * 8 * it, 148 * 2, 256, k_imad_wide<8>, d, 0x7f4a7c15u, it);
  // a row = 4 wide mads (8 lo/hi halves): count wide mads
  run("cc-row (4 wide mads) chains=1 (8 blk/SM x256)", 4.0 * 1 * it * 8, 148 * 8, 256, k_row_cc<1>, d, 0x7f4a7c15u, it * 8);
  run("cc-row (4 wide mads) chains=2 (8 blk/SM x256)", 4.0 * 2 * it * 8, 148 * 8, 256, k_row_cc<2>, d, 0x7f4a7c15u, it * 8);
  run("cc-row (4 wide mads) chains=4 (4 blk/SM x256)", 4.0 * 4 * it * 8, 148 * 4, 256, k_row_cc<4>, d, 0x7f4a7c15u, it * 8);
  for (int bps : {2, 4, 8}) {
    bench_field<Fp<FqBN254Inl>>("FqBN254 inline", bps);
    bench_field<Fp<FqBN254>>("FqBN254 call", bps);
  }
  bench_field<Fp<FqBLS381Inl>>("FqBLS381 inline", 4);
  bench_field<Fp<FqBLS381>>("FqBLS381 call", 4);
  for (int bps : {2, 3, 4, 6}) {
    bench_madd<Fp<FqBN254Inl>>("G1 BN254 inline", bps);
    bench_madd<Fp<FqBN254>>("G1 BN254 call", bps);
  }
  bench_madd<Fp<FqBLS381Inl>>("G1 BLS381 inline", 3);
  bench_madd<Fp<FqBLS381>>("G1 BLS381 call", 3);
  bench_madd<Fp2<FqBN254>>("G2 BN254 call", 2);
#endif
  return 0;
}
