// frvec.cu -- device-resident Fr vector primitives behind the PlonK prover glue: everything
// /root/reference/python/zksnake/plonk/protocol.py:270-466 does with Python loops or whole-vector list marshalling between the
// NTTs and MSMs -- scaled sums of polynomials, Z(omega X), batch inversion (utils.py:42-62), the grand-product accumulator
// (protocol.py:307-313), Horner evaluations (:385-390), division by X - z (:452,459; polynomial.rs:404-438) and by X^n - 1
// (polynomial.rs:466-489), strided / indexed gathers -- as kernels over canonical Fr vectors in HBM.
//
// Convention (as in ntt.cuh): vectors are canonical little-endian 8 x u32; scalars that multiply a vector are converted to
// Montgomery form once on the host, so mont_mul(canonical, montgomery) is canonical again and no conversion pass exists.
#include <cuda_runtime.h>
#include <string.h>
#include <vector>
#include "../../include/zkb200.h"
#include "ntt.cuh"
#include "zkb_internal.h"

namespace zkb {

static inline cudaStream_t S() { return (cudaStream_t)ctx_stream(); }

// out[i] = s * x[i] + y[i]   (x, y zero-extended beyond nx, ny; s in Montgomery form)
template <class F>
__global__ void axpy_kernel(unsigned long long n, F s, const F* __restrict__ x, unsigned long long nx, const F* __restrict__ y,
                            unsigned long long ny, F* __restrict__ out) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    F r = (i < nx) ? ntt_ld(x + i) * s : F::zero();
    if (i < ny) r = r + ntt_ld(y + i);
    ntt_st(out + i, r);
  }
}

// out[i] = x[i] * scale * base^i ; each thread walks CH consecutive exponents (one pow per thread, then one product per step)
#define ZKB_POW_CH 16
template <class F>
__global__ void mul_powers_kernel(unsigned long long n, F base, F scale, const F* __restrict__ x, F* __restrict__ out) {
  unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long lo = t * ZKB_POW_CH;
  if (lo >= n) return;
  F p = pow_u64(base, lo) * scale;   // Montgomery
  for (int k = 0; k < ZKB_POW_CH && lo + k < n; k++) {
    F v = x ? ntt_ld(x + lo + k) * p : from_mont(p);
    ntt_st(out + lo + k, v);
    p = p * base;
  }
}

// out[i] = 1 / x[i] (0 -> 0): Montgomery's trick over CH elements per thread, one Fermat inversion per thread
#define ZKB_INV_CH 8
template <class F>
__global__ void __launch_bounds__(128) inverse_kernel(unsigned long long n, const F* __restrict__ x, F* __restrict__ out) {
  unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long lo = t * ZKB_INV_CH;
  if (lo >= n) return;
  F v[ZKB_INV_CH], pre[ZKB_INV_CH];
  F acc = F::one();
#pragma unroll
  for (int k = 0; k < ZKB_INV_CH; k++) {
    v[k] = (lo + k < n) ? to_mont(ntt_ld(x + lo + k)) : F::one();
    pre[k] = acc;                       // product of the non-zero elements before k
    if (!v[k].is_zero()) acc = acc * v[k];
  }
  F inv_all = inv(acc);
#pragma unroll
  for (int k = ZKB_INV_CH - 1; k >= 0; k--) {
    F r = F::zero();
    if (!v[k].is_zero()) {
      r = inv_all * pre[k];
      inv_all = inv_all * v[k];
    }
    if (lo + k < n) ntt_st(out + lo + k, from_mont(r));
  }
}

// ---- scans --------------------------------------------------------------------------------------------------------------
// op 0: exclusive prefix PRODUCT, n + 1 outputs (out[0] = 1, out[i] = x[0] ... x[i-1]) -- the permutation accumulator.
// op 1: inclusive SUFFIX SUM, n outputs (out[i] = x[i] + x[i+1] + ... + x[n-1]) -- the linear-division recurrence.
// Three kernels: per-chunk totals (a thread owns CH consecutive elements, a CTA 256 threads), scan of the CTA totals by one
// CTA, final pass.  Products are carried in Montgomery form.
#define ZKB_SCAN_CH 8
template <class F, int OP>
struct ScanOp {
  static __device__ __forceinline__ F identity() { return OP == 0 ? F::one() : F::zero(); }
  static __device__ __forceinline__ F combine(const F& a, const F& b) { return OP == 0 ? a * b : a + b; }
  static __device__ __forceinline__ F load(const F* p) { return OP == 0 ? to_mont(ntt_ld(p)) : ntt_ld(p); }
  static __device__ __forceinline__ F store_form(const F& v) { return OP == 0 ? from_mont(v) : v; }
};

// element index of logical position q (suffix scans walk the vector backwards)
template <int OP>
__device__ __forceinline__ unsigned long long scan_index(unsigned long long q, unsigned long long n) { return OP == 0 ? q : n - 1 - q; }

template <class F, int OP>
__device__ __forceinline__ F block_scan_exclusive(F v, F* sh, F* total) {   // 256 threads
  typedef ScanOp<F, OP> O;
  const uint32_t t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
  for (uint32_t off = 1; off < 256; off <<= 1) {
    F a = (t >= off) ? sh[t - off] : O::identity();
    __syncthreads();
    if (t >= off) sh[t] = O::combine(a, sh[t]);
    __syncthreads();
  }
  *total = sh[255];
  F r = t ? sh[t - 1] : O::identity();
  __syncthreads();
  return r;
}

template <class F, int OP>
__global__ void __launch_bounds__(256) scan_partial_fr_kernel(unsigned long long n, const F* __restrict__ x, F* __restrict__ part) {
  typedef ScanOp<F, OP> O;
  __shared__ F sh[256];
  unsigned long long q0 = ((unsigned long long)blockIdx.x * 256 + threadIdx.x) * ZKB_SCAN_CH;
  F acc = O::identity();
  for (int k = 0; k < ZKB_SCAN_CH; k++)
    if (q0 + k < n) acc = O::combine(acc, O::load(x + scan_index<OP>(q0 + k, n)));
  F total;
  block_scan_exclusive<F, OP>(acc, sh, &total);
  if (threadIdx.x == 0) part[blockIdx.x] = total;
}
template <class F, int OP>
__global__ void __launch_bounds__(256) scan_spine_fr_kernel(uint32_t m, F* part) {   // exclusive scan of m CTA totals, in place
  typedef ScanOp<F, OP> O;
  __shared__ F sh[256];
  uint32_t per = (m + 255) / 256;
  uint32_t lo = threadIdx.x * per, hi = lo + per;
  if (hi > m) hi = m;
  F acc = O::identity();
  for (uint32_t i = lo; i < hi; i++) acc = O::combine(acc, part[i]);
  F total;
  F run = block_scan_exclusive<F, OP>(acc, sh, &total);
  for (uint32_t i = lo; i < hi; i++) {
    F v = part[i];
    part[i] = run;
    run = O::combine(run, v);
  }
}
template <class F, int OP>
__global__ void __launch_bounds__(256) scan_final_fr_kernel(unsigned long long n, const F* __restrict__ x, const F* __restrict__ part,
                                                            F* __restrict__ out) {
  typedef ScanOp<F, OP> O;
  __shared__ F sh[256];
  unsigned long long q0 = ((unsigned long long)blockIdx.x * 256 + threadIdx.x) * ZKB_SCAN_CH;
  F v[ZKB_SCAN_CH];
  F acc = O::identity();
  for (int k = 0; k < ZKB_SCAN_CH; k++) {
    v[k] = (q0 + k < n) ? O::load(x + scan_index<OP>(q0 + k, n)) : O::identity();
    acc = O::combine(acc, v[k]);
  }
  F total;
  F run = O::combine(part[blockIdx.x], block_scan_exclusive<F, OP>(acc, sh, &total));
  for (int k = 0; k < ZKB_SCAN_CH; k++) {
    if (q0 + k > n) break;
    if (OP == 0) {                       // exclusive: write the running value BEFORE element q (q == n: the grand total)
      ntt_st(out + q0 + k, O::store_form(run));
      if (q0 + k < n) run = O::combine(run, v[k]);
    } else if (q0 + k < n) {             // inclusive suffix
      run = O::combine(run, v[k]);
      ntt_st(out + scan_index<OP>(q0 + k, n), run);
    }
  }
}

// out[i] = src[offset + i * stride]
template <class F>
__global__ void gather_stride_kernel(unsigned long long n, const F* __restrict__ src, unsigned long long stride,
                                     unsigned long long offset, F* __restrict__ out) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i < n) ntt_st(out + i, ntt_ld(src + offset + i * stride));
}
// out[i] = src[idx[i]]
template <class F>
__global__ void gather_index_kernel(unsigned long long n, const F* __restrict__ src, const uint32_t* __restrict__ idx,
                                    F* __restrict__ out) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i < n) ntt_st(out + i, ntt_ld(src + idx[i]));
}

// partial[block] = sum over the block's chunk of c[i] * z^i : a thread runs Horner over CH coefficients (canonical accumulator,
// Montgomery z), scales by z^(first index) and the CTA adds up.
#define ZKB_EVAL_CH 32
template <class F>
__global__ void __launch_bounds__(256) eval_partial_kernel(unsigned long long n, const F* __restrict__ c, F z, F* __restrict__ partial) {
  __shared__ F sh[256];
  unsigned long long lo = ((unsigned long long)blockIdx.x * 256 + threadIdx.x) * ZKB_EVAL_CH;
  F acc = F::zero();
  if (lo < n) {
    unsigned long long hi = lo + ZKB_EVAL_CH < n ? lo + ZKB_EVAL_CH : n;
    for (unsigned long long i = hi; i-- > lo;) acc = acc * z + ntt_ld(c + i);
    acc = acc * pow_u64(z, lo);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) ntt_st(partial + blockIdx.x, sh[0]);
}

// p / (X^d - 1): q[j] = sum_{k >= 1} p[j + k d] for j < len - d; flag |= 1 when the remainder p[j] + q[j] (j < d) is non-zero
template <class F>
__global__ void div_vanishing_kernel(unsigned long long len, unsigned long long d, const F* __restrict__ p, F* __restrict__ q,
                                     int* flag) {
  unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (j >= len) return;
  F acc = F::zero();
  for (unsigned long long i = j + d; i < len; i += d) acc = acc + ntt_ld(p + i);
  if (j < len - d) ntt_st(q + j, acc);
  if (j < d) {
    F rem = ntt_ld(p + j) + acc;
    if (!rem.is_zero()) atomicOr(flag, 1);
  }
}

// *out = max(*out, i + 1) over the non-zero x[i]: the length of the vector with trailing zeros stripped (ark's
// DensePolynomial invariant, /root/reference/src/bn254/polynomial.rs:132-140 `coeffs()` returns the stripped vector)
template <class F>
__global__ void trim_kernel(unsigned long long n, const F* __restrict__ x, unsigned long long* out) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long best = 0;
  for (; i < n; i += stride)
    if (!ntt_ld(x + i).is_zero()) best = i + 1;
  // warp maximum first: one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
    if (other > best) best = other;
  }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(out, best);
}

// v[idx[k]] += / -= vals[k] for up to 8 entries passed BY VALUE (no staging copy, no synchronisation)
template <class F>
struct SparseArgs {
  unsigned long long idx[8];
  F val[8];
  uint32_t k;
};
template <class F>
__global__ void add_sparse_kernel(SparseArgs<F> a, int subtract, F* v) {
  uint32_t t = threadIdx.x;
  if (t >= a.k) return;
  // entries may repeat an index: serialise through thread 0 when they do (k <= 8)
  bool dup = false;
  for (uint32_t j = 0; j < a.k; j++)
    for (uint32_t i = 0; i < j; i++) dup |= a.idx[i] == a.idx[j];
  if (dup) {
    if (t != 0) return;
    for (uint32_t j = 0; j < a.k; j++) {
      F x = ntt_ld(v + a.idx[j]);
      ntt_st(v + a.idx[j], subtract ? x - a.val[j] : x + a.val[j]);
    }
    return;
  }
  F x = ntt_ld(v + a.idx[t]);
  ntt_st(v + a.idx[t], subtract ? x - a.val[t] : x + a.val[t]);
}


// ---- PlonK quotient on the coset g<w_q> (q = coset size, a multiple `per` = q / n of the gate count) --------------------------
// t[i] = ( gate + alpha (id z - sg z_w) + alpha^2 (z - 1) l1 ) / (x^n - 1)   at x = g w_q^i, with
//   gate = a ql + b qr + c qo + a b qm + qc + pi,
//   id = (a + beta x + gamma)(b + 2 beta x + gamma)(c + 3 beta x + gamma),  sg = (a + beta s1 + gamma)(b + beta s2 + gamma)(c + beta s3 + gamma),
//   z_w = z[(i + per) mod q]  (Z(omega x): omega = w_q^per).
// One pass over 14 input vectors instead of ~45 element-wise launches (python/zksnake/plonk/protocol.py:240-262, 284-300,
// 338-352 compute the same numerator through NTT round trips).  Inputs canonical; wire-like values are lifted to Montgomery
// form in registers so that every product lands in the form its consumer needs (mont_mul(canonical, montgomery) = canonical).
template <class F>
struct QuotientArgs {
  const F *a, *b, *c, *z, *pi, *ql, *qr, *qo, *qm, *qc, *s1, *s2, *s3, *l1;
  F g, wq, beta, beta_r2, gamma, alpha_c, alpha2, zh_inv[8];   // Montgomery except alpha_c (canonical); beta_r2 = beta R^2
  unsigned long long q;
  uint32_t per;
};
#define ZKB_QUOT_CH 8
template <class F>
__global__ void __launch_bounds__(128) plonk_quotient_kernel(QuotientArgs<F> A, F* __restrict__ t) {
  unsigned long long lo = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * ZKB_QUOT_CH;
  if (lo >= A.q) return;
  F x = pow_u64(A.wq, lo) * A.g;   // Montgomery(g w^lo)
  const F r2 = F::r2();
  for (int k = 0; k < ZKB_QUOT_CH; k++) {
    unsigned long long i = lo + k;
    if (i >= A.q) break;
    unsigned long long iw = i + A.per;
    if (iw >= A.q) iw -= A.q;
    F a = ntt_ld(A.a + i) * r2, b = ntt_ld(A.b + i) * r2, c = ntt_ld(A.c + i) * r2;
    F z = ntt_ld(A.z + i) * r2, zw = ntt_ld(A.z + iw) * r2;
    // gate constraint (canonical)
    F gate = a * ntt_ld(A.ql + i) + b * ntt_ld(A.qr + i) + c * ntt_ld(A.qo + i) + (a * b) * ntt_ld(A.qm + i) + ntt_ld(A.qc + i) +
             ntt_ld(A.pi + i);
    // permutation argument (Montgomery)
    F bx = A.beta * x;
    F bx2 = bx + bx;
    F id = ((a + bx + A.gamma) * (b + bx2 + A.gamma)) * (c + (bx2 + bx) + A.gamma);
    F sg = ((a + A.beta_r2 * ntt_ld(A.s1 + i) + A.gamma) * (b + A.beta_r2 * ntt_ld(A.s2 + i) + A.gamma)) *
           (c + A.beta_r2 * ntt_ld(A.s3 + i) + A.gamma);
    F perm = id * z - sg * zw;                              // Montgomery
    F l1t = (z - F::one()) * ntt_ld(A.l1 + i);              // canonical
    F num = gate + perm * A.alpha_c + l1t * A.alpha2;       // canonical
    ntt_st(t + i, num * A.zh_inv[i % A.per]);
    x = x * A.wq;
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
template <class F>
static F mont_from_words(const uint64_t w[4]) {
  F x;
  memcpy(x.v, w, 32);
  // R^2 * x / R = x R mod p: Montgomery form of (x mod p); the product tolerates an unreduced SECOND operand (see reduce_kernel)
  return F::r2() * x;
}

template <class F>
struct FrVecOps {
  static int axpy(size_t n, const uint64_t* s, const void* x, size_t nx, const void* y, size_t ny, void* out) {
    if (n == 0) return ZKB_OK;
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    prof_begin(PROF_VEC);
    axpy_kernel<F><<<blocks, 256, 0, S()>>>(n, mont_from_words<F>(s), (const F*)x, x ? nx : 0, (const F*)y, y ? ny : 0, (F*)out);
    prof_end(PROF_VEC);
    count_launch();
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
  }
  static int mul_powers(size_t n, const uint64_t* base, const uint64_t* scale, const void* x, void* out) {
    if (n == 0) return ZKB_OK;
    size_t threads = (n + ZKB_POW_CH - 1) / ZKB_POW_CH;
    prof_begin(PROF_VEC);
    mul_powers_kernel<F><<<(unsigned)((threads + 127) / 128), 128, 0, S()>>>(n, mont_from_words<F>(base), mont_from_words<F>(scale),
                                                                            (const F*)x, (F*)out);
    prof_end(PROF_VEC);
    count_launch();
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
  }
  static int inverse(size_t n, const void* x, void* out) {
    if (n == 0) return ZKB_OK;
    size_t threads = (n + ZKB_INV_CH - 1) / ZKB_INV_CH;
    prof_begin(PROF_VEC);
    inverse_kernel<F><<<(unsigned)((threads + 127) / 128), 128, 0, S()>>>(n, (const F*)x, (F*)out);
    prof_end(PROF_VEC);
    count_launch();
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
  }
  template <int OP>
  static int scan(size_t n, const void* x, void* out) {
    const size_t tile = 256 * ZKB_SCAN_CH;
    // the exclusive product has n + 1 outputs: the CTA grid covers positions 0..n
    size_t blocks = ((OP == 0 ? n + 1 : n) + tile - 1) / tile;
    if (blocks == 0) return ZKB_OK;
    int rc;
    if ((rc = scratch_reserve((blocks + 1) * sizeof(F) + 4096))) return rc;
    scratch_reset();
    F* part = (F*)scratch_take((blocks + 1) * sizeof(F));
    prof_begin(PROF_VEC);
    scan_partial_fr_kernel<F, OP><<<(unsigned)blocks, 256, 0, S()>>>(n, (const F*)x, part);
    scan_spine_fr_kernel<F, OP><<<1, 256, 0, S()>>>((uint32_t)blocks, part);
    scan_final_fr_kernel<F, OP><<<(unsigned)blocks, 256, 0, S()>>>(n, (const F*)x, part, (F*)out);
    prof_end(PROF_VEC);
    count_launch(3);
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
  }
  static int gather(size_t n, const void* src, size_t stride, size_t offset, const uint32_t* idx, void* out) {
    if (n == 0) return ZKB_OK;
    unsigned blocks = (unsigned)((n + 255) / 256);
    prof_begin(PROF_VEC);
    if (idx) gather_index_kernel<F><<<blocks, 256, 0, S()>>>(n, (const F*)src, idx, (F*)out);
    else gather_stride_kernel<F><<<blocks, 256, 0, S()>>>(n, (const F*)src, stride, offset, (F*)out);
    prof_end(PROF_VEC);
    count_launch();
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
  }
  static int eval(size_t n, const void* c, const uint64_t* point, uint64_t* out) {
    memset(out, 0, 32);
    if (n == 0) return ZKB_OK;
    const size_t tile = 256 * ZKB_EVAL_CH;
    size_t blocks = (n + tile - 1) / tile;
    int rc;
    if ((rc = scratch_reserve(blocks * sizeof(F) + 4096))) return rc;
    scratch_reset();
    F* part = (F*)scratch_take(blocks * sizeof(F));
    prof_begin(PROF_VEC);
    eval_partial_kernel<F><<<(unsigned)blocks, 256, 0, S()>>>(n, (const F*)c, mont_from_words<F>(point), part);
    prof_end(PROF_VEC);
    count_launch();
    ZKB_CUDA(cudaGetLastError());
    std::vector<F> host(blocks);
    ZKB_CUDA(ZKB_D2H(host.data(), part, blocks * sizeof(F)));
    ZKB_CUDA(cudaStreamSynchronize(S()));
    F acc = F::zero();
    for (size_t i = 0; i < blocks; i++) acc = acc + host[i];   // canonical values: plain modular additions
    memcpy(out, acc.v, 32);
    return ZKB_OK;
  }
  static int trim(size_t n, const void* x, size_t* len) {
    *len = 0;
    if (n == 0) return ZKB_OK;
    int rc;
    if ((rc = scratch_reserve(4096))) return rc;
    scratch_reset();
    unsigned long long* d_len = (unsigned long long*)scratch_take(256);
    ZKB_CUDA(cudaMemsetAsync(d_len, 0, sizeof(unsigned long long), S()));
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    prof_begin(PROF_VEC);
    trim_kernel<F><<<blocks, 256, 0, S()>>>(n, (const F*)x, d_len);
    prof_end(PROF_VEC);
    count_launch();
    ZKB_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    ZKB_CUDA(ZKB_D2H(&h, d_len, sizeof(h)));
    ZKB_CUDA(cudaStreamSynchronize(S()));
    *len = (size_t)h;
    return ZKB_OK;
  }
  static int div_vanishing(size_t len, size_t d, const void* p, void* q, int* exact) {
    *exact = 1;
    if (len <= d) return ZKB_OK;   // quotient is zero, remainder is p itself: exact iff p == 0 (caller's business)
    int rc;
    if ((rc = scratch_reserve(4096))) return rc;
    scratch_reset();
    int* flag = (int*)scratch_take(256);
    ZKB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), S()));
    prof_begin(PROF_VEC);
    div_vanishing_kernel<F><<<(unsigned)((len + 255) / 256), 256, 0, S()>>>(len, d, (const F*)p, (F*)q, flag);
    prof_end(PROF_VEC);
    count_launch();
    ZKB_CUDA(cudaGetLastError());
    int h = 0;
    ZKB_CUDA(ZKB_D2H(&h, flag, sizeof(int)));
    ZKB_CUDA(cudaStreamSynchronize(S()));
    *exact = h ? 0 : 1;
    return ZKB_OK;
  }

  static int quotient(size_t q, size_t n, const void* const* in, const uint64_t* g, const uint64_t* wq, const uint64_t* beta,
                      const uint64_t* gamma, const uint64_t* alpha, const uint64_t* zh_inv, void* out) {
    if (n == 0 || q % n || q / n > 8 || q / n < 2) return set_error(ZKB_ERR_ARG, "plonk quotient: coset size must be 2n..8n");
    QuotientArgs<F> A;
    const F** slots[14] = {&A.a, &A.b, &A.c, &A.z, &A.pi, &A.ql, &A.qr, &A.qo, &A.qm, &A.qc, &A.s1, &A.s2, &A.s3, &A.l1};
    for (int k = 0; k < 14; k++) *slots[k] = (const F*)in[k];
    A.g = mont_from_words<F>(g);
    A.wq = mont_from_words<F>(wq);
    A.beta = mont_from_words<F>(beta);
    A.beta_r2 = A.beta * F::r2();
    A.gamma = mont_from_words<F>(gamma);
    F al = mont_from_words<F>(alpha);
    A.alpha_c = from_mont(al);
    A.alpha2 = al * al;
    A.q = q;
    A.per = (uint32_t)(q / n);
    for (uint32_t k = 0; k < 8; k++) A.zh_inv[k] = k < A.per ? mont_from_words<F>(zh_inv + 4 * k) : F::zero();
    size_t threads = (q + ZKB_QUOT_CH - 1) / ZKB_QUOT_CH;
    prof_begin(PROF_VEC);
    plonk_quotient_kernel<F><<<(unsigned)((threads + 127) / 128), 128, 0, S()>>>(A, (F*)out);
    prof_end(PROF_VEC);
    count_launch();
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
  }
  static int add_sparse(void* v, size_t k, const uint64_t* idx, const uint64_t* vals, int subtract) {
    for (size_t done = 0; done < k; done += 8) {
      SparseArgs<F> a;
      memset(&a, 0, sizeof(a));
      a.k = (uint32_t)(k - done < 8 ? k - done : 8);
      for (uint32_t i = 0; i < a.k; i++) {
        a.idx[i] = idx[done + i];
        F x;
        memcpy(x.v, vals + 4 * (done + i), 32);
        a.val[i] = from_mont(F::r2() * x);   // reduce mod r
      }
      add_sparse_kernel<F><<<1, 32, 0, S()>>>(a, subtract, (F*)v);
      count_launch();
    }
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
  }
};

}  // namespace zkb

using namespace zkb;

#define NEED_INIT() \
  if (!ctx_ready()) return set_error(ZKB_ERR_NOINIT, "zkb_init has not been called (no CUDA context; no CPU fallback)"); \
  ZKB_ENTRY_GUARD()
#define BY_CURVE(EXPR_BN, EXPR_BLS)               \
  if (curve == ZKB_BN254) return EXPR_BN;         \
  if (curve == ZKB_BLS12_381) return EXPR_BLS;    \
  return set_error(ZKB_ERR_ARG, "unknown curve id");

extern "C" {

int zkb_fr_axpy_dev(int curve, size_t n, const uint64_t s[4], const void* d_x, size_t nx, const void* d_y, size_t ny, void* d_out) {
  NEED_INIT();
  BY_CURVE(FrVecOps<fr_bn>::axpy(n, s, d_x, nx, d_y, ny, d_out), FrVecOps<fr_bls>::axpy(n, s, d_x, nx, d_y, ny, d_out))
}
int zkb_fr_mul_powers_dev(int curve, size_t n, const uint64_t base[4], const uint64_t scale[4], const void* d_x, void* d_out) {
  NEED_INIT();
  BY_CURVE(FrVecOps<fr_bn>::mul_powers(n, base, scale, d_x, d_out), FrVecOps<fr_bls>::mul_powers(n, base, scale, d_x, d_out))
}
int zkb_fr_inverse_dev(int curve, size_t n, const void* d_x, void* d_out) {
  NEED_INIT();
  BY_CURVE(FrVecOps<fr_bn>::inverse(n, d_x, d_out), FrVecOps<fr_bls>::inverse(n, d_x, d_out))
}
int zkb_fr_scan_dev(int curve, int op, size_t n, const void* d_x, void* d_out) {
  NEED_INIT();
  if (op == 0) { BY_CURVE(FrVecOps<fr_bn>::scan<0>(n, d_x, d_out), FrVecOps<fr_bls>::scan<0>(n, d_x, d_out)) }
  if (op == 1) { BY_CURVE(FrVecOps<fr_bn>::scan<1>(n, d_x, d_out), FrVecOps<fr_bls>::scan<1>(n, d_x, d_out)) }
  return set_error(ZKB_ERR_ARG, "unknown scan op");
}
int zkb_fr_gather_dev(int curve, size_t n, const void* d_src, size_t stride, size_t offset, void* d_out) {
  NEED_INIT();
  BY_CURVE(FrVecOps<fr_bn>::gather(n, d_src, stride, offset, nullptr, d_out),
           FrVecOps<fr_bls>::gather(n, d_src, stride, offset, nullptr, d_out))
}
int zkb_fr_gather_index_dev(int curve, size_t n, const void* d_src, const void* d_idx_u32, void* d_out) {
  NEED_INIT();
  BY_CURVE(FrVecOps<fr_bn>::gather(n, d_src, 0, 0, (const uint32_t*)d_idx_u32, d_out),
           FrVecOps<fr_bls>::gather(n, d_src, 0, 0, (const uint32_t*)d_idx_u32, d_out))
}
int zkb_fr_eval_dev(int curve, size_t n, const void* d_coeffs, const uint64_t point[4], uint64_t out[4]) {
  NEED_INIT();
  BY_CURVE(FrVecOps<fr_bn>::eval(n, d_coeffs, point, out), FrVecOps<fr_bls>::eval(n, d_coeffs, point, out))
}
int zkb_fr_trim_dev(int curve, size_t n, const void* d_x, size_t* len) {
  NEED_INIT();
  if (!len) return set_error(ZKB_ERR_ARG, "trim: null output");
  BY_CURVE(FrVecOps<fr_bn>::trim(n, d_x, len), FrVecOps<fr_bls>::trim(n, d_x, len))
}
int zkb_fr_div_vanishing_dev(int curve, size_t len, size_t d, const void* d_p, void* d_q, int* exact) {
  NEED_INIT();
  if (d == 0) return set_error(ZKB_ERR_ARG, "div_vanishing: d must be positive");
  BY_CURVE(FrVecOps<fr_bn>::div_vanishing(len, d, d_p, d_q, exact), FrVecOps<fr_bls>::div_vanishing(len, d, d_p, d_q, exact))
}
int zkb_plonk_quotient_dev(int curve, size_t q, size_t n, const void* const d_inputs[14], const uint64_t g[4], const uint64_t omega_q[4],
                           const uint64_t beta[4], const uint64_t gamma[4], const uint64_t alpha[4], const uint64_t* zh_inv,
                           void* d_out) {
  NEED_INIT();
  BY_CURVE(FrVecOps<fr_bn>::quotient(q, n, d_inputs, g, omega_q, beta, gamma, alpha, zh_inv, d_out),
           FrVecOps<fr_bls>::quotient(q, n, d_inputs, g, omega_q, beta, gamma, alpha, zh_inv, d_out))
}
int zkb_fr_add_sparse_dev(int curve, void* d_vec, size_t k, const uint64_t* idx, const uint64_t* vals, int subtract) {
  NEED_INIT();
  BY_CURVE(FrVecOps<fr_bn>::add_sparse(d_vec, k, idx, vals, subtract), FrVecOps<fr_bls>::add_sparse(d_vec, k, idx, vals, subtract))
}

}  // extern "C"
