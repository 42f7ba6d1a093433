// ntt.cuh -- Fr NTT / iNTT / coset-NTT for sm_100a.
//
// Replaces ark-poly 0.4.2 `Radix2EvaluationDomain::{fft,ifft}` (+ `get_coset`) as called from
// /root/reference/src/bn254/polynomial.rs:536-585 (bls12_381 twin identical): natural-order input, natural-
// order output, input zero-padded / truncated to N = 2^log_n, inverse includes the 1/N factor, coset variants
// scale coefficient j by offset^j before (forward) or by offset^-j after (inverse).
//
// Decomposition (self-sorting multi-pass Cooley-Tukey, N = R_1 R_2 ... R_P):
//   pass p works on B_p = R_1..R_{p-1} contiguous blocks of M_{p-1} = R_p * M_p elements, viewed as R_p rows of
//   M_p columns.  A CTA stages a tile of R_p rows x C adjacent columns in shared memory (each row segment is
//   C*32 contiguous bytes in HBM), runs the size-R_p sub-transform there (radix-2 DIF, log R_p stages, local
//   twiddles in shared memory) and writes row k back multiplied by the inter-pass twiddle w_N^(col*k*B_p).
//   Passes 1..P-1 are in place; the last pass (M_P = 1) gathers C blocks whose outputs are adjacent and writes
//   the digit-reversed, i.e. natural, order -- again C*32-byte contiguous segments.
// Data is kept in CANONICAL form in HBM; twiddles are in Montgomery form, so mont_mul(data, twiddle) is again
// canonical and no conversion pass exists.  Inter-pass / coset twiddles w^e come from a two-level table
// (w^e = HI[e >> h] * LO[e & (2^h-1)], 2 x 2^h x 32 B, L1/L2 resident) for the coset / scaled tables; the plain twiddle
// tables w^e and w^-e are DIRECT (N entries, one 32-byte load per twiddle) up to N = 2^26: the transform is bound by
// Montgomery products, not by HBM, so 32 B/element/pass of extra traffic buys back one product of
// the ~13 per element.
//
// Shared-memory layout: element slot x = row*C + col is split into two 16-byte planes (plane stride padded by
// 64 B) and the slot index is XOR-swizzled so that the 8 lanes of a quarter-warp (one LDS.128/STS.128
// wavefront) always hit 8 distinct 16-byte bank groups for: contiguous runs, butterfly pairs at any stage
// (a zero bit inserted at any position), and the column-strided scatter of the last pass.
#pragma once
#include "ff.cuh"

namespace zkb {

template <class F>
struct PowTable {      // base^e = hi[e >> h] * lo[e & mask]   (Montgomery form; `hi` may carry a constant factor)
  const F* lo;
  const F* hi;
  uint32_t h;
  uint32_t direct;     // 1: lo holds ALL powers (h = log N, no constant factor): one load, no product
};

struct NttPass {
  uint32_t log_n;      // log2 N
  uint32_t k;          // log2 R_p (rows of the tile = sub-transform size)
  uint32_t log_c;      // log2 C (columns per tile)
  uint32_t log_m;      // log2 M_p (columns per block); 0 on the last pass
  uint32_t log_b;      // log2 B_p (number of blocks)
  uint32_t last;       // 1: gather C blocks, write digit-reversed
  uint32_t k1;         // log2 R_1 (last pass: the fastest output digit)
  uint32_t nmid;       // last pass: number of middle digits (P-2, 0..2)
  uint32_t kmid[2];    // last pass: log2 R_2, log2 R_3
  uint32_t pre;        // 1: multiply input j by pre-table^j (first pass only)
  uint32_t post;       // 0: none, 1: multiply output k by post-table^k (hi carries the constant), 2: constant
  unsigned long long in_len;  // valid input elements (rest read as zero)
};

__device__ __forceinline__ uint32_t ntt_swz(uint32_t x) {
  uint32_t g = (x >> 3) & 7u;
  uint32_t f = ((g >> 1) & 3u) ^ ((g & 1u) ? 7u : 0u);
  return x ^ f;
}

template <class F>
__device__ __forceinline__ F ntt_lds(const uint4* p0, const uint4* p1, uint32_t x) {
  static_assert(F::N == 8, "Fr is 8 x 32-bit limbs");
  uint32_t s = ntt_swz(x);
  uint4 a = p0[s], b = p1[s];
  F r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <class F>
__device__ __forceinline__ void ntt_sts(uint4* p0, uint4* p1, uint32_t x, const F& v) {
  uint32_t s = ntt_swz(x);
  p0[s] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
  p1[s] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}
template <class F>
__device__ __forceinline__ F ntt_ldg(const F* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  F r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <class F>
__device__ __forceinline__ F ntt_ld(const F* p) {  // plain (coherent) load for data another kernel just wrote
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  F r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <class F>
__device__ __forceinline__ void ntt_st(F* p, const F& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
  q[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}
template <class F>
__device__ __forceinline__ F pow_lookup(const PowTable<F>& t, unsigned long long e) {
  if (t.direct) return ntt_ldg(t.lo + e);
  F lo = ntt_ldg(t.lo + (e & ((1ull << t.h) - 1)));
  F hi = ntt_ldg(t.hi + (e >> t.h));
  return lo * hi;
}

__device__ __forceinline__ uint32_t bitrev(uint32_t v, uint32_t bits) { return bits ? (__brev(v) >> (32 - bits)) : 0u; }

// One pass.  grid.x = number of tiles = N / (R*C); dynamic smem = 2 planes * (R*C + 4) * 16 B + (R/2 + 4) * 32 B.
template <class F>
__global__ void __launch_bounds__(512) ntt_pass_kernel(const F* src, F* dst, NttPass pp,
                                                       PowTable<F> tw, PowTable<F> pre, PowTable<F> post, F post_const) {
  extern __shared__ uint4 smem[];
  const uint32_t k = pp.k, log_c = pp.log_c;
  const uint32_t R = 1u << k, C = 1u << log_c;
  const uint32_t tile_elems = R << log_c;
  uint4* p0 = smem;
  uint4* p1 = smem + tile_elems + 4;
  uint4* t0 = p1 + tile_elems + 4;   // local twiddles w_R^j, j < R/2 (plane 0 / plane 1)
  uint4* t1 = t0 + (R >> 1) + 4;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  const unsigned long long tile = blockIdx.x;

  // ---- tile addressing -------------------------------------------------------------------------------
  // non-last: element (row, c) at  base + row * 2^log_m + c ;   last: base_c = base + c * 2^(log_n - k1), + row
  unsigned long long base, col0 = 0, out_base, out_row_stride;
  unsigned long long in_row_stride, in_col_stride;
  if (!pp.last) {
    uint32_t log_cg = pp.log_m - log_c;                 // column groups per block
    unsigned long long blk = tile >> log_cg;
    col0 = (tile & ((1ull << log_cg) - 1)) << log_c;
    base = (blk << (pp.log_m + k)) + col0;
    in_row_stride = 1ull << pp.log_m;
    in_col_stride = 1;
    out_base = base;
    out_row_stride = in_row_stride;
  } else {
    // blocks are indexed (k1, mid digits...) ; a tile takes C consecutive k1 for fixed mid digits
    uint32_t log_g1 = pp.k1 - log_c;                    // groups of C along k1 (k1 == 0 when P == 1)
    unsigned long long k1_0 = (tile & ((1ull << log_g1) - 1)) << log_c;
    unsigned long long rest = tile >> log_g1;           // middle digits, most significant first
    uint32_t log_rest = pp.log_b - pp.k1;               // bits in `rest`
    base = (k1_0 << (pp.log_n - pp.k1)) + (rest << k);
    in_row_stride = 1;
    in_col_stride = 1ull << (pp.log_n - pp.k1);
    // digit-reverse the middle digits: rest = d2 * R3 + d3  ->  d2 + R2 * d3
    unsigned long long rrev = rest;
    if (pp.nmid == 2) {
      unsigned long long d3 = rest & ((1ull << pp.kmid[1]) - 1);
      unsigned long long d2 = rest >> pp.kmid[1];
      rrev = d2 + (d3 << pp.kmid[0]);
    }
    (void)log_rest;
    out_base = k1_0 + (rrev << pp.k1);
    out_row_stride = 1ull << pp.log_b;
  }

  // ---- local twiddle table: w_R^j = w_N^(j * N/R) ------------------------------------------------------
  for (uint32_t j = tid; j < (R >> 1); j += nthr) {
    F w = pow_lookup(tw, (unsigned long long)j << (pp.log_n - k));
    t0[j] = make_uint4(w.v[0], w.v[1], w.v[2], w.v[3]);
    t1[j] = make_uint4(w.v[4], w.v[5], w.v[6], w.v[7]);
  }

  // ---- load tile (optionally pre-scaled by pre^j, j = global input index) -------------------------------
  for (uint32_t u = tid; u < tile_elems; u += nthr) {
    uint32_t row, c;
    if (!pp.last) { row = u >> log_c; c = u & (C - 1); }
    else          { row = u & (R - 1); c = u >> k; }
    unsigned long long g = base + row * in_row_stride + c * in_col_stride;
    F v;
    if (g < pp.in_len) {
      v = ntt_ld(src + g);
      if (pp.pre) v = v * pow_lookup(pre, g);
    } else {
      v = F::zero();
    }
    ntt_sts(p0, p1, (row << log_c) | c, v);
  }
  __syncthreads();

  // ---- radix-2 DIF stages in shared memory ---------------------------------------------------------------
  const uint32_t half_tile = tile_elems >> 1;
  for (int s = (int)k - 1; s >= 0; s--) {
    const uint32_t bp = (uint32_t)s + log_c;
    for (uint32_t q = tid; q < half_tile; q += nthr) {
      uint32_t x_lo = ((q >> bp) << (bp + 1)) | (q & ((1u << bp) - 1));
      uint32_t x_hi = x_lo | (1u << bp);
      F a = ntt_lds<F>(p0, p1, x_lo);
      F b = ntt_lds<F>(p0, p1, x_hi);
      F sum = a + b;
      F dif = a - b;
      if (s > 0) {
        uint32_t row = x_lo >> log_c;
        uint32_t j = (row & ((1u << s) - 1)) << (k - 1 - (uint32_t)s);
        uint4 wa = t0[j], wb = t1[j];
        F w;
        w.v[0] = wa.x; w.v[1] = wa.y; w.v[2] = wa.z; w.v[3] = wa.w;
        w.v[4] = wb.x; w.v[5] = wb.y; w.v[6] = wb.z; w.v[7] = wb.w;
        dif = dif * w;
      }
      ntt_sts(p0, p1, x_lo, sum);
      ntt_sts(p0, p1, x_hi, dif);
    }
    __syncthreads();
  }

  // ---- store: output row kr lives at bit-reversed position ---------------------------------------------
  for (uint32_t u = tid; u < tile_elems; u += nthr) {
    uint32_t kr = u >> log_c, c = u & (C - 1);
    F v = ntt_lds<F>(p0, p1, (bitrev(kr, k) << log_c) | c);
    unsigned long long g = out_base + kr * out_row_stride + c;
    if (!pp.last) {
      unsigned long long e = ((col0 + c) * (unsigned long long)kr) << pp.log_b;
      if (e) v = v * pow_lookup(tw, e);
    } else if (pp.post == 1) {
      v = v * pow_lookup(post, g);
    } else if (pp.post == 2) {
      v = v * post_const;
    }
    ntt_st(dst + g, v);
  }
}

// out[j] = base^j  (Montgomery form), j < n.  Used once per table.
template <class F>
__global__ void pow_table_kernel(F* out, F base, F scale, unsigned long long n, uint32_t shift) {
  unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (j >= n) return;
  F r = pow_u64(base, j << shift) * scale;   // scale is Montgomery(1) for plain tables
  ntt_st(out + j, r);
}

// ------------------------------------------------------------------------------------------------------
// element-wise kernels on canonical Fr vectors
// ------------------------------------------------------------------------------------------------------
enum { VEC_MUL = 0, VEC_ADD = 1, VEC_SUB = 2, VEC_MULSUB_RAW = 3 };

// out = a*b (canonical) | a+b | a-b ;  VEC_MULSUB_RAW: out = mont_mul(a,b) - c   (= (a*b)/R - c, for the H pipeline
// where c arrives pre-divided by R and the factor R is restored by the following inverse transform's constant)
template <class F>
__global__ void vec_op_kernel(int op, unsigned long long n, const F* __restrict__ a, unsigned long long na,
                              const F* __restrict__ b, unsigned long long nb, const F* __restrict__ c, F* __restrict__ out) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    F x = (i < na) ? ntt_ld(a + i) : F::zero();
    F y = (i < nb) ? ntt_ld(b + i) : F::zero();
    F r;
    if (op == VEC_MUL) r = (x * y) * F::r2();
    else if (op == VEC_ADD) r = x + y;
    else if (op == VEC_SUB) r = x - y;
    else r = x * y - ntt_ld(c + i);
    ntt_st(out + i, r);
  }
}

// reduce arbitrary 256-bit little-endian integers mod r (Fr::from(BigUint) semantics, polynomial.rs:537-540)
template <class F>
__global__ void reduce_kernel(unsigned long long n, F* __restrict__ v) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = ntt_ld(v + i);
  // x < 2^256 may exceed p.  The Montgomery product tolerates an unreduced SECOND operand (the row multiplier)
  // when the first is < p:  every partial sum stays < 2p and the result R2*x/R < 2p is reduced by final_sub.
  x = from_mont(F::r2() * x);
  ntt_st(v + i, x);
}

// flag[0] |= 1 if a[i]*b[i] != c[i] for some i  (the R1CS is satisfied iff the QAP remainder is zero;
// mirrors the ValueError of /root/reference/python/zksnake/groth16/qap.py:68-69)
template <class F>
__global__ void check_abc_kernel(unsigned long long n, const F* __restrict__ a, const F* __restrict__ b,
                                 const F* __restrict__ c, int* flag) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  int bad = 0;
  for (; i < n; i += stride) {
    F x = ntt_ld(a + i), y = ntt_ld(b + i), z = ntt_ld(c + i);
    F one_raw = F::zero();
    one_raw.v[0] = 1;
    if ((x * y) != (z * one_raw)) bad = 1;   // compare ab/R with c/R
  }
  if (bad) atomicOr(flag, 1);
}

// out[row] = sum_k val[k] * w[col[k]] over the CSR row -- SparseArray.dot of /root/reference/python/zksnake/array.py:36-43
// (the A.w, B.w, C.w products of qap.py:53-55).  Rows past n_rows_csr (domain padding) are zero.  Canonical in and out:
// the raw Montgomery products val*w/R are summed and one multiplication by R^2 restores the scale.
#define ZKB_SPMV_LONG 64u      // rows with more non-zeros than this are summed by whole CTAs (spmv_long_kernel)
#define ZKB_SPMV_SPLIT 64u     // CTAs per long row

template <class F>
__global__ void spmv_kernel(unsigned long long n_out, unsigned long long n_rows_csr, const unsigned long long* __restrict__ row_ptr,
                            const uint32_t* __restrict__ col, const F* __restrict__ val, const F* __restrict__ w,
                            F* __restrict__ out) {
  unsigned long long row = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (row >= n_out) return;
  F acc = F::zero();
  if (row < n_rows_csr) {
    unsigned long long lo = row_ptr[row], hi = row_ptr[row + 1];
    if (hi - lo > ZKB_SPMV_LONG) return;     // a long row: spmv_long_kernel / spmv_long_finish_kernel write it
    for (unsigned long long k = lo; k < hi; k++) acc = acc + ntt_ldg(val + k) * ntt_ld(w + col[k]);
    acc = acc * F::r2();
  }
  ntt_st(out + row, acc);
}

// Skewed matrices: one column of a circuit that feeds every constraint (the `inp` wire of the chain circuit: a 2^20-entry row of
// B^T in Groth16.setup's L/R/O products, /root/reference/python/zksnake/groth16/protocol.py:64-77) would keep ONE thread busy
// for 2^20 products (270 ms per launch in round 1).  The rows longer than ZKB_SPMV_LONG are listed when the matrix is
// created; CTA (j, r) sums slice j of long row r (raw Montgomery products, block tree), a warp adds the ZKB_SPMV_SPLIT slices.
template <class F>
__global__ void __launch_bounds__(256) spmv_long_kernel(const uint32_t* __restrict__ long_rows,
                                                        const unsigned long long* __restrict__ row_ptr,
                                                        const uint32_t* __restrict__ col, const F* __restrict__ val,
                                                        const F* __restrict__ w, F* __restrict__ partial) {
  __shared__ uint4 sh0[256], sh1[256];
  const uint32_t row = long_rows[blockIdx.y];
  const unsigned long long lo = row_ptr[row], hi = row_ptr[row + 1];
  const unsigned long long per = (hi - lo + ZKB_SPMV_SPLIT - 1) / ZKB_SPMV_SPLIT;
  unsigned long long a = lo + blockIdx.x * per, b = a + per;
  if (b > hi) b = hi;
  F acc = F::zero();
  for (unsigned long long k = a + threadIdx.x; k < b; k += blockDim.x) acc = acc + ntt_ldg(val + k) * ntt_ld(w + col[k]);
  const uint32_t t = threadIdx.x;
  sh0[t] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
  sh1[t] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
  __syncthreads();
  for (uint32_t off = 128; off > 0; off >>= 1) {
    if (t < off) {
      uint4 p = sh0[t + off], q = sh1[t + off];
      F o;
      o.v[0] = p.x; o.v[1] = p.y; o.v[2] = p.z; o.v[3] = p.w;
      o.v[4] = q.x; o.v[5] = q.y; o.v[6] = q.z; o.v[7] = q.w;
      acc = acc + o;
      sh0[t] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
      sh1[t] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
    }
    __syncthreads();
  }
  if (t == 0) ntt_st(partial + (unsigned long long)blockIdx.y * ZKB_SPMV_SPLIT + blockIdx.x, acc);
}
// one warp per long row: sum of its slice sums, scale restored, written to out[row]
template <class F>
__global__ void __launch_bounds__(128) spmv_long_finish_kernel(uint32_t n_long, const uint32_t* __restrict__ long_rows,
                                                               const F* __restrict__ partial, F* __restrict__ out) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_long) return;
  F acc = F::zero();
  for (uint32_t j = lane; j < ZKB_SPMV_SPLIT; j += 32) acc = acc + ntt_ld(partial + (unsigned long long)warp * ZKB_SPMV_SPLIT + j);
  for (uint32_t off = 16; off > 0; off >>= 1) {
    F o;
#pragma unroll
    for (int i = 0; i < F::N; i++) o.v[i] = __shfl_down_sync(0xffffffffu, acc.v[i], off);
    acc = acc + o;
  }
  if (lane == 0) ntt_st(out + long_rows[warp], acc * F::r2());
}

}  // namespace zkb
