// ntt_warp.cuh -- register-resident NTT pass: one warp transforms whole columns with shuffles, no shared-memory round trips.
//
// Same decomposition, addressing and semantics as ntt.cuh (self-sorting multi-pass Cooley-Tukey N = R_1 R_2 ... R_P, natural
// order in and out, canonical data, Montgomery twiddles; /root/reference/src/bn254/polynomial.rs:536-585 via ark-poly's
// Radix2EvaluationDomain).  What changes is the engine that runs the size-R sub-transform of one column:
//
//   * ntt.cuh stages a tile in shared memory and does log2 R radix-2 stages there: one shared-memory round trip and one
//     __syncthreads() PER STAGE (ncu, round 1: sm throughput 60 %, issue-active 40 %, 3.25 M bank conflicts per launch).
//   * here a warp (or a 2^(K-EL)-lane part of it) holds a whole column of R = 2^K elements in registers, E = 2^EL per lane.
//     A radix-2 DIF stage over an index bit that lives in the lanes is turned into a thread-local one by swapping that lane bit
//     with a register-slot bit: every lane sends E/2 elements to its partner (shfl.xor) and receives E/2, after which both
//     members of every butterfly sit in one thread.  Every lane then does exactly E/2 butterflies -- one Montgomery product
//     each, 100 % lane utilisation of the multiplier -- and nothing is ever swapped back: the bit permutation is tracked at
//     compile time and undone by the store addresses.  No shared memory for data, no barriers, no bank conflicts.
//
// Bookkeeping.  Physical bits of an element's place: slot bits 0..EL-1 (which register), lane bits 0..K-EL-1.  Logical bits =
// the row index inside the column.  Initially slot bit i holds logical bit K-EL+i (the top bits) and lane bit i logical bit i.
// Stage t (t = 0..K-1) handles logical bit b = K-1-t.  For t < EL the bit is already a slot bit.  For t >= EL it sits in lane
// bit b and is exchanged with slot bit sigma(t) = (t - EL) mod EL, which holds a bit whose stage is over.  The twiddle of a
// butterfly is w_R^((p mod 2^b) << (K-1-b)); all logical bits below b are still in their original lane / slot positions, so
// for t >= EL it is one table entry per lane per stage.  The DIF output index is the bit reversal of the final logical index.
//
// Work is handed out per column: item = (batch, block, column); a persistent grid strides over the items so that a batch of
// transforms (the three inverse / three coset transforms of the Groth16 quotient) shares one launch and one set of waves.
#pragma once
#include <stdlib.h>
#include "ntt.cuh"

namespace zkb {

// ---- compile-time bit bookkeeping ------------------------------------------------------------------------------------------
template <int EL, int K>
struct WarpNttMap {
  static_assert(EL >= 1 && K > EL && K - EL <= 5, "a column must fit one warp");
  static constexpr int LB = K - EL;   // lane bits of one column
  // logical bit held by slot bit `s` BEFORE stage t
  static constexpr int slot_logical(int s, int t) {
    int cur = K - EL + s;
    for (int i = EL; i < t; i++)
      if ((i - EL) % EL == s) cur = K - 1 - i;   // stage i swapped logical bit K-1-i into this slot
    return cur;
  }
  // logical bit held by lane bit `l` AFTER all stages
  static constexpr int lane_logical_final(int l) {
    // lane bit l is swapped exactly once, at stage t = K-1-l (>= EL), and receives what slot sigma(t) held before that stage
    int t = K - 1 - l;
    return slot_logical((t - EL) % EL, t);
  }
  static constexpr int slot_logical_final(int s) { return slot_logical(s, K); }
};

template <class F>
__device__ __forceinline__ F shfl_xor_field(const F& v, uint32_t mask) {
  F r;
#pragma unroll
  for (int i = 0; i < F::N; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, v.v[i], mask);
  return r;
}
template <class F>
__device__ __forceinline__ F select_field(bool c, const F& a, const F& b) {   // c ? a : b
  F r;
#pragma unroll
  for (int i = 0; i < F::N; i++) r.v[i] = c ? a.v[i] : b.v[i];
  return r;
}
template <class F>
__device__ __forceinline__ F lds_twiddle(const uint4* t0, const uint4* t1, uint32_t j) {
  uint4 a = t0[j], b = t1[j];
  F w;
  w.v[0] = a.x; w.v[1] = a.y; w.v[2] = a.z; w.v[3] = a.w;
  w.v[4] = b.x; w.v[5] = b.y; w.v[6] = b.z; w.v[7] = b.w;
  return w;
}

// ---- the stages ---------------------------------------------------------------------------------------------------------------
// The stages t >= EL run in a loop whose body handles EL consecutive stages (slot bits 0..EL-1 in turn) with the lane bit, the
// shuffle mask and the twiddle index as run-time values, and every butterfly goes through ONE out-of-line routine (Montgomery
// product + add + sub): the kernel is ~2.3 K instructions.  (The first version unrolled all K stages with the multiplier inlined:
// 11.6 K instructions, `no_instruction` + `dispatch` = 22 % of the warp samples in ncu; it ran at the same speed -- see
// DESIGN.md section 5 for what the transform is bound by.)
template <class F>
struct FieldPair {
  F lo, hi;
};
template <class F>
__device__ __noinline__ FieldPair<F> dif_butterfly(F a, F b, F w) {
  FieldPair<F> r;
  r.lo = a + b;
  r.hi = mont_mul(a - b, w);
  return r;
}
template <class F>
__device__ __noinline__ F mul_outline(F a, F b) { return mont_mul(a, b); }
// power-table lookup whose two-level product (coset / scaled tables) goes through the shared out-of-line multiplier
template <class F>
__device__ __forceinline__ F pow_lookup_c(const PowTable<F>& t, unsigned long long e) {
  if (t.direct) return ntt_ldg(t.lo + e);
  F lo = ntt_ldg(t.lo + (e & ((1ull << t.h) - 1)));
  F hi = ntt_ldg(t.hi + (e >> t.h));
  return mul_outline(lo, hi);
}

template <class F, int EL, int SIG>
__device__ __forceinline__ void warp_ntt_lane_stage(F (&x)[1 << EL], uint32_t lane_in_col, uint32_t b, uint32_t kbits, const uint4* t0,
                                                    const uint4* t1) {
  constexpr int E = 1 << EL;
  const bool hi = (lane_in_col >> b) & 1u;
#pragma unroll
  for (int s0 = 0; s0 < E; s0++) {
    if (s0 & (1 << SIG)) continue;
    const int s1 = s0 | (1 << SIG);
    F send = select_field(hi, x[s0], x[s1]);
    F recv = shfl_xor_field(send, 1u << b);
    x[s0] = select_field(hi, recv, x[s0]);
    x[s1] = select_field(hi, x[s1], recv);
  }
  if (b > 0) {
    const F w = lds_twiddle<F>(t0, t1, (lane_in_col & ((1u << b) - 1u)) << (kbits - 1 - b));
#pragma unroll
    for (int s0 = 0; s0 < E; s0++) {
      if (s0 & (1 << SIG)) continue;
      const int s1 = s0 | (1 << SIG);
      FieldPair<F> r = dif_butterfly(x[s0], x[s1], w);
      x[s0] = r.lo;
      x[s1] = r.hi;
    }
  } else {
#pragma unroll
    for (int s0 = 0; s0 < E; s0++) {
      if (s0 & (1 << SIG)) continue;
      const int s1 = s0 | (1 << SIG);
      F a = x[s0], c = x[s1];
      x[s0] = a + c;
      x[s1] = a - c;
    }
  }
}
template <class F, int EL, int SIG>
struct WarpNttLaneGroup {   // slot bits SIG, SIG+1, ... EL-1 for the lane bits b, b-1, ... (stops below bit 0)
  static __device__ __forceinline__ void run(F (&x)[1 << EL], uint32_t lane_in_col, int b, uint32_t kbits, const uint4* t0,
                                             const uint4* t1) {
    if (b < 0) return;
    warp_ntt_lane_stage<F, EL, SIG>(x, lane_in_col, (uint32_t)b, kbits, t0, t1);
    WarpNttLaneGroup<F, EL, SIG + 1>::run(x, lane_in_col, b - 1, kbits, t0, t1);
  }
};
template <class F, int EL>
struct WarpNttLaneGroup<F, EL, EL> {
  static __device__ __forceinline__ void run(F (&)[1 << EL], uint32_t, int, uint32_t, const uint4*, const uint4*) {}
};
// the EL thread-local stages (logical bits K-1 .. K-EL), through the same out-of-line butterfly
template <class F, int EL, int K, int T>
struct WarpNttLocalStages {
  static __device__ __forceinline__ void run(F (&x)[1 << EL], uint32_t lane_in_col, const uint4* t0, const uint4* t1) {
    constexpr int E = 1 << EL;
    constexpr int LB = K - EL;
    constexpr int B = K - 1 - T;
    constexpr int SIG = EL - 1 - T;
#pragma unroll
    for (int s0 = 0; s0 < E; s0++) {
      if (s0 & (1 << SIG)) continue;
      const int s1 = s0 | (1 << SIG);
      const uint32_t low = ((uint32_t)(s0 & ((1 << SIG) - 1)) << LB) | lane_in_col;
      FieldPair<F> r = dif_butterfly(x[s0], x[s1], lds_twiddle<F>(t0, t1, low << (K - 1 - B)));
      x[s0] = r.lo;
      x[s1] = r.hi;
    }
    WarpNttLocalStages<F, EL, K, T + 1>::run(x, lane_in_col, t0, t1);
  }
};
template <class F, int EL, int K>
struct WarpNttLocalStages<F, EL, K, EL> {
  static __device__ __forceinline__ void run(F (&)[1 << EL], uint32_t, const uint4*, const uint4*) {}
};
template <class F, int EL, int K>
__device__ __forceinline__ void warp_ntt_compact(F (&x)[1 << EL], uint32_t lane_in_col, const uint4* t0, const uint4* t1) {
  static_assert(K - EL >= 1, "");
  WarpNttLocalStages<F, EL, K, 0>::run(x, lane_in_col, t0, t1);    // (K - 1 - T > 0 for every local stage: K > EL)
#pragma unroll 1
  for (int b = K - EL - 1; b >= 0; b -= EL) WarpNttLaneGroup<F, EL, 0>::run(x, lane_in_col, b, (uint32_t)K, t0, t1);
}

// final logical index of (lane_in_col, slot), and its bit reversal = the output row
template <int EL, int K>
__device__ __forceinline__ uint32_t warp_ntt_out_row(uint32_t lane_in_col, int slot) {
  typedef WarpNttMap<EL, K> M;
  uint32_t kr = 0;   // bit-reversed logical index: logical bit g contributes to output bit K-1-g
#pragma unroll
  for (int l = 0; l < K - EL; l++) kr |= ((lane_in_col >> l) & 1u) << (K - 1 - M::lane_logical_final(l));
#pragma unroll
  for (int s = 0; s < EL; s++) kr |= (((uint32_t)slot >> s) & 1u) << (K - 1 - M::slot_logical_final(s));
  return kr;
}

// One pass.  Every (batch, block, column) is one column transform handled by 2^(K-EL) lanes; a warp handles 2^(5-(K-EL)) of
// them side by side.  blockDim.x = 128.  Shared memory: the R/2 butterfly twiddles only (static, <= 4 KiB).
#define ZKB_NTT_MAX_BATCH 4
template <class F>
struct PreTables {   // first-pass scaling table per batch member (coset transforms of one batch may differ by a constant factor)
  PowTable<F> t[ZKB_NTT_MAX_BATCH];
};

template <class F, int EL, int K>
__global__ void __launch_bounds__(128, 4)
ntt_warp_pass_kernel(const F* __restrict__ src_all, F* __restrict__ dst_all, NttPass pp, uint32_t batch,
                     unsigned long long src_stride, unsigned long long dst_stride, PowTable<F> tw, PreTables<F> pres,
                     PowTable<F> post, F post_const) {
  constexpr int E = 1 << EL;
  constexpr int LB = K - EL;
  constexpr uint32_t COLS_PER_WARP = 1u << (5 - LB);
  constexpr uint32_t R = 1u << K;
  __shared__ uint4 t0[(R >> 1) + 4];
  __shared__ uint4 t1[(R >> 1) + 4];
  for (uint32_t j = threadIdx.x; j < (R >> 1); j += blockDim.x) {
    F w = pow_lookup(tw, (unsigned long long)j << (pp.log_n - K));   // w_R^j = w_N^(j N/R)
    t0[j] = make_uint4(w.v[0], w.v[1], w.v[2], w.v[3]);
    t1[j] = make_uint4(w.v[4], w.v[5], w.v[6], w.v[7]);
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lane_in_col = lane & ((1u << LB) - 1u);
  const uint32_t sub = lane >> LB;                                        // which of the warp's columns
  const unsigned long long cols_per_batch = 1ull << (pp.log_n - K);       // column transforms of one size-N transform
  const unsigned long long total = cols_per_batch * batch;
  const unsigned long long warp0 = ((unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  const unsigned long long nwarps = (unsigned long long)gridDim.x * (blockDim.x >> 5);

  for (unsigned long long wi = warp0; wi * COLS_PER_WARP < total; wi += nwarps) {
    const unsigned long long item = wi * COLS_PER_WARP + sub;
    const bool live = item < total;                    // (total is a multiple of COLS_PER_WARP whenever log_n - K >= 5 - LB)
    const unsigned long long it = live ? item : 0;
    const uint32_t bi = (uint32_t)(it >> (pp.log_n - K));
    const unsigned long long t = it & (cols_per_batch - 1);
    const F* src = src_all + bi * src_stride;
    F* dst = dst_all + bi * dst_stride;

    // ---- addressing (same conventions as ntt_pass_kernel with C = 1) ----
    unsigned long long base, col = 0, out_base, in_row_stride, out_row_stride;
    if (!pp.last) {
      const unsigned long long blk = t >> pp.log_m;
      col = t & ((1ull << pp.log_m) - 1);
      base = (blk << (pp.log_m + K)) + col;
      in_row_stride = 1ull << pp.log_m;
      out_base = base;
      out_row_stride = in_row_stride;
    } else {
      const unsigned long long k1v = t & ((1ull << pp.k1) - 1);
      const unsigned long long rest = t >> pp.k1;       // middle digits, most significant first
      base = (k1v << (pp.log_n - pp.k1)) + (rest << K);
      in_row_stride = 1;
      unsigned long long rrev = rest;
      if (pp.nmid == 2) {
        const unsigned long long d3 = rest & ((1ull << pp.kmid[1]) - 1);
        const unsigned long long d2 = rest >> pp.kmid[1];
        rrev = d2 + (d3 << pp.kmid[0]);
      }
      out_base = k1v + (rrev << pp.k1);
      out_row_stride = 1ull << pp.log_b;
    }

    // ---- load: slot s of lane l holds row (s << LB) | l ----
    F x[E];
#pragma unroll
    for (int s = 0; s < E; s++) {
      const uint32_t row = ((uint32_t)s << LB) | lane_in_col;
      const unsigned long long g = base + row * in_row_stride;
      if (live && g < pp.in_len) {
        x[s] = ntt_ld(src + g);
        if (pp.pre) x[s] = mul_outline(x[s], pow_lookup_c(pres.t[bi], g));
      } else {
        x[s] = F::zero();
      }
    }

    warp_ntt_compact<F, EL, K>(x, lane_in_col, t0, t1);

    // ---- store: the element in (lane, slot) is output row kr ----
#pragma unroll
    for (int s = 0; s < E; s++) {
      const uint32_t kr = warp_ntt_out_row<EL, K>(lane_in_col, s);
      const unsigned long long g = out_base + kr * out_row_stride;
      F v = x[s];
      // one multiplication site: pick the factor, then one out-of-line product
      bool scale = true;
      F f;
      if (!pp.last) {
        const unsigned long long e = (col * (unsigned long long)kr) << pp.log_b;
        scale = e != 0;
        f = pow_lookup_c(tw, e);
      } else if (pp.post == 1) {
        f = pow_lookup_c(post, g);
      } else if (pp.post == 2) {
        f = post_const;
      } else {
        scale = false;
        f = post_const;
      }
      if (scale) v = mul_outline(v, f);
      if (live) ntt_st(dst + g, v);
    }
  }
}

// host launcher of one pass for a fixed EL; K = pp.k selects the instantiation.  Defined here, INSTANTIATED in ntt_warp_inst_*.cu
// (one translation unit per field and EL: these kernels are the slowest thing in the library to compile).
template <class F, int EL, int K>
static inline void ntt_warp_launch_k(const F* src, F* dst, const NttPass& pp, uint32_t batch, size_t src_stride, size_t dst_stride,
                                     const PowTable<F>& tw, const PreTables<F>& pres, const PowTable<F>& post, const F& post_const,
                                     cudaStream_t st) {
  constexpr int LB = K - EL;
  const unsigned long long cols = ((unsigned long long)batch) << (pp.log_n - K);
  const unsigned long long warps = (cols + (1u << (5 - LB)) - 1) >> (5 - LB);
  unsigned long long ctas = (warps + 3) / 4;
  const unsigned long long resident = 148ull * 4;
  if (ctas > resident) ctas = resident;        // persistent: the warps stride over the columns
  ntt_warp_pass_kernel<F, EL, K><<<(unsigned)ctas, 128, 0, st>>>(src, dst, pp, batch, (unsigned long long)src_stride,
                                                                 (unsigned long long)dst_stride, tw, pres, post, post_const);
}
template <class F, int EL>
int ntt_warp_launch_el(const F* src, F* dst, const NttPass& pp, uint32_t batch, size_t src_stride, size_t dst_stride,
                       const PowTable<F>& tw, const PreTables<F>& pres, const PowTable<F>& post, const F& post_const, void* stream);

#ifdef ZKB_NTT_WARP_INSTANTIATE
template <class F, int EL>
int ntt_warp_launch_el(const F* src, F* dst, const NttPass& pp, uint32_t batch, size_t src_stride, size_t dst_stride,
                       const PowTable<F>& tw, const PreTables<F>& pres, const PowTable<F>& post, const F& post_const, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  switch ((int)pp.k - EL) {
    case 1: ntt_warp_launch_k<F, EL, EL + 1>(src, dst, pp, batch, src_stride, dst_stride, tw, pres, post, post_const, st); return 0;
    case 2: ntt_warp_launch_k<F, EL, EL + 2>(src, dst, pp, batch, src_stride, dst_stride, tw, pres, post, post_const, st); return 0;
    case 3: ntt_warp_launch_k<F, EL, EL + 3>(src, dst, pp, batch, src_stride, dst_stride, tw, pres, post, post_const, st); return 0;
    case 4: ntt_warp_launch_k<F, EL, EL + 4>(src, dst, pp, batch, src_stride, dst_stride, tw, pres, post, post_const, st); return 0;
    case 5: ntt_warp_launch_k<F, EL, EL + 5>(src, dst, pp, batch, src_stride, dst_stride, tw, pres, post, post_const, st); return 0;
  }
  return -1;
}
#endif

}  // namespace zkb
