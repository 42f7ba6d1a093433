// msm_pair.cuh -- G2 bucket accumulation with TWO LANES PER POINT (lane parity = Fp2 component).
//
// The one-thread-per-point kernel of msm.cuh holds a 64-register Fp2 XYZZ sum plus a 32-register affine point and the
// multiplier's operands: 248 registers, 2 warps per scheduler, 0.75 of the multiplier-pipe rate.
// Here the even lane of a pair owns the c0 half of every Fp2 value and the odd lane the c1 half, so a lane carries half
// the state (G1-sized: <= 128 registers on BN254, 4 warps per scheduler) and the Fp2 product becomes one LAZY dot
// product per lane,
//     c0 = a0 b0 + (p - a1) b1        c1 = a1 b0 + a0 b1        (mont_dot2: one interleaved reduction, ff.cuh)
// after exchanging the halves with 2 N shfl.xor -- 2 (3 N^2 + N) wide multiply-adds per Fp2 product, the schoolbook
// count with half of its reductions, against 3 (2 N^2 + N) for Karatsuba over fully reduced products.  The square is
// (a0 + a1)(a0 - a1) on the even lane and (2 a1) a0 on the odd one: a plain product each.
// Everything that shuffles runs warp-converged: the run loop has a warp-uniform trip count, the special cases of the
// group law (identity operands, P - P, P + P) are selects over pair-uniform flags, and the doubling formula is entered
// by the whole warp when any pair needs it.  The exact group law is kept (same results as madd() in ec.cuh).
// Replaces the same reference call as msm.cuh (ark-ec VariableBaseMSM over G2, /root/reference/src/bn254/curve.rs:375-392).
#pragma once
#include "msm.cuh"

namespace zkb {

template <class P>
struct PairFp2 {
  typedef Fp<P> F;
  static constexpr int N = P::N;
  static constexpr uint32_t FULL = 0xffffffffu;

  static __device__ __forceinline__ F xchg(const F& a) {
    F r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = __shfl_xor_sync(FULL, a.v[i], 1);
    return r;
  }
  static __device__ __forceinline__ F sel(bool c, const F& a, const F& b) {
    F r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
  }
  // both halves zero (pair-uniform result)
  static __device__ __forceinline__ bool is_zero(const F& a) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= a.v[i];
    t |= __shfl_xor_sync(FULL, t, 1);
    return t == 0;
  }
  // my half of the Fp2 product of the values whose halves the pair holds in (a, b)
  static __device__ __forceinline__ F mul_body(const F& a, const F& b) {
    const bool odd = threadIdx.x & 1;
    const F ao = xchg(a);
    const F x2 = sel(odd, ao, neg_lazy(ao));
    // even lane: a0 b0 + (p - a1) b1      odd lane: a1 b0 + a0 b1 -- the limbs of b0 / b1 are exchanged row by row
    return mont_dot2_rows<P>(a, x2, [&](int i, uint32_t& ui, uint32_t& vi) {
      const uint32_t mine = b.v[i], other = __shfl_xor_sync(FULL, mine, 1);
      ui = odd ? other : mine;
      vi = odd ? mine : other;
    });
  }
  static __device__ __forceinline__ F sqr_body(const F& a) {
    const bool odd = threadIdx.x & 1;
    F ao = xchg(a);
    F s = a + sel(odd, a, ao);          // even: a0 + a1      odd: 2 a1
    F d = sel(odd, ao, a - ao);         // even: a0 - a1      odd: a0
    return mont_mul(s, d);
  }
  static __device__ __forceinline__ F one(bool odd) { return sel(odd, F::zero(), F::one()); }
};
// The exchange, the sign and the operand selection live INSIDE the out-of-line bodies: a call passes two field elements
// (not four), and the caller keeps no temporaries of the product alive.  (A fully inlined addition is ~4 K instructions in the
// loop and runs slower -- instruction cache; measured 7.85 against 7.46 ms per 2^20-point launch on BN254.)
#if defined(__CUDA_ARCH__)
template <class P>
__device__ __noinline__ Fp<P> pair_mul_call(Fp<P> a, Fp<P> b) { return PairFp2<P>::mul_body(a, b); }
template <class P>
__device__ __noinline__ Fp<P> pair_sqr_call(Fp<P> a) { return PairFp2<P>::sqr_body(a); }
#endif
template <class P>
__device__ __forceinline__ Fp<P> pair_mul(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__)
  if constexpr (P::NOINLINE_MUL) return pair_mul_call<P>(a, b);
  else
#endif
    return PairFp2<P>::mul_body(a, b);
}
template <class P>
__device__ __forceinline__ Fp<P> pair_sqr(const Fp<P>& a) {
#if defined(__CUDA_ARCH__)
  if constexpr (P::NOINLINE_MUL) return pair_sqr_call<P>(a);
  else
#endif
    return PairFp2<P>::sqr_body(a);
}

// acc += (qx, qy) (or -= when negate) for the pair's point; `live` = false leaves acc unchanged (the lane still takes part
// in every shuffle).  Mirrors madd() of ec.cuh case by case; the flags of the special cases are known after the first two
// products, every result is committed as soon as it exists (short live ranges: the kernel sits at the 128-register line), and
// `mid()` is called once the point's coordinates are dead -- the caller loads the next point there.  `again()` must return the
// point once more: only the doubling path (the same point twice in one bucket) uses it.
template <class P, class Mid, class Again>
__device__ __forceinline__ void madd_pair(Fp<P>& X, Fp<P>& Y, Fp<P>& ZZ, Fp<P>& ZZZ, const Fp<P>& qx, const Fp<P>& qy_in,
                                          bool negate, bool live, bool odd, Mid mid, Again again) {
  typedef PairFp2<P> PF;
  typedef Fp<P> F;
  const F qy = negate ? neg(qy_in) : qy_in;
  uint32_t zq = 0, za = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) {
    zq |= qx.v[i] | qy.v[i];
    za |= ZZ.v[i];
  }
  zq |= __shfl_xor_sync(0xffffffffu, zq, 1);
  za |= __shfl_xor_sync(0xffffffffu, za, 1);
  const bool act = live && zq != 0;
  const bool take_q = act && za == 0;                 // identity + Q
  const bool busy = act && za != 0;
  if (take_q) {
    const F one = PF::one(odd);
    X = qx; Y = qy; ZZ = one; ZZZ = one;
  }
  const F Pp = pair_mul<P>(qx, ZZ) - X;
  const F R = pair_mul<P>(qy, ZZZ) - Y;
  mid();
  const bool pz = PF::is_zero(Pp), rz = PF::is_zero(R);
  const bool general = busy && !pz;
  const bool twice = busy && pz && rz;               // P + P
  if (busy && pz && !rz) { X = F::zero(); Y = F::zero(); ZZ = F::zero(); ZZZ = F::zero(); }   // P - P
  const F PP = pair_sqr<P>(Pp);
  const F PPP = pair_mul<P>(Pp, PP);
  {
    const F t = pair_mul<P>(ZZ, PP);
    if (general) ZZ = t;
  }
  {
    const F t = pair_mul<P>(ZZZ, PPP);
    if (general) ZZZ = t;
  }
  const F Q = pair_mul<P>(X, PP);
  const F YP = pair_mul<P>(Y, PPP);
  const F X3 = pair_sqr<P>(R) - PPP - dbl(Q);
  const F Y3 = pair_mul<P>(R, Q - X3) - YP;
  if (general) { X = X3; Y = Y3; }
  if (__any_sync(0xffffffffu, twice)) {   // repeated points ([g] * n keys): the whole warp walks the doubling formula
    F ax, ay;
    again(ax, ay);
    if (negate) ay = neg(ay);
    const F U = dbl(ay);
    const F V = pair_sqr<P>(U);
    const F W = pair_mul<P>(U, V);
    const F S = pair_mul<P>(ax, V);
    const F X2 = pair_sqr<P>(ax);
    const F M = dbl(X2) + X2;
    const F Xd = pair_sqr<P>(M) - dbl(S);
    const F Yd = pair_mul<P>(M, S - Xd) - pair_mul<P>(W, ay);
    if (twice) { X = Xd; Y = Yd; ZZ = V; ZZZ = W; }
  }
}

// Same work decomposition as msm_accumulate_kernel (equal runs of K sorted references, one piece per bucket a run touches),
// one run per LANE PAIR: a warp takes 16 consecutive runs per work item.  `points` is the Affine<Fp2> array seen as 4 Fp per
// point (x.c0, x.c1, y.c0, y.c1), `pieces` the XYZZ<Fp2> array seen as 8 Fp per piece; each lane moves its own halves.
template <class P, int MINB>
__global__ void __launch_bounds__(128, MINB) msm_accumulate_pair_kernel(MsmPlan pl, const Fp<P>* __restrict__ points,
                                                                         const uint32_t* __restrict__ refs,
                                                                         const uint32_t* __restrict__ start,
                                                                         const uint32_t* __restrict__ pstart,
                                                                         const uint32_t* __restrict__ run_bucket,
                                                                         Fp<P>* __restrict__ pieces,
                                                                         unsigned int* __restrict__ work) {
  typedef Fp<P> F;
  const unsigned long long nb = (unsigned long long)pl.bwin * pl.nbuck;
  const uint32_t total = start[nb];
  const uint32_t K = pl.krun;
  const uint32_t nruns = (total + K - 1) / K;
  const uint32_t nitems = (nruns + 15) >> 4;
  const uint32_t lane = threadIdx.x & 31;
  const bool odd = lane & 1;
  const uint32_t half = lane & 1;
  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(work, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nitems) break;
    const uint32_t t = (item << 4) + (lane >> 1);
    const bool valid = t < nruns;
    const uint32_t pos = valid ? t * K : 0u;
    uint32_t end = valid ? pos + K : 0u;
    if (end > total) end = total;
    uint32_t b = 0, bend = 0xffffffffu, slot = 0;
    if (valid) {
      b = run_bucket[t];
      bend = start[b + 1];
      slot = pstart[b] + (t - start[b] / K);
    }
    F X = F::zero(), Y = F::zero(), ZZ = F::zero(), ZZZ = F::zero();
    uint32_t ref_next = 0;
    F px = F::zero(), py = F::zero();
    if (valid) {
      ref_next = refs[pos];
      const F* src = points + (size_t)(ref_next & 0x7fffffffu) * 4 + half;
      px = load_vec(src);
      py = load_vec(src + 2);
    }
    for (uint32_t k = 0; k < K; k++) {   // warp-uniform trip count: the Fp2 products shuffle
      const uint32_t p = pos + k;
      const bool live = p < end;
      const uint32_t ref = ref_next;
      const bool more = p + 1 < end;
      if (more) ref_next = refs[p + 1];
      if (live && p == bend) {   // the run crosses into the next non-empty bucket: emit the finished piece
        F* dst = pieces + (size_t)slot * 8 + half;
        store_vec(dst, X); store_vec(dst + 2, Y); store_vec(dst + 4, ZZ); store_vec(dst + 6, ZZZ);
        X = F::zero(); Y = F::zero(); ZZ = F::zero(); ZZZ = F::zero();
        do {
          b++;
          bend = start[b + 1];
        } while (bend == p);
        slot = pstart[b];
      }
      F nx = px, ny = py;
      madd_pair<P>(X, Y, ZZ, ZZZ, px, py, (ref >> 31) != 0, live, odd,
                   [&]() {   // the point's coordinates are dead: fetch the next one under the remaining eight products
                     if (more) {
                       const F* src = points + (size_t)(ref_next & 0x7fffffffu) * 4 + half;
                       nx = load_vec(src);
                       ny = load_vec(src + 2);
                     }
                   },
                   [&](F& ax, F& ay) {
                     const F* src = points + (size_t)(ref & 0x7fffffffu) * 4 + half;
                     ax = load_vec(src);
                     ay = load_vec(src + 2);
                   });
      px = nx;
      py = ny;
    }
    if (valid) {
      F* dst = pieces + (size_t)slot * 8 + half;
      store_vec(dst, X); store_vec(dst + 2, Y); store_vec(dst + 4, ZZ); store_vec(dst + 6, ZZZ);
    }
  }
}

}  // namespace zkb
