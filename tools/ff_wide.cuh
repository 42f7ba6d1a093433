// ff_wide.cuh -- split Montgomery multiplication: full-width product (one level of subtractive Karatsuba) + separate reduction.
//
// Why: every kernel of this library is bound by the IMAD.WIDE issue rate while the alu pipe has slack.
// The fused CIOS multiplier of ff.cuh spends 2 N^2 wide multiply-adds per product (N^2 for a*b, N^2 for m*p).  Splitting it
//   * lets a*b use Karatsuba: 3 (N/2)^2 instead of N^2 wide multiply-adds, paid for with ~60 IADD3/LOP3 on the idle alu pipe
//     (N = 8: 112 + 8 instead of 128 + 8 multiplier-pipe instructions per product),
//   * lets sums of products share ONE reduction (Fp2 multiplication: 3 products, 2 reductions; the Y coordinate of a mixed
//     addition: 2 products, 1 reduction).
// Same __host__ __device__ text as ff.cuh (the host build emulates the carry flag).
//
// STATUS: EXPERIMENTAL, NOT USED BY ANY KERNEL.  Correct (checked against the fused multiplier on 200 000 random + edge inputs
// per field on the host), but measured SLOWER on the B200: 54 G products/s against 66 G/s for the fused CIOS multiplier (Fq
// BN254, tools/ffbench.cu mode 3, gpurun_out/ffbench_r1w_split.txt).  SASS shows why: ptxas emits 103 IMAD.WIDE per product as
// intended, but fills the alu pipe's deficit with IMAD.MOV.U32 / IMAD.X (41 + 13 per product) which run on the multiplier
// pipe, so the pipe this was meant to relieve ends up busier (578 vs 531 cycles per product).  A parameter struct opts in with
// `static constexpr bool SPLIT_MUL = true;` (ff.cuh: uses_split_mul); nothing in the library does.
#pragma once
#include "ff.cuh"

namespace zkb {

// t[0 .. 2L) = a[0 .. L) * b[0 .. L)   (L even).  Two interleaved accumulators: products whose low word lands on an even column
// and products whose low word lands on an odd column; within a row each class is one uninterrupted carry chain of L/2
// mad.lo.cc / madc.hi.cc pairs (one IMAD.WIDE.X each).  A chain's carry-out always lands in a column that so far holds only
// earlier carry-outs, so it cannot overflow.
template <int L>
ZKB_HD void mul_limbs(const uint32_t* a, const uint32_t* b, uint32_t* t) {
  static_assert(L % 2 == 0, "even limb count required");
  uint32_t ev[2 * L], od[2 * L];
#pragma unroll
  for (int k = 0; k < 2 * L; k++) ev[k] = od[k] = 0;
#pragma unroll
  for (int i = 0; i < L; i++) {
    const uint32_t bi = b[i];
    {  // class "even column": j = i mod 2, i mod 2 + 2, ...
      const int j0 = i & 1, c = i + j0;
      ev[c] = mad_lo_cc(a[j0], bi, ev[c]);
      ev[c + 1] = madc_hi_cc(a[j0], bi, ev[c + 1]);
#pragma unroll
      for (int k = 2; k < L; k += 2) {
        ev[c + k] = madc_lo_cc(a[j0 + k], bi, ev[c + k]);
        ev[c + k + 1] = madc_hi_cc(a[j0 + k], bi, ev[c + k + 1]);
      }
      if (c + L < 2 * L) ev[c + L] = addc(ev[c + L], 0u);   // (beyond 2L the carry is mathematically zero)
    }
    {  // class "odd column"
      const int j0 = (i + 1) & 1, c = i + j0;
      od[c] = mad_lo_cc(a[j0], bi, od[c]);
      od[c + 1] = madc_hi_cc(a[j0], bi, od[c + 1]);
#pragma unroll
      for (int k = 2; k < L; k += 2) {
        od[c + k] = madc_lo_cc(a[j0 + k], bi, od[c + k]);
        od[c + k + 1] = madc_hi_cc(a[j0 + k], bi, od[c + k + 1]);
      }
      if (c + L < 2 * L) od[c + L] = addc(od[c + L], 0u);
    }
  }
  t[0] = ev[0];   // column 0 only ever receives even-class words (od[0] stays 0)
  t[1] = add_cc(ev[1], od[1]);
#pragma unroll
  for (int k = 2; k < 2 * L - 1; k++) t[k] = addc_cc(ev[k], od[k]);
  t[2 * L - 1] = addc(ev[2 * L - 1], od[2 * L - 1]);
}

// |x - y| over H limbs; returns all-ones when x < y
template <int H>
ZKB_HD uint32_t abs_diff(const uint32_t* x, const uint32_t* y, uint32_t* d) {
  d[0] = sub_cc(x[0], y[0]);
#pragma unroll
  for (int k = 1; k < H; k++) d[k] = subc_cc(x[k], y[k]);
  const uint32_t neg = subc(0u, 0u);   // 0xffffffff when x < y
  // conditional two's complement: (d ^ neg) - neg
  d[0] = add_cc(d[0] ^ neg, neg & 1u);
#pragma unroll
  for (int k = 1; k < H - 1; k++) d[k] = addc_cc(d[k] ^ neg, 0u);
  d[H - 1] = addc(d[H - 1] ^ neg, 0u);
  return neg;
}

// t[0 .. 2N) = a * b by one level of subtractive Karatsuba:  a = a0 + a1 X, b = b0 + b1 X  (X = 2^(32 N/2)),
//   a*b = z0 + (z0 + z2 + (a0 - a1)(b1 - b0)) X + z2 X^2,   z0 = a0 b0, z2 = a1 b1.
template <int N>
ZKB_HD void mul_wide(const uint32_t* a, const uint32_t* b, uint32_t* t) {
  constexpr int H = N / 2;
  static_assert(N % 4 == 0, "limb count must be a multiple of four");
  uint32_t da[H], db[H], m[N], s[N];
  mul_limbs<H>(a, b, t);               // z0 -> t[0 .. N)
  mul_limbs<H>(a + H, b + H, t + N);   // z2 -> t[N .. 2N)
  const uint32_t na = abs_diff<H>(a, a + H, da);       // a0 - a1
  const uint32_t nb = abs_diff<H>(b + H, b, db);       // b1 - b0
  mul_limbs<H>(da, db, m);
  const uint32_t neg = na ^ nb;                        // all-ones: the middle product enters with a minus sign
  // s = z0 + z2  (N limbs, carry c1)
  s[0] = add_cc(t[0], t[N]);
#pragma unroll
  for (int k = 1; k < N; k++) s[k] = addc_cc(t[k], t[N + k]);
  uint32_t top = addc(0u, 0u);
  // s +- m: add (m ^ neg) with carry-in (neg != 0) -- the two's complement of m when the sign is negative.  add.cc(neg, neg)
  // sets the carry flag to exactly that bit.  True top limb = c1 + carry-out - (neg & 1).
  (void)add_cc(neg, neg);
#pragma unroll
  for (int k = 0; k < N; k++) s[k] = addc_cc(s[k], m[k] ^ neg);
  top = addc(top, 0u);
  top -= (neg & 1u);
  // t += s X^H  (+ top at limb H + N)
  t[H] = add_cc(t[H], s[0]);
#pragma unroll
  for (int k = 1; k < N; k++) t[H + k] = addc_cc(t[H + k], s[k]);
  t[H + N] = addc_cc(t[H + N], top);
#pragma unroll
  for (int k = H + N + 1; k < 2 * N - 1; k++) t[k] = addc_cc(t[k], 0u);
  t[2 * N - 1] = addc(t[2 * N - 1], 0u);
}

// Montgomery reduction of t (2N limbs, t < p * 2^(32 N)): returns t / 2^(32 N) mod p, fully reduced.  In place on t: row i adds
// m_i * p at limb i as two carry chains (even and odd limbs of p); the chains' carry-outs are collected in small side counters.
template <class P>
ZKB_HD Fp<P> redc(uint32_t* t) {
  constexpr int N = P::N;
  uint32_t kc[N + 2];   // carries into limbs N .. 2N+1
#pragma unroll
  for (int k = 0; k < N + 2; k++) kc[k] = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint32_t m = mul_lo(t[i], P::INV);
    // even limbs of p: columns i, i+1, ..., i+N-1
    t[i] = mad_lo_cc(P::MOD(0), m, t[i]);
    t[i + 1] = madc_hi_cc(P::MOD(0), m, t[i + 1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      t[i + j] = madc_lo_cc(P::MOD(j), m, t[i + j]);
      t[i + j + 1] = madc_hi_cc(P::MOD(j), m, t[i + j + 1]);
    }
    kc[i] = addc(kc[i], 0u);          // carry into limb i + N
    // odd limbs of p: columns i+1 .. i+N
    t[i + 1] = mad_lo_cc(P::MOD(1), m, t[i + 1]);
    if (i + 2 < 2 * N) t[i + 2] = madc_hi_cc(P::MOD(1), m, t[i + 2]);
#pragma unroll
    for (int j = 3; j < N; j += 2) {
      t[i + j] = madc_lo_cc(P::MOD(j), m, t[i + j]);
      if (i + j + 1 < 2 * N) t[i + j + 1] = madc_hi_cc(P::MOD(j), m, t[i + j + 1]);
    }
    kc[i + 1] = addc(kc[i + 1], 0u);  // carry into limb i + N + 1
  }
  // result = t[N .. 2N) + carries  (< 2p)
  Fp<P> r;
  r.v[0] = add_cc(t[N], kc[0]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(t[N + k], kc[k]);
  r.v[N - 1] = addc(t[2 * N - 1], kc[N - 1]);
  final_sub<P>(r.v);
  return r;
}

// ---- dedicated squaring (experimental, unused by the kernels; host-tested through zkb_test_field_op_host op 5) ----
// a^2 needs each off-diagonal product a_i a_j only once: N(N-1)/2 + N wide multiply-adds instead of N^2, then the same
// reduction -- 100 + 8 multiplier-pipe instructions instead of 128 + 8 for N = 8.  SASS of a squaring loop (sm_100a, CUDA 12.9,
// build/sq experiment of round 1): BN254 92 IMAD.WIDE + 34 other multiplier-pipe instructions (IMAD, IMAD.HI, IMAD.MOV,
// IMAD.X) against 120 + 18 for a * a, i.e. ~436 against ~516 pipe cycles per warp (-15 %); BLS12-381 210 + 41 against
// 276 + 26 (-20 %); the price is ~85 more alu-pipe instructions (doubling, merging the column chains), which have slack.
// NOT YET MEASURED on the GPU: 2 of the 10 products of a mixed addition, 3 of the 9 of a doubling and two thirds of a
// square-root / inversion chain are squarings.
// t[0..2N) = a^2 : off-diagonal products once (even/odd column chains), doubled, plus the diagonal squares
template <int N>
ZKB_HD void sqr_limbs(const uint32_t* a, uint32_t* t) {
  uint32_t ev[2 * N], od[2 * N];
#pragma unroll
  for (int k = 0; k < 2 * N; k++) ev[k] = od[k] = 0;
  // row i multiplies a_i with a_j, j > i; products landing on even columns (i + j even) and odd columns separately
#pragma unroll
  for (int i = 0; i < N - 1; i++) {
    const uint32_t ai = a[i];
    {  // j = i + 2, i + 4, ... : column i + j has the parity of 2i = even
      bool first = true;
#pragma unroll
      for (int j = i + 2; j < N; j += 2) {
        const int c = i + j;
        if (first) { ev[c] = mad_lo_cc(a[j], ai, ev[c]); first = false; }
        else ev[c] = madc_lo_cc(a[j], ai, ev[c]);
        ev[c + 1] = madc_hi_cc(a[j], ai, ev[c + 1]);
      }
      if (!first) {
        // carry out of the chain lands on the next even column pair start: column (i + last j) + 2
        int last = i + 2 + ((N - 1 - (i + 2)) / 2) * 2;
        int c = i + last + 2;
        if (c < 2 * N) ev[c] = addc(ev[c], 0u);
      }
    }
    {  // j = i + 1, i + 3, ... : odd columns
      bool first = true;
#pragma unroll
      for (int j = i + 1; j < N; j += 2) {
        const int c = i + j;
        if (first) { od[c] = mad_lo_cc(a[j], ai, od[c]); first = false; }
        else od[c] = madc_lo_cc(a[j], ai, od[c]);
        od[c + 1] = madc_hi_cc(a[j], ai, od[c + 1]);
      }
      if (!first) {
        int last = i + 1 + ((N - 1 - (i + 1)) / 2) * 2;
        int c = i + last + 2;
        if (c < 2 * N) od[c] = addc(od[c], 0u);
      }
    }
  }
  // merge, double
  t[0] = 0;
  t[1] = add_cc(ev[1], od[1]);
#pragma unroll
  for (int k = 2; k < 2 * N - 1; k++) t[k] = addc_cc(ev[k], od[k]);
  t[2 * N - 1] = addc(ev[2 * N - 1], od[2 * N - 1]);
  t[1] = add_cc(t[1], t[1]);
#pragma unroll
  for (int k = 2; k < 2 * N - 1; k++) t[k] = addc_cc(t[k], t[k]);
  t[2 * N - 1] = addc(t[2 * N - 1], t[2 * N - 1]);
  // diagonal
  t[0] = mad_lo_cc(a[0], a[0], t[0]);
  t[1] = madc_hi_cc(a[0], a[0], t[1]);
#pragma unroll
  for (int i = 1; i < N; i++) {
    t[2 * i] = madc_lo_cc(a[i], a[i], t[2 * i]);
    t[2 * i + 1] = madc_hi_cc(a[i], a[i], t[2 * i + 1]);
  }
}
template <class P>
ZKB_HD Fp<P> mont_sqr_split(const Fp<P>& a) {
  uint32_t t[2 * P::N];
  sqr_limbs<P::N>(a.v, t);
  return redc<P>(t);
}

template <class P>
ZKB_HD Fp<P> mont_mul_split(const Fp<P>& a, const Fp<P>& b) {
  uint32_t t[2 * P::N];
  mul_wide<P::N>(a.v, b.v, t);
  return redc<P>(t);
}

}  // namespace zkb
