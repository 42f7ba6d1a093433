// ff.cuh -- multi-limb Montgomery prime-field arithmetic for sm_100a (and a bit-identical host path).
//
// The reference reaches this arithmetic through ark-ff 0.4.2 `Fp<MontBackend<..>>` (Montgomery form,
// 64-bit limbs, CPU) from /root/reference/src/bn254/polynomial.rs and src/bn254/curve.rs.  Here the same
// fields are laid out as N x 32-bit limbs so that every multiply-accumulate is one integer-pipe
// IMAD.WIDE.U32 with carry-in/out: a (mad.lo.cc, madc.hi.cc) PTX pair on the same operands fuses into one
// wide SASS instruction.  The multiplier keeps two interleaved accumulators ("even" columns and "odd"
// columns) so that every row of partial products is a single uninterrupted carry chain.
//
// Everything is __host__ __device__: the host build emulates the PTX carry flag with a thread-local
// variable, so the very same algorithm text is unit-tested on the CPU (tests/test_host_arith.py) and runs on
// the GPU.  The host path is also what the C-ABI uses for the handful of scalar operations after a kernel
// (window recombination, final affine conversion).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZKB_HD __host__ __device__ __forceinline__
#define ZKB_D __device__ __forceinline__
#else
#define ZKB_HD inline
#define ZKB_D inline
#endif

#include "ff_params.cuh"

namespace zkb {

// ----------------------------------------------------------------------------------------------------
// carry-flag primitives
// ----------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define ZKB_ASM_R1(ins, r, a, b)      asm volatile(ins " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define ZKB_ASM_R2(ins, r, a, b, c)   asm volatile(ins " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c))
ZKB_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; ZKB_ASM_R1("add.cc.u32", r, a, b); return r; }
ZKB_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; ZKB_ASM_R1("addc.cc.u32", r, a, b); return r; }
ZKB_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; ZKB_ASM_R1("addc.u32", r, a, b); return r; }
ZKB_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; ZKB_ASM_R1("sub.cc.u32", r, a, b); return r; }
ZKB_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; ZKB_ASM_R1("subc.cc.u32", r, a, b); return r; }
ZKB_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; ZKB_ASM_R1("subc.u32", r, a, b); return r; }
ZKB_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; ZKB_ASM_R1("mul.lo.u32", r, a, b); return r; }
ZKB_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; ZKB_ASM_R1("mul.hi.u32", r, a, b); return r; }
ZKB_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZKB_ASM_R2("mad.lo.cc.u32", r, a, b, c); return r; }
ZKB_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZKB_ASM_R2("madc.lo.cc.u32", r, a, b, c); return r; }
ZKB_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZKB_ASM_R2("mad.hi.cc.u32", r, a, b, c); return r; }
ZKB_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZKB_ASM_R2("madc.hi.cc.u32", r, a, b, c); return r; }
ZKB_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZKB_ASM_R2("madc.hi.u32", r, a, b, c); return r; }
#else
// Host emulation of the PTX condition-code register (one per thread).
inline uint32_t& cf_() { static thread_local uint32_t cf = 0; return cf; }
inline uint32_t add3_(uint64_t a, uint64_t b, uint64_t c, bool out) {
  uint64_t t = a + b + c;
  if (out) cf_() = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t sub3_(uint64_t a, uint64_t b, uint64_t c, bool out) {
  uint64_t t = a - b - c;
  if (out) cf_() = (uint32_t)((t >> 32) & 1);  // borrow
  return (uint32_t)t;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) { return add3_(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return add3_(a, b, cf_(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return add3_(a, b, cf_(), false); }
// PTX: sub.cc writes CC.CF = borrow; subc consumes it as borrow-in.
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return sub3_(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return sub3_(a, b, cf_(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return sub3_(a, b, cf_(), false); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3_(mul_lo(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3_(mul_lo(a, b), c, cf_(), true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3_(mul_hi(a, b), c, 0, true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3_(mul_hi(a, b), c, cf_(), true); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return add3_(mul_hi(a, b), c, cf_(), false); }
#endif

// ----------------------------------------------------------------------------------------------------
// Fp<P>: element of the prime field described by P (ff_params.cuh), Montgomery form, fully reduced
// ----------------------------------------------------------------------------------------------------
template <class P>
struct Fp {
  static constexpr int N = P::N;
  typedef P Params;
  uint32_t v[P::N];

  ZKB_HD static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r;
  }
  ZKB_HD static Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::R1(i);
    return r;
  }
  ZKB_HD static Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::R2(i);
    return r;
  }
  ZKB_HD bool is_zero() const {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= v[i];
    return t == 0;
  }
  ZKB_HD bool operator==(const Fp& o) const {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= v[i] ^ o.v[i];
    return t == 0;
  }
  ZKB_HD bool operator!=(const Fp& o) const { return !(*this == o); }
};

// r = (t >= p) ? t - p : t     (t < 2p)
template <class P>
ZKB_HD void final_sub(uint32_t* t) {
  constexpr int N = P::N;
  uint32_t s[N];
  s[0] = sub_cc(t[0], P::MOD(0));
#pragma unroll
  for (int i = 1; i < N; i++) s[i] = subc_cc(t[i], P::MOD(i));
  uint32_t borrow = subc(0u, 0u);  // 0 - 0 - borrow -> 0xffffffff when t < p
#pragma unroll
  for (int i = 0; i < N; i++) t[i] = borrow ? t[i] : s[i];
}

template <class P>
ZKB_HD Fp<P> operator+(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
  r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
  r.v[N - 1] = addc(a.v[N - 1], b.v[N - 1]);  // p < 2^(32N-1): no carry out
  final_sub<P>(r.v);
  return r;
}

template <class P>
ZKB_HD Fp<P> operator-(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
  r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = subc(0u, 0u);  // all ones when a < b
  r.v[0] = add_cc(r.v[0], P::MOD(0) & borrow);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(r.v[i], P::MOD(i) & borrow);
  r.v[N - 1] = addc(r.v[N - 1], P::MOD(N - 1) & borrow);
  return r;
}

template <class P>
ZKB_HD Fp<P> neg(const Fp<P>& a) {
  constexpr int N = P::N;
  Fp<P> r;
  uint32_t nz = a.is_zero() ? 0u : 0xffffffffu;
  r.v[0] = sub_cc(P::MOD(0) & nz, a.v[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.v[i] = subc_cc(P::MOD(i) & nz, a.v[i]);
  r.v[N - 1] = subc(P::MOD(N - 1) & nz, a.v[N - 1]);
  return r;
}

template <class P>
ZKB_HD Fp<P> dbl(const Fp<P>& a) { return a + a; }

// Montgomery product a*b*R^-1 mod p, interleaved (CIOS-style) with even/odd column accumulators.
//   T = sum_k e[k] 2^(32k) + 2^32 * sum_k o[k] 2^(32k)
// Each row adds a_j*b_i for even j into e (columns j, j+1) and for odd j into o (columns j, j+1 <-> o[j-1], o[j]),
// each as ONE carry chain of N mad instructions (N/2 IMAD.WIDE after ptxas pairs lo/hi); the same for m*p.
// Dividing by 2^32 is a renaming: the old o becomes the new e, the old e (shifted down by two limbs) becomes
// the new o, and the dropped-column carry is injected into the next chain.
template <class P>
ZKB_HD Fp<P> mont_mul(const Fp<P>& A, const Fp<P>& B) {
  constexpr int N = P::N;
  static_assert(N % 2 == 0, "even limb count required");
  const uint32_t* a = A.v;
  const uint32_t* b = B.v;
  uint32_t e[N], o[N];
  // ---- row 0: T = a * b[0]
  {
    uint32_t bi = b[0];
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      e[j] = mul_lo(a[j], bi);
      e[j + 1] = mul_hi(a[j], bi);
      o[j] = mul_lo(a[j + 1], bi);
      o[j + 1] = mul_hi(a[j + 1], bi);
    }
    uint32_t m = mul_lo(e[0], P::INV);
    e[0] = mad_lo_cc(P::MOD(0), m, e[0]);
    e[1] = madc_hi_cc(P::MOD(0), m, e[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      e[j] = madc_lo_cc(P::MOD(j), m, e[j]);
      e[j + 1] = madc_hi_cc(P::MOD(j), m, e[j + 1]);
    }
    o[N - 1] = addc(o[N - 1], 0u);
    o[0] = mad_lo_cc(P::MOD(1), m, o[0]);
    o[1] = madc_hi_cc(P::MOD(1), m, o[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      o[j] = madc_lo_cc(P::MOD(j + 1), m, o[j]);
      o[j + 1] = madc_hi_cc(P::MOD(j + 1), m, o[j + 1]);
    }
  }
  // ---- rows 1..N-1
#pragma unroll
  for (int i = 1; i < N; i++) {
    uint32_t bi = b[i];
    uint32_t E[N], O[N];
    // T/2^32: column 0 of the quotient = o[0] + e[1]; its carry enters the O chain (column 1)
    E[0] = add_cc(o[0], e[1]);
#pragma unroll
    for (int j = 0; j < N; j += 2) {  // O[k] = e[k+2] + odd-j products, carry-in from above
      O[j] = madc_lo_cc(a[j + 1], bi, (j + 2 < N) ? e[j + 2] : 0u);
      O[j + 1] = madc_hi_cc(a[j + 1], bi, (j + 3 < N) ? e[j + 3] : 0u);
    }
    // (top limb got hi(a*b) + carry <= 2^32-1: no carry out)
    E[0] = mad_lo_cc(a[0], bi, E[0]);
    E[1] = madc_hi_cc(a[0], bi, o[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      E[j] = madc_lo_cc(a[j], bi, o[j]);
      E[j + 1] = madc_hi_cc(a[j], bi, o[j + 1]);
    }
    O[N - 1] = addc(O[N - 1], 0u);
    // reduction row
    uint32_t m = mul_lo(E[0], P::INV);
    E[0] = mad_lo_cc(P::MOD(0), m, E[0]);
    E[1] = madc_hi_cc(P::MOD(0), m, E[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      E[j] = madc_lo_cc(P::MOD(j), m, E[j]);
      E[j + 1] = madc_hi_cc(P::MOD(j), m, E[j + 1]);
    }
    O[N - 1] = addc(O[N - 1], 0u);
    O[0] = mad_lo_cc(P::MOD(1), m, O[0]);
    O[1] = madc_hi_cc(P::MOD(1), m, O[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      O[j] = madc_lo_cc(P::MOD(j + 1), m, O[j]);
      O[j + 1] = madc_hi_cc(P::MOD(j + 1), m, O[j + 1]);
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
      e[j] = E[j];
      o[j] = O[j];
    }
  }
  // ---- result = T / 2^32 = o + (e >> 32)      (e[0] == 0)
  Fp<P> r;
  r.v[0] = add_cc(o[0], e[1]);
#pragma unroll
  for (int j = 1; j < N - 1; j++) r.v[j] = addc_cc(o[j], e[j + 1]);
  r.v[N - 1] = addc(o[N - 1], 0u);
  final_sub<P>(r.v);
  return r;
}

// (a*u + b*v) * R^-1 mod p with ONE interleaved reduction: the "lazy" half of an Fp2 product (c0 = a0 b0 + (p - a1) b1,
// c1 = a0 b1 + a1 b0).  3 N^2 + N wide multiply-adds instead of 2 (2 N^2 + N) for two products, and no separate field
// addition.  Same even/odd column accumulators as mont_mul with a third product chain per row.
// Operands may be anything <= p (p itself is allowed, so the caller may negate without a zero test).  Needs p < 2^(32N-2):
// then every running sum T + a u_i + b v_i + m p < 3p (2^32 + 2) fits the two N-limb accumulators (no carry leaves the top
// limb), and the result (a u + b v + M p) / R < p (2p / R + 1) < 2p is reduced by one conditional subtraction.
// (the limbs u_i, v_i are asked for row by row -- rows(i, ui, vi) -- so that a caller which has to fetch them from another lane
// keeps only one limb of each alive; mont_dot2 below is the plain-operand form)
template <class P, class Rows>
ZKB_HD Fp<P> mont_dot2_rows(const Fp<P>& A, const Fp<P>& B, Rows rows) {
  constexpr int N = P::N;
  static_assert(N % 2 == 0, "even limb count required");
  static_assert((P::MOD(N - 1) >> 30) == 0, "mont_dot2 needs two spare bits in the top limb");
  const uint32_t* a = A.v;
  const uint32_t* b = B.v;
  uint32_t e[N], o[N];
  // ---- row 0: T = a * u[0] + b * v[0]
  {
    uint32_t ui, vi;
    rows(0, ui, vi);
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      e[j] = mul_lo(a[j], ui);
      e[j + 1] = mul_hi(a[j], ui);
      o[j] = mul_lo(a[j + 1], ui);
      o[j + 1] = mul_hi(a[j + 1], ui);
    }
    e[0] = mad_lo_cc(b[0], vi, e[0]);
    e[1] = madc_hi_cc(b[0], vi, e[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      e[j] = madc_lo_cc(b[j], vi, e[j]);
      e[j + 1] = madc_hi_cc(b[j], vi, e[j + 1]);
    }
    o[N - 1] = addc(o[N - 1], 0u);
    o[0] = mad_lo_cc(b[1], vi, o[0]);
    o[1] = madc_hi_cc(b[1], vi, o[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      o[j] = madc_lo_cc(b[j + 1], vi, o[j]);
      o[j + 1] = madc_hi_cc(b[j + 1], vi, o[j + 1]);
    }
    uint32_t m = mul_lo(e[0], P::INV);
    e[0] = mad_lo_cc(P::MOD(0), m, e[0]);
    e[1] = madc_hi_cc(P::MOD(0), m, e[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      e[j] = madc_lo_cc(P::MOD(j), m, e[j]);
      e[j + 1] = madc_hi_cc(P::MOD(j), m, e[j + 1]);
    }
    o[N - 1] = addc(o[N - 1], 0u);
    o[0] = mad_lo_cc(P::MOD(1), m, o[0]);
    o[1] = madc_hi_cc(P::MOD(1), m, o[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      o[j] = madc_lo_cc(P::MOD(j + 1), m, o[j]);
      o[j + 1] = madc_hi_cc(P::MOD(j + 1), m, o[j + 1]);
    }
  }
  // ---- rows 1..N-1
#pragma unroll
  for (int i = 1; i < N; i++) {
    uint32_t ui, vi;
    rows(i, ui, vi);
    uint32_t E[N], O[N];
    // T / 2^32 (see mont_mul), then += a * u[i]
    E[0] = add_cc(o[0], e[1]);
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      O[j] = madc_lo_cc(a[j + 1], ui, (j + 2 < N) ? e[j + 2] : 0u);
      O[j + 1] = madc_hi_cc(a[j + 1], ui, (j + 3 < N) ? e[j + 3] : 0u);
    }
    E[0] = mad_lo_cc(a[0], ui, E[0]);
    E[1] = madc_hi_cc(a[0], ui, o[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      E[j] = madc_lo_cc(a[j], ui, o[j]);
      E[j + 1] = madc_hi_cc(a[j], ui, o[j + 1]);
    }
    O[N - 1] = addc(O[N - 1], 0u);
    // += b * v[i]
    E[0] = mad_lo_cc(b[0], vi, E[0]);
    E[1] = madc_hi_cc(b[0], vi, E[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      E[j] = madc_lo_cc(b[j], vi, E[j]);
      E[j + 1] = madc_hi_cc(b[j], vi, E[j + 1]);
    }
    O[N - 1] = addc(O[N - 1], 0u);
    O[0] = mad_lo_cc(b[1], vi, O[0]);
    O[1] = madc_hi_cc(b[1], vi, O[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      O[j] = madc_lo_cc(b[j + 1], vi, O[j]);
      O[j + 1] = madc_hi_cc(b[j + 1], vi, O[j + 1]);
    }
    // reduction row
    uint32_t m = mul_lo(E[0], P::INV);
    E[0] = mad_lo_cc(P::MOD(0), m, E[0]);
    E[1] = madc_hi_cc(P::MOD(0), m, E[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      E[j] = madc_lo_cc(P::MOD(j), m, E[j]);
      E[j + 1] = madc_hi_cc(P::MOD(j), m, E[j + 1]);
    }
    O[N - 1] = addc(O[N - 1], 0u);
    O[0] = mad_lo_cc(P::MOD(1), m, O[0]);
    O[1] = madc_hi_cc(P::MOD(1), m, O[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      O[j] = madc_lo_cc(P::MOD(j + 1), m, O[j]);
      O[j + 1] = madc_hi_cc(P::MOD(j + 1), m, O[j + 1]);
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
      e[j] = E[j];
      o[j] = O[j];
    }
  }
  Fp<P> r;
  r.v[0] = add_cc(o[0], e[1]);
#pragma unroll
  for (int j = 1; j < N - 1; j++) r.v[j] = addc_cc(o[j], e[j + 1]);
  r.v[N - 1] = addc(o[N - 1], 0u);
  final_sub<P>(r.v);
  return r;
}
template <class P>
ZKB_HD Fp<P> mont_dot2(const Fp<P>& A, const Fp<P>& U, const Fp<P>& B, const Fp<P>& V) {
  return mont_dot2_rows<P>(A, B, [&](int i, uint32_t& ui, uint32_t& vi) { ui = U.v[i]; vi = V.v[i]; });
}
// p - a without reducing (a <= p; a = 0 gives p): the negated operand of mont_dot2
template <class P>
ZKB_HD Fp<P> neg_lazy(const Fp<P>& a) {
  constexpr int N = P::N;
  Fp<P> r;
  r.v[0] = sub_cc(P::MOD(0), a.v[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.v[i] = subc_cc(P::MOD(i), a.v[i]);
  r.v[N - 1] = subc(P::MOD(N - 1), a.v[N - 1]);
  return r;
}

// hook for the experimental split multiplier of tools/ff_wide.cuh (measured slower; no product kernel opts in)
template <class P>
ZKB_HD Fp<P> mont_mul_split(const Fp<P>& a, const Fp<P>& b);
// a parameter struct opts in with `static constexpr bool SPLIT_MUL = true;`
template <class P, class = void>
struct uses_split_mul { static constexpr bool value = false; };
template <class P>
struct uses_split_mul<P, decltype((void)P::SPLIT_MUL)> { static constexpr bool value = P::SPLIT_MUL; };

// The base fields of the curves (Fq) call the multiplier out of line on the device: a fully inlined XYZZ addition
// would be 10-40 copies of a 300-600 instruction body, far beyond the instruction cache, and ptxas time explodes.
// The scalar fields (Fr, one multiply per butterfly) keep it inline.
#if defined(__CUDA_ARCH__)
template <class P>
__device__ __noinline__ Fp<P> mont_mul_call(Fp<P> a, Fp<P> b) {
  if constexpr (uses_split_mul<P>::value) return mont_mul_split(a, b);
  else return mont_mul(a, b);
}
#endif
template <class P>
ZKB_HD Fp<P> operator*(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__)
  if constexpr (P::NOINLINE_MUL) return mont_mul_call<P>(a, b);
  else if constexpr (uses_split_mul<P>::value) return mont_mul_split(a, b);
  else return mont_mul(a, b);
#else
  if constexpr (uses_split_mul<P>::value) return mont_mul_split(a, b);
  else return mont_mul(a, b);
#endif
}

// hook for the experimental dedicated squaring of tools/ff_wide.cuh (measured: no gain in the accumulation kernels, DESIGN.md
// section 3); a parameter struct would opt in with `static constexpr bool SPLIT_SQR = true;` -- none does
template <class P>
ZKB_HD Fp<P> mont_sqr_split(const Fp<P>& a);
template <class P, class = void>
struct uses_split_sqr { static constexpr bool value = false; };
template <class P>
struct uses_split_sqr<P, decltype((void)P::SPLIT_SQR)> { static constexpr bool value = P::SPLIT_SQR; };
// A parameter struct with `static constexpr bool NOINLINE_SQR = true;` keeps its products inline but calls ONE shared body for
// the squarings: two of the ten multiplier bodies of a mixed addition leave the loop, which brings it under the 32 KB L1.5
// instruction cache (msm_host.cuh).
template <class P, class = void>
struct sqr_out_of_line { static constexpr bool value = false; };
template <class P>
struct sqr_out_of_line<P, decltype((void)P::NOINLINE_SQR)> { static constexpr bool value = P::NOINLINE_SQR; };
#if defined(__CUDA_ARCH__)
template <class P>
__device__ __noinline__ Fp<P> mont_sqr_call(Fp<P> a) { return mont_mul(a, a); }
#endif
template <class P>
ZKB_HD Fp<P> sqr(const Fp<P>& a) {
  if constexpr (uses_split_sqr<P>::value) return mont_sqr_split(a);
  else {
#if defined(__CUDA_ARCH__)
    if constexpr (sqr_out_of_line<P>::value) return mont_sqr_call<P>(a);
#endif
    return a * a;
  }
}

// canonical <-> Montgomery
template <class P>
ZKB_HD Fp<P> to_mont(const Fp<P>& a) { return a * Fp<P>::r2(); }
template <class P>
ZKB_HD Fp<P> from_mont(const Fp<P>& a) {
  Fp<P> one_raw = Fp<P>::zero();
  one_raw.v[0] = 1;
  return a * one_raw;
}

// a^e for a little-endian limb exponent (variable time; used for inversion and root powers)
template <class P>
ZKB_HD Fp<P> pow_limbs(const Fp<P>& a, const uint32_t* e, int nlimbs) {
  Fp<P> r = Fp<P>::one();
  bool started = false;
  for (int i = nlimbs - 1; i >= 0; i--) {
    for (int bit = 31; bit >= 0; bit--) {
      if (started) r = sqr(r);
      if ((e[i] >> bit) & 1) {
        r = started ? r * a : a;
        started = true;
      }
    }
  }
  return r;
}

template <class P>
ZKB_HD Fp<P> pow_u64(const Fp<P>& a, uint64_t e) {
  uint32_t l[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
  return pow_limbs(a, l, 2);
}

// Fermat inversion a^(p-2); inv(0) = 0
template <class P>
ZKB_HD Fp<P> inv(const Fp<P>& a) {
  uint32_t e[P::N];
#pragma unroll
  for (int i = 0; i < P::N; i++) e[i] = P::PM2(i);
  return pow_limbs(a, e, P::N);
}

// a * k for a small non-negative k (double-and-add on the field adder)
template <class P>
ZKB_HD Fp<P> mul_small(const Fp<P>& a, uint32_t k) {
  Fp<P> r = Fp<P>::zero();
  Fp<P> t = a;
  while (k) {
    if (k & 1) r = r + t;
    t = t + t;
    k >>= 1;
  }
  return r;
}

typedef Fp<FrBN254> fr_bn;
typedef Fp<FqBN254> fq_bn;
typedef Fp<FrBLS381> fr_bls;
typedef Fp<FqBLS381> fq_bls;

// ----------------------------------------------------------------------------------------------------
// Fp2 = Fp[u]/(u^2+1)  (both curves' G2 coordinate field)
// ----------------------------------------------------------------------------------------------------
template <class P>
struct Fp2 {
  typedef P Params;
  Fp<P> c0, c1;
  ZKB_HD static Fp2 zero() { Fp2 r; r.c0 = Fp<P>::zero(); r.c1 = Fp<P>::zero(); return r; }
  ZKB_HD static Fp2 one() { Fp2 r; r.c0 = Fp<P>::one(); r.c1 = Fp<P>::zero(); return r; }
  ZKB_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  ZKB_HD bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
  ZKB_HD bool operator!=(const Fp2& o) const { return !(*this == o); }
};
template <class P>
ZKB_HD Fp2<P> operator+(const Fp2<P>& a, const Fp2<P>& b) { Fp2<P> r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
template <class P>
ZKB_HD Fp2<P> operator-(const Fp2<P>& a, const Fp2<P>& b) { Fp2<P> r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
template <class P>
ZKB_HD Fp2<P> neg(const Fp2<P>& a) { Fp2<P> r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); return r; }
template <class P>
ZKB_HD Fp2<P> dbl(const Fp2<P>& a) { return a + a; }
template <class P>
ZKB_HD Fp2<P> operator*(const Fp2<P>& a, const Fp2<P>& b) {
  // Karatsuba: 3 base-field products
  Fp<P> t0 = a.c0 * b.c0;
  Fp<P> t1 = a.c1 * b.c1;
  Fp<P> t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
  Fp2<P> r;
  r.c0 = t0 - t1;
  r.c1 = t2 - t0 - t1;
  return r;
}
template <class P>
ZKB_HD Fp2<P> sqr(const Fp2<P>& a) {
  // (a0+a1)(a0-a1), 2 a0 a1
  Fp<P> s = a.c0 + a.c1;
  Fp<P> d = a.c0 - a.c1;
  Fp<P> m = a.c0 * a.c1;
  Fp2<P> r;
  r.c0 = s * d;
  r.c1 = m + m;
  return r;
}
template <class P>
ZKB_HD Fp2<P> inv(const Fp2<P>& a) {
  Fp<P> d = inv(sqr(a.c0) + sqr(a.c1));
  Fp2<P> r;
  r.c0 = a.c0 * d;
  r.c1 = neg(a.c1 * d);
  return r;
}
template <class P>
ZKB_HD Fp2<P> to_mont(const Fp2<P>& a) { Fp2<P> r; r.c0 = to_mont(a.c0); r.c1 = to_mont(a.c1); return r; }
template <class P>
ZKB_HD Fp2<P> from_mont(const Fp2<P>& a) { Fp2<P> r; r.c0 = from_mont(a.c0); r.c1 = from_mont(a.c1); return r; }

typedef Fp2<FqBN254> fq2_bn;
typedef Fp2<FqBLS381> fq2_bls;

}  // namespace zkb

#ifdef ZKB_EXPERIMENTAL_WIDE   // tools/ffbench.cu only: the measured-and-rejected split multiplier / dedicated squaring (tools/ff_wide.cuh)
#include "ff_wide.cuh"
#endif
