// ff_host.h -- host-only Montgomery field with 64-bit limbs (unsigned __int128 products).
//
// Same memory layout as Fp<P> (little-endian limbs, Montgomery form with R = 2^(32 N)), so device results can be
// reinterpreted directly.  Used by the host side of the C-ABI for the handful of scalar operations that follow
// a kernel (window recombination of an MSM, final affine conversion, proof assembly) where a kernel launch would
// cost more than the arithmetic.  Plugs into the ec.cuh templates (same operator set as Fp<P>).
#pragma once
#include <stdint.h>
#include <string.h>
#include "ff.cuh"

namespace zkb {

template <class P>
struct Fh {
  static constexpr int N = P::N / 2;  // 64-bit limbs
  typedef P Params;
  uint64_t v[N];

  static uint64_t mod(int i) { return (uint64_t)P::MOD(2 * i) | ((uint64_t)P::MOD(2 * i + 1) << 32); }
  static uint64_t inv64() {
    // -p^-1 mod 2^64 by Newton iteration from the 32-bit constant
    uint64_t p0 = mod(0);
    uint64_t x = (uint64_t)(0u - P::INV);  // p^-1 mod 2^32
    x *= 2 - p0 * x;                       // mod 2^64
    return 0 - x;
  }
  static Fh zero() { Fh r; memset(r.v, 0, sizeof(r.v)); return r; }
  static Fh one() {
    Fh r;
    for (int i = 0; i < N; i++) r.v[i] = (uint64_t)P::R1(2 * i) | ((uint64_t)P::R1(2 * i + 1) << 32);
    return r;
  }
  static Fh r2() {
    Fh r;
    for (int i = 0; i < N; i++) r.v[i] = (uint64_t)P::R2(2 * i) | ((uint64_t)P::R2(2 * i + 1) << 32);
    return r;
  }
  bool is_zero() const {
    uint64_t t = 0;
    for (int i = 0; i < N; i++) t |= v[i];
    return t == 0;
  }
  bool operator==(const Fh& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
  bool operator!=(const Fh& o) const { return !(*this == o); }
};

template <class P>
inline bool geq_mod(const uint64_t* t) {
  for (int i = Fh<P>::N - 1; i >= 0; i--) {
    uint64_t m = Fh<P>::mod(i);
    if (t[i] != m) return t[i] > m;
  }
  return true;
}
template <class P>
inline void sub_mod(uint64_t* t) {
  unsigned __int128 br = 0;
  for (int i = 0; i < Fh<P>::N; i++) {
    unsigned __int128 d = (unsigned __int128)t[i] - Fh<P>::mod(i) - (uint64_t)br;
    t[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
}

template <class P>
inline Fh<P> operator+(const Fh<P>& a, const Fh<P>& b) {
  Fh<P> r;
  unsigned __int128 c = 0;
  for (int i = 0; i < Fh<P>::N; i++) {
    c += (unsigned __int128)a.v[i] + b.v[i];
    r.v[i] = (uint64_t)c;
    c >>= 64;
  }
  if (geq_mod<P>(r.v)) sub_mod<P>(r.v);
  return r;
}
template <class P>
inline Fh<P> operator-(const Fh<P>& a, const Fh<P>& b) {
  Fh<P> r;
  uint64_t br = 0;
  for (int i = 0; i < Fh<P>::N; i++) {
    unsigned __int128 d = (unsigned __int128)a.v[i] - b.v[i] - br;
    r.v[i] = (uint64_t)d;
    br = (uint64_t)(d >> 64) & 1;
  }
  if (br) {
    unsigned __int128 c = 0;
    for (int i = 0; i < Fh<P>::N; i++) {
      c += (unsigned __int128)r.v[i] + Fh<P>::mod(i);
      r.v[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  return r;
}
template <class P>
inline Fh<P> neg(const Fh<P>& a) { return a.is_zero() ? a : (Fh<P>::zero() - a); }
template <class P>
inline Fh<P> dbl(const Fh<P>& a) { return a + a; }

template <class P>
inline Fh<P> operator*(const Fh<P>& a, const Fh<P>& b) {
  constexpr int N = Fh<P>::N;
  static const uint64_t ninv = Fh<P>::inv64();
  uint64_t t[N + 2];
  memset(t, 0, sizeof(t));
  for (int i = 0; i < N; i++) {
    unsigned __int128 c = 0;
    for (int j = 0; j < N; j++) {
      c += (unsigned __int128)a.v[j] * b.v[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[N];
    t[N] = (uint64_t)c;
    t[N + 1] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * ninv;
    c = (unsigned __int128)m * Fh<P>::mod(0) + t[0];
    c >>= 64;
    for (int j = 1; j < N; j++) {
      c += (unsigned __int128)m * Fh<P>::mod(j) + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[N];
    t[N - 1] = (uint64_t)c;
    t[N] = t[N + 1] + (uint64_t)(c >> 64);
  }
  Fh<P> r;
  if (t[N] || geq_mod<P>(t)) sub_mod<P>(t);
  memcpy(r.v, t, sizeof(r.v));
  return r;
}
template <class P>
inline Fh<P> sqr(const Fh<P>& a) { return a * a; }
template <class P>
inline Fh<P> to_mont(const Fh<P>& a) { return a * Fh<P>::r2(); }
template <class P>
inline Fh<P> from_mont(const Fh<P>& a) {
  Fh<P> o = Fh<P>::zero();
  o.v[0] = 1;
  return a * o;
}
template <class P>
inline Fh<P> inv(const Fh<P>& a) {
  Fh<P> r = Fh<P>::one();
  bool started = false;
  for (int i = P::N - 1; i >= 0; i--) {
    uint32_t e = P::PM2(i);
    for (int bit = 31; bit >= 0; bit--) {
      if (started) r = sqr(r);
      if ((e >> bit) & 1) {
        r = started ? r * a : a;
        started = true;
      }
    }
  }
  return r;
}

// quadratic extension on the host type
template <class P>
struct Fh2 {
  typedef P Params;
  Fh<P> c0, c1;
  static Fh2 zero() { Fh2 r; r.c0 = Fh<P>::zero(); r.c1 = Fh<P>::zero(); return r; }
  static Fh2 one() { Fh2 r; r.c0 = Fh<P>::one(); r.c1 = Fh<P>::zero(); return r; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fh2& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fh2& o) const { return !(*this == o); }
};
template <class P>
inline Fh2<P> operator+(const Fh2<P>& a, const Fh2<P>& b) { Fh2<P> r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
template <class P>
inline Fh2<P> operator-(const Fh2<P>& a, const Fh2<P>& b) { Fh2<P> r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
template <class P>
inline Fh2<P> neg(const Fh2<P>& a) { Fh2<P> r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); return r; }
template <class P>
inline Fh2<P> dbl(const Fh2<P>& a) { return a + a; }
template <class P>
inline Fh2<P> operator*(const Fh2<P>& a, const Fh2<P>& b) {
  Fh<P> t0 = a.c0 * b.c0, t1 = a.c1 * b.c1, t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
  Fh2<P> r;
  r.c0 = t0 - t1;
  r.c1 = t2 - t0 - t1;
  return r;
}
template <class P>
inline Fh2<P> sqr(const Fh2<P>& a) { return a * a; }
template <class P>
inline Fh2<P> inv(const Fh2<P>& a) {
  Fh<P> d = inv(sqr(a.c0) + sqr(a.c1));
  Fh2<P> r;
  r.c0 = a.c0 * d;
  r.c1 = neg(a.c1 * d);
  return r;
}
template <class P>
inline Fh2<P> to_mont(const Fh2<P>& a) { Fh2<P> r; r.c0 = to_mont(a.c0); r.c1 = to_mont(a.c1); return r; }
template <class P>
inline Fh2<P> from_mont(const Fh2<P>& a) { Fh2<P> r; r.c0 = from_mont(a.c0); r.c1 = from_mont(a.c1); return r; }

}  // namespace zkb
