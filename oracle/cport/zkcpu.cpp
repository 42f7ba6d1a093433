// zkcpu.cpp -- CPU restatement ("port") of the reference's proving hot path -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.
//
// The reference (Merricx/zksnake) performs this arithmetic inside third-party crates that are NOT vendored under
// /root/reference and cannot be built here (no Rust toolchain): ark-poly 0.4.2 (Cargo.lock:103-104), ark-ec 0.4.2 (:41-42),
// ark-ff 0.4.2 (:59-60), ark-bn254 0.4.0 (:30-31), ark-bls12-381 0.4.0 (:18-19).  This file restates the PUBLISHED algorithms
// of those crates at the reference's call sites, so that (a) results can be cross-checked bit-for-bit against the pure-Python
// oracle (oracle/*.py) and the CUDA product, and (b) bench.py has a CPU arm to time next to the GPU ("kind": "port").
//
//   zkcpu_fft            src/bn254/polynomial.rs:536-585 fft / coset_fft / ifft / coset_ifft
//                        -> ark-poly Radix2EvaluationDomain: forward = DIF butterflies then bit-reversal, inverse =
//                        bit-reversal then DIT butterflies then * N^-1; coset = distribute powers of the offset
//   zkcpu_msm            src/bn254/curve.rs:356-392 multiscalar_mul_g1/g2 -> ark-ec VariableBaseMSM::msm
//                        (msm_bigint_wnaf: signed radix-2^c digits, c = 3 if n < 32 else ceil(log2 n)*69/100 + 2, 2^c Jacobian
//                        buckets per window, running-sum reduction, windows recombined by c doublings); the windows are the
//                        parallel loop exactly as under ark's `parallel` feature (OpenMP here, rayon there)
//   zkcpu_groth16_h      python/zksnake/groth16/qap.py:42-71 QAP.evaluate_witness: 3 ifft(n), mul_over_fft on the 2n domain
//                        (python/zksnake/polynomial.py:126-165), subtract, divide_by_vanishing_poly (polynomial.rs:466-489)
//   zkcpu_groth16_prove  python/zksnake/groth16/protocol.py:115-165 Groth16.prove (SparseArray.dot of array.py:36-43 included)
//
// PARITY STATUS: "parity unpinned" at the numeric level (see oracle/__init__.py): the reference has no golden vectors for
// this path; this port is pinned to the Python oracle, to the published constants and to algebraic identities in tests/.
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load the library built from this
// file.  The product (zksnake_b200/) never does.  Independent of the product's sources: 64-bit limbs with unsigned
// __int128 CIOS Montgomery, Jacobian coordinates (the product uses 32-bit limbs and XYZZ).
//
// Build: make -C oracle/cport   (g++ -O3 -march=native -fopenmp -shared)
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

// ------------------------------------------------------------------------------------------------------------------
// prime fields: N x 64-bit limbs, Montgomery form.  Constants derived at start-up from the modulus alone.
// ------------------------------------------------------------------------------------------------------------------
template <int N_>
struct FieldCtx {
  static constexpr int N = N_;
  u64 p[N_];
  u64 inv;      // -p^-1 mod 2^64
  u64 r1[N_];   // R mod p
  u64 r2[N_];   // R^2 mod p
  u64 pm2[N_];  // p - 2
};

template <int N>
static inline bool geq(const u64* a, const u64* b) {
  for (int i = N - 1; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return true;
}
template <int N>
static inline u64 sub_n(u64* r, const u64* a, const u64* b) {
  u64 borrow = 0;
  for (int i = 0; i < N; i++) {
    u128 d = (u128)a[i] - b[i] - borrow;
    r[i] = (u64)d;
    borrow = (u64)(d >> 64) & 1;
  }
  return borrow;
}
template <int N>
static inline u64 add_n(u64* r, const u64* a, const u64* b) {
  u64 carry = 0;
  for (int i = 0; i < N; i++) {
    u128 s = (u128)a[i] + b[i] + carry;
    r[i] = (u64)s;
    carry = (u64)(s >> 64);
  }
  return carry;
}

template <class C>
struct Fp {
  static constexpr int N = C::N;
  u64 v[C::N];
  static inline Fp zero() { Fp r; memset(r.v, 0, sizeof(r.v)); return r; }
  static inline Fp one() { Fp r; memcpy(r.v, C::ctx.r1, sizeof(r.v)); return r; }
  inline bool is_zero() const { u64 t = 0; for (int i = 0; i < N; i++) t |= v[i]; return t == 0; }
  inline bool operator==(const Fp& o) const { u64 t = 0; for (int i = 0; i < N; i++) t |= v[i] ^ o.v[i]; return t == 0; }
  inline bool operator!=(const Fp& o) const { return !(*this == o); }
};

template <class C>
static inline Fp<C> operator+(const Fp<C>& a, const Fp<C>& b) {
  Fp<C> r;
  u64 carry = add_n<C::N>(r.v, a.v, b.v);
  if (carry || geq<C::N>(r.v, C::ctx.p)) sub_n<C::N>(r.v, r.v, C::ctx.p);
  return r;
}
template <class C>
static inline Fp<C> operator-(const Fp<C>& a, const Fp<C>& b) {
  Fp<C> r;
  if (sub_n<C::N>(r.v, a.v, b.v)) add_n<C::N>(r.v, r.v, C::ctx.p);
  return r;
}
template <class C>
static inline Fp<C> neg(const Fp<C>& a) {
  if (a.is_zero()) return a;
  Fp<C> r;
  sub_n<C::N>(r.v, C::ctx.p, a.v);
  return r;
}
template <class C>
static inline Fp<C> dbl(const Fp<C>& a) { return a + a; }

// Montgomery product, CIOS with the two carry chains (a*b_i and m*p) fused per row.  All four moduli leave the top bit of the
// top limb clear, so the running value stays below 2p and needs no extra carry word ("no-carry" variant, as ark-ff does).
template <class C>
static inline Fp<C> operator*(const Fp<C>& a, const Fp<C>& b) {
  constexpr int N = C::N;
  const u64* p = C::ctx.p;
  const u64 ninv = C::ctx.inv;
  u64 t[N];
  for (int i = 0; i < N; i++) t[i] = 0;
  for (int i = 0; i < N; i++) {
    u128 s = (u128)a.v[0] * b.v[i] + t[0];
    u64 m = (u64)s * ninv;
    u64 c1 = (u64)(s >> 64);
    u128 q = (u128)m * p[0] + (u64)s;
    u64 c2 = (u64)(q >> 64);
    for (int j = 1; j < N; j++) {
      s = (u128)a.v[j] * b.v[i] + t[j] + c1;
      c1 = (u64)(s >> 64);
      q = (u128)m * p[j] + (u64)s + c2;
      c2 = (u64)(q >> 64);
      t[j - 1] = (u64)q;
    }
    t[N - 1] = c1 + c2;
  }
  Fp<C> r;
  u64 d[N];
  u64 borrow = sub_n<N>(d, t, p);
  for (int i = 0; i < N; i++) r.v[i] = borrow ? t[i] : d[i];
  return r;
}
template <class C>
static inline Fp<C> sqr(const Fp<C>& a) { return a * a; }
template <class C>
static inline Fp<C> to_mont(const Fp<C>& a) {
  Fp<C> r2;
  memcpy(r2.v, C::ctx.r2, sizeof(r2.v));
  return a * r2;
}
template <class C>
static inline Fp<C> from_mont(const Fp<C>& a) {
  Fp<C> o = Fp<C>::zero();
  o.v[0] = 1;
  return a * o;
}
template <class C>
static Fp<C> pow_limbs(const Fp<C>& a, const u64* e, int n) {
  Fp<C> r = Fp<C>::one();
  for (int i = n - 1; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      r = sqr(r);
      if ((e[i] >> b) & 1) r = r * a;
    }
  return r;
}
template <class C>
static Fp<C> inv(const Fp<C>& a) { return pow_limbs(a, C::ctx.pm2, C::N); }  // Fermat; inv(0) = 0

// start-up derivation of the Montgomery constants
template <int N>
static void ctx_init(FieldCtx<N>& c, const u64* modulus) {
  memcpy(c.p, modulus, sizeof(c.p));
  u64 x = 1;  // Newton iteration for p^-1 mod 2^64
  for (int i = 0; i < 6; i++) x *= 2 - modulus[0] * x;
  c.inv = (u64)0 - x;
  // R mod p by 64*N doublings of 1, R^2 by 64*N more
  u64 t[N];
  memset(t, 0, sizeof(t));
  t[0] = 1;
  for (int i = 0; i < 2 * 64 * N; i++) {
    u64 carry = add_n<N>(t, t, t);
    if (carry || geq<N>(t, c.p)) sub_n<N>(t, t, c.p);
    if (i == 64 * N - 1) memcpy(c.r1, t, sizeof(t));
  }
  memcpy(c.r2, t, sizeof(t));
  u64 two[N];
  memset(two, 0, sizeof(two));
  two[0] = 2;
  sub_n<N>(c.pm2, c.p, two);
}

struct FrBN { static constexpr int N = 4; static FieldCtx<4> ctx; };
struct FqBN { static constexpr int N = 4; static FieldCtx<4> ctx; };
struct FrBLS { static constexpr int N = 4; static FieldCtx<4> ctx; };
struct FqBLS { static constexpr int N = 6; static FieldCtx<6> ctx; };
FieldCtx<4> FrBN::ctx, FqBN::ctx, FrBLS::ctx;
FieldCtx<6> FqBLS::ctx;

// moduli: /root/reference/python/zksnake/constant.py:5-15
static const u64 FR_BN[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const u64 FQ_BN[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const u64 FR_BLS[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
static const u64 FQ_BLS[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                              0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};

static bool g_ready = false;
static void ensure_init() {
  if (g_ready) return;
  ctx_init(FrBN::ctx, FR_BN);
  ctx_init(FqBN::ctx, FQ_BN);
  ctx_init(FrBLS::ctx, FR_BLS);
  ctx_init(FqBLS::ctx, FQ_BLS);
  g_ready = true;
}

// Fp2 = Fp[u]/(u^2+1)
template <class C>
struct Fp2 {
  Fp<C> c0, c1;
  static inline Fp2 zero() { Fp2 r; r.c0 = Fp<C>::zero(); r.c1 = Fp<C>::zero(); return r; }
  static inline Fp2 one() { Fp2 r; r.c0 = Fp<C>::one(); r.c1 = Fp<C>::zero(); return r; }
  inline bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  inline bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
  inline bool operator!=(const Fp2& o) const { return !(*this == o); }
};
template <class C> static inline Fp2<C> operator+(const Fp2<C>& a, const Fp2<C>& b) { Fp2<C> r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
template <class C> static inline Fp2<C> operator-(const Fp2<C>& a, const Fp2<C>& b) { Fp2<C> r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
template <class C> static inline Fp2<C> neg(const Fp2<C>& a) { Fp2<C> r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); return r; }
template <class C> static inline Fp2<C> dbl(const Fp2<C>& a) { return a + a; }
template <class C> static inline Fp2<C> operator*(const Fp2<C>& a, const Fp2<C>& b) {
  Fp<C> v0 = a.c0 * b.c0, v1 = a.c1 * b.c1;
  Fp2<C> r;
  r.c1 = (a.c0 + a.c1) * (b.c0 + b.c1) - v0 - v1;
  r.c0 = v0 - v1;
  return r;
}
template <class C> static inline Fp2<C> sqr(const Fp2<C>& a) {
  Fp<C> m = a.c0 * a.c1;
  Fp2<C> r;
  r.c0 = (a.c0 + a.c1) * (a.c0 - a.c1);
  r.c1 = m + m;
  return r;
}
template <class C> static Fp2<C> inv(const Fp2<C>& a) {
  Fp<C> d = inv(sqr(a.c0) + sqr(a.c1));
  Fp2<C> r;
  r.c0 = a.c0 * d;
  r.c1 = neg(a.c1 * d);
  return r;
}
template <class C> static inline Fp2<C> to_mont(const Fp2<C>& a) { Fp2<C> r; r.c0 = to_mont(a.c0); r.c1 = to_mont(a.c1); return r; }
template <class C> static inline Fp2<C> from_mont(const Fp2<C>& a) { Fp2<C> r; r.c0 = from_mont(a.c0); r.c1 = from_mont(a.c1); return r; }

// ------------------------------------------------------------------------------------------------------------------
// curves y^2 = x^3 + b, Jacobian coordinates (what ark-ec's short_weierstrass::Projective uses)
// ------------------------------------------------------------------------------------------------------------------
template <class F>
struct Aff {
  F x, y;  // (0, 0) = infinity on the wire
  inline bool is_inf() const { return x.is_zero() && y.is_zero(); }
};
template <class F>
struct Jac {
  F X, Y, Z;  // Z == 0 = identity
  static inline Jac inf() { Jac r; r.X = F::one(); r.Y = F::one(); r.Z = F::zero(); return r; }
  inline bool is_inf() const { return Z.is_zero(); }
};

template <class F>
static inline void jac_double(Jac<F>& p) {  // dbl-2009-l, a = 0
  if (p.is_inf()) return;
  F A = sqr(p.X), B = sqr(p.Y), C = sqr(B);
  F D = sqr(p.X + B) - A - C;
  D = dbl(D);
  F E = dbl(A) + A;
  F Fq = sqr(E);
  F Z3 = dbl(p.Y * p.Z);
  F X3 = Fq - dbl(D);
  F C8 = dbl(dbl(dbl(C)));
  p.Y = E * (D - X3) - C8;
  p.X = X3;
  p.Z = Z3;
}
template <class F>
static inline void jac_add_affine(Jac<F>& p, const Aff<F>& q, bool negate) {  // madd-2007-bl
  if (q.is_inf()) return;
  F qy = negate ? neg(q.y) : q.y;
  if (p.is_inf()) {
    p.X = q.x; p.Y = qy; p.Z = F::one();
    return;
  }
  F Z1Z1 = sqr(p.Z);
  F U2 = q.x * Z1Z1;
  F S2 = qy * p.Z * Z1Z1;
  if (U2 == p.X) {
    if (S2 == p.Y) jac_double(p);
    else p = Jac<F>::inf();
    return;
  }
  F H = U2 - p.X;
  F HH = sqr(H);
  F I = dbl(dbl(HH));
  F J = H * I;
  F r = dbl(S2 - p.Y);
  F V = p.X * I;
  F X3 = sqr(r) - J - dbl(V);
  F Y3 = r * (V - X3) - dbl(p.Y * J);
  F Z3 = sqr(p.Z + H) - Z1Z1 - HH;
  p.X = X3; p.Y = Y3; p.Z = Z3;
}
template <class F>
static inline void jac_add(Jac<F>& p, const Jac<F>& q) {  // add-2007-bl
  if (q.is_inf()) return;
  if (p.is_inf()) { p = q; return; }
  F Z1Z1 = sqr(p.Z), Z2Z2 = sqr(q.Z);
  F U1 = p.X * Z2Z2, U2 = q.X * Z1Z1;
  F S1 = p.Y * q.Z * Z2Z2, S2 = q.Y * p.Z * Z1Z1;
  if (U1 == U2) {
    if (S1 == S2) jac_double(p);
    else p = Jac<F>::inf();
    return;
  }
  F H = U2 - U1;
  F I = sqr(dbl(H));
  F J = H * I;
  F r = dbl(S2 - S1);
  F V = U1 * I;
  F X3 = sqr(r) - J - dbl(V);
  F Y3 = r * (V - X3) - dbl(S1 * J);
  F Z3 = (sqr(p.Z + q.Z) - Z1Z1 - Z2Z2) * H;
  p.X = X3; p.Y = Y3; p.Z = Z3;
}
template <class F>
static Aff<F> jac_to_affine(const Jac<F>& p) {
  Aff<F> r;
  if (p.is_inf()) { r.x = F::zero(); r.y = F::zero(); return r; }
  F zi = inv(p.Z), zi2 = sqr(zi);
  r.x = p.X * zi2;
  r.y = p.Y * zi2 * zi;
  return r;
}
template <class F>
static Jac<F> jac_mul(const Aff<F>& p, const u64* k, int nlimbs) {
  Jac<F> acc = Jac<F>::inf();
  for (int i = nlimbs - 1; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      jac_double(acc);
      if ((k[i] >> b) & 1) jac_add_affine(acc, p, false);
    }
  return acc;
}

// wire <-> internal
template <class C> static inline void load_f(Fp<C>& f, const u64*& p) { memcpy(f.v, p, sizeof(f.v)); p += C::N; f = to_mont(f); }
template <class C> static inline void load_f(Fp2<C>& f, const u64*& p) { load_f(f.c0, p); load_f(f.c1, p); }
template <class C> static inline void store_f(const Fp<C>& f, u64*& p) { Fp<C> t = from_mont(f); memcpy(p, t.v, sizeof(t.v)); p += C::N; }
template <class C> static inline void store_f(const Fp2<C>& f, u64*& p) { store_f(f.c0, p); store_f(f.c1, p); }
template <class F> static inline Aff<F> load_aff(const u64*& p) { Aff<F> a; load_f(a.x, p); load_f(a.y, p); return a; }
template <class F> static inline void store_aff(const Aff<F>& a, u64*& p) { store_f(a.x, p); store_f(a.y, p); }

// ------------------------------------------------------------------------------------------------------------------
// MSM: ark-ec 0.4.2 VariableBaseMSM::msm -> msm_bigint_wnaf
// ------------------------------------------------------------------------------------------------------------------
static inline int ceil_log2(size_t n) { int l = 0; while (((size_t)1 << l) < n) l++; return l; }
static inline int ark_window(size_t n) { return n < 32 ? 3 : ceil_log2(n) * 69 / 100 + 2; }

// signed digits of one canonical scalar (ark make_digits)
static void make_digits(const u64* scalar, int w, int num_bits, int64_t* out) {
  const u64 radix = (u64)1 << w, mask = radix - 1;
  u64 carry = 0;
  int count = (num_bits + w - 1) / w;
  for (int i = 0; i < count; i++) {
    int off = i * w, idx = off / 64, bit = off % 64;
    u64 buf;
    if (bit < 64 - w || idx == 3) buf = scalar[idx] >> bit;
    else buf = (scalar[idx] >> bit) | (scalar[idx + 1] << (64 - bit));
    u64 coef = carry + (buf & mask);
    carry = (coef + radix / 2) >> w;
    int64_t digit = (int64_t)coef - (int64_t)(carry << w);
    if (i == count - 1) digit += (int64_t)(carry << w);
    out[i] = digit;
  }
}

template <class F, class CR>
static void msm_t(const u64* pts, const u64* scalars, size_t n, int num_bits, u64* out_xy, int* out_inf, int threads) {
  ensure_init();
  if (n == 0) {
    memset(out_xy, 0, sizeof(Aff<F>));
    *out_inf = 1;
    return;
  }
  std::vector<Aff<F>> bases(n);
  const int c = ark_window(n);
  const int nd = (num_bits + c - 1) / c;
  std::vector<int64_t> digits(n * (size_t)nd);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (size_t i = 0; i < n; i++) {
    const u64* pp = pts + i * (sizeof(Aff<F>) / 8);
    bases[i] = load_aff<F>(pp);
    // Fr::from(BigUint) reduces mod r (curve.rs:358-361); scalars on this ABI are already < 2^256
    Fp<CR> s;
    memcpy(s.v, scalars + i * 4, 32);
    while (geq<4>(s.v, CR::ctx.p)) sub_n<4>(s.v, s.v, CR::ctx.p);
    make_digits(s.v, c, num_bits, &digits[i * nd]);
  }
  std::vector<Jac<F>> wsum(nd);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int w = 0; w < nd; w++) {
    std::vector<Jac<F>> buckets((size_t)1 << c, Jac<F>::inf());
    for (size_t i = 0; i < n; i++) {
      int64_t d = digits[i * nd + w];
      if (d > 0) jac_add_affine(buckets[d - 1], bases[i], false);
      else if (d < 0) jac_add_affine(buckets[-d - 1], bases[i], true);
    }
    Jac<F> run = Jac<F>::inf(), res = Jac<F>::inf();
    for (size_t b = buckets.size(); b-- > 0;) {
      jac_add(run, buckets[b]);
      jac_add(res, run);
    }
    wsum[w] = res;
  }
  Jac<F> total = Jac<F>::inf();
  for (int w = nd - 1; w >= 1; w--) {
    jac_add(total, wsum[w]);
    for (int k = 0; k < c; k++) jac_double(total);
  }
  jac_add(total, wsum[0]);
  Aff<F> a = jac_to_affine(total);
  *out_inf = a.is_inf() ? 1 : 0;
  u64* o = out_xy;
  store_aff(a, o);
}

typedef Fp<FqBN> fq_bn;
typedef Fp2<FqBN> fq2_bn;
typedef Fp<FqBLS> fq_bls;
typedef Fp2<FqBLS> fq2_bls;

static int default_threads(int t) {
#ifdef _OPENMP
  return t > 0 ? t : omp_get_max_threads();
#else
  (void)t;
  return 1;
#endif
}

static void msm_any(int curve, int group, const u64* pts, const u64* sc, size_t n, u64* out, int* inf, int threads) {
  threads = default_threads(threads);
  if (curve == 0 && group == 1) msm_t<fq_bn, FrBN>(pts, sc, n, 254, out, inf, threads);
  else if (curve == 0) msm_t<fq2_bn, FrBN>(pts, sc, n, 254, out, inf, threads);
  else if (group == 1) msm_t<fq_bls, FrBLS>(pts, sc, n, 255, out, inf, threads);
  else msm_t<fq2_bls, FrBLS>(pts, sc, n, 255, out, inf, threads);
}

// generators (affine, canonical)
static const u64 G1_BN[8] = {1, 0, 0, 0, 2, 0, 0, 0};
static const u64 G2_BN[16] = {
    0x46debd5cd992f6edull, 0x674322d4f75edaddull, 0x426a00665e5c4479ull, 0x1800deef121f1e76ull,
    0x97e485b7aef312c2ull, 0xf1aa493335a9e712ull, 0x7260bfb731fb5d25ull, 0x198e9393920d483aull,
    0x4ce6cc0166fa7daaull, 0xe3d1e7690c43d37bull, 0x4aab71808dcb408full, 0x12c85ea5db8c6debull,
    0x55acdadcd122975bull, 0xbc4b313370b38ef3ull, 0xec9e99ad690c3395ull, 0x090689d0585ff075ull};
static const u64 G1_BLS[12] = {0xfb3af00adb22c6bbull, 0x6c55e83ff97a1aefull, 0xa14e3a3f171bac58ull, 0xc3688c4f9774b905ull,
                               0x2695638c4fa9ac0full, 0x17f1d3a73197d794ull, 0x0caa232946c5e7e1ull, 0xd03cc744a2888ae4ull,
                               0x00db18cb2c04b3edull, 0xfcf5e095d5d00af6ull, 0xa09e30ed741d8ae4ull, 0x08b3f481e3aaa0f1ull};
static const u64 G2_BLS[24] = {
    0xd48056c8c121bdb8ull, 0x0bac0326a805bbefull, 0xb4510b647ae3d177ull, 0xc6e47ad4fa403b02ull, 0x260805272dc51051ull, 0x024aa2b2f08f0a91ull,
    0xe5ac7d055d042b7eull, 0x334cf11213945d57ull, 0xb5da61bbdc7f5049ull, 0x596bd0d09920b61aull, 0x7dacd3a088274f65ull, 0x13e02b6052719f60ull,
    0xe193548608b82801ull, 0x923ac9cc3baca289ull, 0x6d429a695160d12cull, 0xadfd9baa8cbdd3a7ull, 0x8cc9cdc6da2e351aull, 0x0ce5d527727d6e11ull,
    0xaaa9075ff05f79beull, 0x3f370d275cec1da1ull, 0x267492ab572e99abull, 0xcb3e287e85a763afull, 0x32acd2b02bc28b99ull, 0x0606c4a02ea734ccull};

static const u64* generator(int curve, int group) {
  if (curve == 0) return group == 1 ? G1_BN : G2_BN;
  return group == 1 ? G1_BLS : G2_BLS;
}

// out[i] = (k0 + i) * G for i < n, affine canonical (cheap synthetic key vectors for the CPU arm: one mixed addition per point
// and a blocked batch inversion)
template <class F>
static void chain_points_t(const u64* gen, u64 k0, size_t n, u64* out, int threads) {
  ensure_init();
  const u64* gp = gen;
  Aff<F> g = load_aff<F>(gp);
  const size_t BLK = 4096;
  size_t nblk = (n + BLK - 1) / BLK;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (size_t b = 0; b < nblk; b++) {
    size_t lo = b * BLK, hi = lo + BLK < n ? lo + BLK : n;
    u64 k[1] = {k0 + lo};
    Jac<F> cur = jac_mul(g, k, 1);
    std::vector<Jac<F>> js(hi - lo);
    for (size_t i = lo; i < hi; i++) {
      js[i - lo] = cur;
      jac_add_affine(cur, g, false);
    }
    // batch inversion of the Z coordinates (identity entries, Z = 0, are skipped)
    std::vector<F> pref(hi - lo);
    F acc = F::one();
    for (size_t i = 0; i < js.size(); i++) {
      pref[i] = acc;
      if (!js[i].is_inf()) acc = acc * js[i].Z;
    }
    F ia = inv(acc);
    for (size_t i = js.size(); i-- > 0;) {
      Aff<F> a;
      if (js[i].is_inf()) { a.x = F::zero(); a.y = F::zero(); }
      else {
        F zi = ia * pref[i];
        ia = ia * js[i].Z;
        F zi2 = sqr(zi);
        a.x = js[i].X * zi2;
        a.y = js[i].Y * zi2 * zi;
      }
      u64* o = out + (lo + i) * (sizeof(Aff<F>) / 8);
      store_aff(a, o);
    }
  }
}

// sum_i k_i * P_i over a few canonical affine points (proof assembly, scalar multiplication)
template <class F>
static void lincomb_t(int n_terms, const u64* points, const int* infs, const u64* scalars, const int* has_scalar, u64* out_xy,
                      int* out_inf) {
  ensure_init();
  Jac<F> acc = Jac<F>::inf();
  const size_t stride = sizeof(Aff<F>) / 8;
  for (int i = 0; i < n_terms; i++) {
    if (infs && infs[i]) continue;
    const u64* pp = points + i * stride;
    Aff<F> a = load_aff<F>(pp);
    if (has_scalar[i]) {
      Jac<F> t = jac_mul(a, scalars + i * 4, 4);
      jac_add(acc, t);
    } else {
      jac_add_affine(acc, a, false);
    }
  }
  Aff<F> a = jac_to_affine(acc);
  *out_inf = a.is_inf() ? 1 : 0;
  u64* o = out_xy;
  store_aff(a, o);
}
static void lincomb_any(int curve, int group, int n_terms, const u64* points, const int* infs, const u64* scalars,
                        const int* has_scalar, u64* out_xy, int* out_inf) {
  if (curve == 0 && group == 1) lincomb_t<fq_bn>(n_terms, points, infs, scalars, has_scalar, out_xy, out_inf);
  else if (curve == 0) lincomb_t<fq2_bn>(n_terms, points, infs, scalars, has_scalar, out_xy, out_inf);
  else if (group == 1) lincomb_t<fq_bls>(n_terms, points, infs, scalars, has_scalar, out_xy, out_inf);
  else lincomb_t<fq2_bls>(n_terms, points, infs, scalars, has_scalar, out_xy, out_inf);
}

// ------------------------------------------------------------------------------------------------------------------
// Fr NTT: ark-poly 0.4.2 Radix2EvaluationDomain
// ------------------------------------------------------------------------------------------------------------------
template <class CR>
struct FrInfo;
template <> struct FrInfo<FrBN> { static constexpr int two_adicity = 28; static constexpr u64 gen = 5; };
template <> struct FrInfo<FrBLS> { static constexpr int two_adicity = 32; static constexpr u64 gen = 7; };

template <class CR>
static Fp<CR> fr_small(u64 k) {
  Fp<CR> x = Fp<CR>::zero();
  x.v[0] = k;
  return to_mont(x);
}
// group_gen of the size-2^log_n domain: (g^((r-1)/2^s))^(2^(s - log_n))
template <class CR>
static Fp<CR> fr_omega(int log_n) {
  typedef Fp<CR> F;
  const int s = FrInfo<CR>::two_adicity;
  u64 e[4], one[4] = {1, 0, 0, 0};
  sub_n<4>(e, CR::ctx.p, one);
  // e = (r - 1) >> s
  for (int k = 0; k < s; k++) {
    for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (i + 1 < 4 ? e[i + 1] << 63 : 0);
  }
  F w = pow_limbs(fr_small<CR>(FrInfo<CR>::gen), e, 4);
  for (int k = log_n; k < s; k++) w = sqr(w);
  return w;
}

static inline size_t bitrev(size_t x, int bits) {
  size_t r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}
template <class F>
static void derange(F* a, int log_n) {
  size_t n = (size_t)1 << log_n;
  for (size_t i = 1; i < n; i++) {
    size_t j = bitrev(i, log_n);
    if (i < j) { F t = a[i]; a[i] = a[j]; a[j] = t; }
  }
}

// in place on Montgomery-form values.  forward: DIF (in-order in, bit-reversed out) + derange; inverse: derange + DIT.
template <class CR>
static void ntt_core(Fp<CR>* a, int log_n, const Fp<CR>& root, bool dit, int threads) {
  typedef Fp<CR> F;
  size_t n = (size_t)1 << log_n;
  if (n == 1) return;
  std::vector<F> roots(n / 2);  // root^j
  roots[0] = F::one();
  for (size_t j = 1; j < n / 2; j++) roots[j] = roots[j - 1] * root;
  if (!dit) {
    int lg = log_n - 1;
    for (size_t gap = n / 2; gap >= 1; gap >>= 1, lg--) {
      size_t step = (n / 2) >> lg;  // twiddle stride
#pragma omp parallel for num_threads(threads) schedule(static) if (n >= 4096)
      for (size_t q = 0; q < n / 2; q++) {
        size_t blk = q >> lg, j = q & (gap - 1);
        size_t lo = blk * 2 * gap + j, hi = lo + gap;
        F x = a[lo], y = a[hi];
        a[lo] = x + y;
        F d = x - y;
        a[hi] = j ? d * roots[j * step] : d;
      }
      if (gap == 1) break;
    }
    derange(a, log_n);
  } else {
    derange(a, log_n);
    int lg = 0;
    for (size_t gap = 1; gap < n; gap <<= 1, lg++) {
      size_t step = (n / 2) >> lg;
#pragma omp parallel for num_threads(threads) schedule(static) if (n >= 4096)
      for (size_t q = 0; q < n / 2; q++) {
        size_t blk = q >> lg, j = q & (gap - 1);
        size_t lo = blk * 2 * gap + j, hi = lo + gap;
        F x = a[lo], y = j ? a[hi] * roots[j * step] : a[hi];
        a[lo] = x + y;
        a[hi] = x - y;
      }
    }
  }
}

// canonical in -> Montgomery vector of length N (zero-pad / truncate; values >= r reduced)
template <class CR>
static void load_fr_vec(std::vector<Fp<CR>>& v, const u64* in, size_t in_len, size_t n, int threads) {
  if (in_len > n) in_len = n;
  v.assign(n, Fp<CR>::zero());
#pragma omp parallel for num_threads(threads) schedule(static) if (in_len >= 4096)
  for (size_t i = 0; i < in_len; i++) {
    Fp<CR> x;
    memcpy(x.v, in + 4 * i, 32);
    while (geq<4>(x.v, CR::ctx.p)) sub_n<4>(x.v, x.v, CR::ctx.p);
    v[i] = to_mont(x);
  }
}
template <class CR>
static void store_fr_vec(const Fp<CR>* v, size_t n, u64* out, int threads) {
#pragma omp parallel for num_threads(threads) schedule(static) if (n >= 4096)
  for (size_t i = 0; i < n; i++) {
    Fp<CR> x = from_mont(v[i]);
    memcpy(out + 4 * i, x.v, 32);
  }
}

template <class CR>
static void fft_mont(std::vector<Fp<CR>>& v, int log_n, bool inverse, bool coset, int threads) {
  typedef Fp<CR> F;
  size_t n = (size_t)1 << log_n;
  F w = fr_omega<CR>(log_n);
  if (!inverse) {
    if (coset) {  // distribute_powers(offset = group_gen), polynomial.rs:553-556
      F t = F::one();
      for (size_t j = 0; j < n; j++) { v[j] = v[j] * t; t = t * w; }
    }
    ntt_core<CR>(v.data(), log_n, w, false, threads);
  } else {
    F wi = inv(w);
    ntt_core<CR>(v.data(), log_n, wi, true, threads);
    F ninv = inv(fr_small<CR>((u64)n));
    if (coset) {
      F t = ninv;
      for (size_t j = 0; j < n; j++) { v[j] = v[j] * t; t = t * wi; }
    } else {
#pragma omp parallel for num_threads(threads) schedule(static) if (n >= 4096)
      for (size_t j = 0; j < n; j++) v[j] = v[j] * ninv;
    }
  }
}

template <class CR>
static int fft_t(int inverse, int coset, int log_n, const u64* in, size_t in_len, u64* out, int threads) {
  ensure_init();
  if (log_n > FrInfo<CR>::two_adicity) return -4;
  std::vector<Fp<CR>> v;
  load_fr_vec<CR>(v, in, in_len, (size_t)1 << log_n, threads);
  fft_mont<CR>(v, log_n, inverse != 0, coset != 0, threads);
  store_fr_vec<CR>(v.data(), v.size(), out, threads);
  return 0;
}

// QAP.evaluate_witness from the evaluation vectors a, b, c (n = 2^log_n each): returns U, V, W (n coefficients) and H
// (n coefficients, zero padded).  rc -5: non-zero remainder (qap.py:68-69).
template <class CR>
static int groth16_h_t(int log_n, const u64* a, const u64* b, const u64* c, u64* u, u64* v, u64* w, u64* h, int threads) {
  typedef Fp<CR> F;
  ensure_init();
  if (log_n + 1 > FrInfo<CR>::two_adicity) return -4;
  size_t n = (size_t)1 << log_n;
  std::vector<F> U, V, W;
  load_fr_vec<CR>(U, a, n, n, threads);
  load_fr_vec<CR>(V, b, n, n, threads);
  load_fr_vec<CR>(W, c, n, n, threads);
  fft_mont<CR>(U, log_n, true, false, threads);   // qap.py:57-59
  fft_mont<CR>(V, log_n, true, false, threads);
  fft_mont<CR>(W, log_n, true, false, threads);
  if (u) store_fr_vec<CR>(U.data(), n, u, threads);
  if (v) store_fr_vec<CR>(V.data(), n, v, threads);
  if (w) store_fr_vec<CR>(W.data(), n, w, threads);
  // mul_over_fft on the doubled domain (polynomial.py:151-165; _pad_coeffs of two degree-(n-1) operands -> length 2n)
  std::vector<F> eu(U), ev(V);
  eu.resize(2 * n, F::zero());
  ev.resize(2 * n, F::zero());
  int l2 = log_n + 1;
  fft_mont<CR>(eu, l2, false, false, threads);
  fft_mont<CR>(ev, l2, false, false, threads);
#pragma omp parallel for num_threads(threads) schedule(static) if (n >= 2048)
  for (size_t i = 0; i < 2 * n; i++) eu[i] = eu[i] * ev[i];
  fft_mont<CR>(eu, l2, true, false, threads);
  for (size_t i = 0; i < n; i++) eu[i] = eu[i] - W[i];             // U*V - W
  // divide by X^n - 1 (polynomial.rs:466-489): degree < 2n  =>  q = top half, remainder = low half + q
  bool bad = false;
  for (size_t i = 0; i < n; i++)
    if (!(eu[i] + eu[n + i]).is_zero()) bad = true;
  if (h) store_fr_vec<CR>(eu.data() + n, n, h, threads);
  return bad ? -5 : 0;
}

// SparseArray.dot (array.py:36-43) over CSR
template <class CR>
static void spmv_t(size_t n_out, size_t n_rows, const u64* row_ptr, const uint32_t* col, const u64* val, const u64* wit, size_t m,
                   u64* out, int threads) {
  typedef Fp<CR> F;
  ensure_init();
  std::vector<F> w;
  load_fr_vec<CR>(w, wit, m, m, threads);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (size_t row = 0; row < n_out; row++) {
    F acc = F::zero();
    if (row < n_rows)
      for (u64 k = row_ptr[row]; k < row_ptr[row + 1]; k++) {
        F x;
        memcpy(x.v, val + 4 * k, 32);
        while (geq<4>(x.v, CR::ctx.p)) sub_n<4>(x.v, x.v, CR::ctx.p);
        acc = acc + to_mont(x) * w[col[k]];
      }
    F r = from_mont(acc);
    memcpy(out + 4 * row, r.v, 32);
  }
}

template <class CR>
static void fr_mul_canon(const u64* a, const u64* b, u64* out) {
  Fp<CR> x, y;
  memcpy(x.v, a, 32);
  memcpy(y.v, b, 32);
  Fp<CR> r = from_mont(to_mont(x) * to_mont(y));
  memcpy(out, r.v, 32);
}

extern "C" {

int zkcpu_threads(void) { return default_threads(0); }

int zkcpu_fft(int curve, int inverse, int coset, int log_n, const u64* in, size_t in_len, u64* out, int threads) {
  threads = default_threads(threads);
  if (curve == 0) return fft_t<FrBN>(inverse, coset, log_n, in, in_len, out, threads);
  return fft_t<FrBLS>(inverse, coset, log_n, in, in_len, out, threads);
}

int zkcpu_msm(int curve, int group, const u64* pts, size_t n_points, const u64* scalars, size_t n_scalars, u64* out_xy,
              int* out_inf, int threads) {
  if (n_points != n_scalars) return -3;  // "Number of points and scalars mismatch" (curve.rs:369-371)
  msm_any(curve, group, pts, scalars, n_points, out_xy, out_inf, threads);
  return 0;
}

int zkcpu_msm_window(size_t n) { return ark_window(n); }

int zkcpu_chain_points(int curve, int group, u64 k0, size_t n, u64* out, int threads) {
  threads = default_threads(threads);
  const u64* g = generator(curve, group);
  if (curve == 0 && group == 1) chain_points_t<fq_bn>(g, k0, n, out, threads);
  else if (curve == 0) chain_points_t<fq2_bn>(g, k0, n, out, threads);
  else if (group == 1) chain_points_t<fq_bls>(g, k0, n, out, threads);
  else chain_points_t<fq2_bls>(g, k0, n, out, threads);
  return 0;
}

int zkcpu_point_lincomb(int curve, int group, int n_terms, const u64* points, const int* infs, const u64* scalars,
                        const int* has_scalar, u64* out_xy, int* out_inf) {
  lincomb_any(curve, group, n_terms, points, infs, scalars, has_scalar, out_xy, out_inf);
  return 0;
}

int zkcpu_groth16_h(int curve, int log_n, const u64* a, const u64* b, const u64* c, u64* u, u64* v, u64* w, u64* h, int threads) {
  threads = default_threads(threads);
  if (curve == 0) return groth16_h_t<FrBN>(log_n, a, b, c, u, v, w, h, threads);
  return groth16_h_t<FrBLS>(log_n, a, b, c, u, v, w, h, threads);
}

int zkcpu_spmv(int curve, size_t n_out, size_t n_rows, const u64* row_ptr, const uint32_t* col, const u64* val, const u64* wit,
               size_t m, u64* out, int threads) {
  threads = default_threads(threads);
  if (curve == 0) spmv_t<FrBN>(n_out, n_rows, row_ptr, col, val, wit, m, out, threads);
  else spmv_t<FrBLS>(n_out, n_rows, row_ptr, col, val, wit, m, out, threads);
  return 0;
}

// Groth16.prove (protocol.py:115-165) on host buffers.  Key vectors are canonical affine arrays (tau1, target1: n G1 points;
// tau2: n G2 points; kdelta1: n_priv G1 points); alpha1, beta1, delta1 are G1 points, beta2, delta2 G2 points.  witness: m
// canonical scalars (public part first).  phase_ms (optional, 3 floats): SpMV, quotient, MSMs+assembly wall times.
int zkcpu_groth16_prove(int curve, int log_n, size_t n_rows, size_t m, size_t n_public, const u64* const row_ptr[3],
                        const uint32_t* const col[3], const u64* const val[3], const u64* witness, const u64* tau1,
                        const u64* tau2, const u64* target1, const u64* kdelta1, const u64* alpha1, const u64* beta1,
                        const u64* beta2, const u64* delta1, const u64* delta2, const u64* r, const u64* s, u64* out_a,
                        u64* out_b, u64* out_c, int out_inf[3], u64* h_out, int threads) {
  threads = default_threads(threads);
  size_t n = (size_t)1 << log_n;
  size_t g1 = (curve == 0 ? 4 : 6) * 2, g2 = g1 * 2;
  std::vector<u64> ev(3 * n * 4), uvwh(4 * n * 4);
  for (int i = 0; i < 3; i++) zkcpu_spmv(curve, n, n_rows, row_ptr[i], col[i], val[i], witness, m, &ev[i * n * 4], threads);
  u64 *U = &uvwh[0], *V = &uvwh[n * 4], *W = &uvwh[2 * n * 4], *H = &uvwh[3 * n * 4];
  int rc = zkcpu_groth16_h(curve, log_n, &ev[0], &ev[n * 4], &ev[2 * n * 4], U, V, W, H, threads);
  if (rc) return rc;
  if (h_out) memcpy(h_out, H, n * 32);
  std::vector<u64> m_a(g1), m_b1(g1), m_b2(g2), m_hz(g1), m_kw(g1);
  int i_a, i_b1, i_b2, i_hz, i_kw;
  // ecc.py:107-126 multiexp trims the point list to the (stripped) coefficient list; the stripped tail is zero here
  msm_any(curve, 1, tau1, U, n, m_a.data(), &i_a, threads);
  msm_any(curve, 1, tau1, V, n, m_b1.data(), &i_b1, threads);
  msm_any(curve, 2, tau2, V, n, m_b2.data(), &i_b2, threads);
  msm_any(curve, 1, target1, H, n, m_hz.data(), &i_hz, threads);
  size_t n_priv = m - n_public;
  msm_any(curve, 1, kdelta1, witness + 4 * n_public, n_priv, m_kw.data(), &i_kw, threads);
  // A = msm + alpha1 + r*delta1 ; B = msm + beta + s*delta ; C = HZ + KW + s*A + r*B1 - rs*delta1
  std::vector<u64> A(g1), B1(g1);
  int infA, infB1;
  {
    std::vector<u64> pts(3 * g1), sc(12, 0);
    memcpy(&pts[0], m_a.data(), g1 * 8); memcpy(&pts[g1], alpha1, g1 * 8); memcpy(&pts[2 * g1], delta1, g1 * 8);
    memcpy(&sc[8], r, 32);
    int infs[3] = {i_a, 0, 0}, has[3] = {0, 0, 1};
    lincomb_any(curve, 1, 3, pts.data(), infs, sc.data(), has, A.data(), &infA);
  }
  {
    std::vector<u64> pts(3 * g1), sc(12, 0);
    memcpy(&pts[0], m_b1.data(), g1 * 8); memcpy(&pts[g1], beta1, g1 * 8); memcpy(&pts[2 * g1], delta1, g1 * 8);
    memcpy(&sc[8], s, 32);
    int infs[3] = {i_b1, 0, 0}, has[3] = {0, 0, 1};
    lincomb_any(curve, 1, 3, pts.data(), infs, sc.data(), has, B1.data(), &infB1);
  }
  {
    std::vector<u64> pts(3 * g2), sc(12, 0);
    memcpy(&pts[0], m_b2.data(), g2 * 8); memcpy(&pts[g2], beta2, g2 * 8); memcpy(&pts[2 * g2], delta2, g2 * 8);
    memcpy(&sc[8], s, 32);
    int infs[3] = {i_b2, 0, 0}, has[3] = {0, 0, 1};
    lincomb_any(curve, 2, 3, pts.data(), infs, sc.data(), has, out_b, &out_inf[1]);
  }
  {
    u64 rs[4], nrs[4];
    if (curve == 0) fr_mul_canon<FrBN>(r, s, rs); else fr_mul_canon<FrBLS>(r, s, rs);
    const u64* ord = curve == 0 ? FR_BN : FR_BLS;
    if (rs[0] | rs[1] | rs[2] | rs[3]) sub_n<4>(nrs, ord, rs); else memset(nrs, 0, 32);
    std::vector<u64> pts(5 * g1), sc(20, 0);
    memcpy(&pts[0], m_hz.data(), g1 * 8); memcpy(&pts[g1], m_kw.data(), g1 * 8); memcpy(&pts[2 * g1], A.data(), g1 * 8);
    memcpy(&pts[3 * g1], B1.data(), g1 * 8); memcpy(&pts[4 * g1], delta1, g1 * 8);
    memcpy(&sc[8], s, 32); memcpy(&sc[12], r, 32); memcpy(&sc[16], nrs, 32);
    int infs[5] = {i_hz, i_kw, infA, infB1, 0}, has[5] = {0, 0, 1, 1, 1};
    lincomb_any(curve, 1, 5, pts.data(), infs, sc.data(), has, out_c, &out_inf[2]);
  }
  memcpy(out_a, A.data(), g1 * 8);
  out_inf[0] = infA;
  return 0;
}

}  // extern "C"
