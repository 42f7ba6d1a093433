// codec.cu -- bulk (de)serialisation of curve points in the reference's wire format.
//
// The reference stores keys and proofs as concatenated ark-serialize COMPRESSED points (PointG1/PointG2.to_bytes,
// /root/reference/src/bn254/curve.rs:127-141 and :300-314; bls12_381/curve.rs likewise) and reads a key back by calling
// from_hex once per point from a Python loop (/root/reference/python/zksnake/groth16/serialization.py:70-130, 181-206;
// plonk/serialization.py:157-173; ecc.py:128-142) -- one square root (and, for every group but BN254 G1, one subgroup check)
// per point on one core.  Here a whole vector is encoded / decoded by one kernel, one thread per point:
//
//   BN254      x little-endian (G2: c0 || c1); top two bits of the LAST byte: 0x80 = y is the larger of (y, -y), 0x40 = infinity
//   BLS12-381  x big-endian (G2: c1 || c0);    top three bits of the FIRST byte: 0x80 = compressed, 0x40 = infinity, 0x20 = larger y
//   "larger" compares y with -y as integers; in Fq2 lexicographically with c1 the more significant component.
//
// Decoding validates exactly what Validate::Yes does, in this order: flags, x < q, infinity has x = 0, x^3 + b is a square,
// r * P = infinity (skipped for BN254 G1, whose cofactor is 1).  The first offending index and its reason are reported.
// Device points are affine in Montgomery form with (0,0) as the infinity marker, as everywhere in this library.
#include <cuda_runtime.h>
#include <string.h>
#include "../../include/zkb200.h"
#include "ec.cuh"
#include "zkb_internal.h"

namespace zkb {

static inline cudaStream_t S() { return (cudaStream_t)ctx_stream(); }

enum { CODEC_OK = 0, CODEC_FLAGS = 1, CODEC_FIELD = 2, CODEC_INF = 3, CODEC_CURVE = 4, CODEC_SUBGROUP = 5 };

// ------------------------------------------------------------------------------------------------ field helpers
// a > q - a for a canonical (non-Montgomery) value: the sign bit of the encodings
template <class P>
__device__ bool gt_neg(const Fp<P>& a) {
  Fp<P> n = neg(a);   // limb arithmetic only: valid on canonical values too; neg(0) = 0
  for (int i = P::N - 1; i >= 0; i--) {
    if (a.v[i] != n.v[i]) return a.v[i] > n.v[i];
  }
  return false;
}
template <class P>
__device__ bool gt_neg(const Fp2<P>& a) {
  return a.c1.is_zero() ? gt_neg(a.c0) : gt_neg(a.c1);   // (c1, c0) against (-c1, -c0); c1 = -c1 only for c1 = 0
}

template <class P>
__device__ bool below_modulus(const Fp<P>& a) {
  uint32_t t = sub_cc(a.v[0], P::MOD(0));
#pragma unroll
  for (int i = 1; i < P::N; i++) t = subc_cc(a.v[i], P::MOD(i));
  (void)t;
  return subc(0u, 0u) != 0u;   // borrow: a < q
}
template <class P>
__device__ bool below_modulus(const Fp2<P>& a) { return below_modulus(a.c0) && below_modulus(a.c1); }

// a / 2 (works on the Montgomery representative as well: halving is linear)
template <class P>
__device__ Fp<P> half(const Fp<P>& a) {
  constexpr int N = P::N;
  uint32_t odd = (a.v[0] & 1u) ? 0xffffffffu : 0u;
  uint32_t t[N];
  t[0] = add_cc(a.v[0], P::MOD(0) & odd);
#pragma unroll
  for (int i = 1; i < N; i++) t[i] = addc_cc(a.v[i], P::MOD(i) & odd);
  uint32_t top = addc(0u, 0u);
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = (t[i] >> 1) | ((i + 1 < N ? t[i + 1] : top) << 31);
  return r;
}

// square root for q = 3 (mod 4), both base fields: a^((q+1)/4), accepted when it squares back (ark's sqrt for this case)
template <class P>
__device__ bool sqrt_field(const Fp<P>& a, Fp<P>& out) {
  constexpr int N = P::N;
  uint32_t e[N];
  uint32_t carry = 1;
#pragma unroll
  for (int i = 0; i < N; i++) {
    unsigned long long t = (unsigned long long)P::MOD(i) + carry;
    e[i] = (uint32_t)t;
    carry = (uint32_t)(t >> 32);
  }
#pragma unroll
  for (int i = 0; i < N; i++) e[i] = (e[i] >> 2) | ((i + 1 < N ? e[i + 1] : carry) << 30);
  out = pow_limbs(a, e, N);
  return sqr(out) == a;
}

// square root in Fq2 = Fq[u]/(u^2+1) by the norm ("complex") method: for a = a0 + a1 u with a1 != 0,
// n = sqrt(a0^2 + a1^2), x0 = sqrt((a0 +- n)/2), x1 = a1 / (2 x0).  Either root will do: the sign flag picks between +-y.
template <class P>
__device__ bool sqrt_field(const Fp2<P>& a, Fp2<P>& out) {
  typedef Fp<P> B;
  if (a.is_zero()) {
    out = Fp2<P>::zero();
    return true;
  }
  if (a.c1.is_zero()) {
    B s;
    if (sqrt_field(a.c0, s)) {
      out.c0 = s;
      out.c1 = B::zero();
      return true;
    }
    if (sqrt_field(neg(a.c0), s)) {   // (s u)^2 = -s^2
      out.c0 = B::zero();
      out.c1 = s;
      return true;
    }
    return false;
  }
  B n;
  if (!sqrt_field(sqr(a.c0) + sqr(a.c1), n)) return false;
  for (int k = 0; k < 2; k++) {
    B cand = half(k == 0 ? a.c0 + n : a.c0 - n);
    B x0;
    if (!sqrt_field(cand, x0) || x0.is_zero()) continue;
    B x1 = a.c1 * inv(x0 + x0);
    Fp2<P> r;
    r.c0 = x0;
    r.c1 = x1;
    if (sqr(r) == a) {
      out = r;
      return true;
    }
  }
  return false;
}

// ------------------------------------------------------------------------------------------------ byte <-> limb
template <class P, bool BE>
__device__ Fp<P> load_fq(const uint32_t* w) {
  Fp<P> r;
#pragma unroll
  for (int j = 0; j < P::N; j++) r.v[j] = BE ? __byte_perm(w[P::N - 1 - j], 0, 0x0123) : w[j];
  return r;
}
template <class P, bool BE>
__device__ void store_fq(uint32_t* w, const Fp<P>& a) {
#pragma unroll
  for (int j = 0; j < P::N; j++) {
    if (BE) w[P::N - 1 - j] = __byte_perm(a.v[j], 0, 0x0123);
    else w[j] = a.v[j];
  }
}

// One coordinate field's wire layout.  `flagged(x)` is the base-field element whose most significant limb carries the flags
// (c1 in Fq2 on both curves: BN254 writes c0 || c1 little-endian, flags in the last byte; BLS12-381 writes c1 || c0 big-endian,
// flags in the first byte).
template <class P, bool BE>
struct Wire1 {
  typedef Fp<P> F;
  static constexpr int WORDS = P::N;
  __device__ static F load(const uint32_t* w) { return load_fq<P, BE>(w); }
  __device__ static void store(uint32_t* w, const F& x) { store_fq<P, BE>(w, x); }
  __device__ static Fp<P>& flagged(F& x) { return x; }
};
template <class P, bool BE>
struct Wire2 {
  typedef Fp2<P> F;
  static constexpr int WORDS = 2 * P::N;
  __device__ static F load(const uint32_t* w) {
    F x;
    x.c0 = load_fq<P, BE>(w + (BE ? P::N : 0));
    x.c1 = load_fq<P, BE>(w + (BE ? 0 : P::N));
    return x;
  }
  __device__ static void store(uint32_t* w, const F& x) {
    store_fq<P, BE>(w + (BE ? P::N : 0), x.c0);
    store_fq<P, BE>(w + (BE ? 0 : P::N), x.c1);
  }
  __device__ static Fp<P>& flagged(F& x) { return x.c1; }
};

template <class F>
struct CurveB {
  F b;             // Montgomery form
  F psi_x, psi_y;  // G2 only: psi(x, y) = (conj(x) psi_x, conj(y) psi_y), the untwist-Frobenius-twist endomorphism
};

// ------------------------------------------------------------------------------------------------ subgroup membership
// Definition: r * P = infinity.  mode 2 computes exactly that (255 doublings).  mode 1 uses, in G2, the endomorphism criteria
// that replace the 254-bit scalar by the 63/64-bit curve parameter x (psi acts on G2 as multiplication by p):
//   BN254      [x+1]P + psi([x]P) + psi^2([x]P) = psi^3([2x]P)      (El Housni, Guillevic, Piellard, eprint 2022/352, section 4.3)
//   BLS12-381  psi(P) = [x]P, x < 0                                (Scott, eprint 2021/1130)
// Both are proven equivalent to the definition for points of E'(Fq2); tests/test_psi_constants.py pins the constants and checks
// them against r * P on members, random non-members, cofactor-torsion points and mixtures; tests/test_gpu_codec.py checks
// the kernel against the oracle's r * P on the same kinds of points.
template <class P>
__device__ Fp2<P> conj2(const Fp2<P>& a) {
  Fp2<P> r;
  r.c0 = a.c0;
  r.c1 = neg(a.c1);
  return r;
}
template <class P>
__device__ Affine<Fp2<P>> psi(const Affine<Fp2<P>>& p, const CurveB<Fp2<P>>& cb) {
  if (p.is_inf()) return p;
  Affine<Fp2<P>> r;
  r.x = conj2(p.x) * cb.psi_x;
  r.y = conj2(p.y) * cb.psi_y;
  return r;
}
template <class P>
__device__ XYZZ<Fp2<P>> psi(const XYZZ<Fp2<P>>& p, const CurveB<Fp2<P>>& cb) {
  XYZZ<Fp2<P>> r;   // x = X / ZZ, y = Y / ZZZ and conjugation is a field automorphism
  r.X = conj2(p.X) * cb.psi_x;
  r.Y = conj2(p.Y) * cb.psi_y;
  r.ZZ = conj2(p.ZZ);
  r.ZZZ = conj2(p.ZZZ);
  return r;
}

template <class F>
__device__ bool in_subgroup(const Affine<F>& p, const CurveB<F>&, const uint32_t* order, int) {
  uint32_t k[8];
#pragma unroll
  for (int j = 0; j < 8; j++) k[j] = order[j];
  return scalar_mul(p, k, 8).is_inf();
}
__device__ bool in_subgroup(const Affine<Fp2<FqBN254>>& p, const CurveB<Fp2<FqBN254>>& cb, const uint32_t* order, int mode) {
  typedef Fp2<FqBN254> F;
  if (mode == 2) return in_subgroup<F>(p, cb, order, 0);
  const uint32_t x[2] = {0x4A6909F1u, 0x44E992B4u};   // 4965661367192848881
  XYZZ<F> xp = scalar_mul(p, x, 2);
  XYZZ<F> a = xp;
  madd(a, p);                        // [x+1]P
  XYZZ<F> b = psi(xp, cb);           // psi([x]P)
  XYZZ<F> c = psi(b, cb);            // psi^2([x]P)
  XYZZ<F> lhs = add(add(a, b), c);
  XYZZ<F> rhs = psi(psi(psi(dbl(xp), cb), cb), cb);
  return add(lhs, neg(rhs)).is_inf();
}
__device__ bool in_subgroup(const Affine<Fp2<FqBLS381>>& p, const CurveB<Fp2<FqBLS381>>& cb, const uint32_t* order, int mode) {
  typedef Fp2<FqBLS381> F;
  if (mode == 2) return in_subgroup<F>(p, cb, order, 0);
  const uint32_t x[2] = {0x00010000u, 0xd2010000u};   // |x| = 0xd201000000010000, x = -|x|
  XYZZ<F> xp = scalar_mul(p, x, 2);
  madd(xp, psi(p, cb));              // [|x|]P + psi(P) = 0  <=>  psi(P) = [x]P
  return xp.is_inf();
}

// ------------------------------------------------------------------------------------------------ kernels
template <class W, class P, bool BE>
__global__ void __launch_bounds__(128) compress_kernel(unsigned long long n, const Affine<typename W::F>* __restrict__ pts,
                                                       uint32_t* __restrict__ out) {
  typedef typename W::F F;
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = pts[i];
  F x = F::zero();
  uint32_t flags;
  if (p.is_inf()) {
    flags = BE ? 0xC0000000u : 0x40000000u;
  } else {
    x = from_mont(p.x);
    bool larger = gt_neg(from_mont(p.y));
    flags = BE ? (0x80000000u | (larger ? 0x20000000u : 0u)) : (larger ? 0x80000000u : 0u);
  }
  W::flagged(x).v[P::N - 1] |= flags;
  W::store(out + i * W::WORDS, x);
}

// status: atomicMin of (index << 8 | reason) over the offending points
template <class W, class P, bool BE, int SUBGROUP>   // SUBGROUP: 0 none, 1 fast criterion (G2) / r*P (G1), 2 r*P
__global__ void __launch_bounds__(128) decompress_kernel(unsigned long long n, const uint32_t* __restrict__ in,
                                                         CurveB<typename W::F> cb, const uint32_t* __restrict__ order,
                                                         Affine<typename W::F>* __restrict__ pts,
                                                         unsigned long long* __restrict__ status) {
  typedef typename W::F F;
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = W::load(in + i * W::WORDS);
  uint32_t& top = W::flagged(x).v[P::N - 1];
  bool inf, larger;
  int bad = CODEC_OK;
  if (BE) {
    if (!(top & 0x80000000u)) bad = CODEC_FLAGS;   // uncompressed encoding
    inf = (top & 0x40000000u) != 0;
    larger = (top & 0x20000000u) != 0;
    top &= 0x1fffffffu;
  } else {
    inf = (top & 0x40000000u) != 0;
    larger = (top & 0x80000000u) != 0;
    top &= 0x3fffffffu;
    if (inf && larger) bad = CODEC_FLAGS;
  }
  Affine<F> p = Affine<F>::inf();
  if (!bad && !below_modulus(x)) bad = CODEC_FIELD;
  if (!bad) {
    if (inf) {
      if (!x.is_zero()) bad = CODEC_INF;
    } else {
      F xm = to_mont(x);
      F y;
      if (!sqrt_field(sqr(xm) * xm + cb.b, y)) {
        bad = CODEC_CURVE;
      } else {
        if (gt_neg(from_mont(y)) != larger) y = neg(y);
        p.x = xm;
        p.y = y;
        if (SUBGROUP) {
          if (!in_subgroup(p, cb, order, SUBGROUP)) bad = CODEC_SUBGROUP;
        }
      }
    }
  }
  if (bad) {
    atomicMin(status, (i << 8) | (unsigned long long)bad);
    p = Affine<F>::inf();
  }
  pts[i] = p;
}

template <class W, class P, bool BE>
static int compress_run(size_t n, const void* d_pts, void* d_out) {
  compress_kernel<W, P, BE><<<(unsigned)((n + 127) / 128), 128, 0, S()>>>(n, (const Affine<typename W::F>*)d_pts, (uint32_t*)d_out);
  count_launch();
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}

template <class P>
static Fp<P> host_to_mont(const uint64_t* limbs) {
  Fp<P> a;
  for (int i = 0; i < P::N; i++) a.v[i] = (uint32_t)(limbs[i / 2] >> (32 * (i & 1)));
  return to_mont(a);
}

template <class W, class P, bool BE, int SUBGROUP>
static int decompress_run(size_t n, const void* d_in, const CurveB<typename W::F>& cb, const uint32_t* d_order, void* d_pts,
                          unsigned long long* d_status) {
  decompress_kernel<W, P, BE, SUBGROUP><<<(unsigned)((n + 127) / 128), 128, 0, S()>>>(
      n, (const uint32_t*)d_in, cb, d_order, (Affine<typename W::F>*)d_pts, d_status);
  count_launch();
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}

// curve coefficient b (y^2 = x^3 + b), canonical limbs; the G2 twists: BN254 3/(9+u), BLS12-381 4(1+u)
static const uint64_t B_BN_G1[4] = {3, 0, 0, 0};
static const uint64_t B_BN_G2[2][4] = {{0x3267e6dc24a138e5ull, 0xb5b4c5e559dbefa3ull, 0x81be18991be06ac3ull, 0x2b149d40ceb8aaaeull},
                                       {0xe4a2bd0685c315d2ull, 0xa74fa084e52d1852ull, 0xcd2cafadeed8fdf4ull, 0x009713b03af0fed4ull}};
static const uint64_t B_BLS_G1[6] = {4, 0, 0, 0, 0, 0};
static const uint64_t ORDER[2][4] = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                                     {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull}};

static const uint64_t PSI_BN[4][4] = {{0x99e39557176f553dull, 0xb78cc310c2c3330cull, 0x4c0bec3cf559b143ull, 0x2fb347984f7911f7ull}, {0x1665d51c640fcba2ull, 0x32ae2a1d0b7c9dceull, 0x4ba4cc8bd75a0794ull, 0x16c9e55061ebae20ull},
    {0xdc54014671a0135aull, 0xdbaae0eda9c95998ull, 0xdc5ec698b6e2f9b9ull, 0x063cf305489af5dcull}, {0x82d37f632623b0e3ull, 0x21807dc98fa25bd2ull, 0x0704b5a7ec796f2bull, 0x07c03cbcac41049aull}};   // psi_x.c0, psi_x.c1, psi_y.c0, psi_y.c1
static const uint64_t PSI_BLS[4][6] = {{0x0000000000000000ull, 0x0000000000000000ull, 0x0000000000000000ull, 0x0000000000000000ull, 0x0000000000000000ull, 0x0000000000000000ull}, {0x8bfd00000000aaadull, 0x409427eb4f49fffdull, 0x897d29650fb85f9bull, 0xaa0d857d89759ad4ull, 0xec02408663d4de85ull, 0x1a0111ea397fe699ull},
    {0xf1ee7b04121bdea2ull, 0x304466cf3e67fa0aull, 0xef396489f61eb45eull, 0x1c3dedd930b1cf60ull, 0xe2e9c448d77a2cd9ull, 0x135203e60180a68eull}, {0xc81084fbede3cc09ull, 0xee67992f72ec05f4ull, 0x77f76e17009241c5ull, 0x48395dabc2d3435eull, 0x6831e36d6bd17ffeull, 0x06af0e0437ff400bull}};   // psi_x.c0, psi_x.c1, psi_y.c0, psi_y.c1
static uint32_t* g_order = nullptr;               // device copy of the two group orders (8 limbs each)
static unsigned long long* g_status = nullptr;    // device status word

static int codec_prepare() {
  if (!g_order) {
    ZKB_CUDA(cudaMalloc((void**)&g_order, sizeof(ORDER)));
    ZKB_CUDA(cudaMemcpy(g_order, ORDER, sizeof(ORDER), cudaMemcpyHostToDevice));
    ZKB_CUDA(cudaMalloc((void**)&g_status, sizeof(unsigned long long)));
  }
  return ZKB_OK;
}

size_t compressed_bytes(int curve, int group) { return fq_bytes(curve) * (group == 2 ? 2 : 1); }

int points_compress_dev(int curve, int group, const void* d_pts, size_t n, void* d_out) {
  if (n == 0) return ZKB_OK;
  if (curve == ZKB_BN254) {
    return group == 1 ? compress_run<Wire1<FqBN254, false>, FqBN254, false>(n, d_pts, d_out)
                      : compress_run<Wire2<FqBN254, false>, FqBN254, false>(n, d_pts, d_out);
  }
  return group == 1 ? compress_run<Wire1<FqBLS381, true>, FqBLS381, true>(n, d_pts, d_out)
                    : compress_run<Wire2<FqBLS381, true>, FqBLS381, true>(n, d_pts, d_out);
}

// d_in: n compressed encodings (4-byte aligned); *bad = (index << 8 | reason) of the first invalid point, or ~0 when all are fine
int points_decompress_dev(int curve, int group, const void* d_in, size_t n, int validate, void* d_pts, unsigned long long* bad) {
  *bad = ~0ull;
  if (n == 0) return ZKB_OK;
  int rc = codec_prepare();
  if (rc) return rc;
  ZKB_CUDA(cudaMemsetAsync(g_status, 0xff, sizeof(unsigned long long), S()));
  const uint32_t* d_order = g_order + 8 * (curve == ZKB_BN254 ? 0 : 1);
  // validate: 0 = no subgroup check, 1 = fast criterion in G2 (r * P in G1), 2 = r * P everywhere
  if (curve == ZKB_BN254) {
    if (group == 1) {
      CurveB<Fp<FqBN254>> cb;
      memset(&cb, 0, sizeof(cb));
      cb.b = host_to_mont<FqBN254>(B_BN_G1);
      rc = decompress_run<Wire1<FqBN254, false>, FqBN254, false, 0>(n, d_in, cb, d_order, d_pts, g_status);   // cofactor 1
    } else {
      CurveB<Fp2<FqBN254>> cb;
      cb.b.c0 = host_to_mont<FqBN254>(B_BN_G2[0]);
      cb.b.c1 = host_to_mont<FqBN254>(B_BN_G2[1]);
      cb.psi_x.c0 = host_to_mont<FqBN254>(PSI_BN[0]);
      cb.psi_x.c1 = host_to_mont<FqBN254>(PSI_BN[1]);
      cb.psi_y.c0 = host_to_mont<FqBN254>(PSI_BN[2]);
      cb.psi_y.c1 = host_to_mont<FqBN254>(PSI_BN[3]);
      typedef Wire2<FqBN254, false> WW;
      rc = validate == 1   ? decompress_run<WW, FqBN254, false, 1>(n, d_in, cb, d_order, d_pts, g_status)
           : validate == 2 ? decompress_run<WW, FqBN254, false, 2>(n, d_in, cb, d_order, d_pts, g_status)
                           : decompress_run<WW, FqBN254, false, 0>(n, d_in, cb, d_order, d_pts, g_status);
    }
  } else {
    if (group == 1) {
      CurveB<Fp<FqBLS381>> cb;
      memset(&cb, 0, sizeof(cb));
      cb.b = host_to_mont<FqBLS381>(B_BLS_G1);
      typedef Wire1<FqBLS381, true> WW;
      rc = validate ? decompress_run<WW, FqBLS381, true, 2>(n, d_in, cb, d_order, d_pts, g_status)
                    : decompress_run<WW, FqBLS381, true, 0>(n, d_in, cb, d_order, d_pts, g_status);
    } else {
      CurveB<Fp2<FqBLS381>> cb;
      cb.b.c0 = host_to_mont<FqBLS381>(B_BLS_G1);
      cb.b.c1 = cb.b.c0;
      cb.psi_x.c0 = host_to_mont<FqBLS381>(PSI_BLS[0]);
      cb.psi_x.c1 = host_to_mont<FqBLS381>(PSI_BLS[1]);
      cb.psi_y.c0 = host_to_mont<FqBLS381>(PSI_BLS[2]);
      cb.psi_y.c1 = host_to_mont<FqBLS381>(PSI_BLS[3]);
      typedef Wire2<FqBLS381, true> WW;
      rc = validate == 1   ? decompress_run<WW, FqBLS381, true, 1>(n, d_in, cb, d_order, d_pts, g_status)
           : validate == 2 ? decompress_run<WW, FqBLS381, true, 2>(n, d_in, cb, d_order, d_pts, g_status)
                           : decompress_run<WW, FqBLS381, true, 0>(n, d_in, cb, d_order, d_pts, g_status);
    }
  }
  if (rc) return rc;
  ZKB_CUDA(cudaMemcpyAsync(bad, g_status, sizeof(unsigned long long), cudaMemcpyDeviceToHost, S()));
  ZKB_CUDA(cudaStreamSynchronize(S()));
  return ZKB_OK;
}

}  // namespace zkb
