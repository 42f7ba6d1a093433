// ec.cuh -- short-Weierstrass (a = 0) group arithmetic in XYZZ coordinates over any field F from ff.cuh.
//
// Replaces what the reference gets from ark-ec 0.4.2 `Projective` (Jacobian, CPU) behind
// /root/reference/src/bn254/curve.rs:74-118 (point + - * ) and the bucket arithmetic inside
// `VariableBaseMSM::msm` (curve.rs:366,385).  XYZZ (x = X/ZZ, y = Y/ZZZ) makes the bucket update a
// 8M+2S mixed addition with an affine base point and needs no field inversion until the very end.
// All special cases (identity operand, P+P, P-P) are handled exactly -- results are group-law exact, not
// "with overwhelming probability".
#pragma once
#include "ff.cuh"

namespace zkb {

template <class F>
struct Affine {
  F x, y;  // (0,0) encodes the point at infinity (not on either curve since b != 0)
  ZKB_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  ZKB_HD static Affine inf() { Affine r; r.x = F::zero(); r.y = F::zero(); return r; }
};

template <class F>
struct XYZZ {
  F X, Y, ZZ, ZZZ;  // ZZ == 0 encodes the identity
  ZKB_HD bool is_inf() const { return ZZ.is_zero(); }
  ZKB_HD static XYZZ inf() {
    XYZZ r;
    r.X = F::zero(); r.Y = F::zero(); r.ZZ = F::zero(); r.ZZZ = F::zero();
    return r;
  }
  ZKB_HD static XYZZ from_affine(const Affine<F>& p) {
    XYZZ r;
    if (p.is_inf()) return inf();
    r.X = p.x; r.Y = p.y; r.ZZ = F::one(); r.ZZZ = F::one();
    return r;
  }
};

// 2*(x,y) for an affine point, y != 0 guaranteed by prime group order (no 2-torsion in G1/G2)
template <class F>
ZKB_HD XYZZ<F> dbl_affine(const Affine<F>& p) {
  XYZZ<F> r;
  F U = dbl(p.y);
  F V = sqr(U);
  F W = U * V;
  F S = p.x * V;
  F X2 = sqr(p.x);
  F M = dbl(X2) + X2;
  r.X = sqr(M) - dbl(S);
  r.Y = M * (S - r.X) - W * p.y;
  r.ZZ = V;
  r.ZZZ = W;
  return r;
}

template <class F>
ZKB_HD XYZZ<F> dbl(const XYZZ<F>& p) {
  if (p.is_inf()) return p;
  XYZZ<F> r;
  F U = dbl(p.Y);
  F V = sqr(U);
  F W = U * V;
  F S = p.X * V;
  F X2 = sqr(p.X);
  F M = dbl(X2) + X2;
  r.X = sqr(M) - dbl(S);
  r.Y = M * (S - r.X) - W * p.Y;
  r.ZZ = V * p.ZZ;
  r.ZZZ = W * p.ZZZ;
  return r;
}

// acc += (x, y)  or  acc -= (x, y) when negate  (mixed addition, 8M + 2S on the common path)
template <class F>
ZKB_HD void madd(XYZZ<F>& acc, const Affine<F>& q, bool negate = false) {
  if (q.is_inf()) return;
  F qy = negate ? neg(q.y) : q.y;
  if (acc.is_inf()) {
    acc.X = q.x; acc.Y = qy; acc.ZZ = F::one(); acc.ZZZ = F::one();
    return;
  }
  F U2 = q.x * acc.ZZ;
  F S2 = qy * acc.ZZZ;
  F Pp = U2 - acc.X;
  F R = S2 - acc.Y;
  if (Pp.is_zero()) {
    if (R.is_zero()) {
      Affine<F> t; t.x = q.x; t.y = qy;
      acc = dbl_affine(t);
    } else {
      acc = XYZZ<F>::inf();
    }
    return;
  }
  F PP = sqr(Pp);
  F PPP = Pp * PP;
  F Q = acc.X * PP;
  F X3 = sqr(R) - PPP - dbl(Q);
  acc.Y = R * (Q - X3) - acc.Y * PPP;
  acc.X = X3;
  acc.ZZ = acc.ZZ * PP;
  acc.ZZZ = acc.ZZZ * PPP;
}

// full addition XYZZ + XYZZ (12M + 2S)
template <class F>
ZKB_HD XYZZ<F> add(const XYZZ<F>& p, const XYZZ<F>& q) {
  if (p.is_inf()) return q;
  if (q.is_inf()) return p;
  F U1 = p.X * q.ZZ;
  F U2 = q.X * p.ZZ;
  F S1 = p.Y * q.ZZZ;
  F S2 = q.Y * p.ZZZ;
  F Pp = U2 - U1;
  F R = S2 - S1;
  if (Pp.is_zero()) {
    if (R.is_zero()) return dbl(p);
    return XYZZ<F>::inf();
  }
  XYZZ<F> r;
  F PP = sqr(Pp);
  F PPP = Pp * PP;
  F Q = U1 * PP;
  r.X = sqr(R) - PPP - dbl(Q);
  r.Y = R * (Q - r.X) - S1 * PPP;
  r.ZZ = p.ZZ * q.ZZ * PP;
  r.ZZZ = p.ZZZ * q.ZZZ * PPP;
  return r;
}

template <class F>
ZKB_HD XYZZ<F> neg(const XYZZ<F>& p) {
  XYZZ<F> r = p;
  r.Y = neg(p.Y);
  return r;
}

// k * p for a little-endian limb scalar (double-and-add, MSB first)
template <class F>
ZKB_HD XYZZ<F> scalar_mul(const Affine<F>& p, const uint32_t* k, int nlimbs) {
  XYZZ<F> acc = XYZZ<F>::inf();
  bool started = false;
  for (int i = nlimbs - 1; i >= 0; i--) {
    for (int bit = 31; bit >= 0; bit--) {
      if (started) acc = dbl(acc);
      if ((k[i] >> bit) & 1) {
        madd(acc, p);
        started = true;
      }
    }
  }
  return acc;
}

// small-integer multiple of an XYZZ point (bucket-chunk weights)
template <class F>
ZKB_HD XYZZ<F> mul_small(const XYZZ<F>& p, uint32_t k) {
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int bit = 31; bit >= 0; bit--) {
    acc = dbl(acc);
    if ((k >> bit) & 1) acc = add(acc, p);
  }
  return acc;
}

template <class F>
ZKB_HD Affine<F> to_affine(const XYZZ<F>& p) {
  if (p.is_inf()) return Affine<F>::inf();
  // x = X/ZZ, y = Y/ZZZ ; one inversion: i = 1/(ZZ*ZZZ) -> 1/ZZ = i*ZZZ, 1/ZZZ = i*ZZ
  F i = inv(p.ZZ * p.ZZZ);
  Affine<F> r;
  r.x = p.X * (i * p.ZZZ);
  r.y = p.Y * (i * p.ZZ);
  return r;
}

typedef Affine<fq_bn> g1a_bn;
typedef Affine<fq2_bn> g2a_bn;
typedef Affine<fq_bls> g1a_bls;
typedef Affine<fq2_bls> g2a_bls;

}  // namespace zkb
