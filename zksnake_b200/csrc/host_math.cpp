// host_math.cpp -- the few scalar group/field operations that follow a kernel, done on the host with 64-bit limbs.
// Compiled by plain g++ (ZKB_HD expands to `inline`), so the ec.cuh templates instantiate over the host field Fh<P>.
#include <string.h>
#include <vector>
#include "ec.cuh"
#include "ff_host.h"
#include "zkb_internal.h"

namespace zkb {

template <class F>
static void finish_t(const void* sums_, uint32_t nwin, uint32_t win0, uint32_t c, uint32_t nlev, const uint32_t* logk, const uint32_t* parts,
                     uint32_t nbits, uint64_t* out_xy, int* out_inf) {
  const XYZZ<F>* sums = (const XYZZ<F>*)sums_;
  uint32_t first[8], nu = 0;   // first job of level l
  for (uint32_t l = 0; l < nlev; l++) {
    first[l] = nu;
    nu += parts[l];
  }
  const uint32_t njobs = nu + nbits + 1;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int i = (int)nwin - 1; i >= 0; i--) {
    if (!acc.is_inf())
      for (uint32_t k = 0; k < c; k++) acc = dbl(acc);
    const XYZZ<F>* s = sums + (size_t)i * njobs;
    // ws = sum_beta 2^beta A_beta
    XYZZ<F> ws = XYZZ<F>::inf();
    for (int beta = (int)nbits - 1; beta >= 0; beta--) {
      if (!ws.is_inf()) ws = dbl(ws);
      ws = add(ws, s[nu + beta]);
    }
    for (int l = (int)nlev - 1; l >= 0; l--) {
      if (!ws.is_inf())
        for (uint32_t k = 0; k < logk[l]; k++) ws = dbl(ws);
      for (uint32_t p = 0; p < parts[l]; p++) ws = add(ws, s[first[l] + p]);
    }
    ws = add(ws, s[njobs - 1]);
    acc = add(acc, ws);
  }
  if (!acc.is_inf())
    for (uint32_t k = 0; k < c * win0; k++) acc = dbl(acc);   // window shard: the lowest window here is win0
  Affine<F> a = to_affine(acc);
  *out_inf = a.is_inf() ? 1 : 0;
  a.x = from_mont(a.x);
  a.y = from_mont(a.y);
  memcpy(out_xy, &a, sizeof(a));
}

void host_msm_finish(int curve, int group, const void* sums, uint32_t nwin, uint32_t win0, uint32_t c, uint32_t nlev,
                     const uint32_t* logk, const uint32_t* parts, uint32_t nbits, uint64_t* out_xy, int* out_inf) {
  if (curve == ZKB_BN254 && group == 1) finish_t<Fh<FqBN254>>(sums, nwin, win0, c, nlev, logk, parts, nbits, out_xy, out_inf);
  else if (curve == ZKB_BN254) finish_t<Fh2<FqBN254>>(sums, nwin, win0, c, nlev, logk, parts, nbits, out_xy, out_inf);
  else if (group == 1) finish_t<Fh<FqBLS381>>(sums, nwin, win0, c, nlev, logk, parts, nbits, out_xy, out_inf);
  else finish_t<Fh2<FqBLS381>>(sums, nwin, win0, c, nlev, logk, parts, nbits, out_xy, out_inf);
}

// sum_i k_i * P_i ; scalars[i] == nullptr means k_i = 1.  Interleaved (Straus) double-and-add over all terms.
template <class F>
static void lincomb_t(int n_terms, const uint64_t* const* points, const int* infs, const uint64_t* const* scalars,
                      uint64_t* out_xy, int* out_inf) {
  std::vector<Affine<F>> pts(n_terms);
  for (int i = 0; i < n_terms; i++) {
    if (infs && infs[i]) {
      pts[i] = Affine<F>::inf();
    } else {
      memcpy(&pts[i], points[i], sizeof(Affine<F>));
      pts[i].x = to_mont(pts[i].x);
      pts[i].y = to_mont(pts[i].y);
    }
  }
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int bit = 255; bit >= 0; bit--) {
    if (!acc.is_inf()) acc = dbl(acc);
    for (int i = 0; i < n_terms; i++) {
      if (!scalars[i]) continue;
      if ((scalars[i][bit >> 6] >> (bit & 63)) & 1) madd(acc, pts[i]);
    }
  }
  for (int i = 0; i < n_terms; i++)
    if (!scalars[i]) madd(acc, pts[i]);
  Affine<F> a = to_affine(acc);
  *out_inf = a.is_inf() ? 1 : 0;
  a.x = from_mont(a.x);
  a.y = from_mont(a.y);
  memcpy(out_xy, &a, sizeof(a));
}

void host_lincomb(int curve, int group, int n_terms, const uint64_t* const* points, const int* infs,
                  const uint64_t* const* scalars, uint64_t* out_xy, int* out_inf) {
  if (curve == ZKB_BN254 && group == 1) lincomb_t<Fh<FqBN254>>(n_terms, points, infs, scalars, out_xy, out_inf);
  else if (curve == ZKB_BN254) lincomb_t<Fh2<FqBN254>>(n_terms, points, infs, scalars, out_xy, out_inf);
  else if (group == 1) lincomb_t<Fh<FqBLS381>>(n_terms, points, infs, scalars, out_xy, out_inf);
  else lincomb_t<Fh2<FqBLS381>>(n_terms, points, infs, scalars, out_xy, out_inf);
}

template <class F>
static void fr_mul_t(const uint64_t* a, const uint64_t* b, uint64_t* out) {
  F x, y;
  memcpy(x.v, a, 32);
  memcpy(y.v, b, 32);
  F r = from_mont(to_mont(x) * to_mont(y));
  memcpy(out, r.v, 32);
}
void host_fr_mul(int curve, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  if (curve == ZKB_BN254) fr_mul_t<Fh<FrBN254>>(a, b, out);
  else fr_mul_t<Fh<FrBLS381>>(a, b, out);
}

}  // namespace zkb
