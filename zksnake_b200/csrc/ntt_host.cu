// ntt_host.cu -- host drivers for the Fr NTT passes, element-wise kernels and the Groth16 quotient pipeline.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <map>
#include <vector>
#include "ntt.cuh"
#include "ntt_warp.cuh"
#include "zkb_internal.h"

namespace zkb {

static inline cudaStream_t S() { return (cudaStream_t)ctx_stream(); }

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

// ------------------------------------------------------------------------------------------------------
// power tables
// ------------------------------------------------------------------------------------------------------
template <class F>
struct DevTable {
  F* lo = nullptr;
  F* hi = nullptr;
  uint32_t h = 0;
  uint32_t direct = 0;
  PowTable<F> view() const { PowTable<F> t; t.lo = lo; t.hi = hi; t.h = h; t.direct = direct; return t; }
};

// lo[j] = base^j (j < 2^h), hi[j] = scale * base^(j 2^h) (j < 2^(log_n - h))
template <class F>
static int build_table(DevTable<F>& t, const F& base, const F& scale, uint32_t log_n, bool direct = false) {
  uint32_t h = direct ? log_n : (log_n + 1) / 2;
  t.direct = direct ? 1 : 0;
  t.h = h;
  size_t nlo = (size_t)1 << h, nhi = (size_t)1 << (log_n - h);
  ZKB_CUDA(cudaMalloc((void**)&t.lo, nlo * sizeof(F)));
  ZKB_CUDA(cudaMalloc((void**)&t.hi, nhi * sizeof(F)));
  pow_table_kernel<F><<<(unsigned)((nlo + 127) / 128), 128, 0, S()>>>(t.lo, base, F::one(), nlo, 0);
  pow_table_kernel<F><<<(unsigned)((nhi + 127) / 128), 128, 0, S()>>>(t.hi, base, scale, nhi, h);
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}

template <class F>
struct Domain {
  DevTable<F> fwd, inv;          // w^j, w^-j
  DevTable<F> inv_scaled;        // w^-j * N^-1      (coset_ifft post-scale)
  DevTable<F> gen_pre, gen_pre_r;  // g^j, g^j / R     (H pipeline coset, g = multiplicative generator)
  DevTable<F> gen_post;          // g^-j * R / (N (g^N - 1))
  DevTable<F> gen_ipost;         // g^-j / N                  (inverse transform from the coset g<w>)
  F n_inv;                       // Montgomery(N^-1)
  bool have_gen = false;
};

template <class F>
static F host_root(uint32_t log_n, bool inverse) {
  typedef typename F::Params P;
  F w;
  for (int i = 0; i < F::N; i++) w.v[i] = inverse ? P::ROOT_INV(i) : P::ROOT(i);
  for (uint32_t i = log_n; i < (uint32_t)P::TWO_ADICITY; i++) w = sqr(w);
  return w;
}
template <class F>
static F host_small(uint64_t k) {  // Montgomery(k)
  F x = F::zero();
  x.v[0] = (uint32_t)k;
  x.v[1] = (uint32_t)(k >> 32);
  return to_mont(x);
}
template <class F> struct GenOf;
template <> struct GenOf<fr_bn> { static constexpr uint64_t g = 5; };
template <> struct GenOf<fr_bls> { static constexpr uint64_t g = 7; };

template <class F>
static std::map<uint32_t, Domain<F>>& domains() {
  static std::map<uint32_t, Domain<F>> m;
  return m;
}

template <class F>
static int get_domain(uint32_t log_n, bool need_gen, Domain<F>** out) {
  auto& m = domains<F>();
  auto it = m.find(log_n);
  if (it == m.end()) {
    Domain<F> d;
    F w = host_root<F>(log_n, false), wi = host_root<F>(log_n, true);
    d.n_inv = inv(host_small<F>(1ull << log_n));
    int rc;
    const bool direct = log_n <= (uint32_t)env_int("ZKB_NTT_DIRECT_MAX", 26);
    if ((rc = build_table(d.fwd, w, F::one(), log_n, direct))) return rc;
    if ((rc = build_table(d.inv, wi, F::one(), log_n, direct))) return rc;
    if ((rc = build_table(d.inv_scaled, wi, d.n_inv, log_n))) return rc;
    it = m.emplace(log_n, d).first;
  }
  Domain<F>& d = it->second;
  if (need_gen && !d.have_gen) {
    F g = host_small<F>(GenOf<F>::g);
    F gi = inv(g);
    F one_raw = F::zero();
    one_raw.v[0] = 1;
    // a stored (Montgomery-form) value m stands for m/R, so the stored integer 1 stands for 1/R
    F r_inv = one_raw;
    // Z on the coset: g^N - 1
    F gn = g;
    for (uint32_t i = 0; i < log_n; i++) gn = sqr(gn);
    F z = gn - F::one();
    // post constant: R / (N * Z)   (R restores the factor lost by the raw Montgomery product in the pointwise step)
    F r_mont = F::r2();                 // stands for R
    F post_c = r_mont * d.n_inv * inv(z);
    int rc;
    if ((rc = build_table(d.gen_pre, g, F::one(), log_n))) return rc;
    if ((rc = build_table(d.gen_pre_r, g, r_inv, log_n))) return rc;
    if ((rc = build_table(d.gen_post, gi, post_c, log_n))) return rc;
    if ((rc = build_table(d.gen_ipost, gi, d.n_inv, log_n))) return rc;
    d.have_gen = true;
  }
  *out = &d;
  return ZKB_OK;
}

// ------------------------------------------------------------------------------------------------------
// pass planning
// ------------------------------------------------------------------------------------------------------

struct Plan {
  std::vector<uint32_t> k;   // radix bits per pass
  std::vector<uint32_t> log_c;
};

static Plan make_plan(uint32_t log_n) {
  Plan p;
  uint32_t maxk = (uint32_t)env_int("ZKB_NTT_MAXK", 11);
  if (maxk < 3) maxk = 3;
  if (maxk > 11) maxk = 11;
  uint32_t npass = log_n <= 10 ? 1 : (log_n + maxk - 1) / maxk;
  if (npass > 4) npass = 4;
  uint32_t rem = log_n;
  for (uint32_t i = 0; i < npass; i++) {
    uint32_t k = (rem + (npass - i) - 1) / (npass - i);
    p.k.push_back(k);
    rem -= k;
  }
  uint32_t max_tile = (uint32_t)env_int("ZKB_NTT_TILE", 2048);  // elements per tile
  for (uint32_t i = 0; i < npass; i++) {
    uint32_t lc = 0;
    if (npass > 1) {
      // C cannot exceed the column count of the pass (non-last passes) or R_1 (last pass)
      uint32_t limit = 0;
      if (i + 1 < npass) {
        for (uint32_t j = i + 1; j < npass; j++) limit += p.k[j];
      } else {
        limit = p.k[0];
      }
      lc = limit < 3 ? limit : 3;
      while (lc > 0 && ((1u << (p.k[i] + lc)) > max_tile)) lc--;
      while (lc > 0 && (log_n - p.k[i] - lc) < 9) lc--;   // keep >= 512 tiles so that all 148 SMs stay busy
    }
    p.log_c.push_back(lc);
  }
  return p;
}

static size_t pass_smem(uint32_t k, uint32_t log_c) {
  size_t tile = (size_t)1 << (k + log_c);
  return 2 * (tile + 4) * 16 + 2 * (((size_t)1 << k) / 2 + 4) * 16;
}

size_t ntt_scratch_bytes(uint32_t log_n) { return ((size_t)32 << log_n) + 256; }

// ------------------------------------------------------------------------------------------------------------------------
// register-resident passes (ntt_warp.cuh): used from 2^11 up; below that one shared-memory pass of ntt_pass_kernel does it
// ------------------------------------------------------------------------------------------------------------------------
// A lane holds 2^2 elements: a warp transforms columns of up to 2^7 elements (no spills at 128 registers).  Measured on a B200
// (profiles/R2_ntt_probe.md): for ONE transform the register-resident passes tie with the shared-memory passes at 2^20 (0.256 vs
// 0.262 ms) and lose from 2^21 up (an extra pass over the data: 1.11 vs 1.00 ms at 2^22), so zkb_ntt_dev keeps the shared-memory
// kernel; for the BATCHED transforms of the Groth16 quotient (three columns' worth of work per launch, 9 launches instead of 18)
// they win at every size (2^12: 0.116 vs 0.161 ms, 2^16: 0.206 vs 0.222, 2^20: 1.98 vs 2.06), so groth16_h_t uses them.
static inline bool use_warp_ntt(uint32_t log_n) { return log_n >= 11 && log_n <= 28 && env_int("ZKB_NTT_WARP", 1) != 0; }
static inline int warp_ntt_el(uint32_t) { return 2; }

static Plan make_warp_plan(uint32_t log_n, int el) {
  Plan p;
  const uint32_t kmax = (uint32_t)el + 5;
  uint32_t npass = (log_n + kmax - 1) / kmax;
  uint32_t rem = log_n;
  for (uint32_t i = 0; i < npass; i++) {
    uint32_t k = (rem + (npass - i) - 1) / (npass - i);
    p.k.push_back(k);
    p.log_c.push_back(0);
    rem -= k;
  }
  return p;
}

extern template int ntt_warp_launch_el<fr_bn, 2>(const fr_bn*, fr_bn*, const NttPass&, uint32_t, size_t, size_t, const PowTable<fr_bn>&,
                                                 const PreTables<fr_bn>&, const PowTable<fr_bn>&, const fr_bn&, void*);
extern template int ntt_warp_launch_el<fr_bls, 2>(const fr_bls*, fr_bls*, const NttPass&, uint32_t, size_t, size_t,
                                                  const PowTable<fr_bls>&, const PreTables<fr_bls>&, const PowTable<fr_bls>&,
                                                  const fr_bls&, void*);
// `batch` transforms of the same size and direction, member b at in + b * stride / out + b * stride (stride in elements);
// pre[b]: first-pass scaling table of member b (nullptr: none); tmp holds batch * 2^log_n elements.
template <class F>
static int ntt_exec_warp(const F* in, size_t in_len, F* out, uint32_t log_n, uint32_t batch, size_t stride, const PowTable<F>& tw,
                         const PowTable<F>* const* pre, const PowTable<F>* post, const F* post_const, F* tmp) {
  if (batch == 0 || batch > ZKB_NTT_MAX_BATCH) return set_error(ZKB_ERR_ARG, "ntt: batch out of range");
  const int el = warp_ntt_el(log_n);
  Plan plan = make_warp_plan(log_n, el);
  const uint32_t P = (uint32_t)plan.k.size();
  if (P < 2 || P > 4) return set_error(ZKB_ERR_ARG, "ntt: unsupported size for the register-resident path");
  size_t n = (size_t)1 << log_n;
  if (in_len > n) in_len = n;
  PowTable<F> none;
  none.lo = none.hi = nullptr;
  none.h = 0;
  none.direct = 0;
  PreTables<F> pres;
  bool any_pre = false;
  for (uint32_t b = 0; b < ZKB_NTT_MAX_BATCH; b++) {
    pres.t[b] = (pre && b < batch && pre[b]) ? *pre[b] : none;
    any_pre |= pre && b < batch && pre[b];
  }
  if (any_pre)
    for (uint32_t b = 0; b < batch; b++)
      if (!pre[b]) return set_error(ZKB_ERR_ARG, "ntt: a batch mixes scaled and unscaled members");
  uint32_t log_b = 0, log_m = log_n;
  prof_begin(PROF_NTT);
  for (uint32_t p = 0; p < P; p++) {
    NttPass pp;
    memset(&pp, 0, sizeof(pp));
    pp.log_n = log_n;
    pp.k = plan.k[p];
    pp.log_c = 0;
    log_m -= pp.k;
    pp.log_m = log_m;
    pp.log_b = log_b;
    pp.last = (p + 1 == P);
    pp.k1 = plan.k[0];
    pp.nmid = P - 2;
    pp.kmid[0] = P > 2 ? plan.k[1] : 0;
    pp.kmid[1] = P > 3 ? plan.k[2] : 0;
    pp.pre = (p == 0 && any_pre) ? 1 : 0;
    pp.post = 0;
    if (pp.last) pp.post = post ? 1 : (post_const ? 2 : 0);
    pp.in_len = (p == 0) ? in_len : n;
    const F* src = (p == 0) ? in : tmp;
    F* dst = pp.last ? out : tmp;
    const size_t src_stride = (p == 0) ? stride : n, dst_stride = pp.last ? stride : n;
    const PowTable<F>& postv = post ? *post : none;
    const F pc = post_const ? *post_const : F::zero();
    int lrc = ntt_warp_launch_el<F, 2>(src, dst, pp, batch, src_stride, dst_stride, tw, pres, postv, pc, (void*)S());
    if (lrc) {
      prof_end(PROF_NTT);
      return set_error(ZKB_ERR_ARG, "ntt: pass radix outside the instantiated range");
    }
    count_launch();
    log_b += pp.k;
  }
  prof_end(PROF_NTT);
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}

template <class F>
static int ntt_exec(const F* in, size_t in_len, F* out, uint32_t log_n, const PowTable<F>& tw, const PowTable<F>* pre,
                    const PowTable<F>* post, const F* post_const, F* tmp) {
  static bool attr_set = false;
  if (!attr_set) {
    ZKB_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  if (use_warp_ntt(log_n) && (log_n == 11 || env_int("ZKB_NTT_WARP_SINGLE", 0))) {   // (2^11: 0.029 vs 0.045 ms)
    const PowTable<F>* pres[1] = {pre};
    return ntt_exec_warp<F>(in, in_len, out, log_n, 1, (size_t)1 << log_n, tw, pre ? pres : nullptr, post, post_const, tmp);
  }
  Plan plan = make_plan(log_n);
  const uint32_t P = (uint32_t)plan.k.size();
  size_t n = (size_t)1 << log_n;
  if (in_len > n) in_len = n;  // ark truncates inputs longer than the domain
  uint32_t log_b = 0, log_m = log_n;
  PowTable<F> none;
  none.lo = none.hi = nullptr;
  none.h = 0;
  none.direct = 0;
  prof_begin(PROF_NTT);
  for (uint32_t p = 0; p < P; p++) {
    NttPass pp;
    memset(&pp, 0, sizeof(pp));
    pp.log_n = log_n;
    pp.k = plan.k[p];
    pp.log_c = plan.log_c[p];
    log_m -= pp.k;
    pp.log_m = log_m;
    pp.log_b = log_b;
    pp.last = (p + 1 == P);
    pp.k1 = (P > 1) ? plan.k[0] : 0;
    pp.nmid = (P > 2) ? (P - 2) : 0;
    pp.kmid[0] = P > 2 ? plan.k[1] : 0;
    pp.kmid[1] = P > 3 ? plan.k[2] : 0;
    pp.pre = (p == 0 && pre) ? 1 : 0;
    pp.post = 0;
    if (pp.last) pp.post = post ? 1 : (post_const ? 2 : 0);
    pp.in_len = (p == 0) ? in_len : n;
    const F* src = (p == 0) ? in : tmp;
    F* dst = pp.last ? out : tmp;
    size_t tiles = n >> (pp.k + pp.log_c);
    uint32_t half = 1u << (pp.k + pp.log_c) >> 1;
    uint32_t threads = half < 32 ? 32 : (half > 512 ? 512 : half);
    size_t smem = pass_smem(pp.k, pp.log_c);
    ntt_pass_kernel<F><<<(unsigned)tiles, threads, smem, S()>>>(src, dst, pp, tw, pre ? *pre : none, post ? *post : none,
                                                                post_const ? *post_const : F::zero());
    count_launch();
    log_b += pp.k;
  }
  prof_end(PROF_NTT);
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}

template <class F>
static int ntt_dev_t(int inverse, int coset, uint32_t log_n, const void* d_in, size_t in_len, void* d_out) {
  typedef typename F::Params P;
  if (log_n > (uint32_t)P::TWO_ADICITY || log_n > 30)
    return set_error(ZKB_ERR_DOMAIN, "Domain size is too large");
  Domain<F>* d;
  int rc = get_domain<F>(log_n, coset == 2, &d);
  if (rc) return rc;
  F* tmp = nullptr;
  if (log_n > 10) {
    if ((rc = scratch_reserve(ntt_scratch_bytes(log_n)))) return rc;
    scratch_reset();
    tmp = (F*)scratch_take((size_t)32 << log_n);
  }
  PowTable<F> fwd = d->fwd.view(), inv_t = d->inv.view(), inv_s = d->inv_scaled.view();
  if (coset == 2) {   // coset g <w>, g = the field's multiplicative generator (5 / 7)
    PowTable<F> gpre = d->gen_pre.view(), gipost = d->gen_ipost.view();
    if (!inverse) return ntt_exec<F>((const F*)d_in, in_len, (F*)d_out, log_n, fwd, &gpre, nullptr, nullptr, tmp);
    return ntt_exec<F>((const F*)d_in, in_len, (F*)d_out, log_n, inv_t, nullptr, &gipost, nullptr, tmp);
  }
  if (!inverse) return ntt_exec<F>((const F*)d_in, in_len, (F*)d_out, log_n, fwd, coset ? &fwd : nullptr, nullptr, nullptr, tmp);
  if (coset) return ntt_exec<F>((const F*)d_in, in_len, (F*)d_out, log_n, inv_t, nullptr, &inv_s, nullptr, tmp);
  return ntt_exec<F>((const F*)d_in, in_len, (F*)d_out, log_n, inv_t, nullptr, nullptr, &d->n_inv, tmp);
}

int ntt_dev(int curve, int inverse, int coset, uint32_t log_n, const void* d_in, size_t in_len, void* d_out) {
  if (curve == ZKB_BN254) return ntt_dev_t<fr_bn>(inverse, coset, log_n, d_in, in_len, d_out);
  if (curve == ZKB_BLS12_381) return ntt_dev_t<fr_bls>(inverse, coset, log_n, d_in, in_len, d_out);
  return set_error(ZKB_ERR_ARG, "unknown curve id");
}

// ------------------------------------------------------------------------------------------------------
// element-wise
// ------------------------------------------------------------------------------------------------------
template <class F>
static int vec_op_t(int op, size_t n, const void* a, size_t na, const void* b, size_t nb, const void* c, void* out) {
  if (n == 0) return ZKB_OK;
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  prof_begin(PROF_VEC);
  vec_op_kernel<F><<<blocks, 256, 0, S()>>>(op, n, (const F*)a, na, (const F*)b, nb, (const F*)c, (F*)out);
  prof_end(PROF_VEC);
  count_launch();
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}
int vec_op_dev(int curve, int op, size_t n, const void* a, size_t na, const void* b, size_t nb, const void* c, void* out) {
  if (curve == ZKB_BN254) return vec_op_t<fr_bn>(op, n, a, na, b, nb, c, out);
  if (curve == ZKB_BLS12_381) return vec_op_t<fr_bls>(op, n, a, na, b, nb, c, out);
  return set_error(ZKB_ERR_ARG, "unknown curve id");
}

template <class F>
static int reduce_t(size_t n, void* v) {
  if (n == 0) return ZKB_OK;
  prof_begin(PROF_VEC);
  reduce_kernel<F><<<(unsigned)((n + 255) / 256), 256, 0, S()>>>(n, (F*)v);
  prof_end(PROF_VEC);
  count_launch();
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}
int fr_reduce_dev(int curve, size_t n, void* v) {
  if (curve == ZKB_BN254) return reduce_t<fr_bn>(n, v);
  if (curve == ZKB_BLS12_381) return reduce_t<fr_bls>(n, v);
  return set_error(ZKB_ERR_ARG, "unknown curve id");
}

// out[i] = scale * base^i, canonical
template <class F>
__global__ void powers_canonical_kernel(F* out, F base, F scale, unsigned long long n) {
  unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (j >= n) return;
  F r = from_mont(pow_u64(base, j) * scale);
  ntt_st(out + j, r);
}
template <class F>
static int powers_t(const uint64_t* base, const uint64_t* scale, size_t n, void* d_out) {
  F b, s;
  memcpy(b.v, base, 32);
  memcpy(s.v, scale, 32);
  b = to_mont(b);
  s = to_mont(s);
  if (n == 0) return ZKB_OK;
  powers_canonical_kernel<F><<<(unsigned)((n + 127) / 128), 128, 0, S()>>>((F*)d_out, b, s, n);
  count_launch();
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}
int fr_powers_dev(int curve, const uint64_t* base, const uint64_t* scale, size_t n, void* d_out) {
  if (curve == ZKB_BN254) return powers_t<fr_bn>(base, scale, n, d_out);
  if (curve == ZKB_BLS12_381) return powers_t<fr_bls>(base, scale, n, d_out);
  return set_error(ZKB_ERR_ARG, "unknown curve id");
}

// ------------------------------------------------------------------------------------------------------
// Groth16 quotient:  H = (U V - W) / (X^n - 1)
//   reference: QAP.evaluate_witness, /root/reference/python/zksnake/groth16/qap.py:42-71 (3 iNTT(n), 2 NTT(2n), pointwise,
//   iNTT(2n), subtract, divide_by_vanishing_poly).  H is unique, so it is computed here on the coset g*<w> of the SAME size n:
//   3 iNTT(n) + 3 coset-NTT(n) + pointwise + 1 coset-iNTT(n); Z is the constant g^n - 1 on that coset.
// ------------------------------------------------------------------------------------------------------
// One pinned int for the satisfiability flag of check_abc_kernel when the caller defers the test (check == 2): the copy is
// enqueued behind the kernel and nobody waits for it -- a proof that is going to fail fails after its MSMs instead of before them,
// and a proof that is not (every proof but a caller's mistake) never stops the host between the transforms and the MSM batch.
int* groth16_flag_host() {
  static int* p = nullptr;
  if (!p && cudaHostAlloc((void**)&p, 64, cudaHostAllocDefault) != cudaSuccess) p = nullptr;
  return p;
}
template <class F>
static int groth16_h_t(uint32_t log_n, const void* d_a, const void* d_b, const void* d_c, void* d_u, void* d_v, void* d_w,
                       void* d_h, int check, void (*after_interp)(void*), void* arg) {
  typedef typename F::Params P;
  if (log_n > (uint32_t)P::TWO_ADICITY || log_n > 30) return set_error(ZKB_ERR_DOMAIN, "Domain size is too large");
  size_t n = (size_t)1 << log_n;
  Domain<F>* d;
  int rc = get_domain<F>(log_n, true, &d);
  if (rc) return rc;
  size_t need = 6 * ((size_t)32 << log_n) + 8192;
  if ((rc = scratch_reserve(need))) return rc;
  scratch_reset();
  F* tmp = (F*)scratch_take(3 * n * sizeof(F));     // pass-to-pass buffer of up to three batched transforms
  F* ea = (F*)scratch_take(3 * n * sizeof(F));      // coset evaluations of U, V, W (contiguous: one batched transform)
  int* flag = (int*)scratch_take(256);
  if (!tmp || !ea || !flag) return set_error(ZKB_ERR_CUDA, "scratch exhausted");
  F* eb = ea + n;
  F* ec = eb + n;
  if (check) {
    ZKB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), S()));
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    prof_begin(PROF_VEC);
    check_abc_kernel<F><<<blocks, 256, 0, S()>>>(n, (const F*)d_a, (const F*)d_b, (const F*)d_c, flag);
    prof_end(PROF_VEC);
    count_launch();
    if (check == 2) {   // deferred: the caller reads groth16_flag_host() once later work of this stream is known to be complete
      int* hf = groth16_flag_host();
      if (!hf) return set_error(ZKB_ERR_CUDA, "cannot allocate the pinned flag");
      ZKB_CUDA(ZKB_D2H(hf, flag, sizeof(int)));
    }
  }
  PowTable<F> fwd = d->fwd.view(), inv_t = d->inv.view();
  PowTable<F> gpre = d->gen_pre.view(), gpre_r = d->gen_pre_r.view(), gpost = d->gen_post.view();
  F* U = (F*)d_u; F* V = (F*)d_v; F* W = (F*)d_w; F* H = (F*)d_h;
  const bool contiguous = (const F*)d_b == (const F*)d_a + n && (const F*)d_c == (const F*)d_b + n && V == U + n && W == V + n;
  if (use_warp_ntt(log_n) && contiguous) {
    // the three interpolations as ONE batched transform per pass, then the three coset evaluations likewise: 9 launches instead
    // of 18, and three times the columns per launch to fill the waves (W goes through g^j / R, see vec_op_kernel)
    if ((rc = ntt_exec_warp<F>((const F*)d_a, n, U, log_n, 3, n, inv_t, nullptr, nullptr, &d->n_inv, tmp))) return rc;
    if (after_interp) after_interp(arg);
    const PowTable<F>* pres[3] = {&gpre, &gpre, &gpre_r};
    if ((rc = ntt_exec_warp<F>(U, n, ea, log_n, 3, n, fwd, pres, nullptr, nullptr, tmp))) return rc;
  } else {
    if ((rc = ntt_exec<F>((const F*)d_a, n, U, log_n, inv_t, nullptr, nullptr, &d->n_inv, tmp))) return rc;
    if ((rc = ntt_exec<F>((const F*)d_b, n, V, log_n, inv_t, nullptr, nullptr, &d->n_inv, tmp))) return rc;
    if ((rc = ntt_exec<F>((const F*)d_c, n, W, log_n, inv_t, nullptr, nullptr, &d->n_inv, tmp))) return rc;
    if (after_interp) after_interp(arg);
    if ((rc = ntt_exec<F>(U, n, ea, log_n, fwd, &gpre, nullptr, nullptr, tmp))) return rc;
    if ((rc = ntt_exec<F>(V, n, eb, log_n, fwd, &gpre, nullptr, nullptr, tmp))) return rc;
    if ((rc = ntt_exec<F>(W, n, ec, log_n, fwd, &gpre_r, nullptr, nullptr, tmp))) return rc;   // W(g w^i) / R
  }
  if ((rc = vec_op_t<F>(VEC_MULSUB_RAW, n, ea, n, eb, n, ec, ea))) return rc;                  // (U V - W)/R on the coset
  if ((rc = ntt_exec<F>(ea, n, H, log_n, inv_t, nullptr, &gpost, nullptr, tmp))) return rc;
  if (check == 1) {
    int hflag = 0;
    ZKB_CUDA(ZKB_D2H(&hflag, flag, sizeof(int)));
    ZKB_CUDA(cudaStreamSynchronize(S()));
    if (hflag) return set_error(ZKB_ERR_NOT_DIVISIBLE, "(U * V - W) did not divided by Z to zero");
  }
  return ZKB_OK;
}
// ---- the same pipeline in separable steps: on several GPUs the three interpolation -> coset-evaluation chains run on different
// ranks and the results are broadcast (zkb_groth16_spread_begin / _finish, api.cu; SURVEY.md section 8e "independent A/B/C NTTs
// spread across GPUs").  Same kernels and tables as groth16_h_t, so every value is the same field element.
template <class F>
static int groth16_check_t(uint32_t log_n, const void* d_a, const void* d_b, const void* d_c, int* d_flag) {
  const size_t n = (size_t)1 << log_n;
  ZKB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), S()));
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  prof_begin(PROF_VEC);
  check_abc_kernel<F><<<blocks, 256, 0, S()>>>(n, (const F*)d_a, (const F*)d_b, (const F*)d_c, d_flag);
  prof_end(PROF_VEC);
  count_launch();
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}
// chain `which` (0 U, 1 V, 2 W): coefficients = iNTT(evaluations on the domain), then their evaluations on the coset g<w>
// (W's come out divided by R, as the fused (U V - W) kernel expects)
template <class F>
static int groth16_chain_t(uint32_t log_n, int which, const void* d_in, void* d_coeff, void* d_eval, void* d_tmp) {
  typedef typename F::Params P;
  if (log_n > (uint32_t)P::TWO_ADICITY || log_n > 30) return set_error(ZKB_ERR_DOMAIN, "Domain size is too large");
  const size_t n = (size_t)1 << log_n;
  Domain<F>* d;
  int rc = get_domain<F>(log_n, true, &d);
  if (rc) return rc;
  PowTable<F> fwd = d->fwd.view(), inv_t = d->inv.view();
  PowTable<F> gpre = d->gen_pre.view(), gpre_r = d->gen_pre_r.view();
  if ((rc = ntt_exec<F>((const F*)d_in, n, (F*)d_coeff, log_n, inv_t, nullptr, nullptr, &d->n_inv, (F*)d_tmp))) return rc;
  return ntt_exec<F>((const F*)d_coeff, n, (F*)d_eval, log_n, fwd, which == 2 ? &gpre_r : &gpre, nullptr, nullptr, (F*)d_tmp);
}
// H from the three coset evaluation vectors (d_eu is overwritten with (U V - W) / R)
template <class F>
static int groth16_hfin_t(uint32_t log_n, void* d_eu, const void* d_ev, const void* d_ew, void* d_h, void* d_tmp) {
  const size_t n = (size_t)1 << log_n;
  Domain<F>* d;
  int rc = get_domain<F>(log_n, true, &d);
  if (rc) return rc;
  PowTable<F> inv_t = d->inv.view(), gpost = d->gen_post.view();
  if ((rc = vec_op_t<F>(VEC_MULSUB_RAW, n, d_eu, n, d_ev, n, d_ew, d_eu))) return rc;
  return ntt_exec<F>((const F*)d_eu, n, (F*)d_h, log_n, inv_t, nullptr, &gpost, nullptr, (F*)d_tmp);
}
int groth16_check_dev(int curve, uint32_t log_n, const void* d_a, const void* d_b, const void* d_c, int* d_flag) {
  if (curve == ZKB_BN254) return groth16_check_t<fr_bn>(log_n, d_a, d_b, d_c, d_flag);
  if (curve == ZKB_BLS12_381) return groth16_check_t<fr_bls>(log_n, d_a, d_b, d_c, d_flag);
  return set_error(ZKB_ERR_ARG, "unknown curve id");
}
int groth16_chain_dev(int curve, uint32_t log_n, int which, const void* d_in, void* d_coeff, void* d_eval, void* d_tmp) {
  if (curve == ZKB_BN254) return groth16_chain_t<fr_bn>(log_n, which, d_in, d_coeff, d_eval, d_tmp);
  if (curve == ZKB_BLS12_381) return groth16_chain_t<fr_bls>(log_n, which, d_in, d_coeff, d_eval, d_tmp);
  return set_error(ZKB_ERR_ARG, "unknown curve id");
}
int groth16_hfin_dev(int curve, uint32_t log_n, void* d_eu, const void* d_ev, const void* d_ew, void* d_h, void* d_tmp) {
  if (curve == ZKB_BN254) return groth16_hfin_t<fr_bn>(log_n, d_eu, d_ev, d_ew, d_h, d_tmp);
  if (curve == ZKB_BLS12_381) return groth16_hfin_t<fr_bls>(log_n, d_eu, d_ev, d_ew, d_h, d_tmp);
  return set_error(ZKB_ERR_ARG, "unknown curve id");
}

int groth16_h_dev(int curve, uint32_t log_n, const void* d_a, const void* d_b, const void* d_c, void* d_u, void* d_v,
                  void* d_w, void* d_h, int check, void (*after_interp)(void*), void* arg) {
  if (curve == ZKB_BN254) return groth16_h_t<fr_bn>(log_n, d_a, d_b, d_c, d_u, d_v, d_w, d_h, check, after_interp, arg);
  if (curve == ZKB_BLS12_381) return groth16_h_t<fr_bls>(log_n, d_a, d_b, d_c, d_u, d_v, d_w, d_h, check, after_interp, arg);
  return set_error(ZKB_ERR_ARG, "unknown curve id");
}

template <class F>
static int spmv_t(size_t n_out, size_t n_rows, const void* row_ptr, const void* col, const void* val, const void* w, void* out,
                  const uint32_t* long_rows, uint32_t n_long, void* long_partial) {
  if (n_out == 0) return ZKB_OK;
  prof_begin(PROF_SPMV);
  spmv_kernel<F><<<(unsigned)((n_out + 127) / 128), 128, 0, S()>>>(n_out, n_rows, (const unsigned long long*)row_ptr,
                                                                   (const uint32_t*)col, (const F*)val, (const F*)w, (F*)out);
  count_launch();
  if (n_long) {
    spmv_long_kernel<F><<<dim3(ZKB_SPMV_SPLIT, n_long), 256, 0, S()>>>(long_rows, (const unsigned long long*)row_ptr,
                                                                       (const uint32_t*)col, (const F*)val, (const F*)w,
                                                                       (F*)long_partial);
    spmv_long_finish_kernel<F><<<(n_long * 32 + 127) / 128, 128, 0, S()>>>(n_long, long_rows, (const F*)long_partial, (F*)out);
    count_launch(2);
  }
  prof_end(PROF_SPMV);
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}
int spmv_dev(int curve, size_t n_out, size_t n_rows, const void* row_ptr, const void* col, const void* val, const void* w,
             void* out, const uint32_t* long_rows, uint32_t n_long, void* long_partial) {
  if (n_long > 65535) return set_error(ZKB_ERR_ARG, "r1cs: too many long rows");
  if (curve == ZKB_BN254) return spmv_t<fr_bn>(n_out, n_rows, row_ptr, col, val, w, out, long_rows, n_long, long_partial);
  if (curve == ZKB_BLS12_381) return spmv_t<fr_bls>(n_out, n_rows, row_ptr, col, val, w, out, long_rows, n_long, long_partial);
  return set_error(ZKB_ERR_ARG, "unknown curve id");
}

}  // namespace zkb
