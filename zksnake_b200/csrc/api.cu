// api.cu -- extern "C" entry points of libzkb200.so declared in include/zkb200.h (host-buffer wrappers, Groth16
// prover object, self-test hooks).  Reference call sites are cited in the header.
#include <cuda_runtime.h>
#include <string.h>
#include <thread>
#include <vector>
#include "../../include/zkb200.h"
#include "ec.cuh"
#include "zkb_internal.h"

using namespace zkb;

static inline cudaStream_t S() { return (cudaStream_t)ctx_stream(); }
#define NEED_INIT() \
  if (!ctx_ready()) return set_error(ZKB_ERR_NOINIT, "zkb_init has not been called (no CUDA context; no CPU fallback)"); \
  ZKB_ENTRY_GUARD()
#define CHECK_CURVE(c) \
  if ((c) != ZKB_BN254 && (c) != ZKB_BLS12_381) return set_error(ZKB_ERR_ARG, "unknown curve id")
#define CHECK_GROUP(g) \
  if ((g) != 1 && (g) != 2) return set_error(ZKB_ERR_ARG, "group must be 1 (G1) or 2 (G2)")

// grow-only staging buffers for the host-pointer entry points (separate from the kernel scratch arena)
namespace {
struct Stage {
  char* p = nullptr;
  size_t cap = 0;
};
Stage g_stage[6];
int stage(int slot, size_t bytes, void** out) {
  Stage& s = g_stage[slot];
  if (bytes > s.cap) {
    ZKB_CUDA(cudaStreamSynchronize(S()));
    if (s.p) ZKB_CUDA(cudaFree(s.p));
    s.p = nullptr;
    s.cap = 0;
    size_t cap = bytes + (bytes >> 2) + 4096;
    ZKB_CUDA(cudaMalloc((void**)&s.p, cap));
    s.cap = cap;
  }
  *out = s.p;
  return ZKB_OK;
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------------- Fr vectors
int zkb_ntt_dev(int curve, int inverse, int coset, uint32_t log_n, const void* d_in, size_t in_len, void* d_out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  return ntt_dev(curve, inverse, coset, log_n, d_in, in_len, d_out);
}

int zkb_ntt(int curve, int inverse, int coset, uint32_t log_n, const uint64_t* in, size_t in_len, uint64_t* out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  if (log_n > 30) return set_error(ZKB_ERR_DOMAIN, "Domain size is too large");
  size_t n = (size_t)1 << log_n;
  if (in_len > n) in_len = n;
  void *d_in, *d_out;
  int rc;
  if ((rc = stage(0, (in_len ? in_len : 1) * 32, &d_in))) return rc;
  if ((rc = stage(1, n * 32, &d_out))) return rc;
  if (in_len) ZKB_CUDA(ZKB_H2D(d_in, in, in_len * 32));
  if ((rc = fr_reduce_dev(curve, in_len, d_in))) return rc;
  if ((rc = ntt_dev(curve, inverse, coset, log_n, d_in, in_len, d_out))) return rc;
  ZKB_CUDA(ZKB_D2H(out, d_out, n * 32));
  ZKB_CUDA(cudaStreamSynchronize(S()));
  return ZKB_OK;
}

int zkb_vec_op_dev(int curve, int op, size_t n, const void* d_a, size_t na, const void* d_b, size_t nb, void* d_out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  if (op < 0 || op > 2) return set_error(ZKB_ERR_ARG, "unknown vector op");
  return vec_op_dev(curve, op, n, d_a, na, d_b, nb, nullptr, d_out);
}

int zkb_vec_op(int curve, int op, size_t n, const uint64_t* a, size_t na, const uint64_t* b, size_t nb, uint64_t* out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  if (op < 0 || op > 2) return set_error(ZKB_ERR_ARG, "unknown vector op");
  if (n == 0) return ZKB_OK;
  if (na > n) na = n;
  if (nb > n) nb = n;
  void *d_a, *d_b, *d_o;
  int rc;
  if ((rc = stage(0, (na ? na : 1) * 32, &d_a))) return rc;
  if ((rc = stage(1, (nb ? nb : 1) * 32, &d_b))) return rc;
  if ((rc = stage(2, n * 32, &d_o))) return rc;
  if (na) ZKB_CUDA(ZKB_H2D(d_a, a, na * 32));
  if (nb) ZKB_CUDA(ZKB_H2D(d_b, b, nb * 32));
  if ((rc = fr_reduce_dev(curve, na, d_a))) return rc;
  if ((rc = fr_reduce_dev(curve, nb, d_b))) return rc;
  if ((rc = vec_op_dev(curve, op, n, d_a, na, d_b, nb, nullptr, d_o))) return rc;
  ZKB_CUDA(ZKB_D2H(out, d_o, n * 32));
  ZKB_CUDA(cudaStreamSynchronize(S()));
  return ZKB_OK;
}

int zkb_fr_reduce_dev(int curve, size_t n, void* d_inout) {
  NEED_INIT();
  CHECK_CURVE(curve);
  return fr_reduce_dev(curve, n, d_inout);
}
int zkb_fr_reduce(int curve, size_t n, uint64_t* inout) {
  NEED_INIT();
  CHECK_CURVE(curve);
  if (n == 0) return ZKB_OK;
  void* d;
  int rc;
  if ((rc = stage(0, n * 32, &d))) return rc;
  ZKB_CUDA(ZKB_H2D(d, inout, n * 32));
  if ((rc = fr_reduce_dev(curve, n, d))) return rc;
  ZKB_CUDA(ZKB_D2H(inout, d, n * 32));
  ZKB_CUDA(cudaStreamSynchronize(S()));
  return ZKB_OK;
}
int zkb_fr_powers_dev(int curve, const uint64_t base[4], const uint64_t scale[4], size_t n, void* d_out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  return fr_powers_dev(curve, base, scale, n, d_out);
}

// ---------------------------------------------------------------------------------------------------- points / MSM
size_t zkb_affine_bytes(int curve, int group) { return affine_bytes(curve, group); }

int zkb_points_upload(int curve, int group, const uint64_t* pts, size_t n, void* d_out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  CHECK_GROUP(group);
  if (n == 0) return ZKB_OK;
  ZKB_CUDA(ZKB_H2D(d_out, pts, n * affine_bytes(curve, group)));
  return points_to_mont_dev(curve, group, n, d_out);
}
int zkb_points_download(int curve, int group, const void* d_pts, size_t n, uint64_t* out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  CHECK_GROUP(group);
  if (n == 0) return ZKB_OK;
  size_t bytes = n * affine_bytes(curve, group);
  void* d_tmp;
  int rc;
  if ((rc = stage(3, bytes, &d_tmp))) return rc;
  ZKB_CUDA(cudaMemcpyAsync(d_tmp, d_pts, bytes, cudaMemcpyDeviceToDevice, S()));
  if ((rc = points_from_mont_dev(curve, group, n, d_tmp))) return rc;
  ZKB_CUDA(ZKB_D2H(out, d_tmp, bytes));
  ZKB_CUDA(cudaStreamSynchronize(S()));
  return ZKB_OK;
}

size_t zkb_compressed_bytes(int curve, int group) {
  if ((curve != ZKB_BN254 && curve != ZKB_BLS12_381) || (group != 1 && group != 2)) return 0;
  return compressed_bytes(curve, group);
}

int zkb_points_compress(int curve, int group, const void* d_pts, size_t n, uint8_t* out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  CHECK_GROUP(group);
  if (n == 0) return ZKB_OK;
  size_t bytes = n * compressed_bytes(curve, group);
  void* d_tmp;
  int rc;
  if ((rc = stage(3, bytes, &d_tmp))) return rc;
  if ((rc = points_compress_dev(curve, group, d_pts, n, d_tmp))) return rc;
  ZKB_CUDA(ZKB_D2H(out, d_tmp, bytes));
  ZKB_CUDA(cudaStreamSynchronize(S()));
  return ZKB_OK;
}

int zkb_points_decompress(int curve, int group, const uint8_t* in, size_t n, int validate, void* d_pts, long long* bad_index,
                          int* reason) {
  NEED_INIT();
  CHECK_CURVE(curve);
  CHECK_GROUP(group);
  if (bad_index) *bad_index = -1;
  if (reason) *reason = 0;
  if (n == 0) return ZKB_OK;
  size_t bytes = n * compressed_bytes(curve, group);
  void* d_tmp;
  int rc;
  if ((rc = stage(3, bytes, &d_tmp))) return rc;
  ZKB_CUDA(ZKB_H2D(d_tmp, in, bytes));
  unsigned long long bad = ~0ull;
  if ((rc = points_decompress_dev(curve, group, d_tmp, n, validate, d_pts, &bad))) return rc;
  if (bad != ~0ull) {
    static const char* why[] = {"", "invalid flags", "coordinate not in field", "non-zero infinity", "not on curve",
                                "not in the prime-order subgroup"};
    int r = (int)(bad & 0xff);
    if (bad_index) *bad_index = (long long)(bad >> 8);
    if (reason) *reason = r;
    return set_error(ZKB_ERR_POINT, std::string("Cannot deserialize point: ") + why[r < 6 ? r : 0] + " (index " +
                                        std::to_string(bad >> 8) + ")");
  }
  return ZKB_OK;
}

int zkb_msm_dev_windows(int curve, int group, const void* d_pts, const void* d_scalars, size_t n, uint32_t wrank,
                        uint32_t wworld, uint64_t* out_xy, int* out_inf) {
  NEED_INIT();
  CHECK_CURVE(curve);
  CHECK_GROUP(group);
  static MsmTicket tk;
  MsmJob job = {group, d_pts, d_scalars, n, 0, 0};
  int rc = msm_enqueue(curve, job, wrank, wworld, &tk);
  if (rc) return rc;
  return msm_finish(&tk, out_xy, out_inf);
}

// ---- fixed-base tables --------------------------------------------------------------------------------------------------
struct zkb_msm_table {
  int curve, group;
  size_t n;
  uint32_t c, W;
  void* d_table;   // W * n affine points: table[w * n + i] = 2^(c w) * P_i
};

int zkb_msm_table_create(int curve, int group, const void* d_pts, size_t n, uint32_t window_bits, uint32_t world,
                         zkb_msm_table** out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  CHECK_GROUP(group);
  uint32_t c = window_bits, W = 0;
  int rc;
  if ((rc = msm_table_build(curve, group, d_pts, n, world, &c, &W, nullptr))) return rc;
  zkb_msm_table* t = new zkb_msm_table{curve, group, n, c, W, nullptr};
  cudaError_t e = cudaMalloc(&t->d_table, (size_t)W * n * affine_bytes(curve, group));
  if (e != cudaSuccess) {
    delete t;
    return cuda_fail((int)e, "msm table allocation", __FILE__, __LINE__);
  }
  if ((rc = msm_table_build(curve, group, d_pts, n, world, &c, &W, t->d_table))) {
    cudaFree(t->d_table);
    delete t;
    return rc;
  }
  *out = t;
  return ZKB_OK;
}
void zkb_msm_table_free(zkb_msm_table* t) {
  if (!t) return;
  if (ctx_ready()) {
    cudaStreamSynchronize(S());
    cudaFree(t->d_table);
  }
  delete t;
}
int zkb_msm_table_info(const zkb_msm_table* t, uint32_t* window_bits, uint32_t* windows, size_t* bytes) {
  if (!t) return set_error(ZKB_ERR_ARG, "null table");
  if (window_bits) *window_bits = t->c;
  if (windows) *windows = t->W;
  if (bytes) *bytes = (size_t)t->W * t->n * affine_bytes(t->curve, t->group);
  return ZKB_OK;
}
int zkb_msm_table_batch_dev(zkb_msm_table* t, int count, const void* const* d_scalars, const size_t* n_scalars, uint32_t wrank,
                            uint32_t wworld, uint64_t* out_xy, int* out_inf) {
  NEED_INIT();
  if (!t) return set_error(ZKB_ERR_ARG, "null table");
  if (count < 1 || count > 8) return set_error(ZKB_ERR_ARG, "msm batch: 1..8 jobs");
  static MsmTicket tk[8];
  MsmJob jobs[8];
  for (int i = 0; i < count; i++) {
    if (n_scalars[i] > t->n) return set_error(ZKB_ERR_MISMATCH, "Number of points and scalars mismatch");
    jobs[i] = MsmJob{t->group, t->d_table, d_scalars[i], n_scalars[i], t->c, t->n};
  }
  int rc;
  if ((rc = msm_enqueue_batch(t->curve, jobs, count, wrank, wworld, tk))) return rc;
  const size_t limbs = affine_bytes(t->curve, t->group) / 8;
  for (int i = 0; i < count; i++)
    if ((rc = msm_finish(&tk[i], out_xy + i * limbs, &out_inf[i]))) return rc;
  return ZKB_OK;
}

int zkb_msm_table_dev(zkb_msm_table* t, const void* d_scalars, size_t n_scalars, uint32_t wrank, uint32_t wworld,
                      uint64_t* out_xy, int* out_inf) {
  NEED_INIT();
  if (!t) return set_error(ZKB_ERR_ARG, "null table");
  if (n_scalars > t->n) return set_error(ZKB_ERR_MISMATCH, "Number of points and scalars mismatch");
  static MsmTicket tk;
  MsmJob job = {t->group, t->d_table, d_scalars, n_scalars, t->c, t->n};
  int rc = msm_enqueue(t->curve, job, wrank, wworld, &tk);
  if (rc) return rc;
  return msm_finish(&tk, out_xy, out_inf);
}

int zkb_msm_dev(int curve, int group, const void* d_pts, const void* d_scalars, size_t n, uint64_t* out_xy, int* out_inf) {
  NEED_INIT();
  CHECK_CURVE(curve);
  CHECK_GROUP(group);
  return msm_dev(curve, group, d_pts, d_scalars, n, out_xy, out_inf);
}

int zkb_msm(int curve, int group, const uint64_t* pts, size_t n_points, const uint64_t* scalars, size_t n_scalars,
            uint64_t* out_xy, int* out_inf) {
  NEED_INIT();
  CHECK_CURVE(curve);
  CHECK_GROUP(group);
  if (n_points != n_scalars) return set_error(ZKB_ERR_MISMATCH, "Number of points and scalars mismatch");
  size_t n = n_points;
  if (n == 0) {
    memset(out_xy, 0, affine_bytes(curve, group));
    *out_inf = 1;
    return ZKB_OK;
  }
  void *d_p, *d_s;
  int rc;
  if ((rc = stage(3, n * affine_bytes(curve, group), &d_p))) return rc;
  if ((rc = stage(4, n * 32, &d_s))) return rc;
  ZKB_CUDA(ZKB_H2D(d_p, pts, n * affine_bytes(curve, group)));
  ZKB_CUDA(ZKB_H2D(d_s, scalars, n * 32));
  if ((rc = points_to_mont_dev(curve, group, n, d_p))) return rc;
  if ((rc = fr_reduce_dev(curve, n, d_s))) return rc;
  return msm_dev(curve, group, d_p, d_s, n, out_xy, out_inf);
}

int zkb_batch_mul_dev(int curve, int group, const void* d_bases, int single_base, const void* d_scalars, size_t n,
                      void* d_out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  CHECK_GROUP(group);
  return batch_mul_dev(curve, group, d_bases, single_base, d_scalars, n, d_out);
}
void zkb_msm_set_tuning(int window_bits, int segment, int reduce_chunk) { msm_set_tuning(window_bits, segment, reduce_chunk); }

// ---------------------------------------------------------------------------------------------------- Groth16
int zkb_groth16_h_dev(int curve, uint32_t log_n, const void* d_a, const void* d_b, const void* d_c, void* d_u, void* d_v,
                      void* d_w, void* d_h, int check) {
  NEED_INIT();
  CHECK_CURVE(curve);
  return groth16_h_dev(curve, log_n, d_a, d_b, d_c, d_u, d_v, d_w, d_h, check ? 1 : 0);   // (the deferred mode is internal)
}

int zkb_groth16_h(int curve, uint32_t log_n, const uint64_t* a, const uint64_t* b, const uint64_t* c, uint64_t* u,
                  uint64_t* v, uint64_t* w, uint64_t* h) {
  NEED_INIT();
  CHECK_CURVE(curve);
  if (log_n > 30) return set_error(ZKB_ERR_DOMAIN, "Domain size is too large");
  size_t n = (size_t)1 << log_n, bytes = n * 32;
  void* d;
  int rc;
  if ((rc = stage(5, 7 * bytes, &d))) return rc;
  char* p = (char*)d;
  const uint64_t* src[3] = {a, b, c};
  for (int i = 0; i < 3; i++) {
    ZKB_CUDA(ZKB_H2D(p + i * bytes, src[i], bytes));
    if ((rc = fr_reduce_dev(curve, n, p + i * bytes))) return rc;
  }
  if ((rc = groth16_h_dev(curve, log_n, p, p + bytes, p + 2 * bytes, p + 3 * bytes, p + 4 * bytes, p + 5 * bytes,
                          p + 6 * bytes, 1)))
    return rc;
  uint64_t* dst[4] = {u, v, w, h};
  for (int i = 0; i < 4; i++)
    if (dst[i]) ZKB_CUDA(ZKB_D2H(dst[i], p + (3 + i) * bytes, bytes));
  ZKB_CUDA(cudaStreamSynchronize(S()));
  return ZKB_OK;
}

struct zkb_groth16_pk {
  int curve;
  uint32_t log_n;
  size_t n, n_kdelta;
  size_t off, len, koff, klen;  // this rank's slice of the n-point vectors / of the n_kdelta-point vector
  uint32_t wrank, wworld;       // window shard of every MSM (0/1 = all windows)
  bool kw_custom;               // the [K w] MSM runs the windows [kw_first, kw_first + kw_count) instead (zkb_groth16_pk_set_kw_windows)
  uint32_t kw_first, kw_count;
  zkb_msm_table* tab[4];        // optional fixed-base tables of tau1, tau2, target1, kdelta1 (over this key's slices)
  const void *tau1, *tau2, *target1, *kdelta1;
  uint64_t alpha1[12], beta1[12], beta2[24], delta1[12], delta2[24];
  char* work;  // a, b, c, u, v, w, h (n each) + priv (n_kdelta)
  uint64_t msm_xy[5][24];
  int msm_inf[5];
  struct Groth16Pre* pre;       // host-side products of (r, s) with the key, computed while the GPU runs (zkb_groth16_precompute)
};

// r*delta_1, s*delta_1, -(r s)*delta_1 and s*delta_2 depend on the prover's randomness and the key only, not on the MSMs:
// four host scalar multiplications (~0.2 ms each in G1, ~0.8 ms in G2) that used to sit on the critical path AFTER the last
// MSM.  They are computed on two host threads as soon as r and s are known, i.e. under the ~25 ms of GPU work.
struct Groth16Pre {
  std::thread th1, th2;
  bool running = false, valid = false;
  uint64_t r[4], s[4];
  uint64_t rd1[12], sd1[12], nrsd1[12], sd2[24];
  int inf_rd1 = 1, inf_sd1 = 1, inf_nrsd1 = 1, inf_sd2 = 1;
  // single-GPU provers: A, s*A, B1, r*B1 computed by the host threads that finish the [U] and [V] MSMs, i.e. while the GPU
  // is still busy with the remaining MSMs (groth16_msms); assemble uses them when `early` is set
  bool early = false;
  uint64_t A[12], sA[12], B1[12], rB1[12];
  int infA = 1, inf_sA = 1, infB1 = 1, inf_rB1 = 1;
  // sharded provers (several GPUs): C is linear in the MSM results,
  //   C = [H Z] + [K w] + s [U] + r [V]_1 + P0,   P0 = s alpha + r beta_1 + (r s) delta_1,
  // so every rank multiplies ITS partial sums of [U] and [V]_1 by s and r on the host threads that finish those two MSMs -- under
  // its remaining GPU work -- and hands out  [H Z]_i + s [U]_i + r [V]_i  in the HZ slot (flag bit 1 of that slot says so: `folded`).
  // After the exchange only point additions are left; round 1 did both scalar multiplications after it (~0.25 ms exposed).
  std::thread th3;
  bool has_p0 = false, folded = false;
  uint64_t p0[12], sU[12], rV[12];
  int inf_p0 = 1, inf_sU = 1, inf_rV = 1;
};
#define ZKB_SLOT_FOLDED 2   // msm_inf[3] & 2: the HZ slot carries s [U]_i + r [V]_i as well
static void pre_join(Groth16Pre* p);

int zkb_groth16_pk_create_sharded(int curve, uint32_t log_n, const void* d_tau1, const void* d_tau2, const void* d_target1,
                                  size_t off, size_t len, const void* d_kdelta1, size_t n_kdelta, size_t koff, size_t klen,
                                  const uint64_t* alpha1, const uint64_t* beta1, const uint64_t* beta2, const uint64_t* delta1,
                                  const uint64_t* delta2, zkb_groth16_pk** out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  if (log_n > 28) return set_error(ZKB_ERR_DOMAIN, "Domain size is too large");
  if (off + len > ((size_t)1 << log_n) || koff + klen > n_kdelta) return set_error(ZKB_ERR_ARG, "proving-key slice out of range");
  zkb_groth16_pk* pk = new zkb_groth16_pk();
  memset(pk, 0, sizeof(*pk));
  pk->curve = curve;
  pk->wrank = 0;
  pk->wworld = 1;
  pk->log_n = log_n;
  pk->n = (size_t)1 << log_n;
  pk->n_kdelta = n_kdelta;
  pk->off = off;
  pk->len = len;
  pk->koff = koff;
  pk->klen = klen;
  pk->tau1 = d_tau1;
  pk->tau2 = d_tau2;
  pk->target1 = d_target1;
  pk->kdelta1 = d_kdelta1;
  size_t g1 = affine_bytes(curve, 1), g2 = affine_bytes(curve, 2);
  memcpy(pk->alpha1, alpha1, g1);
  memcpy(pk->beta1, beta1, g1);
  memcpy(pk->beta2, beta2, g2);
  memcpy(pk->delta1, delta1, g1);
  memcpy(pk->delta2, delta2, g2);
  size_t bytes = (7 * pk->n + n_kdelta + 8) * 32;
  cudaError_t e = cudaMalloc((void**)&pk->work, bytes);
  if (e != cudaSuccess) {
    delete pk;
    return cuda_fail((int)e, "cudaMalloc(pk work)", __FILE__, __LINE__);
  }
  *out = pk;
  return ZKB_OK;
}

int zkb_groth16_pk_create(int curve, uint32_t log_n, const void* d_tau1, const void* d_tau2, const void* d_target1,
                          const void* d_kdelta1, size_t n_kdelta, const uint64_t* alpha1, const uint64_t* beta1,
                          const uint64_t* beta2, const uint64_t* delta1, const uint64_t* delta2, zkb_groth16_pk** out) {
  if (log_n > 28) return set_error(ZKB_ERR_DOMAIN, "Domain size is too large");
  return zkb_groth16_pk_create_sharded(curve, log_n, d_tau1, d_tau2, d_target1, 0, (size_t)1 << log_n, d_kdelta1, n_kdelta, 0,
                                       n_kdelta, alpha1, beta1, beta2, delta1, delta2, out);
}

void zkb_groth16_pk_free(zkb_groth16_pk* pk) {
  if (!pk) return;
  for (int i = 0; i < 4; i++) zkb_msm_table_free(pk->tab[i]);
  if (ctx_ready()) {
    cudaStreamSynchronize(S());
    cudaFree(pk->work);
  }
  if (pk->pre) {
    pre_join(pk->pre);
    delete pk->pre;
  }
  delete pk;
}

int zkb_groth16_pk_build_tables(zkb_groth16_pk* pk, uint32_t world) {
  NEED_INIT();
  if (!pk) return set_error(ZKB_ERR_ARG, "null proving key");
  const void* vec[4] = {pk->tau1, pk->tau2, pk->target1, pk->kdelta1};
  const size_t cnt[4] = {pk->len, pk->len, pk->len, pk->klen};
  const int grp[4] = {1, 2, 1, 1};
  for (int i = 0; i < 4; i++) {
    if (pk->tab[i] || cnt[i] == 0) continue;
    // ZKB_TABLE_C: window size of every table (experiments).  The cost model's choice was checked on the B200 at 2^20 BN254
    // (profiles/R2z_n2_c*.json, R2z_n1_c20.json, R3a_n1_kw*.json): 1 GPU c = 19 / 14 windows 23.3 ms, c = 20 23.9; 2 GPUs c = 16 /
    // 2 x 8 windows 14.5 ms, c = 18 15.0, c = 19 (2 x 7 windows over 2^18 buckets) 14.8, c = 20 17.0; a smaller window for the
    // LAST MSM of the batch only (shorter exposed reduction, more additions) 23.23-23.68 against 23.26 ms: nothing to gain.
    static const uint32_t force_c = [] { const char* e = getenv("ZKB_TABLE_C"); return e ? (uint32_t)atoi(e) : 0u; }();
    int rc = zkb_msm_table_create(pk->curve, grp[i], vec[i], cnt[i], force_c, world ? world : 1, &pk->tab[i]);
    if (rc) return rc;
  }
  ZKB_CUDA(cudaStreamSynchronize(S()));
  return ZKB_OK;
}

int zkb_groth16_pk_msm_info(const zkb_groth16_pk* pk, int which, uint32_t* window_bits, uint32_t* windows) {
  if (!pk || which < 0 || which > 3) return set_error(ZKB_ERR_ARG, "bad proving key vector");
  if (pk->tab[which]) return zkb_msm_table_info(pk->tab[which], window_bits, windows, nullptr);
  const size_t cnt[4] = {pk->len, pk->len, pk->len, pk->klen};
  uint32_t c = 0, W = 0;
  msm_plan_info(cnt[which] ? cnt[which] : 1, pk->curve == ZKB_BN254 ? 254 : 255, pk->wworld ? pk->wworld : 1, &c, &W);
  if (window_bits) *window_bits = c;
  if (windows) *windows = W;
  return ZKB_OK;
}

int zkb_msm_kernel_info(int curve, int group, int* lanes_per_point, int* ctas_per_sm) {
  CHECK_CURVE(curve);
  CHECK_GROUP(group);
  int lanes = 1, ctas = 0;
  msm_kernel_info(curve, group, &lanes, &ctas);
  if (lanes_per_point) *lanes_per_point = lanes;
  if (ctas_per_sm) *ctas_per_sm = ctas;
  return ZKB_OK;
}

int zkb_groth16_pk_set_kw_windows(zkb_groth16_pk* pk, uint32_t first, uint32_t count, int enable) {
  if (!pk || count > 0xffff) return set_error(ZKB_ERR_ARG, "bad [K w] window range");
  pk->kw_custom = enable != 0;
  pk->kw_first = first;
  pk->kw_count = count;
  return ZKB_OK;
}

int zkb_groth16_pk_set_window_shard(zkb_groth16_pk* pk, uint32_t rank, uint32_t world) {
  if (!pk || world == 0 || rank >= world) return set_error(ZKB_ERR_ARG, "bad window shard");
  pk->wrank = rank;
  pk->wworld = world;
  return ZKB_OK;
}

static bool is_zero_pt(const uint64_t* p, size_t bytes) {
  for (size_t i = 0; i < bytes / 8; i++)
    if (p[i]) return false;
  return true;
}

// The five MSMs of protocol.py:133-155 as one batch: sort + accumulate back to back on the library stream, every reduction
// (latency-bound) on a side stream as soon as its accumulation is done.  The G2 MSM goes first: its reduction chain is the
// longest (an Fp2 addition is ~40 dependent Fq products) and so hides behind the four G1 accumulations.
// Scalars: the full-length U, V, H in pk->work and d_priv.  The MSMs are named by their index in msm_xy (A, B1, B2, HZ, KW); the
// batch ORDER depends on the prover:
//  * one GPU: B2, A, B1, HZ, KW -- the G2 MSM first (its reduction chain, ~40 dependent Fq products per Fp2 addition, hides behind
//    the four G1 accumulations) and [K w] last: nothing runs under the LAST reduction, and the witness MSM has the cheapest one
//    (measured with H last instead: +0.4 ms of exposed reduction at 2^20);
//  * a window shard of a multi-GPU proof: KW, B2, A, B1, HZ -- the order in which the scalars come into existence there: the private
//    witness is there from the start (a rank that runs none of the transform chains accumulates [K w] while it waits for U and V), H
//    arrives last, from the rank that formed it, while the other four MSMs run (zkb_groth16_spread_*).
static MsmTicket g16_tk[5];
enum { MSM_A = 0, MSM_B1 = 1, MSM_B2 = 2, MSM_HZ = 3, MSM_KW = 4 };
static const int* groth16_order(const zkb_groth16_pk* pk) {   // batch position -> MSM
  static const int single[5] = {MSM_B2, MSM_A, MSM_B1, MSM_HZ, MSM_KW}, shard[5] = {MSM_KW, MSM_B2, MSM_A, MSM_B1, MSM_HZ};
  return pk->wworld > 1 ? shard : single;
}
static int groth16_pos(const zkb_groth16_pk* pk, int msm) {   // MSM -> batch position
  const int* order = groth16_order(pk);
  for (int i = 0; i < 5; i++)
    if (order[i] == msm) return i;
  return 0;
}
static void groth16_jobs(const zkb_groth16_pk* pk, const void* d_priv, MsmJob job[5]) {
  const size_t bytes = pk->n * 32;
  const char* w = pk->work;
  const char *d_u = w + 3 * bytes + pk->off * 32, *d_v = w + 4 * bytes + pk->off * 32, *d_h = w + 6 * bytes + pk->off * 32;
  auto mk = [&](int group, int which, const void* pts, const void* sc, size_t n) {
    const zkb_msm_table* t = pk->tab[which];
    if (t) return MsmJob{group, t->d_table, sc, n, t->c, t->n};
    return MsmJob{group, pts, sc, n, 0, 0};
  };
  MsmJob& kw = job[groth16_pos(pk, MSM_KW)];
  kw = mk(1, 3, pk->kdelta1, (const char*)d_priv + pk->koff * 32, pk->klen);
  if (pk->kw_custom) {
    kw.own_wrank = pk->kw_first;
    kw.own_wworld = ZKB_WINDOW_RANGE | pk->kw_count;
  }
  job[groth16_pos(pk, MSM_B2)] = mk(2, 1, pk->tau2, d_v, pk->len);
  job[groth16_pos(pk, MSM_A)] = mk(1, 0, pk->tau1, d_u, pk->len);
  job[groth16_pos(pk, MSM_B1)] = mk(1, 0, pk->tau1, d_v, pk->len);
  job[groth16_pos(pk, MSM_HZ)] = mk(1, 2, pk->target1, d_h, pk->len);
}

// jobs [first, first + count) of the batch on the library stream; `join`: order the stream behind their reductions
static int groth16_enqueue(zkb_groth16_pk* pk, const void* d_priv, int first, int count, bool join) {
  MsmJob job[5];
  groth16_jobs(pk, d_priv, job);
  return msm_enqueue_batch(pk->curve, job + first, count, pk->wrank, pk->wworld, g16_tk + first, join, first);
}
// The host half of the five MSMs, after every job has been enqueued.
static int groth16_collect(zkb_groth16_pk* pk) {
  const int curve = pk->curve;
  MsmTicket* tk = g16_tk;
  const int* slot = groth16_order(pk);   // batch position -> index in msm_xy (A, B1, B2, HZ, KW)
  // the five host recombinations (~0.1-0.2 ms of 64-bit Montgomery arithmetic each) run on five host threads; each waits
  // for its own ticket's event, so they also overlap the reductions still running on the GPU
  int rcs[5] = {0, 0, 0, 0, 0};
  // whole-key single-GPU proofs: the MSM results are final, so the threads that finish [U] (-> A) and [V] in G1 (-> B1) go on
  // to s*A and r*B1, the two scalar multiplications of the assembly that depend on MSM results (~0.2 ms each), while the
  // GPU still runs the later MSMs.  Sharded proofs: the same threads multiply this rank's PARTIAL sums (Groth16Pre::folded).
  Groth16Pre* pre = pk->pre;
  const bool whole = pk->wworld == 1 && pk->len == pk->n && pk->klen == pk->n_kdelta;
  const bool early = pre && pre->valid && whole;
  const bool fold = pre && pre->valid && !whole;
  if (pre) pre->early = pre->folded = false;
  if (early) pre_join(pre);   // r*delta_1, s*delta_1 (started when r, s arrived; long done by now)
  auto after = [&](int i) {
    if (rcs[i]) return;
    if (fold) {
      if (slot[i] == 0 || slot[i] == 1) {
        const uint64_t* p1[1] = {pk->msm_xy[slot[i]]};
        int i1[1] = {pk->msm_inf[slot[i]]};
        const uint64_t* s1[1] = {slot[i] == 0 ? pre->s : pre->r};
        host_lincomb(curve, 1, 1, p1, i1, s1, slot[i] == 0 ? pre->sU : pre->rV, slot[i] == 0 ? &pre->inf_sU : &pre->inf_rV);
      }
      return;
    }
    if (!early) return;
    const size_t g1 = affine_bytes(curve, 1);
    if (slot[i] == 0) {
      const uint64_t* pts[3] = {pk->msm_xy[0], pk->alpha1, pre->rd1};
      int infs[3] = {pk->msm_inf[0], is_zero_pt(pk->alpha1, g1), pre->inf_rd1};
      const uint64_t* sc[3] = {nullptr, nullptr, nullptr};
      host_lincomb(curve, 1, 3, pts, infs, sc, pre->A, &pre->infA);
      const uint64_t* p1[1] = {pre->A};
      int i1[1] = {pre->infA};
      const uint64_t* s1[1] = {pre->s};
      host_lincomb(curve, 1, 1, p1, i1, s1, pre->sA, &pre->inf_sA);
    } else if (slot[i] == 1) {
      const uint64_t* pts[3] = {pk->msm_xy[1], pk->beta1, pre->sd1};
      int infs[3] = {pk->msm_inf[1], is_zero_pt(pk->beta1, g1), pre->inf_sd1};
      const uint64_t* sc[3] = {nullptr, nullptr, nullptr};
      host_lincomb(curve, 1, 3, pts, infs, sc, pre->B1, &pre->infB1);
      const uint64_t* p1[1] = {pre->B1};
      int i1[1] = {pre->infB1};
      const uint64_t* s1[1] = {pre->r};
      host_lincomb(curve, 1, 1, p1, i1, s1, pre->rB1, &pre->inf_rB1);
    }
  };
  std::thread th[4];
  for (int i = 1; i < 5; i++)
    th[i - 1] = std::thread([&, i]() {
      rcs[i] = msm_finish(&tk[i], pk->msm_xy[slot[i]], &pk->msm_inf[slot[i]]);
      after(i);
    });
  rcs[0] = msm_finish(&tk[0], pk->msm_xy[slot[0]], &pk->msm_inf[slot[0]]);
  after(0);
  for (int i = 0; i < 4; i++) th[i].join();
  for (int i = 0; i < 5; i++)
    if (rcs[i]) return set_error(rcs[i], "msm_finish failed in the Groth16 MSM batch");
  if (early) pre->early = true;
  if (fold) pre->folded = true;
  return ZKB_OK;
}
static int groth16_msms(zkb_groth16_pk* pk, const void* d_priv) {
  int rc = groth16_enqueue(pk, d_priv, 0, 5, true);
  if (rc) return rc;
  return groth16_collect(pk);
}
// the deferred satisfiability answer (groth16_h_dev with check == 2 / the spread path's own copy); call after groth16_collect
static int groth16_flag_result() {
  const int* hf = groth16_flag_host();
  if (hf && *hf) return set_error(ZKB_ERR_NOT_DIVISIBLE, "(U * V - W) did not divided by Z to zero");
  return ZKB_OK;
}
// what a rank hands to the exchange: its five partial sums, the HZ slot carrying s [U]_i + r [V]_i when the fold ran
static void groth16_export_partials(zkb_groth16_pk* pk, uint64_t* msm_xy, int* msm_inf) {
  memcpy(msm_xy, pk->msm_xy, sizeof(pk->msm_xy));
  memcpy(msm_inf, pk->msm_inf, sizeof(pk->msm_inf));
  Groth16Pre* pre = pk->pre;
  if (!pre || !pre->folded) return;
  const uint64_t* pts[3] = {pk->msm_xy[3], pre->sU, pre->rV};
  int infs[3] = {pk->msm_inf[3], pre->inf_sU, pre->inf_rV};
  const uint64_t* sc[3] = {nullptr, nullptr, nullptr};
  uint64_t out[24];
  int inf = 1;
  memset(out, 0, sizeof(out));
  host_lincomb(pk->curve, 1, 3, pts, infs, sc, out, &inf);
  memcpy(msm_xy + 3 * 24, out, sizeof(out));
  msm_inf[3] = (inf ? 1 : 0) | ZKB_SLOT_FOLDED;
  pre->folded = false;
}

static void pre_join(Groth16Pre* p) {
  if (p && p->running) {
    p->th1.join();
    p->th2.join();
    if (p->th3.joinable()) p->th3.join();
    p->running = false;
  }
}

static void pre_start(zkb_groth16_pk* pk, const uint64_t r[4], const uint64_t s[4]) {
  if (!pk->pre) pk->pre = new Groth16Pre();
  Groth16Pre* p = pk->pre;
  pre_join(p);
  p->early = p->folded = p->has_p0 = false;   // whatever was derived from an earlier (r, s) is void
  memcpy(p->r, r, 32);
  memcpy(p->s, s, 32);
  const int curve = pk->curve;
  const size_t g1 = affine_bytes(curve, 1), g2 = affine_bytes(curve, 2);
  const int inf_d1 = is_zero_pt(pk->delta1, g1), inf_d2 = is_zero_pt(pk->delta2, g2);
  p->th1 = std::thread([pk, p, curve, inf_d1]() {
    const uint64_t* pt[1] = {pk->delta1};
    int infs[1] = {inf_d1};
    const uint64_t* sc[1] = {p->r};
    host_lincomb(curve, 1, 1, pt, infs, sc, p->rd1, &p->inf_rd1);
    sc[0] = p->s;
    host_lincomb(curve, 1, 1, pt, infs, sc, p->sd1, &p->inf_sd1);
    // -(r s) as (order - r s)
    uint64_t rs[4], neg_rs[4];
    host_fr_mul(curve, p->r, p->s, rs);
    static const uint64_t ORD[2][4] = {
        {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
        {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull}};
    unsigned __int128 br = 0;
    bool rs_zero = !(rs[0] | rs[1] | rs[2] | rs[3]);
    for (int i = 0; i < 4; i++) {
      unsigned __int128 d = (unsigned __int128)ORD[curve][i] - rs[i] - (uint64_t)br;
      neg_rs[i] = (uint64_t)d;
      br = (d >> 64) & 1;
    }
    if (rs_zero) memset(neg_rs, 0, sizeof(neg_rs));
    sc[0] = neg_rs;
    host_lincomb(curve, 1, 1, pt, infs, sc, p->nrsd1, &p->inf_nrsd1);
  });
  p->th2 = std::thread([pk, p, curve, inf_d2]() {
    const uint64_t* pt[1] = {pk->delta2};
    int infs[1] = {inf_d2};
    const uint64_t* sc[1] = {p->s};
    host_lincomb(curve, 2, 1, pt, infs, sc, p->sd2, &p->inf_sd2);
  });
  if (pk->wworld != 1 || pk->len != pk->n || pk->klen != pk->n_kdelta) {
    // a shard of a multi-GPU proof: P0 = s alpha + r beta_1 + (r s) delta_1 for the folded assembly (Groth16Pre)
    p->th3 = std::thread([pk, p, curve, g1, inf_d1]() {
      uint64_t rs[4];
      host_fr_mul(curve, p->r, p->s, rs);
      const uint64_t* pt[3] = {pk->alpha1, pk->beta1, pk->delta1};
      int infs[3] = {is_zero_pt(pk->alpha1, g1), is_zero_pt(pk->beta1, g1), inf_d1};
      const uint64_t* sc[3] = {p->s, p->r, rs};
      host_lincomb(curve, 1, 3, pt, infs, sc, p->p0, &p->inf_p0);
    });
    p->has_p0 = true;
  }
  p->running = true;
  p->valid = true;
}

int zkb_groth16_precompute(zkb_groth16_pk* pk, const uint64_t r[4], const uint64_t s[4]) {
  if (!pk || !r || !s) return set_error(ZKB_ERR_ARG, "null argument");
  pre_start(pk, r, s);
  return ZKB_OK;
}

// folded: slot 3 is sum_i ([H Z]_i + s [U]_i + r [V]_i) (Groth16Pre): A, B and C are point additions only
static int assemble_folded(zkb_groth16_pk* pk, const uint64_t* msm_xy, const int* msm_inf, const uint64_t r[4], const uint64_t s[4],
                           uint64_t* out_a, uint64_t* out_b, uint64_t* out_c, int out_inf[3]) {
  const int curve = pk->curve;
  if (!pk->pre || !pk->pre->valid || !pk->pre->has_p0 || memcmp(pk->pre->r, r, 32) || memcmp(pk->pre->s, s, 32))
    return set_error(ZKB_ERR_ARG, "folded partial sums need zkb_groth16_precompute with the same (r, s) on a sharded key");
  Groth16Pre* pre = pk->pre;
  pre_join(pre);
  pre->valid = false;   // r and s are one-time values
  const size_t g1 = affine_bytes(curve, 1), g2 = affine_bytes(curve, 2);
  const uint64_t* mx[5];
  for (int i = 0; i < 5; i++) mx[i] = msm_xy + i * 24;
  const uint64_t* none[3] = {nullptr, nullptr, nullptr};
  {
    const uint64_t* pts[3] = {mx[0], pk->alpha1, pre->rd1};
    int infs[3] = {msm_inf[0] & 1, is_zero_pt(pk->alpha1, g1), pre->inf_rd1};
    host_lincomb(curve, 1, 3, pts, infs, none, out_a, &out_inf[0]);
  }
  {
    const uint64_t* pts[3] = {mx[2], pk->beta2, pre->sd2};
    int infs[3] = {msm_inf[2] & 1, is_zero_pt(pk->beta2, g2), pre->inf_sd2};
    host_lincomb(curve, 2, 3, pts, infs, none, out_b, &out_inf[1]);
  }
  {
    const uint64_t* pts[3] = {mx[3], mx[4], pre->p0};
    int infs[3] = {msm_inf[3] & 1, msm_inf[4] & 1, pre->inf_p0};
    host_lincomb(curve, 1, 3, pts, infs, none, out_c, &out_inf[2]);
  }
  return ZKB_OK;
}

int zkb_groth16_assemble(zkb_groth16_pk* pk, const uint64_t* msm_xy, const int* msm_inf, const uint64_t r[4], const uint64_t s[4],
                         uint64_t* out_a, uint64_t* out_b, uint64_t* out_c, int out_inf[3]) {
  if (!pk) return set_error(ZKB_ERR_ARG, "null proving key");
  if (msm_inf[3] & ZKB_SLOT_FOLDED) return assemble_folded(pk, msm_xy, msm_inf, r, s, out_a, out_b, out_c, out_inf);
  const int curve = pk->curve;
  // proof assembly, protocol.py:133-165:
  //   A = [U] + alpha + r delta,  B = [V] + beta + s delta  (in G1 and in G2),
  //   C = [H Z] + [K w] + s A + r B1 - (r s) delta
  // the multiples of delta come from zkb_groth16_precompute (started now if the caller did not); s A and r B1 are the only
  // scalar multiplications left after the MSMs and run side by side.
  if (!pk->pre || !pk->pre->valid || memcmp(pk->pre->r, r, 32) || memcmp(pk->pre->s, s, 32)) pre_start(pk, r, s);
  Groth16Pre* pre = pk->pre;
  pre_join(pre);
  pre->valid = false;   // r and s are one-time values
  const size_t g1 = affine_bytes(curve, 1), g2 = affine_bytes(curve, 2);
  const uint64_t* mx[5];
  for (int i = 0; i < 5; i++) mx[i] = msm_xy + i * 24;
  int inf_alpha = is_zero_pt(pk->alpha1, g1), inf_beta1 = is_zero_pt(pk->beta1, g1), inf_beta2 = is_zero_pt(pk->beta2, g2);
  uint64_t A[12], B1[12], sA[12], rB1[12];
  int infA, infB1, infB2, infC, inf_sA, inf_rB1;
  const bool early = pre->early && msm_xy == &pk->msm_xy[0][0];
  pre->early = false;
  std::thread th_sa, th_rb;
  if (early) {
    memcpy(A, pre->A, sizeof(A));
    memcpy(sA, pre->sA, sizeof(sA));
    memcpy(B1, pre->B1, sizeof(B1));
    memcpy(rB1, pre->rB1, sizeof(rB1));
    infA = pre->infA;
    inf_sA = pre->inf_sA;
    infB1 = pre->infB1;
    inf_rB1 = pre->inf_rB1;
  } else {
    {
      const uint64_t* pts[3] = {mx[0], pk->alpha1, pre->rd1};
      int infs[3] = {msm_inf[0], inf_alpha, pre->inf_rd1};
      const uint64_t* sc[3] = {nullptr, nullptr, nullptr};
      host_lincomb(curve, 1, 3, pts, infs, sc, A, &infA);
    }
    th_sa = std::thread([&]() {
      const uint64_t* pts[1] = {A};
      int infs[1] = {infA};
      const uint64_t* sc[1] = {s};
      host_lincomb(curve, 1, 1, pts, infs, sc, sA, &inf_sA);
    });
    {
      const uint64_t* pts[3] = {mx[1], pk->beta1, pre->sd1};
      int infs[3] = {msm_inf[1], inf_beta1, pre->inf_sd1};
      const uint64_t* sc[3] = {nullptr, nullptr, nullptr};
      host_lincomb(curve, 1, 3, pts, infs, sc, B1, &infB1);
    }
    th_rb = std::thread([&]() {
      const uint64_t* pts[1] = {B1};
      int infs[1] = {infB1};
      const uint64_t* sc[1] = {r};
      host_lincomb(curve, 1, 1, pts, infs, sc, rB1, &inf_rB1);
    });
  }
  {
    const uint64_t* pts[3] = {mx[2], pk->beta2, pre->sd2};
    int infs[3] = {msm_inf[2], inf_beta2, pre->inf_sd2};
    const uint64_t* sc[3] = {nullptr, nullptr, nullptr};
    host_lincomb(curve, 2, 3, pts, infs, sc, out_b, &infB2);
  }
  if (th_sa.joinable()) th_sa.join();
  if (th_rb.joinable()) th_rb.join();
  {
    const uint64_t* pts[5] = {mx[3], mx[4], sA, rB1, pre->nrsd1};
    int infs[5] = {msm_inf[3], msm_inf[4], inf_sA, inf_rB1, pre->inf_nrsd1};
    const uint64_t* sc[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    host_lincomb(curve, 1, 5, pts, infs, sc, out_c, &infC);
  }
  memcpy(out_a, A, g1);
  out_inf[0] = infA;
  out_inf[1] = infB2;
  out_inf[2] = infC;
  return ZKB_OK;
}

int zkb_groth16_assemble_partials(zkb_groth16_pk* pk, int world, const uint64_t* all_xy, const int* all_inf, const uint64_t r[4],
                                  const uint64_t s[4], uint64_t* out_a, uint64_t* out_b, uint64_t* out_c, int out_inf[3]) {
  if (!pk || world < 1 || !all_xy || !all_inf) return set_error(ZKB_ERR_ARG, "bad argument");
  // slot sums over the ranks (exact host group law), the five slots side by side on host threads, then the assembly
  uint64_t sum_xy[5][24];
  int sum_inf[5];
  memset(sum_xy, 0, sizeof(sum_xy));
  const int curve = pk->curve;
  int n_folded = 0;
  for (int k = 0; k < world; k++) n_folded += (all_inf[k * 5 + 3] & ZKB_SLOT_FOLDED) ? 1 : 0;
  if (n_folded != 0 && n_folded != world)
    return set_error(ZKB_ERR_ARG, "partial sums of some ranks carry the folded s [U] + r [V] term and others do not");
  Groth16Pre* pre = pk->pre;
  if (n_folded && pre && pre->valid && pre->has_p0 && !memcmp(pre->r, r, 32) && !memcmp(pre->s, s, 32)) {
    // folded partial sums: A, B and C are three independent sums of points -- one lincomb (one affine conversion) each, side by
    // side:  A = sum_i [U]_i + alpha + r delta_1,  B = sum_i [V]_i + beta_2 + s delta_2,  C = sum_i (HZ'_i + KW_i) + P0
    pre_join(pre);
    pre->valid = false;   // r and s are one-time values
    const int curve = pk->curve;
    const size_t g1 = affine_bytes(curve, 1), g2 = affine_bytes(curve, 2);
    auto sum = [&](int grp, std::initializer_list<int> slots, std::initializer_list<const uint64_t*> extra,
                   std::initializer_list<int> extra_inf, uint64_t* out, int* out_i) {
      std::vector<const uint64_t*> pts;
      std::vector<int> infs;
      for (int slot : slots)
        for (int k = 0; k < world; k++) {
          pts.push_back(all_xy + ((size_t)k * 5 + slot) * 24);
          infs.push_back(all_inf[k * 5 + slot] & 1);
        }
      pts.insert(pts.end(), extra.begin(), extra.end());
      infs.insert(infs.end(), extra_inf.begin(), extra_inf.end());
      std::vector<const uint64_t*> sc(pts.size(), nullptr);
      host_lincomb(curve, grp, (int)pts.size(), pts.data(), infs.data(), sc.data(), out, out_i);
    };
    std::thread tb([&]() { sum(2, {2}, {pk->beta2, pre->sd2}, {is_zero_pt(pk->beta2, g2), pre->inf_sd2}, out_b, &out_inf[1]); });
    std::thread tc([&]() { sum(1, {3, 4}, {pre->p0}, {pre->inf_p0}, out_c, &out_inf[2]); });
    sum(1, {0}, {pk->alpha1, pre->rd1}, {is_zero_pt(pk->alpha1, g1), pre->inf_rd1}, out_a, &out_inf[0]);
    tb.join();
    tc.join();
    return ZKB_OK;
  }
  auto add_slot = [&](int slot) {
    const int grp = slot == 2 ? 2 : 1;
    std::vector<const uint64_t*> pts(world), sc(world, nullptr);
    std::vector<int> infs(world);
    for (int k = 0; k < world; k++) {
      pts[k] = all_xy + ((size_t)k * 5 + slot) * 24;
      infs[k] = all_inf[k * 5 + slot] & 1;
    }
    host_lincomb(curve, grp, world, pts.data(), infs.data(), sc.data(), sum_xy[slot], &sum_inf[slot]);
  };
  std::thread th[4];
  for (int slot = 1; slot < 5; slot++) th[slot - 1] = std::thread(add_slot, slot);
  add_slot(0);
  for (int i = 0; i < 4; i++) th[i].join();
  if (n_folded) sum_inf[3] |= ZKB_SLOT_FOLDED;
  return zkb_groth16_assemble(pk, &sum_xy[0][0], sum_inf, r, s, out_a, out_b, out_c, out_inf);
}

int zkb_groth16_prove_dev(zkb_groth16_pk* pk, const void* d_a, const void* d_b, const void* d_c, const void* d_priv,
                          const uint64_t r[4], const uint64_t s[4], uint64_t* out_a, uint64_t* out_b, uint64_t* out_c,
                          int out_inf[3]) {
  NEED_INIT();
  if (!pk) return set_error(ZKB_ERR_ARG, "null proving key");
  if (pk->len != pk->n || pk->klen != pk->n_kdelta || pk->wworld != 1)
    return set_error(ZKB_ERR_ARG, "this proving key holds one slice only: use zkb_groth16_partial + zkb_groth16_assemble");
  pre_start(pk, r, s);   // host multiples of delta under the GPU work
  const size_t bytes = pk->n * 32;
  char* w = pk->work;
  void *d_u = w + 3 * bytes, *d_v = w + 4 * bytes, *d_w = w + 5 * bytes, *d_h = w + 6 * bytes;
  int rc;
  if ((rc = groth16_h_dev(pk->curve, pk->log_n, d_a, d_b, d_c, d_u, d_v, d_w, d_h, 2))) return rc;
  if ((rc = groth16_msms(pk, d_priv))) return rc;
  if ((rc = groth16_flag_result())) return rc;
  return zkb_groth16_assemble(pk, &pk->msm_xy[0][0], pk->msm_inf, r, s, out_a, out_b, out_c, out_inf);
}

int zkb_groth16_prove(zkb_groth16_pk* pk, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* priv,
                      const uint64_t r[4], const uint64_t s[4], uint64_t* out_a, uint64_t* out_b, uint64_t* out_c,
                      int out_inf[3]) {
  NEED_INIT();
  if (!pk) return set_error(ZKB_ERR_ARG, "null proving key");
  const size_t bytes = pk->n * 32;
  char* w = pk->work;
  char* d_priv = w + 7 * bytes;
  ZKB_CUDA(ZKB_H2D(w, a, bytes));
  ZKB_CUDA(ZKB_H2D(w + bytes, b, bytes));
  ZKB_CUDA(ZKB_H2D(w + 2 * bytes, c, bytes));
  if (pk->n_kdelta) ZKB_CUDA(ZKB_H2D(d_priv, priv, pk->n_kdelta * 32));
  // host values may be any 256-bit integers (header contract: reduced mod r like Fr::from(BigUint)); the signed-digit walk of
  // the MSM assumes canonical scalars
  int rc;
  for (int i = 0; i < 3; i++)
    if ((rc = fr_reduce_dev(pk->curve, pk->n, w + i * bytes))) return rc;
  if ((rc = fr_reduce_dev(pk->curve, pk->n_kdelta, d_priv))) return rc;
  return zkb_groth16_prove_dev(pk, w, w + bytes, w + 2 * bytes, d_priv, r, s, out_a, out_b, out_c, out_inf);
}

struct zkb_r1cs {
  int curve;
  size_t n_rows, n_cols;
  unsigned long long* row_ptr[3];
  uint32_t* col[3];
  void* val[3];
  void* w;  // witness staging (n_cols)
  uint32_t* long_rows[3];   // rows with more than SPMV_LONG_ROW non-zeros (summed by whole CTAs, ntt.cuh:spmv_long_kernel)
  uint32_t n_long[3];
  void* long_partial;       // max(n_long) * 64 slice sums
};

void zkb_r1cs_free(zkb_r1cs* r) {
  if (!r) return;
  if (ctx_ready()) {
    cudaStreamSynchronize(S());
    for (int i = 0; i < 3; i++) {
      cudaFree(r->row_ptr[i]);
      cudaFree(r->col[i]);
      cudaFree(r->val[i]);
      cudaFree(r->long_rows[i]);
    }
    cudaFree(r->w);
    cudaFree(r->long_partial);
  }
  delete r;
}

int zkb_r1cs_create(int curve, size_t n_rows, size_t n_cols, const uint64_t* const row_ptr[3], const uint32_t* const col[3],
                    const uint64_t* const val[3], zkb_r1cs** out) {
  NEED_INIT();
  CHECK_CURVE(curve);
  if (n_cols >= ((size_t)1 << 32)) return set_error(ZKB_ERR_ARG, "r1cs: too many columns");
  zkb_r1cs* r = new zkb_r1cs();
  memset(r, 0, sizeof(*r));
  r->curve = curve;
  r->n_rows = n_rows;
  r->n_cols = n_cols;
  int rc = ZKB_OK;
  for (int i = 0; i < 3 && rc == ZKB_OK; i++) {
    size_t nnz = (size_t)row_ptr[i][n_rows];
    for (size_t k = 0; k < nnz; k++)
      if (col[i][k] >= n_cols) rc = set_error(ZKB_ERR_ARG, "r1cs: column index out of range");
    if (rc) break;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&r->row_ptr[i], (n_rows + 1) * 8)) != cudaSuccess ||
        (e = cudaMalloc((void**)&r->col[i], (nnz ? nnz : 1) * 4)) != cudaSuccess ||
        (e = cudaMalloc((void**)&r->val[i], (nnz ? nnz : 1) * 32)) != cudaSuccess ||
        (e = ZKB_H2D(r->row_ptr[i], row_ptr[i], (n_rows + 1) * 8)) != cudaSuccess ||
        (nnz && (e = ZKB_H2D(r->col[i], col[i], nnz * 4)) != cudaSuccess) ||
        (nnz && (e = ZKB_H2D(r->val[i], val[i], nnz * 32)) != cudaSuccess)) {
      rc = cuda_fail((int)e, "r1cs upload", __FILE__, __LINE__);
      break;
    }
    rc = fr_reduce_dev(curve, nnz, r->val[i]);
    if (rc) break;
    std::vector<uint32_t> longs;
    for (size_t row = 0; row < n_rows; row++)
      if (row_ptr[i][row + 1] - row_ptr[i][row] > SPMV_LONG_ROW) longs.push_back((uint32_t)row);
    r->n_long[i] = (uint32_t)longs.size();
    if (!longs.empty()) {
      if ((e = cudaMalloc((void**)&r->long_rows[i], longs.size() * 4)) != cudaSuccess ||
          (e = ZKB_H2D(r->long_rows[i], longs.data(), longs.size() * 4)) != cudaSuccess ||
          (e = cudaStreamSynchronize(S())) != cudaSuccess) {     // (`longs` dies at the end of this iteration)
        rc = cuda_fail((int)e, "r1cs long-row list", __FILE__, __LINE__);
        break;
      }
    }
  }
  if (rc == ZKB_OK) {
    uint32_t most = 0;
    for (int i = 0; i < 3; i++) most = r->n_long[i] > most ? r->n_long[i] : most;
    if (most) {
      cudaError_t e = cudaMalloc(&r->long_partial, (size_t)most * 64 * 32);
      if (e != cudaSuccess) rc = cuda_fail((int)e, "r1cs long-row partial sums", __FILE__, __LINE__);
    }
  }
  if (rc == ZKB_OK) {
    cudaError_t e = cudaMalloc(&r->w, (n_cols ? n_cols : 1) * 32);
    if (e != cudaSuccess) rc = cuda_fail((int)e, "r1cs witness buffer", __FILE__, __LINE__);
  }
  if (rc == ZKB_OK) {
    cudaError_t e = cudaStreamSynchronize(S());
    if (e != cudaSuccess) rc = cuda_fail((int)e, "r1cs sync", __FILE__, __LINE__);
  }
  if (rc) {
    zkb_r1cs_free(r);
    return rc;
  }
  *out = r;
  return ZKB_OK;
}

// witness (n_cols canonical-or-not 256-bit values, host or device) -> r->w, reduced mod r
static int r1cs_load_witness(zkb_r1cs* r, const void* witness, int on_device) {
  if (!on_device) count_h2d(r->n_cols * 32);
  ZKB_CUDA(cudaMemcpyAsync(r->w, witness, r->n_cols * 32, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, S()));
  return fr_reduce_dev(r->curve, r->n_cols, r->w);
}
static int r1cs_spmv3(zkb_r1cs* r, size_t n_out, void* d_a, void* d_b, void* d_c) {
  if (n_out < r->n_rows) return set_error(ZKB_ERR_ARG, "r1cs: output shorter than the row count");
  void* outs[3] = {d_a, d_b, d_c};
  int rc;
  for (int i = 0; i < 3; i++)
    if ((rc = spmv_dev(r->curve, n_out, r->n_rows, r->row_ptr[i], r->col[i], r->val[i], r->w, outs[i], r->long_rows[i], r->n_long[i],
                       r->long_partial)))
      return rc;
  return ZKB_OK;
}
static int r1cs_eval_dev(zkb_r1cs* r, const uint64_t* witness, size_t n_out, void* d_a, void* d_b, void* d_c) {
  int rc;
  if ((rc = r1cs_load_witness(r, witness, 0))) return rc;
  return r1cs_spmv3(r, n_out, d_a, d_b, d_c);
}

int zkb_r1cs_eval(zkb_r1cs* r, const uint64_t* witness, size_t n_out, uint64_t* a, uint64_t* b, uint64_t* c) {
  NEED_INIT();
  if (!r) return set_error(ZKB_ERR_ARG, "null r1cs");
  void* d;
  int rc;
  if ((rc = stage(5, 3 * n_out * 32, &d))) return rc;
  char* p = (char*)d;
  if ((rc = r1cs_eval_dev(r, witness, n_out, p, p + n_out * 32, p + 2 * n_out * 32))) return rc;
  uint64_t* dst[3] = {a, b, c};
  for (int i = 0; i < 3; i++) ZKB_CUDA(ZKB_D2H(dst[i], p + i * n_out * 32, n_out * 32));
  ZKB_CUDA(cudaStreamSynchronize(S()));
  return ZKB_OK;
}

int zkb_r1cs_eval_dev(zkb_r1cs* r, const void* d_witness, size_t n_out, void* d_a, void* d_b, void* d_c) {
  NEED_INIT();
  if (!r) return set_error(ZKB_ERR_ARG, "null r1cs");
  int rc;
  if ((rc = r1cs_load_witness(r, d_witness, 1))) return rc;
  return r1cs_spmv3(r, n_out, d_a, d_b, d_c);
}

static int prove_witness_checks(zkb_groth16_pk* pk, zkb_r1cs* r1cs, size_t n_public) {
  if (!pk || !r1cs) return set_error(ZKB_ERR_ARG, "null proving key or r1cs");
  if (pk->curve != r1cs->curve) return set_error(ZKB_ERR_ARG, "proving key and r1cs are on different curves");
  if (n_public > r1cs->n_cols || r1cs->n_cols - n_public != pk->n_kdelta)
    return set_error(ZKB_ERR_ARG, "Length of kdelta_1 and private_witness must be equal");
  return ZKB_OK;
}

// witness already canonical and resident in r1cs->w: SpMV x3 -> quotient -> the slice's five MSMs
// Digit sorts under the transforms: the sort of an MSM needs its scalars and nothing else, and the scalars of three of the four
// sorts of a proof exist long before the MSM batch starts -- the private witness from the beginning, U and V once the three
// interpolations are done (four more transforms follow).  Those sorts are enqueued on a side stream at these points and run under
// the SpMV / NTT pipeline (both sides are latency-bound and leave each other room); the batch then starts accumulating at once and
// only H is sorted on the critical path.  One held arena epoch carries the transforms' and all MSM scratch.
struct PresortCtx {
  zkb_groth16_pk* pk;
  MsmJob job[5];
  cudaStream_t sort_st;
  cudaEvent_t ev;
  int rc;
};
static void presort_uv(void* arg) {   // called by groth16_h_dev when U and V are final on the library stream
  PresortCtx* c = (PresortCtx*)arg;
  if (cudaEventRecord(c->ev, S()) != cudaSuccess || cudaStreamWaitEvent(c->sort_st, c->ev, 0) != cudaSuccess) {
    c->rc = set_error(ZKB_ERR_CUDA, "presort fork failed");
    return;
  }
  const int b2 = groth16_pos(c->pk, MSM_B2), a = groth16_pos(c->pk, MSM_A);
  c->rc = msm_presort(c->pk->curve, c->job[b2], c->pk->wrank, c->pk->wworld, &g16_tk[b2], c->sort_st);   // V (B1 shares it)
  if (!c->rc) c->rc = msm_presort(c->pk->curve, c->job[a], c->pk->wrank, c->pk->wworld, &g16_tk[a], c->sort_st);   // U
}
static int partial_from_resident_witness(zkb_groth16_pk* pk, zkb_r1cs* r1cs, size_t n_public) {
  const size_t bytes = pk->n * 32;
  char* w = pk->work;
  const void* d_priv = (char*)r1cs->w + n_public * 32;
  int rc;
  // the satisfiability answer is read AFTER the MSMs (check == 2): nothing stops the host between the transforms and the batch
  static const bool presort_on = [] { const char* e = getenv("ZKB_PRESORT"); return !e || atoi(e) != 0; }();
  if (!presort_on || pk->log_n < 12) {
    if ((rc = r1cs_spmv3(r1cs, pk->n, w, w + bytes, w + 2 * bytes))) return rc;
    if ((rc = groth16_h_dev(pk->curve, pk->log_n, w, w + bytes, w + 2 * bytes, w + 3 * bytes, w + 4 * bytes, w + 5 * bytes,
                            w + 6 * bytes, 2)))
      return rc;
    if ((rc = groth16_msms(pk, d_priv))) return rc;
    return groth16_flag_result();
  }
  static cudaEvent_t ev = nullptr;
  if (!ev) ZKB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  PresortCtx ctx;
  ctx.pk = pk;
  ctx.sort_st = (cudaStream_t)ctx_side_stream(7);
  ctx.ev = ev;
  ctx.rc = ZKB_OK;
  if (!ctx.sort_st) return set_error(ZKB_ERR_CUDA, "cannot create the sort stream");
  groth16_jobs(pk, d_priv, ctx.job);
  size_t need_msm = 0;
  if ((rc = msm_batch_need(pk->curve, ctx.job, 5, pk->wrank, pk->wworld, &need_msm))) return rc;
  if ((rc = scratch_hold_begin(need_msm + 6 * ((size_t)32 << pk->log_n) + (1 << 20)))) return rc;
  auto done = [&](int code) {
    for (int i = 0; i < 5; i++) msm_presort_cancel(&g16_tk[i]);   // (no-op for tickets the batch consumed)
    scratch_hold_end();
    return code;
  };
  // the witness is reduced and resident (r1cs_load_witness, library stream): its sort can start here
  if (cudaEventRecord(ev, S()) != cudaSuccess || cudaStreamWaitEvent(ctx.sort_st, ev, 0) != cudaSuccess)
    return done(set_error(ZKB_ERR_CUDA, "presort fork failed"));
  const int kw = groth16_pos(pk, MSM_KW);
  if ((rc = msm_presort(pk->curve, ctx.job[kw], pk->wrank, pk->wworld, &g16_tk[kw], ctx.sort_st))) return done(rc);
  if ((rc = r1cs_spmv3(r1cs, pk->n, w, w + bytes, w + 2 * bytes))) return done(rc);
  if ((rc = groth16_h_dev(pk->curve, pk->log_n, w, w + bytes, w + 2 * bytes, w + 3 * bytes, w + 4 * bytes, w + 5 * bytes,
                          w + 6 * bytes, 2, presort_uv, &ctx)))
    return done(rc);
  if (ctx.rc) return done(ctx.rc);
  if ((rc = groth16_msms(pk, d_priv))) return done(rc);
  return done(groth16_flag_result());
}

// ---- chain spreading (several GPUs): zkb_groth16_spread_begin .. exchange .. [_quotient] .. exchange .. zkb_groth16_spread_finish ----
// What zkb_groth16_partial does in one call, cut where the ranks exchange data (the steps are spelled out in include/zkb200.h and
// driven by zksnake_b200/dist.py:exchange_chains).  The scratch arena stays held from begin to finish: the private-witness digit
// sort (and, on a rank without a chain, the whole [K w] MSM) is already running when begin returns.
static struct SpreadState {
  bool active = false;
  bool kw_enqueued = false;
  zkb_groth16_pk* pk = nullptr;
  int* flag = nullptr;
  void* tmp = nullptr;
  cudaEvent_t ev = nullptr;
} g_spread;
static void spread_drop() {
  for (int i = 0; i < 5; i++) msm_presort_cancel(&g16_tk[i]);
  for (int i = 0; i < 5; i++)   // error paths: reductions of unjoined batches may still run on the arena (collected tickets are empty)
    if (!g16_tk[i].empty && g16_tk[i].event) cudaStreamWaitEvent(S(), (cudaEvent_t)g16_tk[i].event, 0);
  scratch_hold_end();
  g_spread.active = g_spread.kw_enqueued = false;
}
int zkb_groth16_spread_begin(zkb_groth16_pk* pk, zkb_r1cs* r1cs, const void* witness, int witness_on_device, size_t n_public,
                             unsigned chain_mask, void* d_coeffs, void* d_evals) {
  NEED_INIT();
  int rc;
  if (g_spread.active) spread_drop();
  if ((rc = prove_witness_checks(pk, r1cs, n_public))) return rc;
  if (!d_coeffs || !d_evals || chain_mask > 7) return set_error(ZKB_ERR_ARG, "spread: bad argument");
  if (pk->wworld < 2 || groth16_pos(pk, MSM_KW) != 0 || groth16_pos(pk, MSM_HZ) != 4)
    return set_error(ZKB_ERR_ARG, "spread: the proving key must be a window shard of a multi-GPU proof");
  if ((rc = r1cs_load_witness(r1cs, witness, witness_on_device))) return rc;
  const size_t bytes = pk->n * 32;
  char* w = pk->work;
  const void* d_priv = (char*)r1cs->w + n_public * 32;
  if (!g_spread.ev) ZKB_CUDA(cudaEventCreateWithFlags(&g_spread.ev, cudaEventDisableTiming));
  cudaStream_t sort_st = (cudaStream_t)ctx_side_stream(7);
  if (!sort_st) return set_error(ZKB_ERR_CUDA, "cannot create the sort stream");
  int* hflag = groth16_flag_host();
  if (!hflag) return set_error(ZKB_ERR_CUDA, "cannot allocate the pinned flag");
  MsmJob job[5];
  groth16_jobs(pk, d_priv, job);
  size_t need_msm = 0;
  if ((rc = msm_batch_need(pk->curve, job, 5, pk->wrank, pk->wworld, &need_msm))) return rc;
  if ((rc = scratch_hold_begin(need_msm + bytes + (1 << 20)))) return rc;
  g_spread.active = true;
  g_spread.kw_enqueued = false;
  g_spread.pk = pk;
  auto fail = [&](int code) {
    spread_drop();
    return code;
  };
  if (cudaEventRecord(g_spread.ev, S()) != cudaSuccess || cudaStreamWaitEvent(sort_st, g_spread.ev, 0) != cudaSuccess)
    return fail(set_error(ZKB_ERR_CUDA, "presort fork failed"));
  if ((rc = msm_presort(pk->curve, job[0], pk->wrank, pk->wworld, &g16_tk[0], sort_st))) return fail(rc);   // [K w]: position 0
  g_spread.flag = (int*)scratch_take(256);
  g_spread.tmp = scratch_take(bytes);
  if (!g_spread.flag || !g_spread.tmp) return fail(set_error(ZKB_ERR_CUDA, "scratch exhausted"));
  if ((rc = r1cs_spmv3(r1cs, pk->n, w, w + bytes, w + 2 * bytes))) return fail(rc);
  if ((rc = groth16_check_dev(pk->curve, pk->log_n, w, w + bytes, w + 2 * bytes, g_spread.flag))) return fail(rc);
  if (ZKB_D2H(hflag, g_spread.flag, sizeof(int)) != cudaSuccess) return fail(set_error(ZKB_ERR_CUDA, "spread: flag copy failed"));
  for (int c = 0; c < 3; c++)
    if (chain_mask >> c & 1)
      if ((rc = groth16_chain_dev(pk->curve, pk->log_n, c, w + c * bytes, (char*)d_coeffs + c * bytes, (char*)d_evals + c * bytes,
                                  g_spread.tmp)))
        return fail(rc);
  if (!chain_mask) {
    // nothing to transform on this rank: [K w] needs the witness only, so it runs while the chain owners work
    if ((rc = groth16_enqueue(pk, d_priv, 0, 1, false))) return fail(rc);
    g_spread.kw_enqueued = true;
  }
  return ZKB_OK;
}
int zkb_groth16_spread_quotient(zkb_groth16_pk* pk, void* d_evals, void* d_h) {
  NEED_INIT();
  if (!g_spread.active || g_spread.pk != pk || !d_evals || !d_h) return set_error(ZKB_ERR_ARG, "spread: quotient without begin");
  const size_t bytes = pk->n * 32;
  char* e = (char*)d_evals;
  int rc = groth16_hfin_dev(pk->curve, pk->log_n, e, e + bytes, e + 2 * bytes, d_h, g_spread.tmp);
  if (rc) spread_drop();
  return rc;
}
int zkb_groth16_spread_finish(zkb_groth16_pk* pk, zkb_r1cs* r1cs, size_t n_public, const void* d_coeffs, const void* d_h,
                              void* h_ready_event, uint64_t* msm_xy, int* msm_inf) {
  NEED_INIT();
  if (!g_spread.active || g_spread.pk != pk || !r1cs || !d_coeffs || !d_h)
    return set_error(ZKB_ERR_ARG, "spread: finish without begin");
  int rc;
  const size_t bytes = pk->n * 32;
  char* w = pk->work;
  const void* d_priv = (char*)r1cs->w + n_public * 32;
  auto done = [&](int code) {
    spread_drop();
    return code;
  };
  // U, V (complete on the library stream behind the caller's broadcasts) become the MSM scalars
  if (cudaMemcpyAsync(w + 3 * bytes, d_coeffs, 2 * bytes, cudaMemcpyDeviceToDevice, S()) != cudaSuccess)
    return done(set_error(ZKB_ERR_CUDA, "spread: copy failed"));
  PresortCtx ctx;
  ctx.pk = pk;
  ctx.sort_st = (cudaStream_t)ctx_side_stream(7);
  ctx.ev = g_spread.ev;
  ctx.rc = ZKB_OK;
  groth16_jobs(pk, d_priv, ctx.job);
  presort_uv(&ctx);
  if (ctx.rc) return done(ctx.rc);
  // the MSMs that need U, V and the witness only; H is still on its way on most ranks
  const int first = g_spread.kw_enqueued ? 1 : 0, last = 4;   // shard order: KW, B2, A, B1, HZ
  if ((rc = groth16_enqueue(pk, d_priv, first, last - first, false))) return done(rc);
  // H: complete in d_h once h_ready_event has fired (recorded by the caller on the stream its exchange ran on; null: d_h is already
  // ordered on the library stream)
  // (sorting H on the sort stream as soon as it arrives, under the accumulations, was measured at 8 GPUs: 6.64 against 6.57 ms --
  // the sort only gets the SM slots the persistent accumulation grid leaves, profiles/R2z_n8_hsort.json; it stays in line)
  if (h_ready_event && cudaStreamWaitEvent(S(), (cudaEvent_t)h_ready_event, 0) != cudaSuccess)
    return done(set_error(ZKB_ERR_CUDA, "spread: cannot wait for the quotient"));
  if (cudaMemcpyAsync(w + 6 * bytes, d_h, bytes, cudaMemcpyDeviceToDevice, S()) != cudaSuccess)
    return done(set_error(ZKB_ERR_CUDA, "spread: copy failed"));
  if ((rc = groth16_enqueue(pk, d_priv, last, 1, true))) return done(rc);
  for (int i = 0; i < last; i++)   // (the unjoined batches: order the library stream behind their reductions too)
    if (!g16_tk[i].empty && g16_tk[i].event && cudaStreamWaitEvent(S(), (cudaEvent_t)g16_tk[i].event, 0) != cudaSuccess)
      return done(set_error(ZKB_ERR_CUDA, "spread: join failed"));
  g_spread.kw_enqueued = false;
  if ((rc = groth16_collect(pk))) return done(rc);
  if ((rc = groth16_flag_result())) return done(rc);
  groth16_export_partials(pk, msm_xy, msm_inf);
  return done(ZKB_OK);
}

int zkb_groth16_partial(zkb_groth16_pk* pk, zkb_r1cs* r1cs, const void* witness, int witness_on_device, size_t n_public,
                        uint64_t* msm_xy, int* msm_inf) {
  NEED_INIT();
  int rc;
  if ((rc = prove_witness_checks(pk, r1cs, n_public))) return rc;
  if ((rc = r1cs_load_witness(r1cs, witness, witness_on_device))) return rc;
  if ((rc = partial_from_resident_witness(pk, r1cs, n_public))) return rc;
  groth16_export_partials(pk, msm_xy, msm_inf);
  return ZKB_OK;
}

int zkb_groth16_prove_witness(zkb_groth16_pk* pk, zkb_r1cs* r1cs, const uint64_t* witness, size_t n_public,
                              const uint64_t r[4], const uint64_t s[4], uint64_t* out_a, uint64_t* out_b, uint64_t* out_c,
                              int out_inf[3]) {
  NEED_INIT();
  int rc;
  if ((rc = prove_witness_checks(pk, r1cs, n_public))) return rc;
  if (pk->len != pk->n || pk->klen != pk->n_kdelta || pk->wworld != 1)
    return set_error(ZKB_ERR_ARG, "this proving key holds one slice only: use zkb_groth16_partial + zkb_groth16_assemble");
  pre_start(pk, r, s);   // host multiples of delta under the GPU work
  if ((rc = r1cs_load_witness(r1cs, witness, 0))) return rc;
  if ((rc = partial_from_resident_witness(pk, r1cs, n_public))) return rc;
  return zkb_groth16_assemble(pk, &pk->msm_xy[0][0], pk->msm_inf, r, s, out_a, out_b, out_c, out_inf);
}

int zkb_groth16_prove_witness_dev(zkb_groth16_pk* pk, zkb_r1cs* r1cs, const void* d_witness, size_t n_public,
                                  const uint64_t r[4], const uint64_t s[4], uint64_t* out_a, uint64_t* out_b, uint64_t* out_c,
                                  int out_inf[3]) {
  NEED_INIT();
  int rc;
  if ((rc = prove_witness_checks(pk, r1cs, n_public))) return rc;
  if (pk->len != pk->n || pk->klen != pk->n_kdelta || pk->wworld != 1)
    return set_error(ZKB_ERR_ARG, "this proving key holds one slice only: use zkb_groth16_partial + zkb_groth16_assemble");
  pre_start(pk, r, s);   // host multiples of delta under the GPU work
  if ((rc = r1cs_load_witness(r1cs, d_witness, 1))) return rc;
  if ((rc = partial_from_resident_witness(pk, r1cs, n_public))) return rc;
  return zkb_groth16_assemble(pk, &pk->msm_xy[0][0], pk->msm_inf, r, s, out_a, out_b, out_c, out_inf);
}

int zkb_groth16_last_poly(zkb_groth16_pk* pk, int which, uint64_t* out) {
  NEED_INIT();
  if (!pk || which < 0 || which > 2) return set_error(ZKB_ERR_ARG, "bad argument");
  const size_t bytes = pk->n * 32;
  static const int slot[3] = {3, 4, 6};
  ZKB_CUDA(ZKB_D2H(out, pk->work + slot[which] * bytes, bytes));
  ZKB_CUDA(cudaStreamSynchronize(S()));
  return ZKB_OK;
}
int zkb_groth16_last_msm(zkb_groth16_pk* pk, int which, uint64_t* out_xy, int* out_inf) {
  if (!pk || which < 0 || which > 4) return set_error(ZKB_ERR_ARG, "bad argument");
  memcpy(out_xy, pk->msm_xy[which], affine_bytes(pk->curve, which == 2 ? 2 : 1));
  *out_inf = pk->msm_inf[which];
  return ZKB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------- self tests
namespace {
template <class F>
__host__ __device__ F field_op(int op, const F& a, const F& b) {
  F x = to_mont(a), y = to_mont(b), r;
  if (op == 0) r = x * y;
  else if (op == 1) r = x + y;
  else if (op == 2) r = x - y;
  else if (op == 3) r = inv(x);
  else if (op == 5 || op == 6) {
    // lazy dot product (the pair kernels' Fp2 multiplier): op 5 = x y + (x + y)(x - y); op 6 feeds the unreduced operand p
    if constexpr ((F::Params::MOD(F::N - 1) >> 30) == 0) {
      if (op == 5) r = mont_dot2(x, y, x + y, x - y);
      else r = mont_dot2(x, y, neg_lazy(F::zero()), neg_lazy(y));   // x y + p (p - y) = x y  (mod p)
    } else {
      r = op == 5 ? x * y + (x + y) * (x - y) : x * y;
    }
  }
  else r = neg(x);
  return from_mont(r);
}
template <class F>
__global__ void field_op_kernel(int op, size_t n, const F* a, const F* b, F* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = field_op(op, a[i], b[i]);
}
template <class F>
int field_op_host_t(int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  const F* fa = (const F*)a;
  const F* fb = (const F*)b;
  F* fo = (F*)out;
  for (size_t i = 0; i < n; i++) fo[i] = field_op(op, fa[i], fb[i]);
  return ZKB_OK;
}
template <class F>
int field_op_dev_t(int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  F *da, *db, *dout;
  size_t bytes = n * sizeof(F);
  ZKB_CUDA(cudaMalloc((void**)&da, bytes));
  ZKB_CUDA(cudaMalloc((void**)&db, bytes));
  ZKB_CUDA(cudaMalloc((void**)&dout, bytes));
  ZKB_CUDA(ZKB_H2D(da, a, bytes));
  ZKB_CUDA(ZKB_H2D(db, b, bytes));
  field_op_kernel<F><<<(unsigned)((n + 127) / 128), 128, 0, S()>>>(op, n, da, db, dout);
  count_launch();
  ZKB_CUDA(cudaGetLastError());
  ZKB_CUDA(ZKB_D2H(out, dout, bytes));
  ZKB_CUDA(cudaStreamSynchronize(S()));
  cudaFree(da);
  cudaFree(db);
  cudaFree(dout);
  return ZKB_OK;
}
}  // namespace

extern "C" {
int zkb_test_field_op_host(int field, int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  switch (field) {
    case 0: return field_op_host_t<fr_bn>(op, n, a, b, out);
    case 1: return field_op_host_t<fq_bn>(op, n, a, b, out);
    case 2: return field_op_host_t<fr_bls>(op, n, a, b, out);
    case 3: return field_op_host_t<fq_bls>(op, n, a, b, out);
  }
  return set_error(ZKB_ERR_ARG, "unknown field id");
}
int zkb_test_field_op_dev(int field, int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  NEED_INIT();
  switch (field) {
    case 0: return field_op_dev_t<fr_bn>(op, n, a, b, out);
    case 1: return field_op_dev_t<fq_bn>(op, n, a, b, out);
    case 2: return field_op_dev_t<fr_bls>(op, n, a, b, out);
    case 3: return field_op_dev_t<fq_bls>(op, n, a, b, out);
  }
  return set_error(ZKB_ERR_ARG, "unknown field id");
}

// sum_i k_i P_i on the host path (has_scalar[i] == 0 means k_i = 1); exercises ec.cuh + ff_host.h without a GPU
int zkb_point_lincomb(int curve, int group, int n_terms, const uint64_t* points, const int* infs,
                          const uint64_t* scalars, const int* has_scalar, uint64_t* out_xy, int* out_inf) {
  CHECK_CURVE(curve);
  CHECK_GROUP(group);
  size_t ab = affine_bytes(curve, group) / 8;
  std::vector<const uint64_t*> pp(n_terms), ss(n_terms);
  for (int i = 0; i < n_terms; i++) {
    pp[i] = points + i * ab;
    ss[i] = has_scalar[i] ? scalars + i * 4 : nullptr;
  }
  host_lincomb(curve, group, n_terms, pp.data(), infs, ss.data(), out_xy, out_inf);
  return ZKB_OK;
}
}
