/* zkb200.h -- C ABI of libzkb200.so: the B200 (sm_100a) back end for zksnake's proving hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one call the reference makes from its pyo3 extension
 * module `zksnake._algebra` into arkworks (reference paths are relative to /root/reference):
 *
 *   zkb_ntt / zkb_ntt_dev            src/bn254/polynomial.rs:536-545 fft, :548-559 coset_fft, :562-571 ifft,
 *                                    :574-585 coset_ifft            (src/bls12_381/polynomial.rs: same lines)
 *   zkb_vec_op / zkb_vec_op_dev      src/bn254/polynomial.rs:588-607 add_over_evaluation_domain,
 *                                    :610-634 mul_over_evaluation_domain
 *   zkb_fr_reduce                    the `Fr::from(BigUint)` loops, e.g. src/bn254/polynomial.rs:537-540
 *   zkb_msm / zkb_msm_dev            src/bn254/curve.rs:356-373 multiscalar_mul_g1, :375-392 multiscalar_mul_g2
 *                                    (src/bls12_381/curve.rs:367-384, :386-403)
 *   zkb_batch_mul_dev                src/bn254/curve.rs:326-354 batch_multi_scalar_g1/g2 (setup side)
 *   zkb_groth16_h / _h_dev           python/zksnake/groth16/qap.py:42-71 QAP.evaluate_witness (after the A.w/B.w/C.w dots)
 *   zkb_groth16_pk_* / _prove        python/zksnake/groth16/protocol.py:115-165 Groth16.prove
 *   zkb_groth16_partial / _spread_* / _assemble(_partials)   the same two calls (qap.py:42-71, protocol.py:133-165) cut where the
 *                                    ranks of a multi-GPU proof exchange data (the reference has no multi-device path)
 *   zkb_r1cs_create / _eval(_dev)    python/zksnake/array.py:36-43 SparseArray.dot (x3: groth16/qap.py:53-55); over the transposed
 *                                    matrices: the L/R/O loop of Groth16.setup (groth16/protocol.py:64-77)
 *   zkb_fr_*_dev, zkb_plonk_*        the Polynomial / list glue of python/zksnake/plonk/protocol.py:270-466
 *   zkb_msm_table_*                  the same multiexp over a fixed vector (proving key, KZG SRS: commitment/polynomial/kzg.py:32-51)
 *   zkb_points_compress / _decompress  PointG1/G2.to_bytes / from_bytes over a vector: src/bn254/curve.rs:127-141, 300-314 and
 *                                    the key serialisers python/zksnake/groth16/serialization.py:68-220, plonk/serialization.py:157-353
 *
 * Conventions
 *   - curve: ZKB_BN254 (0) or ZKB_BLS12_381 (1); group: 1 = G1, 2 = G2.
 *   - Fr elements on the wire: 4 x uint64 little-endian limbs, canonical (non-Montgomery).  Values >= r are accepted
 *     wherever the reference accepts arbitrary non-negative ints and are reduced mod r (Fr::from(BigUint)).
 *   - Fq elements: 4 x uint64 (BN254) / 6 x uint64 (BLS12-381), canonical.  G1 affine = (x, y); G2 affine =
 *     (x.c0, x.c1, y.c0, y.c1).  The point at infinity is the all-zero coordinate tuple (0,0 is on neither curve).
 *   - "_dev" entry points take device pointers (from zkb_dev_alloc or any CUDA allocation of the same process) and
 *     are asynchronous on the library stream unless they return host data.  Device point vectors are kept in
 *     Montgomery form (zkb_points_upload / zkb_points_download convert); device Fr vectors are canonical.
 *   - Return value: 0 on success, a negative ZKB_ERR_* code otherwise; zkb_last_error() gives the text.  The host
 *     binding maps codes to the reference's exception types.
 *   - No exceptions and no allocator ownership cross this boundary: callers allocate outputs.
 *   - One host thread per process drives the library (one process per GPU, as torchrun launches them): entry points are
 *     not re-entrant -- they share one stream, one scratch arena and per-call result tickets.
 *   - There is NO CPU fallback: every compute entry point fails with ZKB_ERR_NOINIT / ZKB_ERR_CUDA without a B200.
 */
#ifndef ZKB200_H
#define ZKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKB_BN254 0
#define ZKB_BLS12_381 1

#define ZKB_OK 0
#define ZKB_ERR_CUDA (-1)          /* CUDA failure or no usable device            -> RuntimeError */
#define ZKB_ERR_ARG (-2)           /* bad argument                                -> ValueError   */
#define ZKB_ERR_MISMATCH (-3)      /* "Number of points and scalars mismatch"     -> ValueError   (curve.rs:369-371) */
#define ZKB_ERR_DOMAIN (-4)        /* "Domain size is too large"                  -> ValueError   (polynomial.rs:638-639) */
#define ZKB_ERR_NOT_DIVISIBLE (-5) /* "(U * V - W) did not divided by Z to zero"  -> ValueError   (qap.py:68-69) */
#define ZKB_ERR_NOINIT (-6)        /* zkb_init not called                         -> RuntimeError */
#define ZKB_ERR_POINT (-7)         /* "Cannot deserialize point"                  -> ValueError   (curve.rs:134-141) */

#define ZKB_VEC_MUL 0
#define ZKB_VEC_ADD 1
#define ZKB_VEC_SUB 2

/* ---- context ---------------------------------------------------------------------------------------------- */
int zkb_device_count(void);
int zkb_init(int device);                 /* binds this process to one GPU (one process per GPU) */
void zkb_shutdown(void);
const char* zkb_last_error(void);
void* zkb_stream(void);                   /* the cudaStream_t all work is issued on */
int zkb_sync(void);
unsigned long long zkb_launch_count(void); /* kernels launched by this library so far */
/* bytes the library has copied host->device / device->host so far (bench.py's e2e accounting) */
void zkb_transfer_count(unsigned long long* h2d_bytes, unsigned long long* d2h_bytes);
int zkb_timer_start(void);                /* CUDA events on the library stream */
int zkb_timer_stop(float* ms);
/* Per-kernel-family device time (CUDA events on the library stream around every launch of the family), for bench.py's
 * roofline lines.  tag: 0 NTT (one record per whole transform), 1 MSM digit sort (count/scan/scatter), 2 MSM bucket
 * accumulation G1, 3 the same G2, 4 MSM bucket reduction, 5 R1CS SpMV, 6 element-wise Fr kernels, 7 other. */
int zkb_prof_enable(int on);              /* also clears the records */
int zkb_prof_read(int tag, float* total_ms, unsigned long long* count);
/* Integer-pipe microbenchmark: sustained 32-bit multiply-add lane-operations per second of this GPU (the MSM roofline
 * denominator).  which: 0 = IMAD (mad.lo.u32), 1 = IMAD.WIDE (mad.wide.u32 counted as one op). */
int zkb_imad_peak(int which, double* ops_per_s);
/* Shape of the bucket-accumulation kernel compiled for (curve, group): lanes that share one point (1: one thread per point,
 * msm_accumulate_kernel; 2: an Fp2 half per lane, msm_accumulate_pair_kernel) and resident CTAs per SM.  bench.py derives the
 * executed multiply-add count of its roofline lines from it. */
int zkb_msm_kernel_info(int curve, int group, int* lanes_per_point, int* ctas_per_sm);

/* ---- raw memory ------------------------------------------------------------------------------------------- */
int zkb_dev_alloc(size_t bytes, void** out);
int zkb_dev_free(void* p);
int zkb_host_alloc(size_t bytes, void** out);   /* pinned */
int zkb_host_free(void* p);
int zkb_h2d(void* dst, const void* src, size_t bytes);
/* the same without the synchronisation: enqueued on the library stream; `src` must stay untouched until the stream has passed the
 * copy (pinned memory from zkb_host_alloc makes it a true asynchronous DMA transfer) */
int zkb_h2d_async(void* dst, const void* src, size_t bytes);
int zkb_d2h(void* dst, const void* src, size_t bytes);
int zkb_d2d(void* dst, const void* src, size_t bytes);
int zkb_memset(void* dst, int value, size_t bytes);

/* ---- Fr vectors ------------------------------------------------------------------------------------------- */
/* N = 2^log_n.  in_len <= any; inputs are zero-padded or truncated to N.  coset: 0 = plain, 1 = the reference's
 * coset (offset = generator of the size-N domain), 2 = the coset g<w> with g the multiplicative generator of Fr (5 on BN254,
 * 7 on BLS12-381; used for quotient polynomials: X^n - 1 has no zero on it).  out holds N elements. */
int zkb_ntt(int curve, int inverse, int coset, uint32_t log_n, const uint64_t* in, size_t in_len, uint64_t* out);
int zkb_ntt_dev(int curve, int inverse, int coset, uint32_t log_n, const void* d_in, size_t in_len, void* d_out);
/* out[i] = a[i] op b[i] for i < n; a / b shorter than n are zero-extended (mul_over_evaluation_domain semantics). */
int zkb_vec_op(int curve, int op, size_t n, const uint64_t* a, size_t na, const uint64_t* b, size_t nb, uint64_t* out);
int zkb_vec_op_dev(int curve, int op, size_t n, const void* d_a, size_t na, const void* d_b, size_t nb, void* d_out);
int zkb_fr_reduce(int curve, size_t n, uint64_t* inout);            /* arbitrary 256-bit values -> mod r */
int zkb_fr_reduce_dev(int curve, size_t n, void* d_inout);
int zkb_fr_powers_dev(int curve, const uint64_t base[4], const uint64_t scale[4], size_t n, void* d_out); /* scale*base^i */

/* Device-resident vector primitives for the PlonK prover glue (python/zksnake/plonk/protocol.py:270-466, utils.py:42-62,
 * src/bn254/polynomial.rs:404-489): all pointers are device Fr vectors (canonical), scalars are canonical host words (values
 * >= r are reduced).  Asynchronous on the library stream unless they return host data. */
int zkb_fr_axpy_dev(int curve, size_t n, const uint64_t s[4], const void* d_x, size_t nx, const void* d_y, size_t ny,
                    void* d_out);                       /* out[i] = s*x[i] + y[i], i < n; x / y zero-extended (d_y may be null) */
int zkb_fr_mul_powers_dev(int curve, size_t n, const uint64_t base[4], const uint64_t scale[4], const void* d_x,
                          void* d_out);                 /* out[i] = x[i] * scale * base^i  (d_x null: x = 1) */
int zkb_fr_inverse_dev(int curve, size_t n, const void* d_x, void* d_out);   /* batch inversion, 0 -> 0 (utils.py:42-62) */
/* op 0: exclusive prefix product, n+1 outputs (out[0] = 1, out[n] = product of all) -- the grand-product accumulator of
 * protocol.py:307-313.  op 1: inclusive suffix sum, n outputs (out[i] = x[i] + ... + x[n-1]). */
int zkb_fr_scan_dev(int curve, int op, size_t n, const void* d_x, void* d_out);
int zkb_fr_gather_dev(int curve, size_t n, const void* d_src, size_t stride, size_t offset, void* d_out);  /* src[offset + i*stride] */
int zkb_fr_gather_index_dev(int curve, size_t n, const void* d_src, const void* d_idx_u32, void* d_out);   /* src[idx[i]] */
int zkb_fr_eval_dev(int curve, size_t n, const void* d_coeffs, const uint64_t point[4], uint64_t out[4]);   /* sum c_i z^i (sync) */
/* *len = number of coefficients left after stripping trailing zeros (0 for the zero vector): the DensePolynomial invariant behind
 * Polynomial.coeffs() / degree() / is_zero() (src/bn254/polynomial.rs:132-160, 440) for a device-resident coefficient vector (sync) */
int zkb_fr_trim_dev(int curve, size_t n, const void* d_x, size_t* len);
/* q = p / (X^d - 1) (len - d coefficients, polynomial.rs:466-489); *exact = 0 when the remainder is non-zero (sync) */
int zkb_fr_div_vanishing_dev(int curve, size_t len, size_t d, const void* d_p, void* d_q, int* exact);
/* PlonK quotient evaluations on the coset g<w_q> of size q (2n..8n) in one pass (python/zksnake/plonk/protocol.py:240-262, 284-300,
 * 338-360 reach the same polynomial through ~20 transforms): d_inputs = coset evaluations of a, b, c, z, pi, qL, qR, qO, qM, qC,
 * sigma1, sigma2, sigma3, L1;  out[i] = (gate + alpha (id z - sg z(omega x)) + alpha^2 (z - 1) L1) / (x^n - 1) at x = g w_q^i;
 * zh_inv = the q/n distinct values of 1 / (x^n - 1) on the coset (index i mod q/n). */
int zkb_plonk_quotient_dev(int curve, size_t q, size_t n, const void* const d_inputs[14], const uint64_t g[4], const uint64_t omega_q[4],
                           const uint64_t beta[4], const uint64_t gamma[4], const uint64_t alpha[4], const uint64_t* zh_inv,
                           void* d_out);
int zkb_fr_add_sparse_dev(int curve, void* d_vec, size_t k, const uint64_t* idx, const uint64_t* vals, int subtract); /* async */

/* ---- points and MSM --------------------------------------------------------------------------------------- */
size_t zkb_affine_bytes(int curve, int group);
int zkb_points_upload(int curve, int group, const uint64_t* pts, size_t n, void* d_out);    /* canonical -> device Montgomery */
int zkb_points_download(int curve, int group, const void* d_pts, size_t n, uint64_t* out);  /* device Montgomery -> canonical */

/* Bulk wire format.  The reference's keys and proofs are concatenated ark-serialize COMPRESSED
 * points: PointG1/PointG2.to_bytes / from_bytes (src/bn254/curve.rs:127-141, 300-314; bls12_381/curve.rs likewise), looped
 * over a key one from_hex call per point by python/zksnake/groth16/serialization.py:70-159,181-220,
 * plonk/serialization.py:157-173,255-262 and ecc.py:128-142.  One kernel launch per vector here.
 *   zkb_compressed_bytes: 32 / 64 (BN254 G1 / G2), 48 / 96 (BLS12-381).
 *   zkb_points_compress:   n device points -> n encodings in host memory.
 *   zkb_points_decompress: n encodings in host memory -> n device points.  validate != 0 adds the prime-order subgroup check
 *     (ark's Validate::Yes, what from_bytes does): 1 = by the endomorphism criteria in G2 (psi(P) = [x]P on BLS12-381,
 *     [x+1]P + psi([x]P) + psi^2([x]P) = psi^3([2x]P) on BN254; r * P = infinity in G1), 2 = r * P = infinity everywhere --
 *     the same accept set, about three times the work in G2.  Flags, x < q, "infinity has x = 0" and "x^3 + b is a square"
 *     are always checked.  On an invalid encoding returns ZKB_ERR_POINT with *bad_index = the first offending point and *reason =
 *     1 flags | 2 coordinate not in field | 3 non-zero infinity | 4 not on curve | 5 not in the subgroup (both optional). */
size_t zkb_compressed_bytes(int curve, int group);
int zkb_points_compress(int curve, int group, const void* d_pts, size_t n, uint8_t* out);
int zkb_points_decompress(int curve, int group, const uint8_t* in, size_t n, int validate, void* d_pts, long long* bad_index,
                          int* reason);
/* sum_i scalars[i] * pts[i].  n_points must equal n_scalars (else ZKB_ERR_MISMATCH).  out_xy: one affine point. */
int zkb_msm(int curve, int group, const uint64_t* pts, size_t n_points, const uint64_t* scalars, size_t n_scalars,
            uint64_t* out_xy, int* out_inf);
int zkb_msm_dev(int curve, int group, const void* d_pts, const void* d_scalars, size_t n, uint64_t* out_xy, int* out_inf);
/* Window shard wrank of wworld (multi-GPU): the partial sum over the scalar windows
 * [W*wrank/wworld, W*(wrank+1)/wworld) of ALL n points, already scaled by 2^(c*first window); the wworld partial results add
 * up to zkb_msm_dev's result.  Every phase of the MSM (sort, accumulate, bucket reduction) shrinks by 1/wworld. */
int zkb_msm_dev_windows(int curve, int group, const void* d_pts, const void* d_scalars, size_t n, uint32_t wrank,
                        uint32_t wworld, uint64_t* out_xy, int* out_inf);
/* d_out[i] = scalars[i] * bases[single_base ? 0 : i]  (device Montgomery points in and out) */
int zkb_batch_mul_dev(int curve, int group, const void* d_bases, int single_base, const void* d_scalars, size_t n,
                      void* d_out);
void zkb_msm_set_tuning(int window_bits, int segment, int reduce_chunk);  /* 0 = heuristic */

/* Fixed-base tables: when the same points are used by many MSMs (a proving key, a KZG SRS -- the reference rebuilds nothing
 * either: python/zksnake/groth16/protocol.py:133-155 and commitment/polynomial/kzg.py:32-37 pass the same key vectors to every
 * multiexp), table[w * n + i] = 2^(c w) * P_i is computed once (W * n affine points in HBM; 0.9 GiB for 2^20 BN254 G1 points).
 * Window w of scalar i then gathers its own pre-multiplied point, so ALL windows share ONE set of 2^(c-1) buckets: the bucket
 * reduction shrinks W-fold and the window can be larger (fewer windows, fewer additions).  window_bits 0 = cost model for
 * `world` window-sharding ranks.  zkb_msm_table_dev covers the first n_scalars <= n points (ZKB_ERR_MISMATCH beyond). */
typedef struct zkb_msm_table zkb_msm_table;
int zkb_msm_table_create(int curve, int group, const void* d_pts, size_t n, uint32_t window_bits, uint32_t world,
                         zkb_msm_table** out);
void zkb_msm_table_free(zkb_msm_table* t);
int zkb_msm_table_info(const zkb_msm_table* t, uint32_t* window_bits, uint32_t* windows, size_t* bytes);
int zkb_msm_table_dev(zkb_msm_table* t, const void* d_scalars, size_t n_scalars, uint32_t wrank, uint32_t wworld,
                      uint64_t* out_xy, int* out_inf);
/* `count` (<= 8) MSMs over the same table as one batch: the accumulations run back to back, the latency-bound bucket
 * reductions overlap on side streams (the three commitments of a PlonK round).  out_xy: count affine points, tightly packed.
 * wrank / wworld: window shard as in zkb_msm_table_dev (0, 1 = the whole MSMs). */
int zkb_msm_table_batch_dev(zkb_msm_table* t, int count, const void* const* d_scalars, const size_t* n_scalars, uint32_t wrank,
                            uint32_t wworld, uint64_t* out_xy, int* out_inf);

/* ---- Groth16 ---------------------------------------------------------------------------------------------- */
/* a, b, c: the vectors A.w, B.w, C.w (n = 2^log_n each).  u, v, w, h receive n coefficients each (h[n-1] = 0).
 * Returns ZKB_ERR_NOT_DIVISIBLE when a[i]*b[i] != c[i] somewhere (the reference's non-zero remainder). */
int zkb_groth16_h(int curve, uint32_t log_n, const uint64_t* a, const uint64_t* b, const uint64_t* c, uint64_t* u,
                  uint64_t* v, uint64_t* w, uint64_t* h);
int zkb_groth16_h_dev(int curve, uint32_t log_n, const void* d_a, const void* d_b, const void* d_c, void* d_u, void* d_v,
                      void* d_w, void* d_h, int check);

/* Device-resident proving key (ProvingKey of python/zksnake/groth16/serialization.py:45-66).  The four point vectors are
 * device Montgomery buffers the key takes a reference to (the caller keeps ownership and must keep them alive);
 * the five single points are canonical host coordinates. */
typedef struct zkb_groth16_pk zkb_groth16_pk;
int zkb_groth16_pk_create(int curve, uint32_t log_n, const void* d_tau1, const void* d_tau2, const void* d_target1,
                          const void* d_kdelta1, size_t n_kdelta, const uint64_t* alpha1, const uint64_t* beta1,
                          const uint64_t* beta2, const uint64_t* delta1, const uint64_t* delta2, zkb_groth16_pk** out);
/* One rank's slice of the key for multi-GPU proving: d_tau1 / d_tau2 / d_target1 hold elements
 * [off, off+len) of the n-point vectors, d_kdelta1 holds elements [koff, koff+klen) of the n_kdelta-point vector. */
int zkb_groth16_pk_create_sharded(int curve, uint32_t log_n, const void* d_tau1, const void* d_tau2, const void* d_target1,
                                  size_t off, size_t len, const void* d_kdelta1, size_t n_kdelta, size_t koff, size_t klen,
                                  const uint64_t* alpha1, const uint64_t* beta1, const uint64_t* beta2, const uint64_t* delta1,
                                  const uint64_t* delta2, zkb_groth16_pk** out);
void zkb_groth16_pk_free(zkb_groth16_pk* pk);
/* Window sharding for multi-GPU proving: the key holds the FULL vectors (create with zkb_groth16_pk_create) and every MSM of
 * zkb_groth16_partial covers only window shard `rank` of `world`.  Scales better than point slices (the digit sort and the
 * bucket reduction shrink too); costs the whole key per GPU (320 MiB at 2^20 BN254). */
int zkb_groth16_pk_set_window_shard(zkb_groth16_pk* pk, uint32_t rank, uint32_t world);
/* ... except, when enabled, for the MSM over the private witness ([K w], protocol.py:151-155), which then covers the windows
 * [first, first + count) of its scalars (zkb_groth16_pk_msm_info(pk, 3, ...) says how many there are; count may be 0).  That MSM
 * needs the witness only, so the ranks of a multi-GPU proof that run none of the quotient's transforms take the windows of the
 * ranks that do (zksnake_b200/dist.py:kw_windows); over all ranks every window must be covered exactly once. */
int zkb_groth16_pk_set_kw_windows(zkb_groth16_pk* pk, uint32_t first, uint32_t count, int enable);
/* Optional, host only: start computing the multiples of delta_1 / delta_2 that depend on (r, s) alone on host threads, so that
 * they overlap the GPU work of zkb_groth16_partial; zkb_groth16_assemble picks them up when called with the same r, s (and
 * computes them itself otherwise).  The single-call provers do this internally. */
int zkb_groth16_precompute(zkb_groth16_pk* pk, const uint64_t r[4], const uint64_t s[4]);
/* Window size c and window count W the MSM over key vector `which` (0 tau_1, 1 tau_2, 2 target_1, 3 kdelta_1) runs with --
 * the executed work is W mixed additions per point (bench.py reports it beside the canonical 16-window count). */
int zkb_groth16_pk_msm_info(const zkb_groth16_pk* pk, int which, uint32_t* window_bits, uint32_t* windows);
/* Build fixed-base tables for the key's four point vectors (see zkb_msm_table_create); every later prove uses them. */
int zkb_groth16_pk_build_tables(zkb_groth16_pk* pk, uint32_t world);
/* Groth16.prove from host buffers: a, b, c = A.w, B.w, C.w (n each), priv = private witness (n_kdelta scalars), r, s = the
 * prover's randomness.  Outputs canonical affine A (G1), B (G2), C (G1) and their infinity flags.  The timed e2e region of
 * bench.py is exactly one call of this function. */
int zkb_groth16_prove(zkb_groth16_pk* pk, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* priv,
                      const uint64_t r[4], const uint64_t s[4], uint64_t* out_a, uint64_t* out_b, uint64_t* out_c,
                      int out_inf[3]);
/* same with the four input vectors already resident on the device */
int zkb_groth16_prove_dev(zkb_groth16_pk* pk, const void* d_a, const void* d_b, const void* d_c, const void* d_priv,
                          const uint64_t r[4], const uint64_t s[4], uint64_t* out_a, uint64_t* out_b, uint64_t* out_c,
                          int out_inf[3]);
/* Device-resident R1CS (A, B, C in CSR form: row_ptr has n_rows+1 uint64 entries, col uint32, val canonical Fr), replacing the
 * pure-Python SparseArray.dot of python/zksnake/array.py:36-43 as called from qap.py:53-55.  n_rows <= 2^log_n of the key it is
 * used with; n_cols = witness length m. */
typedef struct zkb_r1cs zkb_r1cs;
int zkb_r1cs_create(int curve, size_t n_rows, size_t n_cols, const uint64_t* const row_ptr[3], const uint32_t* const col[3],
                    const uint64_t* const val[3], zkb_r1cs** out);
void zkb_r1cs_free(zkb_r1cs* r1cs);
/* a = A.w etc. into host buffers (n_out elements each, rows >= n_rows are zero) */
int zkb_r1cs_eval(zkb_r1cs* r1cs, const uint64_t* witness, size_t n_out, uint64_t* a, uint64_t* b, uint64_t* c);
/* the same with the vector and the three products resident on the device (asynchronous).  With the TRANSPOSED matrices and the
 * Lagrange coefficients L_i(tau) as the vector this is the L / R / O loop of Groth16.setup (groth16/protocol.py:64-77). */
int zkb_r1cs_eval_dev(zkb_r1cs* r1cs, const void* d_witness, size_t n_out, void* d_a, void* d_b, void* d_c);
/* Groth16.prove(public + private witness) end to end: witness (m = n_cols canonical scalars, public part first, n_public of
 * them) is the ONLY per-proof host->device traffic; outputs as zkb_groth16_prove. */
int zkb_groth16_prove_witness(zkb_groth16_pk* pk, zkb_r1cs* r1cs, const uint64_t* witness, size_t n_public,
                              const uint64_t r[4], const uint64_t s[4], uint64_t* out_a, uint64_t* out_b, uint64_t* out_c,
                              int out_inf[3]);
/* same with the witness already resident on the device (bench.py's device-resident `value`) */
int zkb_groth16_prove_witness_dev(zkb_groth16_pk* pk, zkb_r1cs* r1cs, const void* d_witness, size_t n_public,
                                  const uint64_t r[4], const uint64_t s[4], uint64_t* out_a, uint64_t* out_b, uint64_t* out_c,
                                  int out_inf[3]);
/* Multi-GPU split of Groth16.prove: every rank evaluates the witness polynomials (SpMV + quotient) and runs the five MSMs of
 * protocol.py:133-155 over ITS key slice (zkb_groth16_partial: msm_xy = 5 x 24 uint64, canonical affine, G2 in slot 2;
 * msm_inf = 5 flags); the per-rank partial sums are exchanged (a few hundred bytes, NCCL all-gather in zksnake_b200/dist.py),
 * added with zkb_point_lincomb and turned into the proof by zkb_groth16_assemble (protocol.py:133-165). */
int zkb_groth16_partial(zkb_groth16_pk* pk, zkb_r1cs* r1cs, const void* witness, int witness_on_device, size_t n_public,
                        uint64_t* msm_xy, int* msm_inf);
/* zkb_groth16_partial cut where the ranks exchange data, so that the three independent interpolation -> coset-evaluation chains of
 * QAP.evaluate_witness (/root/reference/python/zksnake/groth16/qap.py:57-63: the A, B and C transforms) run on DIFFERENT GPUs and
 * the quotient's last step (qap.py:64-69) on ONE of them, instead of everything on every GPU.
 *   begin:    witness -> A.w, B.w, C.w and the satisfiability check on every rank (the SpMVs are cheap and the check needs all
 *             three), then the chains whose bit is set in chain_mask (bit 0 U, 1 V, 2 W): coefficients to d_coeffs + c * n,
 *             evaluations on the coset to d_evals + c * n (caller-owned device buffers of 3 n Fr elements each, n = the key's domain
 *             size).  A rank with chain_mask == 0 starts its [K w] MSM instead (it needs the witness only).
 *   exchange: (caller; zksnake_b200/dist.py:exchange_chains does it with NCCL) coefficient vectors 0 and 1 broadcast from their owners
 *             on the library stream (zkb_stream); evaluation vectors sent to the quotient rank.
 *   quotient: on the quotient rank only, on the library stream: H = coset-iNTT((eU eV - eW) / Z) from d_evals into d_h (n elements).
 *   exchange: d_h broadcast from the quotient rank -- on any stream; h_ready_event (a cudaEvent_t recorded behind that broadcast, or
 *             NULL if it ran on the library stream) tells finish when d_h is complete.
 *   finish:   this rank's share of the MSMs over U, V and the witness, then -- behind h_ready_event -- of [H Z]; msm_xy / msm_inf as
 *             zkb_groth16_partial returns them.
 * Every value is the same field element as in the one-call path, so the proof bytes do not change. */
int zkb_groth16_spread_begin(zkb_groth16_pk* pk, zkb_r1cs* r1cs, const void* witness, int witness_on_device, size_t n_public,
                             unsigned chain_mask, void* d_coeffs, void* d_evals);
int zkb_groth16_spread_quotient(zkb_groth16_pk* pk, void* d_evals, void* d_h);
int zkb_groth16_spread_finish(zkb_groth16_pk* pk, zkb_r1cs* r1cs, size_t n_public, const void* d_coeffs, const void* d_h,
                              void* h_ready_event, uint64_t* msm_xy, int* msm_inf);
/* msm_inf[3] may carry bit 1 (value 2): the rank was given (r, s) beforehand (zkb_groth16_precompute on a sharded key) and its HZ
 * slot is [H Z]_i + s [U]_i + r [V]_i -- C is linear in the partial sums, so the two scalar multiplications of protocol.py:157-160
 * are done per rank under its GPU work and only point additions follow the exchange.  Either every rank's slot carries the bit or
 * none does. */
int zkb_groth16_assemble(zkb_groth16_pk* pk, const uint64_t* msm_xy, const int* msm_inf, const uint64_t r[4],
                         const uint64_t s[4], uint64_t* out_a, uint64_t* out_b, uint64_t* out_c, int out_inf[3]);
/* The last two steps in one call: all_xy = world x 5 x 24 uint64 (every rank's msm_xy, rank-major), all_inf = world x 5 flags; the
 * slot sums over the ranks are formed on host threads and assembled (protocol.py:133-165). */
int zkb_groth16_assemble_partials(zkb_groth16_pk* pk, int world, const uint64_t* all_xy, const int* all_inf, const uint64_t r[4],
                                  const uint64_t s[4], uint64_t* out_a, uint64_t* out_b, uint64_t* out_c, int out_inf[3]);
/* intermediate results of the last prove on this key, for parity tests: which = 0 U, 1 V, 2 H (n coefficients each) */
int zkb_groth16_last_poly(zkb_groth16_pk* pk, int which, uint64_t* out);
/* the five raw MSM results of the last prove (A, B1, B2, HZ, KW) as affine canonical points + infinity flags */
int zkb_groth16_last_msm(zkb_groth16_pk* pk, int which, uint64_t* out_xy, int* out_inf);

/* ---- self test hooks (used by tests/ to compare host and device arithmetic paths) ------------------------- */
/* field: 0 FrBN254, 1 FqBN254, 2 FrBLS381, 3 FqBLS381.  op: 0 mul, 1 add, 2 sub, 3 inv(a), 4 neg(a).  Canonical in/out. */
int zkb_test_field_op_host(int field, int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out);
int zkb_test_field_op_dev(int field, int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out);

/* sum_i k_i P_i over a handful of canonical affine points, on the host (has_scalar[i] == 0: k_i = 1).  Backs the
 * PointG1/PointG2 operators of src/bn254/curve.rs:74-118, :249-292 (+ - neg * int); needs no GPU. */
int zkb_point_lincomb(int curve, int group, int n_terms, const uint64_t* points, const int* infs,
                          const uint64_t* scalars, const int* has_scalar, uint64_t* out_xy, int* out_inf);

#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
