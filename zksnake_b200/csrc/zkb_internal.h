// zkb_internal.h -- internal (non-ABI) interfaces shared by the translation units of libzkb200.so.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string>

#define ZKB_BN254 0
#define ZKB_BLS12_381 1

#define ZKB_OK 0
#define ZKB_ERR_CUDA (-1)
#define ZKB_ERR_ARG (-2)
#define ZKB_ERR_MISMATCH (-3)     // "Number of points and scalars mismatch"  (curve.rs:369-371)
#define ZKB_ERR_DOMAIN (-4)       // "Domain size is too large"              (polynomial.rs:638-639)
#define ZKB_ERR_NOT_DIVISIBLE (-5)  // "(U * V - W) did not divided by Z to zero" (qap.py:68-69)
#define ZKB_ERR_NOINIT (-6)
#define ZKB_ERR_POINT (-7)        // "Cannot deserialize point" (ecc.py:128-142 -> curve.rs:134-141)

namespace zkb {

// sizes in bytes of one element for (curve, group)
inline size_t fq_bytes(int curve) { return curve == ZKB_BN254 ? 32 : 48; }
inline size_t affine_bytes(int curve, int group) { return fq_bytes(curve) * 2 * (group == 2 ? 2 : 1); }
inline size_t xyzz_bytes(int curve, int group) { return affine_bytes(curve, group) * 2; }

int set_error(int code, const std::string& msg);
int cuda_fail(int cuda_err, const char* what, const char* file, int line);
#define ZKB_CUDA(x)                                                           \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) return zkb::cuda_fail((int)e_, #x, __FILE__, __LINE__); \
  } while (0)

// One host thread drives the library (include/zkb200.h, "Conventions"): entry points share one stream, one scratch arena and
// per-call result tickets.  EntryGuard makes that contract checked instead of assumed: the first thread inside an entry point
// owns the library until it leaves (nested entry-point calls of the same thread are fine); a second thread gets ZKB_ERR_ARG.
struct EntryGuard {
  bool ok;
  EntryGuard();
  ~EntryGuard();
};
#define ZKB_ENTRY_GUARD()     \
  zkb::EntryGuard guard_;      \
  if (!guard_.ok) return zkb::set_error(ZKB_ERR_ARG, "libzkb200 entry points are not re-entrant: another host thread is inside the library")

void* ctx_stream();                 // cudaStream_t
void* ctx_stream_swap(void* s);     // make `s` the stream every helper enqueues on; returns the previous one (core.cu)
void* ctx_side_stream(int i);       // cudaStream_t, i < 8 (created on first use); nullptr on failure
bool ctx_ready();
// grow-only device scratch arena; `scratch_reset` starts a new allocation epoch (pointers from the previous epoch die)
int scratch_reserve(size_t bytes);  // make sure the arena holds at least `bytes` (may reallocate; sync)
void scratch_reset();
void* scratch_take(size_t bytes);   // 256-byte aligned bump allocation; nullptr when exhausted
// one epoch shared by several stages: reserve `total`, reset once, then the stages' own reserve / reset calls only check / do nothing
int scratch_hold_begin(size_t total);
void scratch_hold_end();
void count_launch(int n = 1);
void count_h2d(size_t bytes);   // host<->device traffic issued by the library (zkb_transfer_count)
void count_d2h(size_t bytes);
// counted copies on the library stream (the translation unit provides S())
#define ZKB_H2D(dst, src, bytes) (zkb::count_h2d(bytes), cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, S()))
#define ZKB_D2H(dst, src, bytes) (zkb::count_d2h(bytes), cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, S()))
// device-time accounting per kernel family (zkb_prof_enable / zkb_prof_read); no-ops unless enabled
enum { PROF_NTT = 0, PROF_MSM_SORT = 1, PROF_MSM_ACCUM_G1 = 2, PROF_MSM_ACCUM_G2 = 3, PROF_MSM_REDUCE = 4, PROF_SPMV = 5,
       PROF_VEC = 6, PROF_MISC = 7, PROF_NTAGS = 8 };
void prof_begin(int tag);
void prof_end(int tag);
unsigned long long launches();

// ---- NTT (ntt_host.cu) ----
// coset: 0 none | 1 reference coset (offset = group generator of the same domain, polynomial.rs:553-556, 579-582)
//        | 2 coset g<w> with g the multiplicative generator of Fr (5 for BN254, 7 for BLS12-381): quotient computations
int ntt_dev(int curve, int inverse, int coset, uint32_t log_n, const void* d_in, size_t in_len, void* d_out);
int vec_op_dev(int curve, int op, size_t n, const void* a, size_t na, const void* b, size_t nb, const void* c, void* out);
int fr_reduce_dev(int curve, size_t n, void* v);
int fr_powers_dev(int curve, const uint64_t* base, const uint64_t* scale, size_t n, void* d_out);
// H = (U*V - W)/Z from the evaluation vectors a,b,c (canonical, n = 2^log_n each, device).  Writes U,V (n coeffs each, the
// MSM scalars) and H (n coeffs, top one zero).  d_w is scratch for W.  Returns ZKB_ERR_NOT_DIVISIBLE when a.b != c.
// after_interp(arg), when given, is called once the three interpolations are enqueued: U and V (the MSM scalars) are final on the
// library stream from there on, while four more transforms follow -- the caller forks the digit sorts of the U / V MSMs there.
// check: 0 none | 1 test a.b == c and wait for the answer | 2 test it, leave the answer in *groth16_flag_host() (pinned; non-zero =
// not satisfied) and do not wait: valid once any later work of the library stream is known to have completed.
int groth16_h_dev(int curve, uint32_t log_n, const void* d_a, const void* d_b, const void* d_c, void* d_u, void* d_v,
                  void* d_w, void* d_h, int check, void (*after_interp)(void*) = nullptr, void* arg = nullptr);
int* groth16_flag_host();
// the same pipeline step by step (multi-GPU chain spreading): witness check, one interpolation -> coset-evaluation chain
// (which: 0 U, 1 V, 2 W), and H from the three coset evaluation vectors; d_tmp holds 2^log_n elements
int groth16_check_dev(int curve, uint32_t log_n, const void* d_a, const void* d_b, const void* d_c, int* d_flag);
int groth16_chain_dev(int curve, uint32_t log_n, int which, const void* d_in, void* d_coeff, void* d_eval, void* d_tmp);
int groth16_hfin_dev(int curve, uint32_t log_n, void* d_eu, const void* d_ev, const void* d_ew, void* d_h, void* d_tmp);
size_t ntt_scratch_bytes(uint32_t log_n);
// long_rows: the n_long rows with more than SPMV_LONG_ROW non-zeros (device array, listed when the matrix is created); they are
// summed by whole CTAs into long_partial (n_long * 64 elements) instead of by one thread each
#define SPMV_LONG_ROW 64
int spmv_dev(int curve, size_t n_out, size_t n_rows, const void* row_ptr, const void* col, const void* val, const void* w,
             void* out, const uint32_t* long_rows, uint32_t n_long, void* long_partial);

// window size / window count the plain (table-less) MSM heuristic picks for n points (msm_common.cu)
void msm_plan_info(size_t n, uint32_t scalar_bits, uint32_t wworld, uint32_t* c, uint32_t* W);
void msm_kernel_info(int curve, int group, int* lanes_per_point, int* ctas_per_sm);

// ---- point codec (codec.cu) ----
size_t compressed_bytes(int curve, int group);
int points_compress_dev(int curve, int group, const void* d_pts, size_t n, void* d_out);
int points_decompress_dev(int curve, int group, const void* d_in, size_t n, int validate, void* d_pts, unsigned long long* bad);

// ---- MSM (msm_host.cuh instantiations) ----
// d_points: affine, Montgomery form; d_scalars: canonical 4 x u64.  Result: affine canonical coordinates on the host.
// An MSM is enqueued on the library stream (no synchronisation) against a ticket that owns a pinned result buffer and an
// event; msm_finish waits for that event and recombines the per-window sums on the host -- so the host part of MSM i
// overlaps the kernels of MSM i+1.  A ticket can be reused after msm_finish.
// what one MSM's digit sort produced; another MSM of the same batch over the SAME scalars can reuse it
struct MsmSorted {
  const void* scalars;
  size_t n;
  uint32_t table_c;
  size_t table_n;
  uint32_t wrank, wworld;
  size_t nb;
  uint32_t krun;
  uint32_t *start, *pstart, *npieces, *refs, *run_bucket, *hot_list, *vhot_list, *counters;
};
struct MsmTicket {
  MsmSorted sorted = {};
  int curve = 0, group = 1;
  bool empty = true;
  uint32_t nwin = 0, win0 = 0, c = 0, nlev = 0, nbits = 0, logk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, parts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned char* host = nullptr;   // pinned
  size_t host_cap = 0;
  void* event = nullptr;           // cudaEvent_t
  bool event_blocking = false;     // created with cudaEventBlockingSync (msm_common.cu:ticket_reserve)
  bool presorted = false;          // msm_presort ran this ticket's digit sort on a side stream; sort_event marks its end
  void* sort_event = nullptr;      // cudaEvent_t
  alignas(16) unsigned char dev[512];   // MsmDev<X>: plan + device pointers between phase 1 and phase 2
};
int ticket_reserve(MsmTicket* tk, size_t bytes);
void ticket_release(MsmTicket* tk);
// wrank/wworld: window shard handled by this call (0/1 = the whole MSM); the result is then the partial sum over those windows.
// wworld = ZKB_WINDOW_RANGE | count selects the explicit window range [wrank, wrank + count) instead of an even share (count may be
// 0: the job is empty): ranks of a multi-GPU proof that have idle time take more windows of an MSM than the busy ones.
#define ZKB_WINDOW_RANGE 0x80000000u
struct MsmJob {
  int group;
  const void* points;     // n affine points, or a fixed-base table (table_n != 0)
  const void* scalars;
  size_t n;
  uint32_t table_c;       // window size the table was built for (0 = plain points)
  size_t table_n;         // points per window of the table
  uint32_t own_wrank, own_wworld;   // own_wworld != 0: this job's window shard, overriding the batch's
};
inline uint32_t job_wrank(const MsmJob& j, uint32_t batch_wrank) { return j.own_wworld ? j.own_wrank : batch_wrank; }
inline uint32_t job_wworld(const MsmJob& j, uint32_t batch_wworld) { return j.own_wworld ? j.own_wworld : batch_wworld; }
// Scratch bytes a batch of these jobs takes from the arena (what msm_enqueue_batch reserves).
int msm_batch_need(int curve, const MsmJob* jobs, int njobs, uint32_t wrank, uint32_t wworld, size_t* total);
// The digit sort of one job of a LATER msm_enqueue_batch call, enqueued on `stream` now (its scalars must be complete there); the
// arena must be held (scratch_hold_begin) from here to that call.  The batch then only waits for the ticket's sort_event.
int msm_presort(int curve, const MsmJob& job, uint32_t wrank, uint32_t wworld, MsmTicket* tk, void* stream);
// Orders the library stream behind a presorted ticket's sort and clears the flag (error paths: the sort may still be running).
void msm_presort_cancel(MsmTicket* tk);
// One MSM, both phases on the library stream.
int msm_enqueue(int curve, const MsmJob& job, uint32_t wrank, uint32_t wworld, MsmTicket* tk);
// fixed-base table (see msm_host.cuh:msm_table_run)
int msm_table_build(int curve, int group, const void* d_pts, size_t n, uint32_t world, uint32_t* c, uint32_t* W, void* d_table);
// A batch: phase 1 (sort + accumulate) of every job back to back on the library stream, then all phase 2s (the latency-bound
// folds / bucket reductions) concurrently on side streams, joined back into the library stream.
// join = false: the library stream is NOT ordered behind the reductions (the caller goes on enqueueing -- another batch, an exchange
// step -- and waits for the tickets' events later); side_base: first of the seven side streams the reductions use, so
// that consecutive unjoined batches do not queue their reductions behind each other.
int msm_enqueue_batch(int curve, const MsmJob* jobs, int njobs, uint32_t wrank, uint32_t wworld, MsmTicket* tickets,
                      bool join = true, int side_base = 0);
int msm_finish(MsmTicket* tk, uint64_t* out_xy, int* out_inf);
int msm_dev(int curve, int group, const void* d_points, const void* d_scalars, size_t n, uint64_t* out_xy, int* out_inf);
int points_to_mont_dev(int curve, int group, size_t n, void* d_points);
int points_from_mont_dev(int curve, int group, size_t n, void* d_points);
int batch_mul_dev(int curve, int group, const void* d_bases, int single_base, const void* d_scalars, size_t n, void* d_out);
void msm_set_tuning(int c, int seg, int kchunk);

// ---- host math (host_math.cpp, plain g++) ----
// recombine the per-window sums of the bucket reduction (XYZZ, Montgomery; per window: parts[0] partial sums of U_0, ...,
// parts[nlev-1] of U_{nlev-1}, then A_0..A_{nbits-1}, R_top):
//   window sum = R_top + U_0 + 2^logk[0] (U_1 + ... + 2^logk[nlev-1] (sum_beta 2^beta A_beta)),
// then Horner over the windows with c doublings each
// (the windows are win0 .. win0+nwin-1 of the scalar: the result is multiplied by 2^(c*win0))
void host_msm_finish(int curve, int group, const void* sums, uint32_t nwin, uint32_t win0, uint32_t c, uint32_t nlev,
                     const uint32_t* logk, const uint32_t* parts, uint32_t nbits, uint64_t* out_xy, int* out_inf);
// out = sum_i k_i * P_i + sum_j Q_j over a handful of canonical affine points (proof assembly)
void host_lincomb(int curve, int group, int n_terms, const uint64_t* const* points, const int* infs,
                  const uint64_t* const* scalars, uint64_t* out_xy, int* out_inf);
void host_fr_mul(int curve, const uint64_t* a, const uint64_t* b, uint64_t* out);

}  // namespace zkb
