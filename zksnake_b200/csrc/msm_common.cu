// msm_common.cu -- (curve, group) dispatch for the MSM and point-vector entry points.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include "zkb_internal.h"

namespace zkb {

int g_msm_c = 0, g_msm_seg = 0, g_msm_kchunk = 0;
void msm_set_tuning(int c, int seg, int kchunk) {
  g_msm_c = c;
  g_msm_seg = seg;
  g_msm_kchunk = kchunk;
}

#define DECL(SUFFIX)                                                                      \
  int msm_need_##SUFFIX(size_t n, uint32_t wr, uint32_t ww, uint32_t tc, size_t tn, size_t* need);                     \
  int msm_phase1_##SUFFIX(const void* p, const void* s, size_t n, uint32_t wr, uint32_t ww, uint32_t tc, size_t tn,     \
                          const MsmTicket* share, MsmTicket* tk);                                                        \
  int msm_sort_##SUFFIX(const void* s, size_t n, uint32_t wr, uint32_t ww, uint32_t tc, size_t tn, const MsmTicket* share,  \
                        MsmTicket* tk, void* stream);                                                                    \
  int msm_accum_##SUFFIX(const void* p, MsmTicket* tk, int leave_room);                                                 \
  int msm_phase2_##SUFFIX(MsmTicket* tk, void* stream);                                                                 \
  int msm_table_##SUFFIX(const void* pts, size_t n, uint32_t world, uint32_t* c, uint32_t* W, void* table);             \
  int points_conv_##SUFFIX(int to, size_t n, void* p);                                   \
  int batch_mul_##SUFFIX(const void* b, int single, const void* s, size_t n, void* o);
DECL(g1bn) DECL(g2bn) DECL(g1bls) DECL(g2bls)

#define DISPATCH(CALL_BN1, CALL_BN2, CALL_BL1, CALL_BL2)                  \
  if (curve == ZKB_BN254 && group == 1) return CALL_BN1;                  \
  if (curve == ZKB_BN254 && group == 2) return CALL_BN2;                  \
  if (curve == ZKB_BLS12_381 && group == 1) return CALL_BL1;              \
  if (curve == ZKB_BLS12_381 && group == 2) return CALL_BL2;              \
  return set_error(ZKB_ERR_ARG, "unknown (curve, group)");

// Ranks of a multi-GPU proof share one host: five finishing threads per rank spin in cudaEventSynchronize (the default for events)
// while other ranks' main threads may still be enqueueing kernels.  ZKB_BLOCKING_EVENTS=1 creates the ticket events of window
// shards with cudaEventBlockingSync, so that the waiters sleep.  OFF by default -- measured on the B200 boxes (32 host cores): 8
// ranks 6.58 ms either way, 2 ranks 13.87 ms spinning against 14.27 ms sleeping (the wake-up latency lands on the proof's tail,
// profiles/R3g_n2_*.json); for hosts with fewer cores than 6 x ranks.
static bool g_blocking_events = false;
int ticket_reserve(MsmTicket* tk, size_t bytes) {
  if (tk->event && tk->event_blocking != g_blocking_events) {
    cudaEventDestroy((cudaEvent_t)tk->event);
    tk->event = nullptr;
  }
  if (!tk->event) {
    cudaEvent_t e;
    ZKB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | (g_blocking_events ? cudaEventBlockingSync : 0)));
    tk->event = (void*)e;
    tk->event_blocking = g_blocking_events;
  }
  if (bytes > tk->host_cap) {
    if (tk->host) ZKB_CUDA(cudaFreeHost(tk->host));
    tk->host = nullptr;
    tk->host_cap = 0;
    size_t cap = bytes + 4096;
    ZKB_CUDA(cudaHostAlloc((void**)&tk->host, cap, cudaHostAllocDefault));
    tk->host_cap = cap;
  }
  return ZKB_OK;
}
void ticket_release(MsmTicket* tk) {
  if (tk->host) cudaFreeHost(tk->host);
  if (tk->event) cudaEventDestroy((cudaEvent_t)tk->event);
  if (tk->sort_event) cudaEventDestroy((cudaEvent_t)tk->sort_event);
  tk->sort_event = nullptr;
  tk->presorted = false;
  tk->host = nullptr;
  tk->host_cap = 0;
  tk->event = nullptr;
}
static int msm_need(int curve, const MsmJob& j, uint32_t wr, uint32_t ww, size_t* need) {
  const int group = j.group;
  wr = job_wrank(j, wr), ww = job_wworld(j, ww);
  DISPATCH(msm_need_g1bn(j.n, wr, ww, j.table_c, j.table_n, need), msm_need_g2bn(j.n, wr, ww, j.table_c, j.table_n, need),
           msm_need_g1bls(j.n, wr, ww, j.table_c, j.table_n, need), msm_need_g2bls(j.n, wr, ww, j.table_c, j.table_n, need))
}
static int msm_phase1(int curve, const MsmJob& j, uint32_t wr, uint32_t ww, const MsmTicket* share, MsmTicket* tk) {
  const int group = j.group;
  wr = job_wrank(j, wr), ww = job_wworld(j, ww);
  static const bool blocking_ok = [] { const char* e = getenv("ZKB_BLOCKING_EVENTS"); return e && atoi(e) != 0; }();
  if (ww != 1 && blocking_ok) g_blocking_events = true;   // a window shard: this process is one of several on the host (ticket_reserve)
  DISPATCH(msm_phase1_g1bn(j.points, j.scalars, j.n, wr, ww, j.table_c, j.table_n, share, tk),
           msm_phase1_g2bn(j.points, j.scalars, j.n, wr, ww, j.table_c, j.table_n, share, tk),
           msm_phase1_g1bls(j.points, j.scalars, j.n, wr, ww, j.table_c, j.table_n, share, tk),
           msm_phase1_g2bls(j.points, j.scalars, j.n, wr, ww, j.table_c, j.table_n, share, tk))
}
static int msm_sort(int curve, const MsmJob& j, uint32_t wr, uint32_t ww, const MsmTicket* share, MsmTicket* tk, void* stream) {
  const int group = j.group;
  wr = job_wrank(j, wr), ww = job_wworld(j, ww);
  static const bool blocking_ok = [] { const char* e = getenv("ZKB_BLOCKING_EVENTS"); return e && atoi(e) != 0; }();
  if (ww != 1 && blocking_ok) g_blocking_events = true;   // a window shard: this process is one of several on the host (ticket_reserve)
  DISPATCH(msm_sort_g1bn(j.scalars, j.n, wr, ww, j.table_c, j.table_n, share, tk, stream),
           msm_sort_g2bn(j.scalars, j.n, wr, ww, j.table_c, j.table_n, share, tk, stream),
           msm_sort_g1bls(j.scalars, j.n, wr, ww, j.table_c, j.table_n, share, tk, stream),
           msm_sort_g2bls(j.scalars, j.n, wr, ww, j.table_c, j.table_n, share, tk, stream))
}
static int msm_accum(int curve, const MsmJob& j, MsmTicket* tk, int leave_room) {
  const int group = j.group;
  DISPATCH(msm_accum_g1bn(j.points, tk, leave_room), msm_accum_g2bn(j.points, tk, leave_room),
           msm_accum_g1bls(j.points, tk, leave_room), msm_accum_g2bls(j.points, tk, leave_room))
}
int msm_table_build(int curve, int group, const void* p, size_t n, uint32_t world, uint32_t* c, uint32_t* W, void* t) {
  DISPATCH(msm_table_g1bn(p, n, world, c, W, t), msm_table_g2bn(p, n, world, c, W, t), msm_table_g1bls(p, n, world, c, W, t),
           msm_table_g2bls(p, n, world, c, W, t))
}
static int msm_phase2(MsmTicket* tk, void* stream) {
  const int curve = tk->curve, group = tk->group;
  DISPATCH(msm_phase2_g1bn(tk, stream), msm_phase2_g2bn(tk, stream), msm_phase2_g1bls(tk, stream), msm_phase2_g2bls(tk, stream))
}
int msm_enqueue(int curve, const MsmJob& job, uint32_t wr, uint32_t ww, MsmTicket* tk) {
  size_t need = 0;
  int rc;
  if ((rc = msm_need(curve, job, wr, ww, &need))) return rc;
  if ((rc = scratch_reserve(need))) return rc;
  scratch_reset();
  if ((rc = msm_phase1(curve, job, wr, ww, nullptr, tk))) return rc;
  prof_begin(PROF_MSM_REDUCE);
  rc = msm_phase2(tk, ctx_stream());
  prof_end(PROF_MSM_REDUCE);
  return rc;
}
int msm_batch_need(int curve, const MsmJob* jobs, int njobs, uint32_t wr, uint32_t ww, size_t* total) {
  *total = 0;
  int rc;
  for (int i = 0; i < njobs; i++) {
    size_t need = 0;
    if ((rc = msm_need(curve, jobs[i], wr, ww, &need))) return rc;
    *total += need + 8192;
  }
  return ZKB_OK;
}
int msm_presort(int curve, const MsmJob& job, uint32_t wr, uint32_t ww, MsmTicket* tk, void* stream) {
  tk->presorted = false;
  int rc = msm_sort(curve, job, wr, ww, nullptr, tk, stream);
  if (rc) return rc;
  if (!tk->sort_event) {
    cudaEvent_t e;
    ZKB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    tk->sort_event = (void*)e;
  }
  ZKB_CUDA(cudaEventRecord((cudaEvent_t)tk->sort_event, (cudaStream_t)stream));
  tk->presorted = true;
  return ZKB_OK;
}
void msm_presort_cancel(MsmTicket* tk) {
  if (tk->presorted && tk->sort_event) cudaStreamWaitEvent((cudaStream_t)ctx_stream(), (cudaEvent_t)tk->sort_event, 0);
  tk->presorted = false;
}
int msm_enqueue_batch(int curve, const MsmJob* jobs, int njobs, uint32_t wr, uint32_t ww, MsmTicket* tickets, bool join,
                      int side_base) {
  if (njobs > 8 || side_base < 0) return set_error(ZKB_ERR_ARG, "msm batch: at most 8 jobs");
  size_t total = 0;
  int rc;
  // a ticket whose digit sort was enqueued ahead of this call (msm_presort) must be for exactly this job; it took its scratch then
  for (int i = 0; i < njobs; i++)
    if (tickets[i].presorted && !(tickets[i].sorted.scalars == jobs[i].scalars && tickets[i].sorted.n == jobs[i].n) && !tickets[i].empty)
      return set_error(ZKB_ERR_ARG, "msm batch: presorted ticket does not match its job");
  for (int i = 0; i < njobs; i++) {
    if (tickets[i].presorted) continue;
    size_t need = 0;
    if ((rc = msm_batch_need(curve, jobs + i, 1, wr, ww, &need))) return rc;
    total += need;
  }
  if ((rc = scratch_reserve(total))) return rc;
  scratch_reset();
  cudaStream_t main_st = (cudaStream_t)ctx_stream();
  // fork after every phase 1: the job's reduction runs on its own side stream while the library stream goes on with the next
  // job's sort and accumulation (the accumulation is a work-stealing persistent grid, so it tolerates the few CTAs slots the
  // short reduction kernels borrow); join at the end so later library work is ordered after all of them
  static cudaEvent_t fork_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool forked[8] = {false, false, false, false, false, false, false, false};
  // Sorts ahead (experiment, ZKB_SORT_AHEAD=1): when the batch opens with a G2 job, the digit sorts of the later jobs go to a side
  // stream under that job's accumulation instead of running one by one between the accumulations.  Measured on 2^20 BN254
  // (profiles/R2p_*, R2q_*): the persistent accumulation grid owns every register of every SM, so the sorts only get the scraps --
  // 23.65 -> 23.13 ms, and with one CTA slot per SM kept free for them (ZKB_G2_ROOM_CTAS=3) the accumulation itself loses more than
  // the sorts gain.  The sorts that can start EARLY (api.cu: under the transforms) do better; that is the default, this is off.
  static const bool sort_ahead_on = [] { const char* e = getenv("ZKB_SORT_AHEAD"); return e && atoi(e) != 0; }();
  const bool ahead = sort_ahead_on && njobs > 2 && jobs[0].group == 2;
  static cudaEvent_t sort_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  static cudaEvent_t batch_ev = nullptr;
  bool sorted_ahead[8] = {false, false, false, false, false, false, false, false};
  auto share_of = [&](int i) {   // an earlier job over the same scalars (the sort stage checks that the geometry matches too)
    const MsmTicket* share = nullptr;
    for (int j = 0; j < i && !share; j++)
      if (!tickets[j].empty && jobs[j].scalars == jobs[i].scalars && jobs[j].n == jobs[i].n) share = &tickets[j];
    return share;
  };
  // on a mid-batch failure the reductions already forked keep running on their side streams over the shared scratch arena:
  // order the library stream after them before handing the error back, so that the caller's next scratch epoch cannot race them
  auto fail = [&](int code) {
    for (int j = 0; j < njobs; j++)
      if (forked[j] && tickets[j].event) cudaStreamWaitEvent(main_st, (cudaEvent_t)tickets[j].event, 0);
    for (int j = 0; j < njobs; j++)
      if (sorted_ahead[j] && sort_ev[j]) cudaStreamWaitEvent(main_st, sort_ev[j], 0);
    for (int j = 0; j < njobs; j++) msm_presort_cancel(&tickets[j]);
    return code;
  };
  for (int i = 0; i < njobs; i++) {
    bool accumulate_only = false;
    if (ahead && i == 0) {
      // job 0: sort on the library stream (unless presorted); the other jobs' sorts -- those that are neither presorted nor able
      // to share one -- queue up on the sort stream behind the point where the library stream stands then (their scalars are
      // complete there, and they start when job 0's sort has finished, together with job 0's accumulation: two sorts at once would
      // only slow the one the accumulation is waiting for)
      if (!tickets[0].presorted && (rc = msm_sort(curve, jobs[0], wr, ww, nullptr, &tickets[0], main_st))) return fail(rc);
      cudaStream_t sort_st = (cudaStream_t)ctx_side_stream(7);
      cudaError_t ce = cudaSuccess;
      if (!batch_ev) ce = cudaEventCreateWithFlags(&batch_ev, cudaEventDisableTiming);
      if (ce != cudaSuccess || !sort_st) return fail(set_error(ZKB_ERR_CUDA, "msm batch: cannot create the sort stream"));
      if ((ce = cudaEventRecord(batch_ev, main_st)) != cudaSuccess || (ce = cudaStreamWaitEvent(sort_st, batch_ev, 0)) != cudaSuccess)
        return fail(cuda_fail((int)ce, "msm batch sort fork", __FILE__, __LINE__));
      for (int k = 1; k < njobs; k++) {
        bool shares = tickets[k].presorted;
        for (int j = 0; j < k && !shares; j++) shares = jobs[j].scalars == jobs[k].scalars && jobs[j].n == jobs[k].n;
        if (shares) continue;
        if ((rc = msm_sort(curve, jobs[k], wr, ww, nullptr, &tickets[k], sort_st))) return fail(rc);
        if (!sort_ev[k]) ce = cudaEventCreateWithFlags(&sort_ev[k], cudaEventDisableTiming);
        if (ce != cudaSuccess || (ce = cudaEventRecord(sort_ev[k], sort_st)) != cudaSuccess)
          return fail(cuda_fail((int)ce, "msm batch sort event", __FILE__, __LINE__));
        sorted_ahead[k] = true;
      }
      accumulate_only = true;
    }
    if (tickets[i].presorted) {
      cudaError_t ce = cudaStreamWaitEvent(main_st, (cudaEvent_t)tickets[i].sort_event, 0);
      tickets[i].presorted = false;
      if (ce != cudaSuccess) return fail(cuda_fail((int)ce, "msm batch presort join", __FILE__, __LINE__));
      accumulate_only = true;
    } else if (sorted_ahead[i]) {
      cudaError_t ce = cudaStreamWaitEvent(main_st, sort_ev[i], 0);
      if (ce != cudaSuccess) return fail(cuda_fail((int)ce, "msm batch sort join", __FILE__, __LINE__));
      accumulate_only = true;
    }
    if (accumulate_only) {
      if ((rc = msm_accum(curve, jobs[i], &tickets[i], (ahead && i == 0) ? 1 : 0))) return fail(rc);
    } else {
      if ((rc = msm_phase1(curve, jobs[i], wr, ww, share_of(i), &tickets[i]))) return fail(rc);
    }
    if (tickets[i].empty) continue;
    cudaError_t ce = cudaSuccess;
    if (!fork_ev[i]) ce = cudaEventCreateWithFlags(&fork_ev[i], cudaEventDisableTiming);
    cudaStream_t side = (cudaStream_t)ctx_side_stream((side_base + i) % 7);   // (side stream 7 is the sort stream)
    if (ce != cudaSuccess || !side) return fail(set_error(ZKB_ERR_CUDA, "msm batch: cannot create a side stream"));
    if ((ce = cudaEventRecord(fork_ev[i], main_st)) != cudaSuccess || (ce = cudaStreamWaitEvent(side, fork_ev[i], 0)) != cudaSuccess)
      return fail(cuda_fail((int)ce, "msm batch fork", __FILE__, __LINE__));
    rc = msm_phase2(&tickets[i], side);
    forked[i] = true;   // (whatever phase 2 enqueued before failing is on the side stream; its event may not be recorded yet)
    if (rc) {
      cudaEventRecord((cudaEvent_t)tickets[i].event, side);
      return fail(rc);
    }
  }
  if (!join) return ZKB_OK;       // the caller enqueues more work first and joins the tickets' events itself
  prof_begin(PROF_MSM_REDUCE);   // what is left of the reductions after the last accumulation
  for (int i = 0; i < njobs; i++)
    if (!tickets[i].empty) ZKB_CUDA(cudaStreamWaitEvent(main_st, (cudaEvent_t)tickets[i].event, 0));
  prof_end(PROF_MSM_REDUCE);
  return ZKB_OK;
}
int msm_finish(MsmTicket* tk, uint64_t* out_xy, int* out_inf) {
  if (tk->empty) {
    memset(out_xy, 0, affine_bytes(tk->curve, tk->group));
    *out_inf = 1;
    return ZKB_OK;
  }
  ZKB_CUDA(cudaEventSynchronize((cudaEvent_t)tk->event));
  host_msm_finish(tk->curve, tk->group, tk->host, tk->nwin, tk->win0, tk->c, tk->nlev, tk->logk, tk->parts, tk->nbits, out_xy, out_inf);
  tk->empty = true;
  return ZKB_OK;
}
int msm_dev(int curve, int group, const void* p, const void* s, size_t n, uint64_t* o, int* inf) {
  static MsmTicket tk;
  MsmJob job = {group, p, s, n, 0, 0};
  int rc = msm_enqueue(curve, job, 0, 1, &tk);
  if (rc) return rc;
  return msm_finish(&tk, o, inf);
}
int points_to_mont_dev(int curve, int group, size_t n, void* p) {
  DISPATCH(points_conv_g1bn(1, n, p), points_conv_g2bn(1, n, p), points_conv_g1bls(1, n, p), points_conv_g2bls(1, n, p))
}
int points_from_mont_dev(int curve, int group, size_t n, void* p) {
  DISPATCH(points_conv_g1bn(0, n, p), points_conv_g2bn(0, n, p), points_conv_g1bls(0, n, p), points_conv_g2bls(0, n, p))
}
int batch_mul_dev(int curve, int group, const void* b, int single, const void* s, size_t n, void* o) {
  DISPATCH(batch_mul_g1bn(b, single, s, n, o), batch_mul_g2bn(b, single, s, n, o), batch_mul_g1bls(b, single, s, n, o),
           batch_mul_g2bls(b, single, s, n, o))
}

}  // namespace zkb
