// msm_host.cuh -- host driver of the Pippenger pipeline, instantiated once per (curve, group) in msm_inst_*.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "msm.cuh"
#include "msm_pair.cuh"
#include "zkb_internal.h"

namespace zkb {

extern int g_msm_c, g_msm_seg, g_msm_kchunk;   // tuning overrides (0 = heuristic), msm_common.cu

static inline cudaStream_t MS() { return (cudaStream_t)ctx_stream(); }

// the inline-multiplier twin of a stored field (identical layout; only the device code generation differs)
template <class P> struct InlineMul : P { static constexpr bool NOINLINE_MUL = false; };
template <class P> struct InlineMulCallSqr : P { static constexpr bool NOINLINE_MUL = false; static constexpr bool NOINLINE_SQR = true; };
#ifndef ZKB_G1_BN_BASE
#define ZKB_G1_BN_BASE InlineMul<FqBN254>
#endif
// PAIR: two lanes per point (msm_pair.cuh) -- the accumulation grid then runs 148 x MINB x 64 runs at a time
template <class F> struct AccumField { typedef F type; static constexpr int MINB = 2; static constexpr bool PAIR = false; };
// G1 BN254: the fully inlined XYZZ addition is ~9-15 % faster than calling the multiplier out of line (tools/ffbench.cu).
// G1 BLS12-381: inlined it is 7.2 K instructions (115 KB, far beyond the 32 KB L1.5 instruction cache) -- out of line at 4 CTAs
// per SM the 2^20-point launch goes 6.43 -> 6.08 ms and the proof 50.7 -> 48.0 ms (profiles/R2s_*).
// G2: two lanes per point, multiplier bodies out of line (msm_pair.cuh).
template <> struct AccumField<Fp<FqBN254>> { typedef Fp<ZKB_G1_BN_BASE> type; static constexpr int MINB = 4; static constexpr bool PAIR = false; };
#ifndef ZKB_G1_BLS_BASE
#define ZKB_G1_BLS_BASE FqBLS381
#endif
#ifndef ZKB_G1_BLS_MINB
#define ZKB_G1_BLS_MINB 4
#endif
template <> struct AccumField<Fp<FqBLS381>> { typedef Fp<ZKB_G1_BLS_BASE> type; static constexpr int MINB = ZKB_G1_BLS_MINB; static constexpr bool PAIR = false; };
// G2: lane pairs over the base field (PairBase = the multiplier flavour the pair kernel is generated with)
#ifndef ZKB_G2_PAIR
#define ZKB_G2_PAIR 1
#endif
#ifndef ZKB_G2_PAIR_BASE_BN
#define ZKB_G2_PAIR_BASE_BN FqBN254
#endif
#ifndef ZKB_G2_PAIR_BASE_BLS
#define ZKB_G2_PAIR_BASE_BLS FqBLS381
#endif
#ifndef ZKB_G2_PAIR_MINB_BN
#define ZKB_G2_PAIR_MINB_BN 4
#endif
#ifndef ZKB_G2_PAIR_MINB_BLS
#define ZKB_G2_PAIR_MINB_BLS 3
#endif
#if ZKB_G2_PAIR
template <> struct AccumField<Fp2<FqBN254>> {
  typedef Fp2<FqBN254> type; typedef ZKB_G2_PAIR_BASE_BN PairBase;
  static constexpr int MINB = ZKB_G2_PAIR_MINB_BN; static constexpr bool PAIR = true;
};
template <> struct AccumField<Fp2<FqBLS381>> {
  typedef Fp2<FqBLS381> type; typedef ZKB_G2_PAIR_BASE_BLS PairBase;
  static constexpr int MINB = ZKB_G2_PAIR_MINB_BLS; static constexpr bool PAIR = true;
};
#endif
template <class F> constexpr uint32_t accum_run_threads() { return 148u * AccumField<F>::MINB * (AccumField<F>::PAIR ? 64u : 128u); }

// wrank/wworld: this launch handles the windows [W*wrank/wworld, W*(wrank+1)/wworld) of every scalar (multi-GPU window
// sharding); 0/1 = all windows.
// table_c / table_n: non-zero when the points come from a fixed-base table built for window size table_c over table_n points.
inline int msm_choose_window(size_t n, uint32_t scalar_bits, uint32_t wworld, bool table) {
  if (wworld & ZKB_WINDOW_RANGE) wworld = 1;   // an explicit window range: plan as for the whole MSM
  uint32_t logn = 0;
  while (((size_t)1 << logn) < n) logn++;
  // cost model in units of one mixed addition: every window costs n additions; every bucket ~5 (the bucket reduction is ~2.3
  // full additions per bucket plus the piece folds); with a table there is ONE set of buckets, otherwise one per window; with
  // window sharding the slowest rank has ceil(W / world) windows, which favours a W that divides evenly.
  double best = 0;
  int c = 0;
  int lo = (int)logn - 7 < 3 ? 3 : (int)logn - 7, hi = (int)logn - (table ? 1 : 3) < 3 ? 3 : (int)logn - (table ? 1 : 3);
  if (hi > 22) hi = 22;
  if (lo > hi) lo = hi;
  for (int cc = lo; cc <= hi; cc++) {
    uint32_t W = (scalar_bits + 1 + cc - 1) / cc;
    uint32_t per_rank = (W + wworld - 1) / wworld;
    // (below 2^20 points the bucket reduction is latency-bound and partly hidden: measured optimum sits one or two windows
    // sizes higher than a cost of 5 per bucket predicts -- tools/perf_probe.py small, gpurun_out/small_r2g.txt)
    double buckets = ((!table && logn <= 19) ? 3.0 : 5.0) * (double)(1u << (cc - 1));
    double cost = table ? (double)per_rank * (double)n + buckets : (double)per_rank * ((double)n + buckets);
    if (c == 0 || cost < best) {
      best = cost;
      c = cc;
    }
  }
  return c;
}

// accum_threads: resident threads of the accumulation grid (148 SMs x MINB CTAs x 128)
inline MsmPlan msm_make_plan(size_t n, uint32_t scalar_bits, uint32_t wrank, uint32_t wworld, uint32_t table_c = 0,
                             size_t table_n = 0, uint32_t accum_threads = 148 * 4 * 128) {
  MsmPlan pl;
  memset(&pl, 0, sizeof(pl));
  uint32_t logn = 0;
  while (((size_t)1 << logn) < n) logn++;
  int c = table_c ? (int)table_c : (g_msm_c ? g_msm_c : msm_choose_window(n, scalar_bits, wworld, false));
  if (c < 3) c = 3;
  if (c > 22) c = 22;
  pl.c = (uint32_t)c;
  pl.nwin_total = (scalar_bits + 1 + pl.c - 1) / pl.c;
  if (wworld & ZKB_WINDOW_RANGE) {   // explicit range [wrank, wrank + count), clipped to the windows there are
    const uint32_t count = wworld & ~ZKB_WINDOW_RANGE;
    pl.win0 = wrank < pl.nwin_total ? wrank : pl.nwin_total;
    pl.nwin = pl.win0 + count <= pl.nwin_total ? count : pl.nwin_total - pl.win0;
  } else {
    pl.win0 = pl.nwin_total * wrank / wworld;
    pl.nwin = pl.nwin_total * (wrank + 1) / wworld - pl.win0;
  }
  pl.table_n = (uint32_t)table_n;
  pl.bwin = table_n ? (pl.nwin ? 1u : 0u) : pl.nwin;
  pl.nbuck = 1u << (pl.c - 1);
  // run length: a bucket with r references is cut into ~r/K + 1 pieces, and more than ZKB_MSM_HOT pieces send it to the
  // (lane-inefficient) warp fold -- keep the typical bucket at <= 3 pieces
  pl.krun = 32u;
  while (pl.krun < 256u && (table_n ? n * pl.nwin : n) / pl.nbuck > 2 * (size_t)pl.krun) pl.krun *= 2;
  // ... but never so long that the runs cannot fill the grid: a run is one thread's serial chain, so with fewer runs than
  // resident threads the accumulation time is set by the chain length (2^16 points: 0.80 -> 0.20 ms)
  {
    size_t per_thread = n * pl.nwin / accum_threads;
    uint32_t fill = 8;
    while (fill < 256u && 3 * (size_t)fill <= 2 * per_thread) fill *= 2;   // power of two nearest to refs / threads, >= 8
    if (pl.krun > fill) pl.krun = fill;
    // ... and trimmed so that the runs come out as a whole number of grid-wide rounds: work is handed out 32 runs at a time to
    // 148 x MINB x 4 warps, so with T references the kernel lasts ceil(T / (K P)) rounds of K additions; K = 32 at 2^20 points
    // is 6.06 rounds -- a seventh round that keeps 6 % of the warps busy.  Pick the round count k nearest to the heuristic K
    // and the smallest K that fits T into k rounds (2 % slack for the runs that straddle bucket boundaries).
    double pt = (double)n * pl.nwin / accum_threads;
    if (pt > 8.0) {
      uint32_t k = (uint32_t)(pt / pl.krun + 0.5);
      if (k < 1) k = 1;
      uint32_t fit = (uint32_t)(pt * 1.02 / k) + 1;
      if (fit < 8) fit = 8;
      if (fit > 2 * pl.krun) fit = 2 * pl.krun;
      if (k <= 12) pl.krun = fit;   // (with many rounds the last one costs little and K's integer steps are too coarse to aim)
    }
  }
  if (g_msm_seg) pl.krun = (uint32_t)g_msm_seg;
  uint32_t maxlog = g_msm_kchunk ? (uint32_t)g_msm_kchunk : 3u;   // log2 of the reduction radix
  if (maxlog < 1) maxlog = 1;
  if (maxlog > 5) maxlog = 5;
  // level 0 always runs (it is what reads the pieces); further levels while a window has more than 512 inputs
  pl.lsize[0] = pl.nbuck;
  uint32_t l = 0;
  do {
    uint32_t lg = 0;
    while ((2u << lg) <= pl.lsize[l] && lg + 1 <= maxlog) lg++;
    pl.logk[l] = lg;
    pl.lsize[l + 1] = pl.lsize[l] >> lg;
    l++;
  } while (pl.lsize[l] > 512 && l < ZKB_MSM_MAXLEV);
  pl.nlev = l;
  pl.n = n;
  pl.max_runs = ((unsigned long long)n * pl.nwin + pl.krun - 1) / pl.krun;
  return pl;
}

// exclusive scan of n u32 -> out[0..n]; `part` holds ceil(n / SCAN_TILE) + 1 words
static inline void scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* part, cudaStream_t st) {
  uint32_t nparts = (uint32_t)((n + ZKB_SCAN_TILE - 1) / ZKB_SCAN_TILE);
  scan_partial_kernel<<<nparts, ZKB_SCAN_THREADS, 0, st>>>(in, part, n);
  scan_spine_kernel<<<1, ZKB_SCAN_THREADS, 0, st>>>(part, nparts);
  scan_final_kernel<<<nparts, ZKB_SCAN_THREADS, 0, st>>>(in, out, part, n, nparts);
}

// Sizes of one MSM launch, derived from the plan only (so the scratch of a whole batch can be reserved up front).
// level split of the digit sort (msm.cuh: msm_part_*_kernel / msm_bucket_sort_kernel)
inline MsmSortGeom msm_sort_geometry(size_t n, size_t nb) {
  MsmSortGeom sg;
  uint32_t lognb = 0;
  while (((size_t)1 << lognb) < nb) lognb++;
  // level 2 resolves L <= 7 bits per CTA; keep at least ~1024 partitions when the key space allows it (one CTA each)
  int L = (int)lognb - 10;
  if (L < 3) L = 3;
  if (L > 7) L = 7;
  sg.L = (uint32_t)L;
  sg.np = (uint32_t)((nb + ((size_t)1 << L) - 1) >> L);
  size_t ctas = (n + 511) / 512;              // >= 512 scalars per level-1 CTA
  if (ctas > 148 * 8) ctas = 148 * 8;         // (8 CTAs of 256 threads per SM: the walk is latency-bound, occupancy pays)
  if (ctas < 1) ctas = 1;
  sg.G = (uint32_t)ctas;
  sg.per_cta = (uint32_t)((n + ctas - 1) / ctas);
  return sg;
}

template <class X>
struct MsmGeom {
  MsmPlan pl;
  MsmSortGeom sg;
  size_t nb, nrefs, max_pieces, nparts, max_vhot, lev_elems, out_bytes, need;
  uint32_t m, nbits, njobs, parts[ZKB_MSM_MAXLEV];
  bool skip;   // nothing to do (n == 0, or a window shard beyond the last window)
};
template <class X>
inline int msm_geometry(size_t n, uint32_t scalar_bits, uint32_t wrank, uint32_t wworld, uint32_t table_c, size_t table_n,
                        uint32_t run_threads, MsmGeom<X>* g) {
  memset(g, 0, sizeof(*g));
  if (wworld == 0 || (!(wworld & ZKB_WINDOW_RANGE) && wrank >= wworld)) return set_error(ZKB_ERR_ARG, "msm: bad window shard");
  if (n >= ((size_t)1 << 31)) return set_error(ZKB_ERR_ARG, "msm: more than 2^31-1 points");
  if (n == 0) {
    g->skip = true;
    return ZKB_OK;
  }
  if (table_n && n > table_n) return set_error(ZKB_ERR_ARG, "msm: more scalars than table points");
  g->pl = msm_make_plan(n, scalar_bits, wrank, wworld, table_c, table_n, run_threads);
  const MsmPlan& pl = g->pl;
  if (pl.nwin == 0) {   // more ranks than windows
    g->skip = true;
    return ZKB_OK;
  }
  g->nb = (size_t)pl.bwin * pl.nbuck;
  g->nrefs = n * pl.nwin;
  if (g->nrefs >= ((size_t)1 << 32)) return set_error(ZKB_ERR_ARG, "msm: n * windows exceeds 2^32 references");
  if (table_n && (size_t)pl.nwin_total * table_n >= ((size_t)1 << 31))
    return set_error(ZKB_ERR_ARG, "msm: table index exceeds 31 bits");
  g->max_pieces = pl.max_runs + g->nb + 1;
  g->sg = msm_sort_geometry(n, g->nb);
  if ((size_t)g->sg.np * 4 > 96 * 1024) return set_error(ZKB_ERR_ARG, "msm: key space too large for the partition histogram");
  {
    size_t scan_max = (size_t)g->sg.np * g->sg.G + 1;          // the per-CTA partition histogram is the longest scan input
    if (scan_max < g->nb) scan_max = g->nb;
    g->nparts = (scan_max + ZKB_SCAN_TILE - 1) / ZKB_SCAN_TILE + 2;
  }
  g->max_vhot = pl.max_runs / ZKB_MSM_VHOT + 1;
  for (uint32_t l = 0; l < pl.nlev; l++) g->lev_elems += (size_t)pl.bwin * pl.lsize[l + 1];
  g->m = pl.lsize[pl.nlev];
  while ((1u << g->nbits) < g->m) g->nbits++;
  g->njobs = g->nbits + 1;
  for (uint32_t l = 0; l < pl.nlev; l++) {   // U_l is summed in parts of ~512 elements (at most ZKB_MSM_MAXPARTS)
    uint32_t p = (pl.lsize[l + 1] + 511) / 512;
    g->parts[l] = p > ZKB_MSM_MAXPARTS ? ZKB_MSM_MAXPARTS : p;
    g->njobs += g->parts[l];
  }
  if (g->njobs > ZKB_MSM_MAXJOBS) return set_error(ZKB_ERR_ARG, "msm: too many reduction jobs");
  g->out_bytes = (size_t)pl.bwin * g->njobs * sizeof(X);
  g->need = (g->nb + 1) * 4 * 6 + g->nrefs * 4 + (pl.max_runs + 1) * 4 + g->nb * 4 + g->max_vhot * 4 + g->nparts * 4 +
            g->max_pieces * sizeof(X) + 2 * g->lev_elems * sizeof(X) + g->max_vhot * ZKB_MSM_VHOT_SPLIT * sizeof(X) +
            g->out_bytes + 32 * 256 +
            ((size_t)g->sg.np * g->sg.G + 2) * 4 * 2 + g->nrefs * 8 + (size_t)g->sg.np * 4 + 2048;   // two-level sort: CTA histograms (+ scanned) and elements
  return ZKB_OK;
}

// device pointers an MSM keeps between its two phases (stored in the ticket)
template <class X>
struct MsmDev {
  MsmGeom<X> g;
  uint32_t *np_eff, *pstart, *hot_list, *vhot_list, *counters;
  X *pieces, *lev_t, *lev_r, *side, *sums;
};

template <class F, int SCALAR_BITS>
int msm_need_t(size_t n, uint32_t wrank, uint32_t wworld, uint32_t table_c, size_t table_n, size_t* need) {
  typedef XYZZ<typename AccumField<F>::type> X;
  MsmGeom<X> g;
  int rc = msm_geometry<X>(n, SCALAR_BITS, wrank, wworld, table_c, table_n, accum_run_threads<F>(), &g);
  *need = g.need + 4096;
  return rc;
}

// Phase 1 = digit sort (msm_sort_t, on stream `st`) + bucket accumulation (msm_accum_t, on the library stream).  Both take their
// scratch from the arena WITHOUT resetting it (the caller reserved and reset once for the whole batch), so several MSMs can be
// between their stages at once; a batch may run the sorts of its later jobs on a side stream under an earlier accumulation
// (msm_common.cu:msm_enqueue_batch).
template <class F, int SCALAR_BITS>
int msm_sort_t(int curve, int group, const void* d_scalars, size_t n, uint32_t wrank, uint32_t wworld,
               uint32_t table_c, size_t table_n, const MsmTicket* share, MsmTicket* tk, cudaStream_t st) {
  typedef typename AccumField<F>::type FA;   // G1: inline-multiplier twin (same layout); G2: F itself
  typedef XYZZ<FA> X;
  static_assert(sizeof(FA) == sizeof(F), "inline twin must share the layout");
  static_assert(sizeof(MsmDev<X>) <= sizeof(tk->dev), "ticket device-state storage too small");
  tk->curve = curve;
  tk->group = group;
  MsmDev<X>* d = reinterpret_cast<MsmDev<X>*>(tk->dev);
  int rc;
  if ((rc = msm_geometry<X>(n, SCALAR_BITS, wrank, wworld, table_c, table_n, accum_run_threads<F>(), &d->g))) return rc;
  tk->empty = d->g.skip;
  if (d->g.skip) return ZKB_OK;
  const MsmGeom<X>& g = d->g;
  const MsmPlan& pl = g.pl;
  if ((rc = ticket_reserve(tk, g.out_bytes))) return rc;
  const size_t nb = g.nb;
  // A job whose scalars, length, table geometry and window shard equal an earlier job's of the batch (Groth16: [B]_1 and [B]_2
  // are both MSMs over V) reuses that job's digit sort: same references, bucket offsets, piece plan and hot lists.
  const MsmSorted* sh = share ? &share->sorted : nullptr;
  if (sh && !(sh->scalars == d_scalars && sh->n == n && sh->table_c == table_c && sh->table_n == table_n && sh->wrank == wrank &&
              sh->wworld == wworld && sh->nb == nb && sh->krun == pl.krun))
    sh = nullptr;
  uint32_t *cnt = nullptr, *start, *cursor = nullptr, *npieces, *refs, *run_bucket, *part = nullptr;
  d->np_eff = (uint32_t*)scratch_take((nb + 1) * 4);
  d->counters = (uint32_t*)scratch_take(256);   // [0] hot count, [1] accumulate work counter, [2] very hot count
  if (sh) {
    start = sh->start;
    npieces = sh->npieces;
    refs = sh->refs;
    run_bucket = sh->run_bucket;
    d->pstart = sh->pstart;
    d->hot_list = sh->hot_list;
    d->vhot_list = sh->vhot_list;
  } else {
    cnt = (uint32_t*)scratch_take((nb + 1) * 4);
    start = (uint32_t*)scratch_take((nb + 1) * 4);
    cursor = (uint32_t*)scratch_take((nb + 1) * 4);
    npieces = (uint32_t*)scratch_take((nb + 1) * 4);
    d->pstart = (uint32_t*)scratch_take((nb + 1) * 4);
    refs = (uint32_t*)scratch_take(g.nrefs * 4);
    run_bucket = (uint32_t*)scratch_take((pl.max_runs + 1) * 4);
    d->hot_list = (uint32_t*)scratch_take(nb * 4);
    d->vhot_list = (uint32_t*)scratch_take(g.max_vhot * 4);
    part = (uint32_t*)scratch_take(g.nparts * 4);
  }
  d->pieces = (X*)scratch_take(g.max_pieces * sizeof(X));
  d->lev_t = (X*)scratch_take(g.lev_elems * sizeof(X));
  d->lev_r = (X*)scratch_take(g.lev_elems * sizeof(X));
  d->side = (X*)scratch_take(g.max_vhot * ZKB_MSM_VHOT_SPLIT * sizeof(X));
  d->sums = (X*)scratch_take(g.out_bytes);
  if ((!sh && (!cnt || !cursor || !part)) || !start || !npieces || !d->np_eff || !d->pstart || !refs || !run_bucket ||
      !d->hot_list || !d->vhot_list || !d->counters || !d->pieces || !d->lev_t || !d->lev_r || !d->side || !d->sums)
    return set_error(ZKB_ERR_CUDA, "msm: scratch exhausted");

  const bool timed = st == MS();   // the per-family timers bracket the library stream only
  const uint32_t* sc = (const uint32_t*)d_scalars;
  if (timed) prof_begin(PROF_MSM_SORT);
  if (sh) {
    // own copies of what the folds rewrite (np_eff) and of the counters (hot counts kept, work counter cleared)
    ZKB_CUDA(cudaMemcpyAsync(d->np_eff, npieces, (nb + 1) * 4, cudaMemcpyDeviceToDevice, st));
    ZKB_CUDA(cudaMemcpyAsync(d->counters, sh->counters, 256, cudaMemcpyDeviceToDevice, st));
    ZKB_CUDA(cudaMemsetAsync(d->counters + 1, 0, 4, st));
  } else {
    ZKB_CUDA(cudaMemsetAsync(d->counters, 0, 256, st));
    unsigned bblocks = (unsigned)((nb + 255) / 256);
    static const bool sort2 = [] { const char* e = getenv("ZKB_MSM_SORT2"); return !e || atoi(e) != 0; }();
    if (sort2) {
      // two-level sort, all histograms in shared memory (msm.cuh)
      const MsmSortGeom& sg = g.sg;
      const size_t hist_len = (size_t)sg.np * sg.G;
      uint32_t* cta_hist = (uint32_t*)scratch_take((hist_len + 2) * 4);
      uint32_t* cta_off = (uint32_t*)scratch_take((hist_len + 2) * 4);
      uint2* elems = (uint2*)scratch_take(g.nrefs * 8);
      if (!cta_hist || !cta_off || !elems) return set_error(ZKB_ERR_CUDA, "msm: scratch exhausted");
      static bool sort_attr = false;
      if (!sort_attr) {
        ZKB_CUDA(cudaFuncSetAttribute(msm_part_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        ZKB_CUDA(cudaFuncSetAttribute(msm_part_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        sort_attr = true;
      }
      const size_t hsm = (size_t)sg.np * 4;
      msm_part_hist_kernel<<<sg.G, 256, hsm, st>>>(pl, sg, sc, cta_hist);
      scan_u32(cta_hist, cta_off, hist_len, part, st);
      msm_part_scatter_kernel<<<sg.G, 256, hsm, st>>>(pl, sg, sc, cta_off, elems);
      // oversized partitions (top window, skewed witnesses) are listed by the level-2 kernel and spread over the grid
      uint32_t* big_count = (uint32_t*)scratch_take(256);
      uint32_t* big_list = (uint32_t*)scratch_take((size_t)sg.np * 4);
      if (!big_count || !big_list) return set_error(ZKB_ERR_CUDA, "msm: scratch exhausted");
      ZKB_CUDA(cudaMemsetAsync(big_count, 0, 4, st));
      msm_bucket_sort_kernel<<<sg.np, 256, 0, st>>>(sg, (unsigned long long)nb, cta_off, elems, cnt, start, refs, big_count, big_list);
      msm_big_count_kernel<<<148 * 2, 256, 0, st>>>(sg, cta_off, elems, big_count, big_list, cnt);
      msm_big_scan_kernel<<<32, 128, 0, st>>>(sg, (unsigned long long)nb, cta_off, big_count, big_list, cnt, start, cursor);
      msm_big_scatter_kernel<<<148 * 2, 256, 0, st>>>(sg, cta_off, elems, big_count, big_list, cursor, refs);
    } else {
      ZKB_CUDA(cudaMemsetAsync(cnt, 0, (nb + 1) * 4, st));
      unsigned pblocks = (unsigned)((n + 255) / 256);
      msm_count_kernel<<<pblocks, 256, 0, st>>>(pl, sc, cnt);
      scan_u32(cnt, start, nb, part, st);
      ZKB_CUDA(cudaMemcpyAsync(cursor, start, (nb + 1) * 4, cudaMemcpyDeviceToDevice, st));
      msm_scatter_kernel<<<pblocks, 256, 0, st>>>(pl, sc, cursor, refs);
    }
    msm_piece_plan_kernel<<<bblocks, 256, 0, st>>>(pl, cnt, start, npieces, d->np_eff, run_bucket, d->hot_list, d->vhot_list,
                                                   d->counters);
    scan_u32(npieces, d->pstart, nb, part, st);
    count_launch(10);
  }
  if (timed) prof_end(PROF_MSM_SORT);
  tk->sorted = MsmSorted{d_scalars, n, table_c, table_n, wrank, wworld, nb, pl.krun, start, d->pstart, npieces, refs, run_bucket,
                         d->hot_list, d->vhot_list, d->counters};
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}

// Bucket accumulation of a sorted job on the library stream.  leave_room (experiment, off unless ZKB_G2_ROOM_CTAS is set): run
// the pair kernel with fewer CTAs per SM than fit, pinned with dynamic shared memory, so that sort kernels of later jobs find room
// on every SM while it runs.  Measured on 2^20 BN254: the accumulation goes 7.25 -> 7.84 ms at 3 CTAs and the proof gains
// nothing (profiles/R2p_*), so the default leaves the grid alone.
template <class F, int SCALAR_BITS>
int msm_accum_t(const void* d_points, MsmTicket* tk, bool leave_room) {
  typedef typename AccumField<F>::type FA;
  typedef XYZZ<FA> X;
  if (tk->empty) return ZKB_OK;
  MsmDev<X>* d = reinterpret_cast<MsmDev<X>*>(tk->dev);
  const MsmPlan& pl = d->g.pl;
  const MsmSorted& so = tk->sorted;
  cudaStream_t st = MS();
  const int acc_tag = tk->group == 2 ? PROF_MSM_ACCUM_G2 : PROF_MSM_ACCUM_G1;
  prof_begin(acc_tag);
  constexpr int MINB = AccumField<F>::MINB;
  if constexpr (AccumField<F>::PAIR) {
    typedef typename AccumField<F>::PairBase PB;
    int ctas = MINB;
    size_t dyn = 0;
    if (leave_room && MINB > 2) {
      static const int room_ctas = [] { const char* e = getenv("ZKB_G2_ROOM_CTAS"); return e ? atoi(e) : 0; }();
      if (room_ctas > 0 && room_ctas < MINB) {
        ctas = room_ctas;
        dyn = (size_t)(200 * 1024 / ctas) & ~(size_t)1023;   // ctas CTAs fill ~200 KiB of the SM's 227: no further one fits, a sort CTA does
        static bool attr = false;
        if (!attr) {
          ZKB_CUDA(cudaFuncSetAttribute(msm_accumulate_pair_kernel<PB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          attr = true;
        }
      }
    }
    msm_accumulate_pair_kernel<PB, MINB><<<148 * ctas, 128, dyn, st>>>(pl, (const Fp<PB>*)d_points, so.refs, so.start, so.pstart,
                                                                       so.run_bucket, (Fp<PB>*)d->pieces, d->counters + 1);
  } else {
    msm_accumulate_kernel<FA, MINB><<<148 * MINB, 128, 0, st>>>(pl, (const Affine<FA>*)d_points, so.refs, so.start, so.pstart,
                                                                so.run_bucket, d->pieces, d->counters + 1);
  }
  prof_end(acc_tag);
  count_launch(1);
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}
template <class F, int SCALAR_BITS>
int msm_phase1_t(int curve, int group, const void* d_points, const void* d_scalars, size_t n, uint32_t wrank, uint32_t wworld,
                 uint32_t table_c, size_t table_n, const MsmTicket* share, MsmTicket* tk) {
  int rc = msm_sort_t<F, SCALAR_BITS>(curve, group, d_scalars, n, wrank, wworld, table_c, table_n, share, tk, MS());
  if (rc) return rc;
  return msm_accum_t<F, SCALAR_BITS>(d_points, tk, false);
}

// Phase 2 on stream `st` (the library stream, or a side stream so that the latency-bound reductions of several MSMs overlap):
// folds, bucket reduction, plain sums, device->host copy of the per-window sums into the ticket's pinned buffer, event record.
template <class F, int SCALAR_BITS>
int msm_phase2_t(MsmTicket* tk, cudaStream_t st) {
  typedef typename AccumField<F>::type FA;
  typedef XYZZ<FA> X;
  if (tk->empty) return ZKB_OK;
  MsmDev<X>* d = reinterpret_cast<MsmDev<X>*>(tk->dev);
  const MsmGeom<X>& g = d->g;
  const MsmPlan& pl = g.pl;
  constexpr int SUM_THREADS = sizeof(X) <= 192 ? 256 : 128;
  static bool attr_set = false;
  if (!attr_set) {
    ZKB_CUDA(cudaFuncSetAttribute(msm_fold_cta1_kernel<FA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * (int)sizeof(X)));
    ZKB_CUDA(cudaFuncSetAttribute(msm_fold_cta2_kernel<FA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)ZKB_MSM_VHOT_SPLIT * (int)sizeof(X)));
    ZKB_CUDA(cudaFuncSetAttribute(msm_sums_kernel<FA, SUM_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  SUM_THREADS * (int)sizeof(X)));
    attr_set = true;
  }
  msm_fold_cta1_kernel<FA><<<dim3(ZKB_MSM_VHOT_SPLIT, 8), 128, 128 * sizeof(X), st>>>(d->vhot_list, d->counters, d->np_eff,
                                                                                      d->pstart, d->pieces, d->side);
  msm_fold_cta2_kernel<FA><<<8, ZKB_MSM_VHOT_SPLIT, ZKB_MSM_VHOT_SPLIT * sizeof(X), st>>>(d->vhot_list, d->counters, d->np_eff,
                                                                                         d->pstart, d->pieces, d->side);
  msm_fold_warp_kernel<FA><<<148 * 4, 128, 0, st>>>(d->hot_list, d->counters, d->np_eff, d->pstart, d->pieces);
  SumJobs<FA> jobs;
  memset(&jobs, 0, sizeof(jobs));
  size_t off = 0;
  uint32_t q = 0;
  const X* r_prev = nullptr;
  for (uint32_t l = 0; l < pl.nlev; l++) {
    size_t outs = (size_t)pl.bwin * pl.lsize[l + 1];
    unsigned blocks = (unsigned)((outs + 127) / 128);
    if (l == 0) msm_level0_kernel<FA><<<blocks, 128, 0, st>>>(pl, d->np_eff, d->pstart, d->pieces, d->lev_t + off, d->lev_r + off);
    else msm_level_kernel<FA><<<blocks, 128, 0, st>>>(pl.bwin, pl.lsize[l], pl.logk[l], r_prev, d->lev_t + off, d->lev_r + off);
    uint32_t per = (pl.lsize[l + 1] + g.parts[l] - 1) / g.parts[l];
    for (uint32_t p = 0; p < g.parts[l]; p++, q++) {
      jobs.base[q] = d->lev_t + off;
      jobs.stride[q] = pl.lsize[l + 1];
      jobs.offset[q] = p * per;
      jobs.count[q] = (p + 1) * per <= pl.lsize[l + 1] ? per : pl.lsize[l + 1] - p * per;
      jobs.bit[q] = -1;
    }
    r_prev = d->lev_r + off;
    off += outs;
  }
  for (uint32_t k = 0; k <= g.nbits; k++, q++) {
    jobs.base[q] = r_prev;
    jobs.stride[q] = g.m;
    jobs.offset[q] = 0;
    jobs.count[q] = g.m;
    jobs.bit[q] = (k == g.nbits) ? -1 : (int)k;
  }
  msm_sums_kernel<FA, SUM_THREADS><<<dim3(g.njobs, pl.bwin), SUM_THREADS, SUM_THREADS * sizeof(X), st>>>(jobs, g.njobs, d->sums);
  count_launch(4 + (int)pl.nlev);
  ZKB_CUDA(cudaGetLastError());
  ZKB_CUDA(cudaMemcpyAsync(tk->host, d->sums, g.out_bytes, cudaMemcpyDeviceToHost, st));
  count_d2h(g.out_bytes);
  ZKB_CUDA(cudaEventRecord((cudaEvent_t)tk->event, st));
  tk->nwin = pl.bwin;
  tk->win0 = pl.table_n ? 0 : pl.win0;   // table entries already carry the factor 2^(c w)
  tk->c = pl.c;
  tk->nlev = pl.nlev;
  tk->nbits = g.nbits;
  for (uint32_t l = 0; l < ZKB_MSM_MAXLEV; l++) {
    tk->logk[l] = pl.logk[l];
    tk->parts[l] = l < pl.nlev ? g.parts[l] : 0;
  }
  return ZKB_OK;
}

// Fixed-base table of the n points: chooses the window size (cost model with ONE bucket set) when *c == 0, reports W, and --
// when `table` is non-null -- fills table[w * n + i] = 2^(c w) * pts[i] (W * n affine points).  Call once with table == nullptr
// to size the allocation.
template <class F, int SCALAR_BITS>
int msm_table_run(const void* d_pts, size_t n, uint32_t world, uint32_t* c, uint32_t* W, void* d_table) {
  if (n == 0) return set_error(ZKB_ERR_ARG, "msm table: no points");
  if (!*c) *c = (uint32_t)msm_choose_window(n, SCALAR_BITS, world ? world : 1, true);
  if (*c < 3 || *c > 22) return set_error(ZKB_ERR_ARG, "msm table: window size out of range");
  *W = (SCALAR_BITS + 1 + *c - 1) / *c;
  if ((size_t)*W * n >= ((size_t)1 << 31)) return set_error(ZKB_ERR_ARG, "msm table: W * n exceeds 31 bits");
  if (!d_table) return ZKB_OK;
  msm_table_kernel<F><<<(unsigned)((n + 127) / 128), 128, 0, MS()>>>(n, *c, *W, (const Affine<F>*)d_pts, (Affine<F>*)d_table);
  count_launch();
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}

template <class F>
int points_conv_run(bool to, size_t n, void* d_points) {
  if (n == 0) return ZKB_OK;
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (to) points_to_mont_kernel<F><<<blocks, 128, 0, MS()>>>(n, (Affine<F>*)d_points);
  else points_from_mont_kernel<F><<<blocks, 128, 0, MS()>>>(n, (Affine<F>*)d_points);
  count_launch();
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}

template <class F>
int batch_mul_run(const void* d_bases, int single_base, const void* d_scalars, size_t n, void* d_out) {
  if (n == 0) return ZKB_OK;
  batch_scalar_mul_kernel<F><<<(unsigned)((n + 127) / 128), 128, 0, MS()>>>(n, (const Affine<F>*)d_bases, single_base,
                                                                          (const uint32_t*)d_scalars, (Affine<F>*)d_out);
  count_launch();
  ZKB_CUDA(cudaGetLastError());
  return ZKB_OK;
}

// every (curve, group) translation unit exports these three with a unique suffix
#define ZKB_MSM_INSTANTIATE(SUFFIX, FIELD, BITS, CURVE, GROUP)                                                           \
  int msm_need_##SUFFIX(size_t n, uint32_t wr, uint32_t ww, uint32_t tc, size_t tn, size_t* need) {                     \
    return msm_need_t<FIELD, BITS>(n, wr, ww, tc, tn, need);                                                            \
  }                                                                                                                      \
  int msm_phase1_##SUFFIX(const void* p, const void* s, size_t n, uint32_t wr, uint32_t ww, uint32_t tc, size_t tn,     \
                          const MsmTicket* share, MsmTicket* tk) {                                                       \
    return msm_phase1_t<FIELD, BITS>(CURVE, GROUP, p, s, n, wr, ww, tc, tn, share, tk);                                 \
  }                                                                                                                      \
  int msm_sort_##SUFFIX(const void* s, size_t n, uint32_t wr, uint32_t ww, uint32_t tc, size_t tn, const MsmTicket* share,  \
                        MsmTicket* tk, void* stream) {                                                                   \
    return msm_sort_t<FIELD, BITS>(CURVE, GROUP, s, n, wr, ww, tc, tn, share, tk, (cudaStream_t)stream);                \
  }                                                                                                                      \
  int msm_accum_##SUFFIX(const void* p, MsmTicket* tk, int leave_room) { return msm_accum_t<FIELD, BITS>(p, tk, leave_room != 0); } \
  int msm_phase2_##SUFFIX(MsmTicket* tk, void* stream) { return msm_phase2_t<FIELD, BITS>(tk, (cudaStream_t)stream); }  \
  int msm_table_##SUFFIX(const void* pts, size_t n, uint32_t world, uint32_t* c, uint32_t* W, void* table) {            \
    return msm_table_run<FIELD, BITS>(pts, n, world, c, W, table);                                                      \
  }                                                                                                                      \
  int points_conv_##SUFFIX(int to, size_t n, void* p) { return points_conv_run<FIELD>(to != 0, n, p); }                 \
  int batch_mul_##SUFFIX(const void* b, int single, const void* s, size_t n, void* o) {                                 \
    return batch_mul_run<FIELD>(b, single, s, n, o);                                                                    \
  }

}  // namespace zkb
