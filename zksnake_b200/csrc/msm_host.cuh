// msm_host.cuh -- host driver of the Pippenger pipeline, instantiated once per (curve, group) in msm_inst_*.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <vector>
#include "msm.cuh"
#include "zkb_internal.h"

namespace zkb {

extern int g_msm_c, g_msm_seg, g_msm_kchunk;   // tuning overrides (0 = heuristic), msm_common.cu

static inline cudaStream_t MS() { return (cudaStream_t)ctx_stream(); }

inline MsmPlan msm_make_plan(size_t n, uint32_t scalar_bits) {
  MsmPlan pl;
  uint32_t logn = 0;
  while (((size_t)1 << logn) < n) logn++;
  int c = g_msm_c ? g_msm_c : (int)logn - 4;
  if (c < 3) c = 3;
  if (c > 16) c = 16;
  pl.c = (uint32_t)c;
  pl.nwin = (scalar_bits + 1 + pl.c - 1) / pl.c;
  pl.nbuck = 1u << (pl.c - 1);
  pl.seg = g_msm_seg ? (uint32_t)g_msm_seg : 32u;
  uint32_t k = g_msm_kchunk ? (uint32_t)g_msm_kchunk : 16u;
  while (k > pl.nbuck) k >>= 1;
  // keep at least ~16k threads in the bucket reduction when the bucket count allows it
  while (k > 2 && (size_t)pl.nwin * pl.nbuck / k < 16384) k >>= 1;
  pl.kchunk = k;
  pl.n = n;
  pl.max_segs = (unsigned long long)pl.nwin * pl.nbuck + ((unsigned long long)n * pl.nwin) / pl.seg + 1;
  return pl;
}

template <class F, int SCALAR_BITS>
int msm_run(int curve, int group, const void* d_points, const void* d_scalars, size_t n, uint64_t* out_xy, int* out_inf) {
  typedef XYZZ<F> X;
  size_t abytes = sizeof(Affine<F>);
  if (n == 0) {
    memset(out_xy, 0, abytes);
    *out_inf = 1;
    return ZKB_OK;
  }
  if (n >= ((size_t)1 << 31)) return set_error(ZKB_ERR_ARG, "msm: more than 2^31-1 points");
  static bool attr_set = false;
  if (!attr_set) {
    ZKB_CUDA(cudaFuncSetAttribute(msm_hot_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * (int)sizeof(X)));
    ZKB_CUDA(cudaFuncSetAttribute(msm_window_sum_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * (int)sizeof(X)));
    attr_set = true;
  }
  MsmPlan pl = msm_make_plan(n, SCALAR_BITS);
  const size_t nb = (size_t)pl.nwin * pl.nbuck;
  const size_t nrefs = n * pl.nwin;
  if (nrefs >= ((size_t)1 << 32)) return set_error(ZKB_ERR_ARG, "msm: n * windows exceeds 2^32 references");
  const uint32_t chunks_per_win = pl.nbuck / pl.kchunk;
  const size_t nchunks = (size_t)pl.nwin * chunks_per_win;

  size_t need = (nb + 1) * 4 * 5 + nrefs * 4 + pl.max_segs * 4 + nb * 4 + pl.max_segs * sizeof(X) + nchunks * sizeof(X) +
                pl.nwin * sizeof(X) + 64 * 256;
  int rc;
  if ((rc = scratch_reserve(need))) return rc;
  scratch_reset();
  uint32_t* cnt = (uint32_t*)scratch_take((nb + 1) * 4);
  uint32_t* start = (uint32_t*)scratch_take((nb + 1) * 4);
  uint32_t* cursor = (uint32_t*)scratch_take((nb + 1) * 4);
  uint32_t* nseg = (uint32_t*)scratch_take((nb + 1) * 4);
  uint32_t* segstart = (uint32_t*)scratch_take((nb + 1) * 4);
  uint32_t* refs = (uint32_t*)scratch_take(nrefs * 4);
  uint32_t* seg_bucket = (uint32_t*)scratch_take(pl.max_segs * 4);
  uint32_t* hot_list = (uint32_t*)scratch_take(nb * 4);
  uint32_t* hot_count = (uint32_t*)scratch_take(256);
  X* seg_sum = (X*)scratch_take(pl.max_segs * sizeof(X));
  X* contrib = (X*)scratch_take(nchunks * sizeof(X));
  X* win_sum = (X*)scratch_take(pl.nwin * sizeof(X));
  if (!cnt || !start || !cursor || !nseg || !segstart || !refs || !seg_bucket || !hot_list || !hot_count || !seg_sum ||
      !contrib || !win_sum)
    return set_error(ZKB_ERR_CUDA, "msm: scratch exhausted");

  cudaStream_t st = MS();
  const uint32_t* sc = (const uint32_t*)d_scalars;
  ZKB_CUDA(cudaMemsetAsync(cnt, 0, (nb + 1) * 4, st));
  ZKB_CUDA(cudaMemsetAsync(hot_count, 0, 4, st));
  unsigned pblocks = (unsigned)((n + 255) / 256);
  prof_begin(PROF_MSM_SORT);
  msm_count_kernel<<<pblocks, 256, 0, st>>>(pl, sc, cnt);
  scan_kernel<<<1, 1024, 0, st>>>(cnt, start, nb);
  ZKB_CUDA(cudaMemcpyAsync(cursor, start, (nb + 1) * 4, cudaMemcpyDeviceToDevice, st));
  msm_scatter_kernel<<<pblocks, 256, 0, st>>>(pl, sc, cursor, refs);
  unsigned bblocks = (unsigned)((nb + 255) / 256);
  msm_nseg_kernel<<<bblocks, 256, 0, st>>>(pl, cnt, nseg);
  scan_kernel<<<1, 1024, 0, st>>>(nseg, segstart, nb);
  msm_segfill_kernel<<<bblocks, 256, 0, st>>>(pl, nseg, segstart, seg_bucket, hot_list, hot_count);
  prof_end(PROF_MSM_SORT);
  unsigned ablocks = (unsigned)((pl.max_segs + 127) / 128);
  const int acc_tag = group == 2 ? PROF_MSM_ACCUM_G2 : PROF_MSM_ACCUM_G1;
  prof_begin(acc_tag);
  msm_accumulate_kernel<F><<<ablocks, 128, 0, st>>>(pl, (const Affine<F>*)d_points, refs, cnt, start, segstart, seg_bucket,
                                                    seg_sum);
  prof_end(acc_tag);
  prof_begin(PROF_MSM_REDUCE);
  msm_hot_kernel<F><<<296, 128, 128 * sizeof(X), st>>>(hot_list, hot_count, nseg, segstart, seg_sum);
  msm_bucket_reduce_kernel<F><<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>(pl, nseg, segstart, seg_sum, contrib);
  msm_window_sum_kernel<F><<<pl.nwin, 128, 128 * sizeof(X), st>>>(chunks_per_win, contrib, win_sum);
  prof_end(PROF_MSM_REDUCE);
  count_launch(9);
  ZKB_CUDA(cudaGetLastError());
  std::vector<unsigned char> host(pl.nwin * sizeof(X));
  ZKB_CUDA(cudaMemcpyAsync(host.data(), win_sum, pl.nwin * sizeof(X), cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(cudaStreamSynchronize(st));
  host_msm_finish(curve, group, host.data(), pl.nwin, pl.c, out_xy, out_inf);
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
  int msm_run_##SUFFIX(const void* p, const void* s, size_t n, uint64_t* o, int* inf) {                                 \
    return msm_run<FIELD, BITS>(CURVE, GROUP, p, s, n, o, inf);                                                         \
  }                                                                                                                      \
  int points_conv_##SUFFIX(int to, size_t n, void* p) { return points_conv_run<FIELD>(to != 0, n, p); }                 \
  int batch_mul_##SUFFIX(const void* b, int single, const void* s, size_t n, void* o) {                                 \
    return batch_mul_run<FIELD>(b, single, s, n, o);                                                                    \
  }

}  // namespace zkb
