// msm.cuh -- Pippenger multi-scalar multiplication kernels (G1 and G2, BN254 and BLS12-381) for sm_100a.
//
// Replaces ark-ec 0.4.2 `VariableBaseMSM::msm` as called from /root/reference/src/bn254/curve.rs:356-392
// (bls12_381 twin :367-403).  The result is the group element sum_i s_i * P_i, independent of the algorithm;
// the pipeline here is
//   1. signed-digit decomposition of every scalar into W windows of c bits (digits in [-2^(c-1), 2^(c-1)]),
//   2. a counting sort (one radix pass keyed by (window, |digit|)) of the N*W (point, sign) references:
//      histogram -> multi-CTA exclusive scan -> warp-aggregated scatter,
//   3. bucket accumulation over the SORTED reference array cut into equal runs of K references: every lane of every
//      warp adds exactly K points (XYZZ mixed additions, 8M+2S), whatever the bucket sizes are, and emits one partial
//      sum ("piece") per bucket its run touches.  Bucket b owns the contiguous piece slots [pstart[b], pstart[b+1]):
//      slot pstart[b] + (t - floor(start[b]/K)) for the run t.  Warps fetch runs from a global counter (persistent grid).
//   4. buckets with more than ZKB_MSM_HOT pieces (skewed scalars, e.g. an all-ones witness vector) are folded to one
//      piece: one warp per bucket (lane-strided sums + shuffle tree), 64 CTAs + 1 for buckets beyond ZKB_MSM_VHOT,
//   5. bucket reduction sum_b (b+1) S_b per window.  Every XYZZ addition is ~14 dependent field multiplications, so the
//      reduction is organised for DEPTH, not operation count: radix-8 running-sum levels while a window still has more
//      than 512 inputs (a thread turns 8 adjacent inputs into their total R_j and the zero-based weighted sum T_j;
//      the totals feed the next level), then one kernel of plain tree sums: U_l = sum_j T_j for every level, the
//      bit sums A_beta = sum_{j : bit beta of j set} R_j of the last level's totals, and R_top = sum_j R_j:
//        sum_b (b+1) S_b = R_top + U_0 + 8 (U_1 + ... + 8 (sum_beta 2^beta A_beta)),
//   6. the per-window sums (a few KiB) go to the host, which applies the Horner steps and the c doublings per window
//      (host_math.cpp) while the GPU already runs the next MSM.
// Exact group law everywhere (identity operands, P+P, P-P), so results are bit-exact after affine conversion.
#pragma once
#include "ec.cuh"

namespace zkb {

#define ZKB_MSM_MAXLEV 8
#define ZKB_MSM_MAXJOBS 96     // plain-sum jobs per window: <= 8 parts per level of U sums + <= 16 bit sums + R_top
#define ZKB_MSM_MAXPARTS 8
#define ZKB_MSM_HOT 6u         // buckets with more pieces are folded by a warp
#define ZKB_MSM_VHOT 2048u     // ... by 64 CTAs
#define ZKB_MSM_VHOT_SPLIT 64u
struct MsmPlan {
  uint32_t c;        // window bits
  uint32_t nwin;     // windows handled by THIS launch: [win0, win0 + nwin) of the scalar's nwin_total windows
  uint32_t win0;
  uint32_t nwin_total;
  uint32_t bwin;     // bucket sets: nwin, or 1 with a fixed-base table (every window's digits share one set of buckets)
  uint32_t table_n;  // 0, or the point count N of the fixed-base table: window w of point i gathers table[w * N + i] = 2^(c w) P_i
  uint32_t nbuck;    // buckets per window = 2^(c-1)
  uint32_t krun;     // K: references per accumulation run
  uint32_t nlev;     // levels of the bucket reduction
  uint32_t logk[ZKB_MSM_MAXLEV];   // log2 of the chunk size of level l
  uint32_t lsize[ZKB_MSM_MAXLEV + 1];  // inputs per window at level l (lsize[0] = nbuck); lsize[nlev] <= 512 are bit-summed
  unsigned long long n;         // points
  unsigned long long max_runs;  // upper bound on the run count  ceil(n*W / K)
};

__device__ __forceinline__ uint32_t scalar_bits(const uint32_t* s, uint32_t pos, uint32_t c) {
  // bits [pos, pos+c) of a 256-bit value, c <= 24
  uint32_t limb = pos >> 5, off = pos & 31;
  if (limb >= 8) return 0;
  unsigned long long v = s[limb];
  if (limb + 1 < 8) v |= (unsigned long long)s[limb + 1] << 32;
  return (uint32_t)(v >> off) & ((1u << c) - 1);
}

__device__ __forceinline__ void load_scalar(const uint32_t* p, uint32_t* s) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
  s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
}

// pass 1: histogram of (window, bucket).  The signed digits are produced by walking the windows with a carry.
// Window 0 and the top two windows are where skewed inputs pile up (0/1 or small witnesses live in window 0; when the top
// window only holds the final carry, ~half of all scalars share its bucket 1 -- BLS12-381 with c = 15), so there the atomics
// are warp-aggregated: one atomicAdd per distinct bucket per warp instead of up to 2^20 on one address.
static __global__ void msm_count_kernel(MsmPlan pl, const uint32_t* __restrict__ scalars, uint32_t* __restrict__ cnt) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const bool live = i < pl.n;
  uint32_t s[8];
  if (live) load_scalar(scalars + i * 8, s);
  else { for (int j = 0; j < 8; j++) s[j] = 0; }
  uint32_t carry = 0;
  const uint32_t half = 1u << (pl.c - 1);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t wend = pl.win0 + pl.nwin;
  for (uint32_t w = 0; w < wend; w++) {   // the carry has to be walked up from window 0 even when win0 > 0
    uint32_t d = scalar_bits(s, w * pl.c, pl.c) + carry;
    carry = d > half;
    uint32_t mag = carry ? ((1u << pl.c) - d) : d;
    if (w < pl.win0) continue;   // warp-uniform
    const bool have = live && mag != 0;
    const uint32_t key = (pl.table_n ? 0u : (w - pl.win0) * pl.nbuck) + mag - 1;
    if (w == 0 || w + 2 >= pl.nwin_total) {
      uint32_t peers = __match_any_sync(0xffffffffu, have ? key : 0xffffffffu);
      if (have && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&cnt[key], (uint32_t)__popc(peers));
    } else if (have) {
      atomicAdd(&cnt[key], 1u);
    }
  }
}

// pass 2: scatter (point index | sign << 31) to its bucket's slice; `cursor` starts as a copy of the exclusive scan
static __global__ void msm_scatter_kernel(MsmPlan pl, const uint32_t* __restrict__ scalars, uint32_t* __restrict__ cursor,
                                   uint32_t* __restrict__ refs) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  bool live = i < pl.n;
  uint32_t s[8];
  if (live) load_scalar(scalars + i * 8, s);
  else { for (int j = 0; j < 8; j++) s[j] = 0; }
  uint32_t carry = 0;
  const uint32_t half = 1u << (pl.c - 1);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t wend = pl.win0 + pl.nwin;
  for (uint32_t w = 0; w < wend; w++) {
    uint32_t d = scalar_bits(s, w * pl.c, pl.c) + carry;
    carry = d > half;
    uint32_t mag = carry ? ((1u << pl.c) - d) : d;
    if (w < pl.win0) continue;   // warp-uniform
    bool have = live && mag != 0;
    // warp-aggregated atomics: lanes with the same bucket take consecutive slots from one atomicAdd
    uint32_t key = have ? ((pl.table_n ? 0u : (w - pl.win0) * pl.nbuck) + mag - 1) : 0xffffffffu;
    uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (have) {
      uint32_t leader = __ffs(peers) - 1;
      uint32_t rank = __popc(peers & ((1u << lane) - 1));
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      refs[base + rank] = ((uint32_t)i + w * pl.table_n) | (carry << 31);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// Digit sort, two levels, every histogram in shared memory (round 2; replaces one global atomic per reference in each of
// msm_count_kernel / msm_scatter_kernel: 2 x 14.7 M global atomics per 2^20-point MSM, 0.7 ms).
//   level 1 partitions the (window, |digit|) keys by their high bits: key >> L selects one of NP partitions;
//     msm_part_hist_kernel    -- CTA g walks ITS slice of the scalars, histogram of the partitions in shared memory, written to
//                                cta_hist[partition * G + g] (partition-major, so ONE exclusive scan yields every CTA's write offset
//                                for every partition and partition p is the contiguous range [scan[p G], scan[(p+1) G]) );
//     msm_part_scatter_kernel -- the same walk; cursor[partition] (shared memory, seeded with the scanned offsets) hands out the
//                                slots; element = (low key bits, reference);
//   level 2: msm_bucket_sort_kernel -- one CTA per partition: histogram of the 2^L low key values in shared memory, exclusive scan
//     -> cnt[] and start[] of its buckets (start = partition offset + local prefix), second walk scatters the references.
// Equal keys in a warp (skewed scalars: a 0/1 witness puts every reference of window 0 into ONE bucket) are combined before the
// shared-memory atomic when the whole warp agrees, which is the case that would otherwise serialise 32-fold.
// The order of the references inside a bucket is arbitrary; the bucket sum does not depend on it (exact group law).
// ------------------------------------------------------------------------------------------------------
struct MsmSortGeom {
  uint32_t L;        // low key bits resolved by level 2
  uint32_t np;       // partitions = ceil(nb / 2^L)
  uint32_t G;        // CTAs of level 1
  uint32_t per_cta;  // scalars per level-1 CTA
};

// one warp-wide shared-memory increment; returns the slot each lane got (count-only callers ignore it)
__device__ __forceinline__ uint32_t smem_take(uint32_t* counters, uint32_t key, bool have) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t k0 = __shfl_sync(0xffffffffu, key, 0);
  const uint32_t ballot = __ballot_sync(0xffffffffu, have);
  if (__all_sync(0xffffffffu, !have || key == k0) && __shfl_sync(0xffffffffu, (uint32_t)have, 0)) {
    // every live lane wants the same counter (and lane 0 is live, so k0 is that counter): one atomic for the warp
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&counters[k0], (uint32_t)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
  }
  return have ? atomicAdd(&counters[key], 1u) : 0u;
}

// walks the windows of one scalar; calls f(key, ref) for every non-zero digit of the windows this launch handles
template <class Fn>
__device__ __forceinline__ void msm_walk_digits(const MsmPlan& pl, const uint32_t* s, uint32_t index, bool live, Fn f) {
  uint32_t carry = 0;
  const uint32_t half = 1u << (pl.c - 1);
  const uint32_t wend = pl.win0 + pl.nwin;
  for (uint32_t w = 0; w < wend; w++) {   // the carry has to be walked up from window 0 even when win0 > 0
    uint32_t d = scalar_bits(s, w * pl.c, pl.c) + carry;
    carry = d > half;
    uint32_t mag = carry ? ((1u << pl.c) - d) : d;
    if (w < pl.win0) continue;   // warp-uniform
    const bool have = live && mag != 0;
    const uint32_t key = (pl.table_n ? 0u : (w - pl.win0) * pl.nbuck) + mag - 1;
    const uint32_t ref = (index + w * pl.table_n) | (carry << 31);
    f(have, key, ref);
  }
}

static __global__ void __launch_bounds__(256) msm_part_hist_kernel(MsmPlan pl, MsmSortGeom sg, const uint32_t* __restrict__ scalars,
                                                                   uint32_t* __restrict__ cta_hist) {
  extern __shared__ uint32_t part_smem[];
  for (uint32_t j = threadIdx.x; j < sg.np; j += blockDim.x) part_smem[j] = 0;
  __syncthreads();
  const unsigned long long lo = (unsigned long long)blockIdx.x * sg.per_cta;
  const unsigned long long hi = lo + sg.per_cta < pl.n ? lo + sg.per_cta : pl.n;
  for (unsigned long long i0 = lo; i0 < hi; i0 += blockDim.x) {     // (warp-uniform trip count: shuffles inside)
    const unsigned long long i = i0 + threadIdx.x;
    const bool live = i < hi;
    uint32_t s[8];
    if (live) load_scalar(scalars + i * 8, s);
    else { for (int j = 0; j < 8; j++) s[j] = 0; }
    msm_walk_digits(pl, s, (uint32_t)i, live, [&](bool have, uint32_t key, uint32_t) { smem_take(part_smem, key >> sg.L, have); });
  }
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < sg.np; j += blockDim.x) cta_hist[(size_t)j * sg.G + blockIdx.x] = part_smem[j];
}

static __global__ void __launch_bounds__(256) msm_part_scatter_kernel(MsmPlan pl, MsmSortGeom sg, const uint32_t* __restrict__ scalars,
                                                                      const uint32_t* __restrict__ cta_offset,
                                                                      uint2* __restrict__ parts) {
  extern __shared__ uint32_t part_smem[];
  for (uint32_t j = threadIdx.x; j < sg.np; j += blockDim.x) part_smem[j] = cta_offset[(size_t)j * sg.G + blockIdx.x];
  __syncthreads();
  const unsigned long long lo = (unsigned long long)blockIdx.x * sg.per_cta;
  const unsigned long long hi = lo + sg.per_cta < pl.n ? lo + sg.per_cta : pl.n;
  const uint32_t lowmask = (1u << sg.L) - 1u;
  for (unsigned long long i0 = lo; i0 < hi; i0 += blockDim.x) {
    const unsigned long long i = i0 + threadIdx.x;
    const bool live = i < hi;
    uint32_t s[8];
    if (live) load_scalar(scalars + i * 8, s);
    else { for (int j = 0; j < 8; j++) s[j] = 0; }
    msm_walk_digits(pl, s, (uint32_t)i, live, [&](bool have, uint32_t key, uint32_t ref) {
      uint32_t pos = smem_take(part_smem, key >> sg.L, have);
      if (have) parts[pos] = make_uint2(key & lowmask, ref);
    });
  }
}

// level 2.  grid.x = np; partition p = elements [off[p G], off[(p+1) G]) of `parts` (off = the scanned cta_hist, np G + 1 entries).
// Writes cnt / start of the buckets [p 2^L, (p+1) 2^L) (clipped to nb) and the sorted references; the last CTA also writes start[nb].
// A partition with more than ZKB_SORT_BIG elements is NOT sorted here: it is appended to big_list (its buckets' cnt zeroed) and
// the three msm_big_* kernels spread it over the whole grid.  Such partitions are systematic, not exotic: the top window of a
// 254-bit scalar holds 7 significant bits at c = 19, so ALL N of its digits fall into the first 128 buckets, and a 0/1 witness
// does the same to window 0.  (One CTA walking 2^20 elements twice took 3.3 ms in the first version of this kernel.)
#define ZKB_SORT_BIG 24576u
static __global__ void __launch_bounds__(256) msm_bucket_sort_kernel(MsmSortGeom sg, unsigned long long nb, const uint32_t* __restrict__ off,
                                                                     const uint2* __restrict__ parts, uint32_t* __restrict__ cnt,
                                                                     uint32_t* __restrict__ start, uint32_t* __restrict__ refs,
                                                                     uint32_t* __restrict__ big_count, uint32_t* __restrict__ big_list) {
  __shared__ uint32_t hist[128];
  __shared__ uint32_t cursor[128];
  __shared__ uint32_t warp_tot[8];
  const uint32_t p = blockIdx.x;
  const uint32_t nkeys = 1u << sg.L;     // <= 128
  const uint32_t begin = off[(size_t)p * sg.G], end = off[(size_t)(p + 1) * sg.G];
  if (p + 1 == gridDim.x && threadIdx.x == 0) start[nb] = end;
  const uint32_t span = end - begin;
  if (span > ZKB_SORT_BIG) {
    const unsigned long long b = (unsigned long long)p * nkeys + threadIdx.x;
    if (threadIdx.x < nkeys && b < nb) cnt[b] = 0;
    if (threadIdx.x == 0) big_list[atomicAdd(big_count, 1u)] = p;
    return;
  }
  if (threadIdx.x < 128) hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t trips = (span + blockDim.x - 1) / blockDim.x;
  for (uint32_t it = 0; it < trips; it++) {
    const uint32_t e = begin + it * blockDim.x + threadIdx.x;
    const bool have = e < end;
    uint2 el = have ? parts[e] : make_uint2(0, 0);
    smem_take(hist, el.x, have);
  }
  __syncthreads();
  // exclusive scan of the (<= 128) counters by the first 128 threads: warp scans + 4 warp totals
  uint32_t v = 0, incl = 0;
  if (threadIdx.x < 128) {
    v = threadIdx.x < nkeys ? hist[threadIdx.x] : 0;
    incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= (uint32_t)o) incl += y;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    uint32_t before = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) before += warp_tot[w];
    const uint32_t excl = before + incl - v;
    cursor[threadIdx.x] = begin + excl;
    const unsigned long long b = (unsigned long long)p * nkeys + threadIdx.x;
    if (threadIdx.x < nkeys && b < nb) {
      cnt[b] = v;
      start[b] = begin + excl;
    }
  }
  __syncthreads();
  for (uint32_t it = 0; it < trips; it++) {
    const uint32_t e = begin + it * blockDim.x + threadIdx.x;
    const bool have = e < end;
    uint2 el = have ? parts[e] : make_uint2(0, 0);
    uint32_t pos = smem_take(cursor, el.x, have);
    if (have) refs[pos] = el.y;
  }
}

// ---- oversized partitions: every CTA of the grid takes a strided share of each listed partition ---------------------------------
// element e of partition q belongs to CTA (e - begin) / 256 mod gridDim.x (whole 256-element tiles, the same in count and scatter)
// count: shared-memory histogram of the CTA's tiles, then ONE global atomic per non-empty bucket
static __global__ void __launch_bounds__(256) msm_big_count_kernel(MsmSortGeom sg, const uint32_t* __restrict__ off,
                                                                   const uint2* __restrict__ parts, const uint32_t* __restrict__ big_count,
                                                                   const uint32_t* __restrict__ big_list, uint32_t* __restrict__ cnt) {
  __shared__ uint32_t hist[128];
  const uint32_t nbig = *big_count;
  const uint32_t nkeys = 1u << sg.L;
  for (uint32_t q = 0; q < nbig; q++) {
    const uint32_t p = big_list[q];
    const uint32_t begin = off[(size_t)p * sg.G], end = off[(size_t)(p + 1) * sg.G];
    if (threadIdx.x < 128) hist[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t t0 = begin + blockIdx.x * 256u; t0 < end; t0 += gridDim.x * 256u) {
      const uint32_t e = t0 + threadIdx.x;
      const bool have = e < end;
      uint2 el = have ? parts[e] : make_uint2(0, 0);
      smem_take(hist, el.x, have);
    }
    __syncthreads();
    if (threadIdx.x < nkeys && hist[threadIdx.x]) atomicAdd(&cnt[(size_t)p * nkeys + threadIdx.x], hist[threadIdx.x]);
    __syncthreads();
  }
}
// scan: one CTA (128 threads) per listed partition: start[] of its buckets, and the global cursors the scatter draws from
static __global__ void __launch_bounds__(128) msm_big_scan_kernel(MsmSortGeom sg, unsigned long long nb, const uint32_t* __restrict__ off,
                                                                  const uint32_t* __restrict__ big_count,
                                                                  const uint32_t* __restrict__ big_list, const uint32_t* __restrict__ cnt,
                                                                  uint32_t* __restrict__ start, uint32_t* __restrict__ gcursor) {
  __shared__ uint32_t warp_tot[4];
  const uint32_t nbig = *big_count;
  const uint32_t nkeys = 1u << sg.L;
  for (uint32_t q = blockIdx.x; q < nbig; q += gridDim.x) {
    const uint32_t p = big_list[q];
    const uint32_t begin = off[(size_t)p * sg.G];
    const unsigned long long b = (unsigned long long)p * nkeys + threadIdx.x;
    const bool live = threadIdx.x < nkeys && b < nb;
    const uint32_t v = live ? cnt[b] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= (uint32_t)o) incl += y;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) before += warp_tot[w];
    if (live) {
      start[b] = begin + before + incl - v;
      gcursor[b] = begin + before + incl - v;
    }
    __syncthreads();
  }
}
// scatter: the CTA recounts its tiles, reserves a contiguous range per bucket with ONE global atomic each, fills it
static __global__ void __launch_bounds__(256) msm_big_scatter_kernel(MsmSortGeom sg, const uint32_t* __restrict__ off,
                                                                     const uint2* __restrict__ parts, const uint32_t* __restrict__ big_count,
                                                                     const uint32_t* __restrict__ big_list, uint32_t* __restrict__ gcursor,
                                                                     uint32_t* __restrict__ refs) {
  __shared__ uint32_t hist[128];
  const uint32_t nbig = *big_count;
  const uint32_t nkeys = 1u << sg.L;
  for (uint32_t q = 0; q < nbig; q++) {
    const uint32_t p = big_list[q];
    const uint32_t begin = off[(size_t)p * sg.G], end = off[(size_t)(p + 1) * sg.G];
    if (threadIdx.x < 128) hist[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t t0 = begin + blockIdx.x * 256u; t0 < end; t0 += gridDim.x * 256u) {
      const uint32_t e = t0 + threadIdx.x;
      const bool have = e < end;
      uint2 el = have ? parts[e] : make_uint2(0, 0);
      smem_take(hist, el.x, have);
    }
    __syncthreads();
    if (threadIdx.x < nkeys) {
      const uint32_t mine = hist[threadIdx.x];
      hist[threadIdx.x] = mine ? atomicAdd(&gcursor[(size_t)p * nkeys + threadIdx.x], mine) : 0u;   // becomes this CTA's cursor
    }
    __syncthreads();
    for (uint32_t t0 = begin + blockIdx.x * 256u; t0 < end; t0 += gridDim.x * 256u) {
      const uint32_t e = t0 + threadIdx.x;
      const bool have = e < end;
      uint2 el = have ? parts[e] : make_uint2(0, 0);
      uint32_t pos = smem_take(hist, el.x, have);
      if (have) refs[pos] = el.y;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------
// exclusive scan of n u32 into out[0..n] (out[n] = total): per-CTA partial sums, one CTA scans the partials, per-CTA
// rescan with its offset.  SCAN_TILE elements per CTA.
// ------------------------------------------------------------------------------------------------------
#define ZKB_SCAN_THREADS 256
#define ZKB_SCAN_ITEMS 8
#define ZKB_SCAN_TILE (ZKB_SCAN_THREADS * ZKB_SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* total, uint32_t* sh /* >= 9 words */) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  if (lane == 31) sh[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t s = lane < (blockDim.x >> 5) ? sh[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= (uint32_t)o) s += y;
    }
    if (lane < (blockDim.x >> 5)) sh[lane] = s;   // inclusive over warps
  }
  __syncthreads();
  uint32_t warp_off = warp ? sh[warp - 1] : 0;
  *total = sh[(blockDim.x >> 5) - 1];
  uint32_t r = warp_off + x - v;
  __syncthreads();
  return r;
}

static __global__ void __launch_bounds__(ZKB_SCAN_THREADS) scan_partial_kernel(const uint32_t* __restrict__ in,
                                                                        uint32_t* __restrict__ part, unsigned long long n) {
  __shared__ uint32_t sh[32];
  unsigned long long base = (unsigned long long)blockIdx.x * ZKB_SCAN_TILE + threadIdx.x * ZKB_SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ZKB_SCAN_ITEMS; k++)
    if (base + k < n) s += in[base + k];
  uint32_t total;
  block_exclusive_scan_256(s, &total, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = total;
}

// one CTA of 256 threads: exclusive scan of `m` partials in place (m <= 65536), part[m] = grand total
static __global__ void __launch_bounds__(ZKB_SCAN_THREADS) scan_spine_kernel(uint32_t* part, uint32_t m) {
  __shared__ uint32_t sh[32];
  uint32_t per = (m + ZKB_SCAN_THREADS - 1) / ZKB_SCAN_THREADS;
  uint32_t lo = threadIdx.x * per, hi = lo + per;
  if (hi > m) hi = m;
  uint32_t s = 0;
  for (uint32_t i = lo; i < hi; i++) s += part[i];
  uint32_t total;
  uint32_t run = block_exclusive_scan_256(s, &total, sh);
  for (uint32_t i = lo; i < hi; i++) {
    uint32_t v = part[i];
    part[i] = run;
    run += v;
  }
  if (threadIdx.x == 0) part[m] = total;
}

static __global__ void __launch_bounds__(ZKB_SCAN_THREADS) scan_final_kernel(const uint32_t* __restrict__ in,
                                                                      uint32_t* __restrict__ out,
                                                                      const uint32_t* __restrict__ part,
                                                                      unsigned long long n, uint32_t nparts) {
  __shared__ uint32_t sh[32];
  unsigned long long base = (unsigned long long)blockIdx.x * ZKB_SCAN_TILE + threadIdx.x * ZKB_SCAN_ITEMS;
  uint32_t v[ZKB_SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ZKB_SCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = block_exclusive_scan_256(s, &total, sh) + part[blockIdx.x];
#pragma unroll
  for (int k = 0; k < ZKB_SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = part[nparts];
}

// ------------------------------------------------------------------------------------------------------
// piece planning: bucket b with references [start, start+cnt) is touched by the runs floor(start/K) .. floor((start+cnt-1)/K)
//   npieces[b] = that count (0 for an empty bucket), np_eff[b] = the same (rewritten to 1 by the folds),
//   run_bucket[t] = b for every run t whose first reference t*K lies in b,
//   buckets with more than ZKB_MSM_HOT / ZKB_MSM_VHOT pieces are appended to the hot / very hot list.
//   counters: [0] hot count, [1] accumulate work counter, [2] very hot count
// ------------------------------------------------------------------------------------------------------
static __global__ void msm_piece_plan_kernel(MsmPlan pl, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ start,
                                      uint32_t* __restrict__ npieces, uint32_t* __restrict__ np_eff,
                                      uint32_t* __restrict__ run_bucket, uint32_t* __restrict__ hot_list,
                                      uint32_t* __restrict__ vhot_list, uint32_t* __restrict__ counters) {
  unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (b >= (unsigned long long)pl.bwin * pl.nbuck) return;
  uint32_t n = cnt[b], st = start[b];
  uint32_t np = 0;
  if (n) {
    uint32_t first = st / pl.krun, last = (st + n - 1) / pl.krun;
    np = last - first + 1;
    uint32_t t0 = (st % pl.krun) ? first + 1 : first;   // runs that BEGIN inside this bucket
    for (uint32_t t = t0; t <= last; t++) run_bucket[t] = (uint32_t)b;
    if (np > ZKB_MSM_VHOT) vhot_list[atomicAdd(counters + 2, 1u)] = (uint32_t)b;
    else if (np > ZKB_MSM_HOT) hot_list[atomicAdd(counters, 1u)] = (uint32_t)b;
  }
  npieces[b] = np;
  np_eff[b] = np;
}

template <class F>
__device__ __forceinline__ Affine<F> load_affine(const Affine<F>* p) {
  // whole-struct copy through 16-byte vector loads (sizeof(Affine<F>) is a multiple of 16)
  Affine<F> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(Affine<F>) / 16); i++) d[i] = __ldg(q + i);
  return r;
}
template <class T>
__device__ __forceinline__ void store_vec(T* dst, const T& v) {
  const uint4* s = reinterpret_cast<const uint4*>(&v);
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}
template <class T>
__device__ __forceinline__ T load_vec(const T* src) {
  T r;
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
  return r;
}

// pass 3: persistent grid; a warp takes 32 consecutive runs at a time from `work`, lane l owns run 32*item + l, i.e. the
// sorted references [t*K, t*K + K).  F may be the inline-multiplier twin of the stored field (same layout).
template <class F, int MINB>
__global__ void __launch_bounds__(128, MINB) msm_accumulate_kernel(MsmPlan pl, const Affine<F>* __restrict__ points,
                                                                    const uint32_t* __restrict__ refs,
                                                                    const uint32_t* __restrict__ start,
                                                                    const uint32_t* __restrict__ pstart,
                                                                    const uint32_t* __restrict__ run_bucket,
                                                                    XYZZ<F>* __restrict__ pieces,
                                                                    unsigned int* __restrict__ work) {
  const unsigned long long nb = (unsigned long long)pl.bwin * pl.nbuck;
  const uint32_t total = start[nb];
  const uint32_t K = pl.krun;
  const uint32_t nruns = (total + K - 1) / K;
  const uint32_t nitems = (nruns + 31) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(work, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nitems) break;
    uint32_t t = (item << 5) + lane;
    if (t >= nruns) continue;
    uint32_t pos = t * K;
    uint32_t end = pos + K;
    if (end > total) end = total;
    uint32_t b = run_bucket[t];
    uint32_t bend = start[b + 1];
    uint32_t slot = pstart[b] + (t - start[b] / K);
    XYZZ<F> acc = XYZZ<F>::inf();
    // one-ahead software prefetch of the gathered point where the register file allows it (G1)
    constexpr bool PREFETCH = true;
    uint32_t ref_next = refs[pos];
    Affine<F> p_next;
    if (PREFETCH) p_next = load_affine(points + (ref_next & 0x7fffffffu));
    for (uint32_t p = pos; p < end; p++) {
      uint32_t ref = ref_next;
      Affine<F> pt;
      if (PREFETCH) {
        pt = p_next;
        if (p + 1 < end) {
          ref_next = refs[p + 1];
          p_next = load_affine(points + (ref_next & 0x7fffffffu));
        }
      } else {
        pt = load_affine(points + (ref & 0x7fffffffu));
        if (p + 1 < end) ref_next = refs[p + 1];
      }
      if (p == bend) {   // the run crosses into the next non-empty bucket: emit the finished piece
        store_vec(pieces + slot, acc);
        acc = XYZZ<F>::inf();
        do {
          b++;
          bend = start[b + 1];
        } while (bend == p);
        slot = pstart[b];
      }
      madd(acc, pt, (ref >> 31) != 0);
    }
    store_vec(pieces + slot, acc);
  }
}

// warp-shuffle of a whole struct (sizeof multiple of 4)
template <class T>
__device__ __forceinline__ T shfl_down_struct(const T& v, uint32_t delta) {
  T r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 4); i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta);
  return r;
}

// CTA-wide sum of one XYZZ per thread through shared memory (blockDim.x a power of two); result valid in thread 0
template <class F>
__device__ __forceinline__ XYZZ<F> block_sum(XYZZ<F> acc, XYZZ<F>* sh) {
  const uint32_t t = threadIdx.x;
  store_vec(sh + t, acc);
  __syncthreads();
  for (uint32_t off = blockDim.x >> 1; off > 0; off >>= 1) {
    if (t < off) {
      acc = add(acc, load_vec(sh + t + off));
      store_vec(sh + t, acc);
    }
    __syncthreads();
  }
  return acc;
}

// pass 4a: one warp per hot bucket: lane-strided sums, shuffle tree, result into the bucket's first piece slot
template <class F>
__global__ void __launch_bounds__(128) msm_fold_warp_kernel(const uint32_t* __restrict__ hot_list,
                                                            const uint32_t* __restrict__ counters,
                                                            uint32_t* __restrict__ np_eff, const uint32_t* __restrict__ pstart,
                                                            XYZZ<F>* __restrict__ pieces) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t nh = counters[0];
  for (uint32_t h = warp; h < nh; h += nwarps) {
    uint32_t b = hot_list[h];
    uint32_t ns = np_eff[b], st = pstart[b];
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t k = lane; k < ns; k += 32) acc = add(acc, load_vec(pieces + st + k));
#pragma unroll 1
    for (uint32_t off = 16; off > 0; off >>= 1) {
      XYZZ<F> other = shfl_down_struct(acc, off);
      if (lane < off) acc = add(acc, other);
    }
    if (lane == 0) {
      store_vec(pieces + st, acc);
      np_eff[b] = 1;
    }
  }
}

// pass 4b: very hot buckets: CTA (j, .) sums the j-th of ZKB_MSM_VHOT_SPLIT slices of the bucket's pieces into side[h][j]
template <class F>
__global__ void __launch_bounds__(128) msm_fold_cta1_kernel(const uint32_t* __restrict__ vhot_list,
                                                            const uint32_t* __restrict__ counters,
                                                            const uint32_t* __restrict__ np_eff,
                                                            const uint32_t* __restrict__ pstart,
                                                            const XYZZ<F>* __restrict__ pieces, XYZZ<F>* __restrict__ side) {
  extern __shared__ uint4 fold_smem[];
  XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(fold_smem);
  const uint32_t nv = counters[2];
  for (uint32_t h = blockIdx.y; h < nv; h += gridDim.y) {
    uint32_t b = vhot_list[h];
    uint32_t ns = np_eff[b], st = pstart[b];
    uint32_t per = (ns + ZKB_MSM_VHOT_SPLIT - 1) / ZKB_MSM_VHOT_SPLIT;
    uint32_t lo = blockIdx.x * per, hi = lo + per;
    if (hi > ns) hi = ns;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t k = lo + threadIdx.x; k < hi; k += blockDim.x) acc = add(acc, load_vec(pieces + st + k));
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) store_vec(side + (size_t)h * ZKB_MSM_VHOT_SPLIT + blockIdx.x, acc);
    __syncthreads();
  }
}
// pass 4c: one CTA of ZKB_MSM_VHOT_SPLIT threads per very hot bucket adds the slice sums
template <class F>
__global__ void __launch_bounds__(ZKB_MSM_VHOT_SPLIT) msm_fold_cta2_kernel(const uint32_t* __restrict__ vhot_list,
                                                                           const uint32_t* __restrict__ counters,
                                                                           uint32_t* __restrict__ np_eff,
                                                                           const uint32_t* __restrict__ pstart,
                                                                           XYZZ<F>* __restrict__ pieces,
                                                                           const XYZZ<F>* __restrict__ side) {
  extern __shared__ uint4 fold_smem[];
  XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(fold_smem);
  const uint32_t nv = counters[2];
  for (uint32_t h = blockIdx.x; h < nv; h += gridDim.x) {
    uint32_t b = vhot_list[h];
    XYZZ<F> acc = load_vec(side + (size_t)h * ZKB_MSM_VHOT_SPLIT + threadIdx.x);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) {
      store_vec(pieces + pstart[b], acc);
      np_eff[b] = 1;
    }
    __syncthreads();
  }
}

// pass 5, level 0: thread (w, j) folds the buckets [j*k, (j+1)*k) of window w:  R = sum_i S_i,  T = sum_i i * S_i
// (S_i = the bucket's <= ZKB_MSM_HOT pieces added up).  Outputs at [w * (nbuck/k) + j].
template <class F>
__global__ void __launch_bounds__(128, 3) msm_level0_kernel(MsmPlan pl, const uint32_t* __restrict__ np_eff,
                                                         const uint32_t* __restrict__ pstart,
                                                         const XYZZ<F>* __restrict__ pieces, XYZZ<F>* __restrict__ t_out,
                                                         XYZZ<F>* __restrict__ r_out) {
  unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const uint32_t logk = pl.logk[0];
  const uint32_t per_win = pl.nbuck >> logk;
  if (t >= (unsigned long long)pl.bwin * per_win) return;
  uint32_t w = (uint32_t)(t / per_win), j = (uint32_t)(t % per_win);
  unsigned long long b0 = (unsigned long long)w * pl.nbuck + ((unsigned long long)j << logk);
  XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
  for (int i = (1 << logk) - 1; i >= 0; i--) {
    uint32_t ns = np_eff[b0 + i], st = pstart[b0 + i];
    for (uint32_t k = 0; k < ns; k++) run = add(run, load_vec(pieces + st + k));
    if (i > 0) acc = add(acc, run);
  }
  store_vec(t_out + t, acc);
  store_vec(r_out + t, run);
}

// pass 5, level l > 0: the same fold over the previous level's totals
template <class F>
__global__ void __launch_bounds__(128, 3) msm_level_kernel(uint32_t nwin, uint32_t in_per_win, uint32_t logk,
                                                        const XYZZ<F>* __restrict__ r_in, XYZZ<F>* __restrict__ t_out,
                                                        XYZZ<F>* __restrict__ r_out) {
  unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const uint32_t per_win = in_per_win >> logk;
  if (t >= (unsigned long long)nwin * per_win) return;
  uint32_t w = (uint32_t)(t / per_win), j = (uint32_t)(t % per_win);
  const XYZZ<F>* src = r_in + (unsigned long long)w * in_per_win + ((unsigned long long)j << logk);
  XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
  for (int i = (1 << logk) - 1; i >= 0; i--) {
    run = add(run, load_vec(src + i));
    if (i > 0) acc = add(acc, run);
  }
  store_vec(t_out + t, acc);
  store_vec(r_out + t, run);
}

// pass 5b: plain sums.  Job q of window w adds base[w * stride + offset + j] over the j < count with bit `bit` of j set
// (bit < 0: all j) -> out[w * njobs + q].  grid (njobs, nwin).  Long U sums are cut into <= ZKB_MSM_MAXPARTS jobs (the host
// adds the parts) because the cost of a job is its DEPTH: count/THREADS serial additions plus log2(THREADS) tree levels.
template <class F>
struct SumJobs {
  const XYZZ<F>* base[ZKB_MSM_MAXJOBS];
  uint32_t stride[ZKB_MSM_MAXJOBS];
  uint32_t offset[ZKB_MSM_MAXJOBS];
  uint32_t count[ZKB_MSM_MAXJOBS];
  int bit[ZKB_MSM_MAXJOBS];
};
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS) msm_sums_kernel(SumJobs<F> jobs, uint32_t njobs, XYZZ<F>* __restrict__ out) {
  extern __shared__ uint4 ws_smem[];
  XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(ws_smem);
  const uint32_t t = threadIdx.x, q = blockIdx.x, w = blockIdx.y;
  const uint32_t cntq = jobs.count[q];
  const int bit = jobs.bit[q];
  const XYZZ<F>* src = jobs.base[q] + (unsigned long long)w * jobs.stride[q] + jobs.offset[q];
  XYZZ<F> acc = XYZZ<F>::inf();
  if (bit < 0) {
    for (uint32_t k = t; k < cntq; k += blockDim.x) acc = add(acc, load_vec(src + k));
  } else {
    // the indices with the bit set, enumerated densely: insert a 1 at position `bit` of m < count/2
    const uint32_t lowmask = (1u << bit) - 1;
    for (uint32_t m = t; m < (cntq >> 1); m += blockDim.x) {
      uint32_t k = ((m & ~lowmask) << 1) | (1u << bit) | (m & lowmask);
      acc = add(acc, load_vec(src + k));
    }
  }
  acc = block_sum(acc, sh);
  if (t == 0) store_vec(out + w * njobs + q, acc);
}

// ------------------------------------------------------------------------------------------------------
// point-vector kernels (setup side / format conversion)
// ------------------------------------------------------------------------------------------------------
// canonical affine coordinates -> Montgomery form, in place ((0,0) stays the infinity marker)
template <class F>
__global__ void points_to_mont_kernel(unsigned long long n, Affine<F>* pts) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = pts[i];
  p.x = to_mont(p.x);
  p.y = to_mont(p.y);
  pts[i] = p;
}
template <class F>
__global__ void points_from_mont_kernel(unsigned long long n, Affine<F>* pts) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = pts[i];
  p.x = from_mont(p.x);
  p.y = from_mont(p.y);
  pts[i] = p;
}

// fixed-base table: table[w * n + i] = 2^(c w) * pts[i] for w < W (affine, Montgomery; infinity stays (0,0))
template <class F>
__global__ void __launch_bounds__(128) msm_table_kernel(unsigned long long n, uint32_t c, uint32_t W,
                                                        const Affine<F>* __restrict__ pts, Affine<F>* __restrict__ table) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = pts[i];
  table[i] = p;
  XYZZ<F> acc = XYZZ<F>::from_affine(p);
  for (uint32_t w = 1; w < W; w++) {
    for (uint32_t k = 0; k < c; k++) acc = dbl(acc);
    table[(unsigned long long)w * n + i] = to_affine(acc);
  }
}

// out[i] = scalars[i] * bases[i or 0]  as affine Montgomery points -- replaces the per-point `g.point * Fr` loop of
// batch_multi_scalar_g1/g2 (/root/reference/src/bn254/curve.rs:326-354) used by Groth16/PlonK setup.
template <class F>
__global__ void __launch_bounds__(128) batch_scalar_mul_kernel(unsigned long long n, const Affine<F>* __restrict__ bases,
                                                               int single_base, const uint32_t* __restrict__ scalars,
                                                               Affine<F>* __restrict__ out) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8];
  load_scalar(scalars + i * 8, s);
  Affine<F> base = bases[single_base ? 0 : i];
  XYZZ<F> r = scalar_mul(base, s, 8);
  out[i] = to_affine(r);
}

}  // namespace zkb
