// msm.cuh -- Pippenger multi-scalar multiplication kernels (G1 and G2, BN254 and BLS12-381) for sm_100a.
//
// Replaces ark-ec 0.4.2 `VariableBaseMSM::msm` as called from /root/reference/src/bn254/curve.rs:356-392
// (bls12_381 twin :367-403).  The result is the group element sum_i s_i * P_i, independent of the algorithm;
// the pipeline here is
//   1. signed-digit decomposition of every scalar into W windows of c bits (digits in [-2^(c-1), 2^(c-1)]),
//   2. a counting sort (one radix pass keyed by (window, |digit|)) of the N*W (point, sign) references:
//      histogram -> exclusive scan -> warp-aggregated scatter,
//   3. bucket accumulation: every bucket is cut into segments of <= SEG references; one thread per segment
//      runs XYZZ mixed additions (8M+2S) over its gathered affine points,
//   4. hot buckets (many segments, e.g. all-ones witness vectors) are pre-reduced by one CTA each,
//   5. running-sum reduction of buckets to one point per window (chunks of K buckets per thread + a tree),
//   6. the W window sums are recombined (c doublings per window) by the host side of the C-ABI.
// Exact group law everywhere (identity operands, P+P, P-P), so results are bit-exact after affine conversion.
#pragma once
#include "ec.cuh"

namespace zkb {

struct MsmPlan {
  uint32_t c;        // window bits
  uint32_t nwin;     // W
  uint32_t nbuck;    // buckets per window = 2^(c-1)
  uint32_t seg;      // max references per segment
  uint32_t kchunk;   // buckets per thread in the running-sum reduction (power of two)
  unsigned long long n;         // points
  unsigned long long max_segs;  // upper bound on the segment count
};

// signed digit of window w for a canonical little-endian scalar (8 x u32); carry handled by recomputation:
// digit_w = raw_w + carry_{w-1}, carry_w = digit_w > 2^(c-1).  carry_{w-1} depends only on lower bits, and can be
// computed without a sequential scan: carry_{w-1} = 1 iff the scalar's low (w*c) bits, read as a number, are
// > 2^(w*c-1) ... which is NOT equivalent in general, so the kernels below walk the windows sequentially instead.
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

// pass 1: histogram of (window, bucket)
static __global__ void msm_count_kernel(MsmPlan pl, const uint32_t* __restrict__ scalars, uint32_t* __restrict__ cnt) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= pl.n) return;
  uint32_t s[8];
  load_scalar(scalars + i * 8, s);
  uint32_t carry = 0;
  const uint32_t half = 1u << (pl.c - 1);
  for (uint32_t w = 0; w < pl.nwin; w++) {
    uint32_t d = scalar_bits(s, w * pl.c, pl.c) + carry;
    carry = d > half;
    uint32_t mag = carry ? ((1u << pl.c) - d) : d;
    if (mag) atomicAdd(&cnt[w * pl.nbuck + mag - 1], 1u);
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
  for (uint32_t w = 0; w < pl.nwin; w++) {
    uint32_t d = scalar_bits(s, w * pl.c, pl.c) + carry;
    carry = d > half;
    uint32_t mag = carry ? ((1u << pl.c) - d) : d;
    bool have = live && mag != 0;
    // warp-aggregated atomics: lanes with the same bucket take consecutive slots from one atomicAdd
    uint32_t key = have ? (w * pl.nbuck + mag - 1) : 0xffffffffu;
    uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (have) {
      uint32_t leader = __ffs(peers) - 1;
      uint32_t rank = __popc(peers & ((1u << lane) - 1));
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      refs[base + rank] = (uint32_t)i | (carry << 31);
    }
  }
}

// single-CTA exclusive scan of n u32 (n up to a few million); also writes the total to out[n]
static __global__ void __launch_bounds__(1024) scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                    unsigned long long n) {
  __shared__ unsigned long long part[1024];
  const uint32_t t = threadIdx.x;
  unsigned long long per = (n + 1023) / 1024;
  unsigned long long lo = t * per, hi = lo + per;
  if (hi > n) hi = n;
  unsigned long long sum = 0;
  for (unsigned long long i = lo; i < hi; i++) sum += in[i];
  part[t] = sum;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials
  for (uint32_t off = 1; off < 1024; off <<= 1) {
    unsigned long long v = (t >= off) ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  unsigned long long run = part[t] - sum;
  for (unsigned long long i = lo; i < hi; i++) {
    uint32_t v = in[i];
    out[i] = (uint32_t)run;
    run += v;
  }
  if (t == 1023) out[n] = (uint32_t)part[1023];
}

// nseg[b] = ceil(cnt[b] / seg)
static __global__ void msm_nseg_kernel(MsmPlan pl, const uint32_t* __restrict__ cnt, uint32_t* __restrict__ nseg) {
  unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (b >= (unsigned long long)pl.nwin * pl.nbuck) return;
  nseg[b] = (cnt[b] + pl.seg - 1) / pl.seg;
}

// seg_bucket[segstart[b] + k] = b ; buckets with more than HOT segments are appended to the hot list
#define ZKB_MSM_HOT 8u
static __global__ void msm_segfill_kernel(MsmPlan pl, const uint32_t* __restrict__ nseg, const uint32_t* __restrict__ segstart,
                                   uint32_t* __restrict__ seg_bucket, uint32_t* __restrict__ hot_list,
                                   uint32_t* __restrict__ hot_count) {
  unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (b >= (unsigned long long)pl.nwin * pl.nbuck) return;
  uint32_t ns = nseg[b], st = segstart[b];
  for (uint32_t k = 0; k < ns; k++) seg_bucket[st + k] = (uint32_t)b;
  if (ns > ZKB_MSM_HOT) hot_list[atomicAdd(hot_count, 1u)] = (uint32_t)b;
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

// pass 3: one thread per segment
template <class F>
__global__ void __launch_bounds__(128) msm_accumulate_kernel(MsmPlan pl, const Affine<F>* __restrict__ points,
                                                             const uint32_t* __restrict__ refs,
                                                             const uint32_t* __restrict__ cnt,
                                                             const uint32_t* __restrict__ start,
                                                             const uint32_t* __restrict__ segstart,
                                                             const uint32_t* __restrict__ seg_bucket,
                                                             XYZZ<F>* __restrict__ seg_sum) {
  unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const unsigned long long nb = (unsigned long long)pl.nwin * pl.nbuck;
  uint32_t total = segstart[nb];
  if (s >= total) return;
  uint32_t b = seg_bucket[s];
  uint32_t k = (uint32_t)s - segstart[b];
  uint32_t lo = start[b] + k * pl.seg;
  uint32_t hi = start[b] + cnt[b];
  if (hi > lo + pl.seg) hi = lo + pl.seg;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t r = lo; r < hi; r++) {
    uint32_t ref = refs[r];
    Affine<F> p = load_affine(points + (ref & 0x7fffffffu));
    madd(acc, p, (ref >> 31) != 0);
  }
  seg_sum[s] = acc;
}

// pass 4: hot buckets -- one CTA folds all segment sums of a bucket into its first segment slot and
// rewrites nseg[b] = 1.  Tree in shared memory.
template <class F>
__global__ void __launch_bounds__(128) msm_hot_kernel(const uint32_t* __restrict__ hot_list,
                                                      const uint32_t* __restrict__ hot_count,
                                                      uint32_t* __restrict__ nseg, const uint32_t* __restrict__ segstart,
                                                      XYZZ<F>* __restrict__ seg_sum) {
  extern __shared__ uint4 hot_smem[];
  XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(hot_smem);
  const uint32_t t = threadIdx.x;
  uint32_t nh = *hot_count;
  for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
    uint32_t b = hot_list[h];
    uint32_t ns = nseg[b], st = segstart[b];
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t k = t; k < ns; k += blockDim.x) acc = add(acc, seg_sum[st + k]);
    sh[t] = acc;
    __syncthreads();
    for (uint32_t off = blockDim.x >> 1; off > 0; off >>= 1) {
      if (t < off) sh[t] = add(sh[t], sh[t + off]);
      __syncthreads();
    }
    if (t == 0) {
      seg_sum[st] = sh[0];
      nseg[b] = 1;
    }
    __syncthreads();
  }
}

// pass 5a: per chunk of K buckets: sum_b (b+1) * B_b  restricted to the chunk  ->  contrib[chunk]
template <class F>
__global__ void __launch_bounds__(128) msm_bucket_reduce_kernel(MsmPlan pl, const uint32_t* __restrict__ nseg,
                                                                const uint32_t* __restrict__ segstart,
                                                                const XYZZ<F>* __restrict__ seg_sum,
                                                                XYZZ<F>* __restrict__ contrib) {
  unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const uint32_t chunks_per_win = pl.nbuck / pl.kchunk;
  if (t >= (unsigned long long)pl.nwin * chunks_per_win) return;
  uint32_t w = (uint32_t)(t / chunks_per_win), j = (uint32_t)(t % chunks_per_win);
  uint32_t b0 = j * pl.kchunk;
  XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
  for (int idx = (int)pl.kchunk - 1; idx >= 0; idx--) {
    unsigned long long b = (unsigned long long)w * pl.nbuck + b0 + idx;
    uint32_t ns = nseg[b], st = segstart[b];
    for (uint32_t k = 0; k < ns; k++) run = add(run, seg_sum[st + k]);
    acc = add(acc, run);
  }
  // acc = sum (idx+1) B ; chunk offset adds b0 * run
  if (b0) acc = add(acc, mul_small(run, b0));
  contrib[t] = acc;
}

// pass 5b: one CTA per window: sum of its chunk contributions
template <class F>
__global__ void __launch_bounds__(128) msm_window_sum_kernel(uint32_t chunks_per_win, const XYZZ<F>* __restrict__ contrib,
                                                             XYZZ<F>* __restrict__ win_sum) {
  extern __shared__ uint4 ws_smem[];
  XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(ws_smem);
  const uint32_t t = threadIdx.x, w = blockIdx.x;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t k = t; k < chunks_per_win; k += blockDim.x) acc = add(acc, contrib[(unsigned long long)w * chunks_per_win + k]);
  sh[t] = acc;
  __syncthreads();
  for (uint32_t off = blockDim.x >> 1; off > 0; off >>= 1) {
    if (t < off) sh[t] = add(sh[t], sh[t + off]);
    __syncthreads();
  }
  if (t == 0) win_sum[w] = sh[0];
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
