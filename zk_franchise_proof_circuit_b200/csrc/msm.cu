// Batched fixed-base Pippenger MSM for the Groth16 prover.
//
// The bases of a proving key never change, so every base is expanded once into its window multiples
// 2^(c j) P (affine).  A 254-bit scalar then contributes ceil(254/c) signed c-bit digits that all land in
// ONE set of 2^(c-1) buckets:  sum_k s_k P_k = sum_b (b+1) * B_b,
// B_b = sum of +-2^(c j) P_k over the (k, j) whose |digit| is b+1.  c = 16 for the dense H MSM, c = 13
// for the four witness MSMs (few thousand non-zero scalars after the template difference).  Per batch item:
//   1. digits + counting sort of (k, j) by bucket      (k_digits<false>, k_scan, k_digits<true>)
//   2. one thread per bucket sums its list in XYZZ     (k_accumulate) - the IMAD-bound bulk
//   3. weighted bucket reduction, running sums over chunks of 32, then a warp (k_reduce1/2)
// A/B1/B2/C share the witness as scalars, so they share one sort; bases at infinity ((0,0)) are
// skipped by the mixed add.  Integer pipe (IMAD.WIDE carry chains) only - no tensor-core work here.
#include "msm.cuh"
#include "ec_coop.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

namespace zkb {

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

// level-1 reduction chunk: 2^5 buckets per thread in batch shape, 2^3 when a launch is small (latency shape)
static constexpr int RED_LOG = 5, RED_LOG_LAT = 3;

template <class T>
__device__ __forceinline__ T ldg_pod(const T *p) {
  static_assert(sizeof(T) % 16 == 0, "16-byte multiples");
  T r;
  const uint4 *s = reinterpret_cast<const uint4 *>(p);
  uint4 *d = reinterpret_cast<uint4 *>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
  return r;
}
template <class T>
__device__ __forceinline__ void stg_pod(T *p, const T &v) {
  const uint4 *s = reinterpret_cast<const uint4 *>(&v);
  uint4 *d = reinterpret_cast<uint4 *>(p);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}

// ---------------------------------------------------------------------------------------------
// table build
// ---------------------------------------------------------------------------------------------
template <class F>
__global__ void k_table(Affine<F> *tab, const Affine<F> *bases, uint32_t n, int c, int windows) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  tab += (size_t)blockIdx.y * windows * n;      // sub-range blockIdx.y: its own [windows][n] block
  bases += (size_t)blockIdx.y * n;
  Affine<F> p = ldg_pod(bases + k);
  XYZZ<F> acc = XYZZ<F>::from_affine(p);
  for (int j = 0; j < windows; j++) {
    Affine<F> a;
    xyzz_to_affine_ni(&acc, &a);
    stg_pod(tab + (size_t)j * n + k, a);
    if (j + 1 < windows)
      for (int i = 0; i < c; i++) xyzz_dbl_ni(&acc);
  }
}

template <class F>
cudaError_t msm_build_table(MsmTable<F> &t, const Affine<F> *bases, uint32_t n, MsmCfg cfg, cudaStream_t st,
                            uint32_t subs) {
  t.n = n;
  t.cfg = cfg;
  CK(cudaMalloc(&t.tab, (size_t)subs * n * cfg.windows * sizeof(Affine<F>)));
  k_table<F><<<dim3((n + 63) / 64, subs), 64, 0, st>>>(t.tab, bases, n, cfg.c, cfg.windows);
  return cudaGetLastError();
}
template cudaError_t msm_build_table<Fq>(MsmTable<Fq> &, const Affine<Fq> *, uint32_t, MsmCfg, cudaStream_t, uint32_t);
template cudaError_t msm_build_table<Fq2>(MsmTable<Fq2> &, const Affine<Fq2> *, uint32_t, MsmCfg, cudaStream_t, uint32_t);

// ---------------------------------------------------------------------------------------------
// digits + counting sort
// ---------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void k_digits(const Fr *scalars, size_t scalar_stride, uint32_t n, int c, int windows, uint32_t buckets,
                         uint32_t *counts_or_cursor, uint32_t *entries) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t b = blockIdx.y;
  if (k >= n) return;
  const uint4 *sp = reinterpret_cast<const uint4 *>(scalars + (size_t)b * scalar_stride + k);
  uint4 lo = sp[0], hi = sp[1];
  uint32_t w[9] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w, 0u};
  if ((w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) == 0) return;
  uint32_t *cc = counts_or_cursor + (size_t)b * buckets;
  uint32_t *ent = entries + (size_t)b * n * windows;
  const uint32_t mask = (1u << c) - 1, full = 1u << c;
  uint32_t carry = 0;
  for (int j = 0; j < windows; j++) {
    int bit = j * c, wd = bit >> 5, sh = bit & 31;
    uint64_t two = (uint64_t)w[wd] | ((uint64_t)(wd < 8 ? w[wd + 1] : 0u) << 32);
    uint32_t v = ((uint32_t)(two >> sh) & mask) + carry;     // in [0, 2^c]
    uint32_t neg = v > buckets;
    uint32_t mag = neg ? full - v : v;                         // |digit| in [0, 2^(c-1)]
    carry = neg;
    if (mag) {
      if (SCATTER) {
        uint32_t pos = atomicAdd(cc + (mag - 1), 1u);
        ent[pos] = ((uint32_t)j * n + k) | (neg << 31);
      } else {
        atomicAdd(cc + (mag - 1), 1u);
      }
    }
  }
}

// Variable-base form (no window table): batch item j * nsets + s holds window j of scalar set s, i.e. every scalar files
// ONE entry per item (its digit j, index k into the set's bases), so each window has its own bucket set and the window
// sums are combined afterwards by Horner's rule (k_horner).  Same signed digits as k_digits.
template <bool SCATTER>
__global__ void k_digits_vb(const Fr *scalars, size_t scalar_stride, uint32_t n, int c, int windows, uint32_t buckets,
                            uint32_t nsets, uint32_t *counts_or_cursor, uint32_t *entries) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t sset = blockIdx.y;
  if (k >= n) return;
  const uint4 *sp = reinterpret_cast<const uint4 *>(scalars + (size_t)sset * scalar_stride + k);
  uint4 lo = sp[0], hi = sp[1];
  uint32_t w[9] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w, 0u};
  if ((w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) == 0) return;
  const uint32_t mask = (1u << c) - 1, full = 1u << c;
  uint32_t carry = 0;
  for (int j = 0; j < windows; j++) {
    int bit = j * c, wd = bit >> 5, sh = bit & 31;
    uint64_t two = (uint64_t)w[wd] | ((uint64_t)(wd < 8 ? w[wd + 1] : 0u) << 32);
    uint32_t v = ((uint32_t)(two >> sh) & mask) + carry;
    uint32_t neg = v > buckets;
    uint32_t mag = neg ? full - v : v;
    carry = neg;
    if (mag) {
      const size_t item = (size_t)j * nsets + sset;
      uint32_t *cc = counts_or_cursor + item * buckets;
      if (SCATTER) {
        uint32_t pos = atomicAdd(cc + (mag - 1), 1u);
        entries[item * n + pos] = k | (neg << 31);
      } else {
        atomicAdd(cc + (mag - 1), 1u);
      }
    }
  }
}

// out[s] = sum_j 2^(c j) * sum_t in[j * nsets * fold + s * fold + t]   (one thread per set; fold = sub-ranges per set)
__global__ void k_horner(const XYZZ<Fq> *in, XYZZ<Fq> *out, uint32_t nsets, uint32_t fold, int windows, int c) {
  uint32_t sidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= nsets) return;
  XYZZ<Fq> acc = XYZZ<Fq>::infinity();
  for (int j = windows - 1; j >= 0; j--) {
    for (int i = 0; i < c; i++) xyzz_dbl_ni(&acc);
    for (uint32_t t = 0; t < fold; t++) xyzz_add_ni(&acc, in + ((size_t)j * nsets * fold + (size_t)sidx * fold + t));
  }
  stg_pod(out + sidx, acc);
}
cudaError_t msm_horner(const XYZZ<Fq> *in, XYZZ<Fq> *out, uint32_t nsets, uint32_t fold, int windows, int c, cudaStream_t st) {
  k_horner<<<(nsets + 31) / 32, 32, 0, st>>>(in, out, nsets, fold, windows, c);
  return cudaGetLastError();
}

// offsets[b][0..buckets] = exclusive scan of counts[b]; cursor[b] = offsets[b][0..buckets).  One CTA of 1024
// threads per batch item, each thread owning buckets/1024 consecutive counters (<= 32).
__global__ void __launch_bounds__(1024) k_scan(const uint32_t *counts, uint32_t *offsets, uint32_t *cursor,
                                               uint32_t buckets) {
  __shared__ uint32_t warp_tot[32];
  const uint32_t b = blockIdx.x, t = threadIdx.x, per = buckets / 1024;
  const uint32_t *c = counts + (size_t)b * buckets + t * per;
  uint32_t loc[32], sum = 0;
#pragma unroll
  for (int i = 0; i < 32; i++)
    if ((uint32_t)i < per) { loc[i] = sum; sum += c[i]; }
  uint32_t x = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if ((t & 31) >= d) x += y;
  }
  if ((t & 31) == 31) warp_tot[t >> 5] = x;
  __syncthreads();
  if (t < 32) {
    uint32_t v = warp_tot[t], z = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, z, d);
      if (t >= d) z += y;
    }
    warp_tot[t] = z - v;
  }
  __syncthreads();
  uint32_t base = warp_tot[t >> 5] + x - sum;
  uint32_t *o = offsets + (size_t)b * (buckets + 1) + t * per;
  uint32_t *cu = cursor + (size_t)b * buckets + t * per;
#pragma unroll
  for (int i = 0; i < 32; i++)
    if ((uint32_t)i < per) { o[i] = base + loc[i]; cu[i] = base + loc[i]; }
  if (t == 1023) offsets[(size_t)b * (buckets + 1) + buckets] = base + sum;
}

// order[b][i] = bucket ids of batch item b sorted by descending size (counting sort on the size, sizes >= 1023
// share the first bin).  The accumulate kernel walks buckets in this order, so the 32 lanes of a warp get lists of
// (nearly) equal length and the longest lists start first.
// Lists longer than `long_bucket` are summed by a whole warp, at most `max_long` of them per batch item (the rest
// stay with one thread).  Throughput shape (many batch items): 128 / 256 - only the pathological lists.  Latency
// shape (fewer than 2^18 buckets in the launch): 24 / 2048 - a warp costs 5 extra full additions for its shuffle fold but the longest serial
// chain drops from 128 mixed adds to ~1.
static constexpr uint32_t LONG_BUCKET = 128, MAX_LONG = 256;
static constexpr uint32_t LONG_BUCKET_LAT = 24, MAX_LONG_LAT = 2048;

__global__ void __launch_bounds__(1024) k_order(const uint32_t *counts, uint32_t *order, uint32_t *n_long,
                                                uint32_t buckets, uint32_t long_bucket, uint32_t max_long) {
  __shared__ uint32_t hist[1024];
  const uint32_t b = blockIdx.x, t = threadIdx.x;
  const uint32_t *c = counts + (size_t)b * buckets;
  hist[t] = 0;
  __syncthreads();
  for (uint32_t i = t; i < buckets; i += 1024) atomicAdd(&hist[1023u - min(c[i], 1023u)], 1u);   // bin 0 = largest
  __syncthreads();
  // exclusive scan of hist (1024 entries, one per thread)
  uint32_t v = hist[t], x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if ((t & 31) >= d) x += y;
  }
  __shared__ uint32_t wt[32];
  if ((t & 31) == 31) wt[t >> 5] = x;
  __syncthreads();
  if (t < 32) {
    uint32_t w = wt[t], z = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, z, d);
      if (t >= d) z += y;
    }
    wt[t] = z - w;
  }
  __syncthreads();
  hist[t] = wt[t >> 5] + x - v;
  __syncthreads();
  if (t == 1023u - long_bucket) n_long[b] = min(hist[t], max_long);   // buckets in bins before t: size > long_bucket
  __syncthreads();
  uint32_t *o = order + (size_t)b * buckets;
  for (uint32_t i = t; i < buckets; i += 1024) {
    uint32_t pos = atomicAdd(&hist[1023u - min(c[i], 1023u)], 1u);
    o[pos] = i;
  }
}

// kernels one MsmSort::run launches.  (A single-CTA-per-item variant with the 2^(c-1) counters in shared memory was
// tried: on B200 the 2 x 2.1 M shared-memory atomics per H sort made it 2x slower than these L2-atomic passes.)
int msm_sort_launches() { return 4; }

cudaError_t MsmSort::alloc(uint32_t n_, uint32_t batch_, MsmCfg cfg_) {
  n = n_;
  batch = batch_;
  cfg = cfg_;
  if (cfg.buckets < 2048 || cfg.buckets > 32768) return cudaErrorInvalidValue;
  CK(cudaMalloc(&counts, (size_t)batch * cfg.buckets * 4));
  CK(cudaMalloc(&offsets, (size_t)batch * (cfg.buckets + 1) * 4));
  CK(cudaMalloc(&cursor, (size_t)batch * cfg.buckets * 4));
  CK(cudaMalloc(&entries, (size_t)batch * n * cfg.windows * 4));
  CK(cudaMalloc(&order, (size_t)batch * cfg.buckets * 4));
  CK(cudaMalloc(&n_long, (size_t)batch * 4));
  return cudaSuccess;
}
void MsmSort::free_all() {
  cudaFree(counts); cudaFree(offsets); cudaFree(cursor); cudaFree(entries); cudaFree(order); cudaFree(n_long);
  counts = offsets = cursor = entries = order = n_long = nullptr;
}
// variable-base digits: nsets scalar sets -> nsets * windows batch items (item j * nsets + s = window j of set s);
// this sort's cfg must have windows == 1 (one entry per scalar per item)
cudaError_t MsmSort::run_vb(const Fr *scalars, size_t scalar_stride, uint32_t nsets, int windows, cudaStream_t st) {
  const uint32_t nitems = nsets * (uint32_t)windows;
  if (nitems > batch || cfg.windows != 1) return cudaErrorInvalidValue;
  CK(cudaMemsetAsync(counts, 0, (size_t)nitems * cfg.buckets * 4, st));
  dim3 grid((n + 255) / 256, nsets);
  k_digits_vb<false><<<grid, 256, 0, st>>>(scalars, scalar_stride, n, cfg.c, windows, cfg.buckets, nsets, counts, nullptr);
  k_scan<<<nitems, 1024, 0, st>>>(counts, offsets, cursor, cfg.buckets);
  const bool lat = (uint64_t)nitems * cfg.buckets < 262144;
  max_long = lat ? MAX_LONG_LAT : MAX_LONG;
  if (max_long > cfg.buckets / 2) max_long = cfg.buckets / 2;
  k_order<<<nitems, 1024, 0, st>>>(counts, order, n_long, cfg.buckets, lat ? LONG_BUCKET_LAT : LONG_BUCKET, max_long);
  k_digits_vb<true><<<grid, 256, 0, st>>>(scalars, scalar_stride, n, cfg.c, windows, cfg.buckets, nsets, cursor, entries);
  return cudaGetLastError();
}

cudaError_t MsmSort::run(const Fr *scalars, size_t scalar_stride, uint32_t nbatch, cudaStream_t st) {
  if (nbatch > batch) return cudaErrorInvalidValue;
  CK(cudaMemsetAsync(counts, 0, (size_t)nbatch * cfg.buckets * 4, st));
  dim3 grid((n + 255) / 256, nbatch);
  k_digits<false><<<grid, 256, 0, st>>>(scalars, scalar_stride, n, cfg.c, cfg.windows, cfg.buckets, counts, nullptr);
  k_scan<<<nbatch, 1024, 0, st>>>(counts, offsets, cursor, cfg.buckets);
  const bool lat = (uint64_t)nbatch * cfg.buckets < 262144;   // fewer bucket threads than ~1 wave of the GPU
  max_long = lat ? MAX_LONG_LAT : MAX_LONG;
  if (max_long > cfg.buckets / 2) max_long = cfg.buckets / 2;
  k_order<<<nbatch, 1024, 0, st>>>(counts, order, n_long, cfg.buckets, lat ? LONG_BUCKET_LAT : LONG_BUCKET, max_long);
  k_digits<true><<<grid, 256, 0, st>>>(scalars, scalar_stride, n, cfg.c, cfg.windows, cfg.buckets, cursor, entries);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// bucket accumulation: thread = (bucket, table, batch item)
// ---------------------------------------------------------------------------------------------
template <class F>
struct TablePtrs { const Affine<F> *tab[4]; };

// Warp-convergent: every lane of a warp runs the same number of iterations (the longest list in the warp) and the
// mixed add is computed branch-free with the result selected per lane, so lanes never drift apart.  (The first
// version returned early from the add for empty accumulators / infinity bases; with independent thread scheduling
// the lanes then ran their lists one after another - ncu showed 2.4 of 32 threads active per instruction.)
// The P == +-acc cases are handled in a rare slow path taken only when some lane of the warp needs it.
// INL: 0 = products through the out-of-line calls, 1 = all inlined, 2 = all inlined with squares as plain products,
// 3 = products inlined, squares through the out-of-line 36-product squaring (smaller loop body; the default)
template <int INL, class F> __device__ __forceinline__ F mulx(const F &a, const F &b) {
  if constexpr (INL != 0) return a * b; else return a.mulc(b);
}
// squares: Fq2 has a 2-product complex squaring (3 for a general product); Fq has a 36-limb-product square (fp.cuh)
template <int INL, class F> __device__ __forceinline__ F sqrx(const F &a) {
  if constexpr (INL == 2) return a * a; else if constexpr (INL == 1) return a.sqr(); else return a.sqrc();
}

template <class F, int THREADS, int MINB, int INL>
__global__ void __launch_bounds__(THREADS, MINB) k_accumulate(TablePtrs<F> tabs, int ntab, uint32_t n, int windows,
                                                        uint32_t nbuckets, const uint32_t *__restrict__ offsets,
                                                        const uint32_t *__restrict__ entries,
                                                        const uint32_t *__restrict__ order,
                                                        const uint32_t *__restrict__ n_long, XYZZ<F> *buckets,
                                                        size_t tab_batch_stride, uint32_t tab_mod) {
  uint32_t t = blockIdx.y, b = blockIdx.z;
  const uint32_t pos = blockIdx.x * THREADS + threadIdx.x;
  if ((pos | 31u) < n_long[b]) return;     // this whole warp's buckets belong to k_accumulate_long
  const bool mine = pos >= n_long[b];
  const uint32_t bucket = order[(size_t)b * nbuckets + pos];
  const Affine<F> *tab = (t == 0 ? tabs.tab[0] : (t == 1 ? tabs.tab[1] : (t == 2 ? tabs.tab[2] : tabs.tab[3]))) +
                         (size_t)(b % tab_mod) * tab_batch_stride;
  const uint32_t *off = offsets + (size_t)b * (nbuckets + 1);
  const uint32_t *ent = entries + (size_t)b * n * windows;
  const uint32_t beg = off[bucket], len = mine ? off[bucket + 1] - beg : 0u;
  const uint32_t maxlen = __reduce_max_sync(0xffffffffu, len);
  XYZZ<F> acc = XYZZ<F>::infinity();
  bool acc_inf = true;
  for (uint32_t i = 0; i < maxlen; i++) {
    const bool active = i < len;
    const uint32_t x = active ? ent[beg + i] : 0u;
    Affine<F> p = ldg_pod(tab + (x & 0x7fffffffu));
    const bool use = active && !p.is_inf();
    p.y = F::select((x >> 31) != 0, p.y.neg(), p.y);
    // madd-2008-s, unconditionally
    F U2 = mulx<INL>(p.x, acc.ZZ), S2 = mulx<INL>(p.y, acc.ZZZ);
    F P = U2 - acc.X, R = S2 - acc.Y;
    F PP = sqrx<INL>(P), PPP = mulx<INL>(P, PP), Q = mulx<INL>(acc.X, PP);
    F X3 = sqrx<INL>(R) - PPP - Q.dbl();
    F Y3 = mulx<INL>(R, Q - X3) - mulx<INL>(acc.Y, PPP);
    F ZZ3 = mulx<INL>(acc.ZZ, PP), ZZZ3 = mulx<INL>(acc.ZZZ, PPP);
    const bool special = use && !acc_inf && P.is_zero();
    const bool normal = use && !acc_inf && !special, first = use && acc_inf;
    acc.X = F::select(normal, X3, F::select(first, p.x, acc.X));
    acc.Y = F::select(normal, Y3, F::select(first, p.y, acc.Y));
    acc.ZZ = F::select(normal, ZZ3, F::select(first, F::one(), acc.ZZ));
    acc.ZZZ = F::select(normal, ZZZ3, F::select(first, F::one(), acc.ZZZ));
    acc_inf = acc_inf && !first;
    if (__any_sync(0xffffffffu, special)) {
      if (special) {
        if (R.is_zero()) { acc = XYZZ<F>::dbl_affine(p); }
        else { acc = XYZZ<F>::infinity(); acc_inf = true; }
      }
      __syncwarp();
    }
  }
  if (acc_inf) acc = XYZZ<F>::infinity();
  if (mine) stg_pod(buckets + ((size_t)(b * ntab + t) * nbuckets + bucket), acc);
}

// G2 bucket accumulation on lane PAIRS: lane 0 of a pair holds (X, ZZ) of the bucket's XYZZ<Fq2> accumulator and the x
// coordinate of the incoming point, lane 1 holds (Y, ZZZ) and the y coordinate.  The ten Fq2 products of madd-2008-s
// split 5 / 5 with both lanes running the SAME instruction stream on different operands:
//   step 1   U2 = x ZZ, P = U2 - X        |  S2 = y ZZZ, R = S2 - Y
//   step 2   PP = P^2                     |  RR = R^2
//   step 3   PPP = P PP                   |  Q = X PP                (X, PP from the partner)
//   step 4   ZZ' = ZZ PP                  |  ZZZ' = ZZZ PPP          (PPP from the partner)
//   step 5   Y PPP                        |  X' = RR - PPP - 2Q,  R (Q - X')
//   finish   X' from the partner          |  Y' = R (Q - X') - Y PPP
// Four 16-word exchanges per addition, no extra products (28 Fq products per addition in total, as on one lane), and
// each lane keeps 2 instead of 4 Fq2 accumulators: the one-lane kernel needed 128 registers + 576 B of stack, i.e. its
// accumulator lived in local memory (ncu: 11 GB of DRAM writes per launch against 0.13 GB of bucket stores).
__device__ __forceinline__ Fq2 pair_xchg(const Fq2 &v, uint32_t mask = 0xffffffffu) {
  Fq2 r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.a.v[i] = __shfl_xor_sync(mask, v.a.v[i], 1);
    r.b.v[i] = __shfl_xor_sync(mask, v.b.v[i], 1);
  }
  return r;
}

template <int THREADS, int MINB, int INL, bool SMB>
__global__ void __launch_bounds__(THREADS, MINB) k_accumulate_g2pair(TablePtrs<Fq2> tabs, int ntab, uint32_t n, int windows,
                                                                   uint32_t nbuckets, const uint32_t *__restrict__ offsets,
                                                                   const uint32_t *__restrict__ entries,
                                                                   const uint32_t *__restrict__ order,
                                                                   const uint32_t *__restrict__ n_long, XYZZ<Fq2> *buckets,
                                                                   size_t tab_batch_stride, uint32_t tab_mod) {
  const uint32_t t = blockIdx.y, b = blockIdx.z;
  const uint32_t tid = blockIdx.x * THREADS + threadIdx.x, pos = tid >> 1;
  const bool h = tid & 1u;                      // 0: (X, ZZ, x)   1: (Y, ZZZ, y)
  if ((pos | 15u) < n_long[b]) return;          // the 16 buckets of this warp all belong to k_accumulate_long
  const bool mine = pos >= n_long[b];
  const uint32_t bucket = order[(size_t)b * nbuckets + pos];
  const Affine<Fq2> *tab = (t == 0 ? tabs.tab[0] : (t == 1 ? tabs.tab[1] : (t == 2 ? tabs.tab[2] : tabs.tab[3]))) +
                           (size_t)(b % tab_mod) * tab_batch_stride;
  const uint32_t *off = offsets + (size_t)b * (nbuckets + 1);
  const uint32_t *ent = entries + (size_t)b * n * windows;
  const uint32_t beg = off[bucket], len = mine ? off[bucket + 1] - beg : 0u;
  const uint32_t maxlen = __reduce_max_sync(0xffffffffu, len);
  Fq2 A = Fq2::zero(), B = Fq2::zero();         // (X | Y), (ZZ | ZZZ)
  // SMB: (ZZ | ZZZ) lives in shared memory between its two uses of an iteration (word-interleaved, conflict-free),
  // which takes 16 registers off the live set
  __shared__ uint32_t smB[SMB ? 16 * THREADS : 1];
  auto putB = [&](const Fq2 &v) {
#pragma unroll
    for (int w = 0; w < 8; w++) { smB[w * THREADS + threadIdx.x] = v.a.v[w]; smB[(8 + w) * THREADS + threadIdx.x] = v.b.v[w]; }
  };
  auto getB = [&]() {
    Fq2 v;
#pragma unroll
    for (int w = 0; w < 8; w++) { v.a.v[w] = smB[w * THREADS + threadIdx.x]; v.b.v[w] = smB[(8 + w) * THREADS + threadIdx.x]; }
    return v;
  };
  if (SMB) putB(B);
  bool acc_inf = true;
  for (uint32_t i = 0; i < maxlen; i++) {
    const bool active = i < len;
    const uint32_t x = active ? ent[beg + i] : 0u;
    Fq2 pc = ldg_pod(reinterpret_cast<const Fq2 *>(tab + (x & 0x7fffffffu)) + (h ? 1 : 0));
    if (SMB) B = getB();
    const bool myzero = pc.is_zero();
    const int other_zero = __shfl_xor_sync(0xffffffffu, (int)myzero, 1);   // (not inside the &&: every lane must shuffle)
    const bool p_inf = myzero && other_zero != 0;
    const bool use = active && !p_inf;
    pc = Fq2::select(h && (x >> 31) != 0, pc.neg(), pc);
    // step 1
    const Fq2 D = mulx<INL>(pc, B) - A;         // P | R
    const bool dz = D.is_zero();
    const bool Pz = __shfl_sync(0xffffffffu, (int)dz, (threadIdx.x & 31u) & ~1u);
    const bool special = use && !acc_inf && Pz;
    const bool normal = use && !acc_inf && !special, first = use && acc_inf;
    // an empty accumulator takes the point itself; what the steps below compute from it is discarded (normal = false)
    A = Fq2::select(first, pc, A);
    // step 2
    const Fq2 DD = sqrx<INL>(D);                // PP | RR
    Fq2 M3;
    {
      const Fq2 oDD = pair_xchg(DD), oA = pair_xchg(A);     // lane 1 now has PP and X
      // step 3
      M3 = mulx<INL>(Fq2::select(h, oA, D), Fq2::select(h, oDD, DD));           // PPP | Q
    }
    Fq2 X3;
    {
      const Fq2 oM3 = pair_xchg(M3);            // Q | PPP
      // step 4
      if (SMB) B = getB();
      const Fq2 B3 = mulx<INL>(B, Fq2::select(h, oM3, DD));                      // ZZ PP | ZZZ PPP
      B = Fq2::select(normal, B3, Fq2::select(first, Fq2::one(), B));
      if (SMB) putB(B);
      X3 = DD - oM3 - M3.dbl();                 // lane 1: RR - PPP - 2Q
    }
    // step 5
    Fq2 M5;
    {
      const Fq2 oA = pair_xchg(A);              // lane 0 now has Y (exchanged a second time: 16 registers fewer live)
      M5 = mulx<INL>(Fq2::select(h, D, oA), Fq2::select(h, M3 - X3, M3));       // Y PPP | R (Q - X')
    }
    {
      const Fq2 o5 = pair_xchg(Fq2::select(h, X3, M5));     // lane 0 receives X', lane 1 receives Y PPP
      A = Fq2::select(normal, Fq2::select(h, M5 - o5, o5), A);                  // X' | Y'
    }
    acc_inf = acc_inf && !first;
    if (__any_sync(0xffffffffu, special)) {     // P == +-acc: rare, pair-uniform
      if (special) {
        const uint32_t pm = 3u << ((threadIdx.x & 31u) & ~1u);
        const bool Rz = __shfl_sync(pm, (int)dz, (threadIdx.x & 31u) | 1u);
        Fq2 pcr = ldg_pod(reinterpret_cast<const Fq2 *>(tab + (x & 0x7fffffffu)) + (h ? 1 : 0));   // the point again
        pcr = Fq2::select(h && (x >> 31) != 0, pcr.neg(), pcr);
        const Fq2 po = pair_xchg(pcr, pm);
        if (Rz) {
          Affine<Fq2> pt;
          pt.x = h ? po : pcr;
          pt.y = h ? pcr : po;
          XYZZ<Fq2> d = XYZZ<Fq2>::dbl_affine(pt);
          A = h ? d.Y : d.X;
          B = h ? d.ZZZ : d.ZZ;
        } else {
          A = Fq2::zero(); B = Fq2::zero(); acc_inf = true;
        }
        if (SMB) putB(B);
      }
      __syncwarp();
    }
  }
  if (SMB) B = getB();
  if (acc_inf) { A = Fq2::zero(); B = Fq2::zero(); }
  if (mine) {
    Fq2 *dst = reinterpret_cast<Fq2 *>(buckets + ((size_t)(b * ntab + t) * nbuckets + bucket));
    stg_pod(dst + (h ? 1 : 0), A);
    stg_pod(dst + (h ? 3 : 2), B);
  }
}

// One warp per long bucket (the n_long[b] largest): lanes take entries l, l + 32, ... with the same branch-free
// mixed add, then the 32 partial sums are folded with shuffles.  Long lists come from many scalars sharing a digit
// (e.g. hundreds of 0/1-valued wires that flip with respect to the template all have the difference +-1).
template <class F>
__global__ void __launch_bounds__(32) k_accumulate_long(TablePtrs<F> tabs, int ntab, uint32_t n, int windows,
                                                        uint32_t nbuckets, const uint32_t *__restrict__ offsets,
                                                        const uint32_t *__restrict__ entries,
                                                        const uint32_t *__restrict__ order,
                                                        const uint32_t *__restrict__ n_long, XYZZ<F> *buckets,
                                                        size_t tab_batch_stride, uint32_t tab_mod) {
  const uint32_t t = blockIdx.y, b = blockIdx.z, lane = threadIdx.x;
  if (blockIdx.x >= n_long[b]) return;
  const Affine<F> *tab = (t == 0 ? tabs.tab[0] : (t == 1 ? tabs.tab[1] : (t == 2 ? tabs.tab[2] : tabs.tab[3]))) +
                         (size_t)(b % tab_mod) * tab_batch_stride;
  const uint32_t bucket = order[(size_t)b * nbuckets + blockIdx.x];
  const uint32_t *off = offsets + (size_t)b * (nbuckets + 1);
  const uint32_t *ent = entries + (size_t)b * n * windows;
  const uint32_t beg = off[bucket], len = off[bucket + 1] - beg;
  XYZZ<F> acc = XYZZ<F>::infinity();
  bool acc_inf = true;
  for (uint32_t i = lane; i < len + lane; i += 32) {      // same trip count on every lane
    const bool active = i < len;
    const uint32_t x = active ? ent[beg + i] : 0u;
    Affine<F> p = ldg_pod(tab + (x & 0x7fffffffu));
    const bool use = active && !p.is_inf();
    p.y = F::select((x >> 31) != 0, p.y.neg(), p.y);
    F U2 = p.x.mulc(acc.ZZ), S2 = p.y.mulc(acc.ZZZ);
    F P = U2 - acc.X, R = S2 - acc.Y;
    const bool special = use && !acc_inf && P.is_zero();
    F PP = P.sqrc(), PPP = P.mulc(PP), Q = acc.X.mulc(PP);
    F X3 = R.sqrc() - PPP - Q.dbl();
    F Y3 = R.mulc(Q - X3) - acc.Y.mulc(PPP);
    F ZZ3 = acc.ZZ.mulc(PP), ZZZ3 = acc.ZZZ.mulc(PPP);
    const bool normal = use && !acc_inf && !special, first = use && acc_inf;
    acc.X = F::select(normal, X3, F::select(first, p.x, acc.X));
    acc.Y = F::select(normal, Y3, F::select(first, p.y, acc.Y));
    acc.ZZ = F::select(normal, ZZ3, F::select(first, F::one(), acc.ZZ));
    acc.ZZZ = F::select(normal, ZZZ3, F::select(first, F::one(), acc.ZZZ));
    acc_inf = acc_inf && !first;
    if (__any_sync(0xffffffffu, special)) {
      if (special) {
        if (R.is_zero()) { acc = XYZZ<F>::dbl_affine(p); }
        else { acc = XYZZ<F>::infinity(); acc_inf = true; }
      }
      __syncwarp();
    }
  }
  if (acc_inf) acc = XYZZ<F>::infinity();
  constexpr int WORDS = sizeof(XYZZ<F>) / 4;
  for (int d = 16; d >= 1; d >>= 1) {
    XYZZ<F> other;
    uint32_t *src = reinterpret_cast<uint32_t *>(&acc), *dst = reinterpret_cast<uint32_t *>(&other);
#pragma unroll
    for (int w = 0; w < WORDS; w++) dst[w] = __shfl_down_sync(0xffffffffu, src[w], d);
    if (lane < (uint32_t)d) xyzz_add_ni(&acc, &other);
    __syncwarp();
  }
  if (lane == 0) stg_pod(buckets + ((size_t)(b * ntab + t) * nbuckets + bucket), acc);
}

// ---------------------------------------------------------------------------------------------
// batched-affine pair tree (see MsmAffineWs in msm.cuh)
// ---------------------------------------------------------------------------------------------
// One CTA per batch item: lvl_off[l][b][0..B] = exclusive scan of len_{l+1}[j] = ceil(len_l[j] / 2), l < levels,
// starting from the sorted-entry offsets (len_0).  1024 threads, buckets / 1024 <= 32 counters each.
__global__ void __launch_bounds__(1024) k_affine_scans(const uint32_t *offsets, uint32_t *lvl_off, uint32_t buckets,
                                                       uint32_t nbatch, int levels) {
  __shared__ uint32_t warp_tot[32];
  const uint32_t b = blockIdx.x, t = threadIdx.x, per = buckets / 1024;
  const uint32_t *o0 = offsets + (size_t)b * (buckets + 1) + t * per;
  uint32_t len[32];
#pragma unroll
  for (int i = 0; i < 32; i++)
    if ((uint32_t)i < per) len[i] = o0[i + 1] - o0[i];
  for (int l = 0; l < levels; l++) {
    uint32_t loc[32], sum = 0;
#pragma unroll
    for (int i = 0; i < 32; i++)
      if ((uint32_t)i < per) { len[i] = (len[i] + 1) >> 1; loc[i] = sum; sum += len[i]; }
    uint32_t x = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if ((t & 31) >= d) x += y;
    }
    __syncthreads();                         // warp_tot of the previous level is no longer read
    if ((t & 31) == 31) warp_tot[t >> 5] = x;
    __syncthreads();
    if (t < 32) {
      uint32_t v = warp_tot[t], z = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, z, d);
        if (t >= d) z += y;
      }
      warp_tot[t] = z - v;
    }
    __syncthreads();
    const uint32_t base = warp_tot[t >> 5] + x - sum;
    uint32_t *o = lvl_off + ((size_t)l * nbatch + b) * (buckets + 1) + t * per;
#pragma unroll
    for (int i = 0; i < 32; i++)
      if ((uint32_t)i < per) o[i] = base + loc[i];
    if (t == 1023) lvl_off[((size_t)l * nbatch + b) * (buckets + 1) + buckets] = base + sum;
  }
}

__device__ __forceinline__ Fq ldg_fq(const Fq *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  const uint4 a = __ldg(q), b = __ldg(q + 1);
  Fq r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void stg_fq(Fq *p, const Fq &x) {
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

// What one output of a level is made of.  kind: 0 = sum of two distinct points (divisor x2 - x1), 1 = copy of a single
// point (odd tail of a list), 2 = first operand at infinity (take the second), 3 = second at infinity (take the first),
// 4 = doubling (equal points, divisor 2 y1), 5 = opposite points (result at infinity).  Kinds 2..5 only occur with
// degenerate bases (repeated or at infinity); they cost a branch that is never taken for a real proving key.
struct PairSrc {
  const Affine<Fq> *p1, *p2;
  bool neg1, neg2, has2;
};
template <bool FIRST>
__device__ __forceinline__ PairSrc pair_src(const Affine<Fq> *tab, const uint32_t *ent, const Affine<Fq> *pin, uint32_t in0,
                                            bool has2) {
  PairSrc s;
  s.has2 = has2;
  if (FIRST) {
    const uint32_t e1 = ent[in0], e2 = has2 ? ent[in0 + 1] : 0u;
    s.p1 = tab + (e1 & 0x7fffffffu);
    s.p2 = tab + (e2 & 0x7fffffffu);
    s.neg1 = (e1 >> 31) != 0;
    s.neg2 = (e2 >> 31) != 0;
  } else {
    s.p1 = pin + in0;
    s.p2 = pin + in0 + (has2 ? 1 : 0);
    s.neg1 = s.neg2 = false;
  }
  return s;
}
// divisor of the pair (Montgomery one when the output needs no division) and its kind
__device__ __forceinline__ Fq pair_divisor(const PairSrc &s, int &kind) {
  if (!s.has2) { kind = 1; return Fq::one(); }
  const Fq x1 = ldg_fq(&s.p1->x), x2 = ldg_fq(&s.p2->x);
  Fq d = x2 - x1;
  kind = 0;
  if (x1.is_zero() || x2.is_zero() || d.is_zero()) {       // rare: look at y
    Fq y1 = ldg_fq(&s.p1->y), y2 = ldg_fq(&s.p2->y);
    if (s.neg1) y1 = y1.neg();
    if (s.neg2) y2 = y2.neg();
    if (x1.is_zero() && y1.is_zero()) { kind = 2; d = Fq::one(); }
    else if (x2.is_zero() && y2.is_zero()) { kind = 3; d = Fq::one(); }
    else if (d.is_zero()) {
      if (y1 == y2 && !y1.is_zero()) { kind = 4; d = y1.dbl(); }
      else { kind = 5; d = Fq::one(); }
    }
  }
  return d;
}

// thread = `group` consecutive outputs of level l + 1 of batch item blockIdx.y
template <bool FIRST>
__global__ void __launch_bounds__(128) k_affine_level(const Affine<Fq> *tab, size_t tab_batch_stride, uint32_t tab_mod,
                                                      uint32_t item0, const uint32_t *__restrict__ entries, size_t ent_stride,
                                                      const Affine<Fq> *pin, size_t pin_stride,
                                                      const uint32_t *__restrict__ off_in_all,
                                                      const uint32_t *__restrict__ off_out_all, Affine<Fq> *pout_all,
                                                      size_t pout_stride, Fq *park_all, size_t park_stride,
                                                      uint32_t nbuckets, uint32_t group) {
  const uint32_t b = blockIdx.y;
  const uint32_t *off_in = off_in_all + (size_t)b * (nbuckets + 1), *off_out = off_out_all + (size_t)b * (nbuckets + 1);
  const uint32_t total = off_out[nbuckets];
  const uint32_t o0 = (blockIdx.x * 128u + threadIdx.x) * group;
  if (o0 >= total) return;
  const uint32_t cnt = min(group, total - o0);
  const Affine<Fq> *tb = FIRST ? tab + (size_t)((item0 + b) % tab_mod) * tab_batch_stride : nullptr;
  const uint32_t *ent = FIRST ? entries + (size_t)b * ent_stride : nullptr;
  const Affine<Fq> *pi = FIRST ? nullptr : pin + (size_t)b * pin_stride;
  Affine<Fq> *pout = pout_all + (size_t)b * pout_stride;
  Fq *park = park_all + (size_t)b * park_stride;
  // bucket of the first output: the j with off_out[j] <= o0 < off_out[j + 1]
  uint32_t lo = 0, hi = nbuckets;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (off_out[mid] <= o0) lo = mid; else hi = mid;
  }
  uint32_t j = lo, k = o0 - off_out[lo];
  const uint32_t j0 = j, k0 = k;
  uint32_t out_end = off_out[j + 1] - off_out[j];          // outputs of bucket j
  uint32_t in_beg = off_in[j], in_len = off_in[j + 1] - in_beg;
  // pass 1: prefix products of the divisors
  Fq acc = Fq::one();
  for (uint32_t i = 0; i < cnt; i++) {
    while (k == out_end) {                                 // next non-empty bucket
      j++; k = 0;
      out_end = off_out[j + 1] - off_out[j];
      in_beg = off_in[j]; in_len = off_in[j + 1] - in_beg;
    }
    const PairSrc s = pair_src<FIRST>(tb, ent, pi, in_beg + 2 * k, 2 * k + 1 < in_len);
    int kind;
    const Fq d = pair_divisor(s, kind);
    acc = acc * d;
    stg_fq(park + o0 + i, acc);
    k++;
  }
  Fq inv = acc.inv();
  // pass 2, backwards: 1 / d_i = inv(d_0 .. d_i) * (d_0 .. d_{i-1})
  k--;                                                     // (j, k) = last output of the range
  for (uint32_t i = cnt; i-- > 0;) {
    const PairSrc s = pair_src<FIRST>(tb, ent, pi, in_beg + 2 * k, 2 * k + 1 < in_len);
    int kind;
    const Fq d = pair_divisor(s, kind);
    const Fq prev = i ? park[o0 + i - 1] : Fq::one();     // written by this thread in pass 1: plain (coherent) load
    const Fq dinv = inv * prev;
    inv = inv * d;
    Affine<Fq> r;
    Fq x1 = ldg_fq(&s.p1->x), y1 = ldg_fq(&s.p1->y);
    if (s.neg1) y1 = y1.neg();
    if (kind == 0 || kind == 4) {
      Fq x2, lam;
      if (kind == 0) {
        x2 = ldg_fq(&s.p2->x);
        Fq y2 = ldg_fq(&s.p2->y);
        if (s.neg2) y2 = y2.neg();
        lam = (y2 - y1) * dinv;
      } else {
        x2 = x1;
        const Fq xx = x1 * x1;
        lam = (xx.dbl() + xx) * dinv;
      }
      r.x = lam * lam - x1 - x2;
      r.y = lam * (x1 - r.x) - y1;
    } else if (kind == 1 || kind == 3) {
      r.x = x1; r.y = y1;
    } else if (kind == 2) {
      r.x = ldg_fq(&s.p2->x);
      r.y = ldg_fq(&s.p2->y);
      if (s.neg2) r.y = r.y.neg();
    } else {
      r.x = Fq::zero(); r.y = Fq::zero();
    }
    stg_pod(pout + o0 + i, r);
    if (i) {
      if (k == 0) {                                        // previous non-empty bucket (never before (j0, k0))
        do {
          j--;
          out_end = off_out[j + 1] - off_out[j];
        } while (out_end == 0);
        k = out_end;
        in_beg = off_in[j]; in_len = off_in[j + 1] - in_beg;
      }
      k--;
    }
  }
  (void)j0; (void)k0;
}

// XYZZ tail: thread = bucket, summing its remaining affine points (list of the last level)
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_accumulate_pts(const Affine<Fq> *pts_all, size_t pts_stride,
                                                                  const uint32_t *__restrict__ off_all,
                                                                  const uint32_t *__restrict__ order, uint32_t nbuckets,
                                                                  XYZZ<Fq> *buckets) {
  const uint32_t b = blockIdx.y;
  const uint32_t pos = blockIdx.x * THREADS + threadIdx.x;
  const uint32_t bucket = order[(size_t)b * nbuckets + pos];
  const uint32_t *off = off_all + (size_t)b * (nbuckets + 1);
  const Affine<Fq> *pts = pts_all + (size_t)b * pts_stride;
  const uint32_t beg = off[bucket], len = off[bucket + 1] - beg;
  const uint32_t maxlen = __reduce_max_sync(0xffffffffu, len);
  XYZZ<Fq> acc = XYZZ<Fq>::infinity();
  bool acc_inf = true;
  for (uint32_t i = 0; i < maxlen; i++) {
    const bool active = i < len;
    Affine<Fq> p = ldg_pod(pts + (active ? beg + i : 0u));
    const bool use = active && !p.is_inf();
    Fq U2 = p.x * acc.ZZ, S2 = p.y * acc.ZZZ;
    Fq P = U2 - acc.X, R = S2 - acc.Y;
    const bool special = use && !acc_inf && P.is_zero();
    Fq PP = P * P, PPP = P * PP, Q = acc.X * PP;
    Fq X3 = R * R - PPP - Q.dbl();
    Fq Y3 = R * (Q - X3) - acc.Y * PPP;
    Fq ZZ3 = acc.ZZ * PP, ZZZ3 = acc.ZZZ * PPP;
    const bool normal = use && !acc_inf && !special, first = use && acc_inf;
    acc.X = Fq::select(normal, X3, Fq::select(first, p.x, acc.X));
    acc.Y = Fq::select(normal, Y3, Fq::select(first, p.y, acc.Y));
    acc.ZZ = Fq::select(normal, ZZ3, Fq::select(first, Fq::one(), acc.ZZ));
    acc.ZZZ = Fq::select(normal, ZZZ3, Fq::select(first, Fq::one(), acc.ZZZ));
    acc_inf = acc_inf && !first;
    if (__any_sync(0xffffffffu, special)) {
      if (special) {
        if (R.is_zero()) { acc = XYZZ<Fq>::dbl_affine(p); }
        else { acc = XYZZ<Fq>::infinity(); acc_inf = true; }
      }
      __syncwarp();
    }
  }
  if (acc_inf) acc = XYZZ<Fq>::infinity();
  stg_pod(buckets + ((size_t)b * nbuckets + bucket), acc);
}

cudaError_t MsmAffineWs::alloc(uint32_t n_entries, uint32_t batch_, MsmCfg cfg) {
  batch = batch_;
  buckets = cfg.buckets;
  if (buckets < 1024 || buckets % 1024 || buckets / 1024 > 32) return cudaErrorInvalidValue;
  cap_a = ((size_t)n_entries + buckets) / 2 + 1;
  cap_b = (cap_a + buckets) / 2 + 1;
  const char *lv = getenv("ZKB_AFFINE_LEVELS");
  if (lv && atoi(lv) >= 1 && atoi(lv) <= MAX_LEVELS) levels = atoi(lv);
  const char *gr = getenv("ZKB_AFFINE_GROUP");   // "512" or "512,256,128"
  if (gr) {
    uint32_t last = 512;
    for (int l = 0; l < MAX_LEVELS; l++) {
      if (gr && *gr) { last = (uint32_t)atoi(gr); gr = strchr(gr, ','); if (gr) gr++; }
      group[l] = last >= 8 ? last : 8;
    }
  }
  CK(cudaMalloc(&pa, (size_t)batch * cap_a * sizeof(Affine<Fq>)));
  CK(cudaMalloc(&pb, (size_t)batch * cap_b * sizeof(Affine<Fq>)));
  CK(cudaMalloc(&park, (size_t)batch * cap_a * sizeof(Fq)));
  CK(cudaMalloc(&lvl_off, (size_t)MAX_LEVELS * batch * (buckets + 1) * 4));
  return cudaSuccess;
}
void MsmAffineWs::free_all() {
  cudaFree(pa); cudaFree(pb); cudaFree(park); cudaFree(lvl_off);
  pa = pb = nullptr; park = nullptr; lvl_off = nullptr;
}
int msm_affine_launches(const MsmAffineWs &ws, uint32_t nbatch) {
  const uint32_t subs = ws.batch ? (nbatch + ws.batch - 1) / ws.batch : 1;
  return (int)subs * (1 + ws.levels + 1);
}

cudaError_t msm_accumulate_affine(const MsmSort &sort, const MsmTable<Fq> &table, uint32_t nbatch, MsmWork<Fq> &work,
                                  uint32_t slot0, MsmAffineWs &ws, cudaStream_t st, size_t tab_batch_stride,
                                  uint32_t tab_mod) {
  if (tab_mod == 0) tab_mod = nbatch;
  if (!ws.batch || slot0 + nbatch > work.slots || ws.buckets != sort.cfg.buckets) return cudaErrorInvalidValue;
  if (table.n != sort.n || table.cfg.c != sort.cfg.c || work.cfg.c != sort.cfg.c) return cudaErrorInvalidValue;
  const uint32_t nb = sort.cfg.buckets;
  const size_t ent_stride = (size_t)sort.n * sort.cfg.windows, lstride = (size_t)ws.batch * (nb + 1);
  // the workspace holds ws.batch items: a larger launch runs as consecutive sub-batches on the same stream
  for (uint32_t b0 = 0; b0 < nbatch; b0 += ws.batch) {
    const uint32_t cnt = nbatch - b0 < ws.batch ? nbatch - b0 : ws.batch;
    const uint32_t *offs = sort.offsets + (size_t)b0 * (nb + 1);
    k_affine_scans<<<cnt, 1024, 0, st>>>(offs, ws.lvl_off, nb, ws.batch, ws.levels);
    size_t cap_in = ent_stride;
    for (int l = 0; l < ws.levels; l++) {
      const uint32_t *off_in = l ? ws.lvl_off + (size_t)(l - 1) * lstride : offs;
      const size_t off_in_stride = nb + 1;
      (void)off_in_stride;
      const uint32_t *off_out = ws.lvl_off + (size_t)l * lstride;
      const Affine<Fq> *pin = (l & 1) ? ws.pa : ws.pb;
      Affine<Fq> *pout = (l & 1) ? ws.pb : ws.pa;
      const size_t pin_stride = (l & 1) ? ws.cap_a : ws.cap_b, pout_stride = (l & 1) ? ws.cap_b : ws.cap_a;
      const size_t cap_out = (cap_in + nb) / 2 + 1;           // upper bound of this level's outputs per item
      const uint32_t g = ws.group[l];
      dim3 grid((unsigned)((cap_out + (size_t)g * 128 - 1) / ((size_t)g * 128)), cnt);
      if (l == 0)
        k_affine_level<true><<<grid, 128, 0, st>>>(table.tab, tab_batch_stride, tab_mod, b0, sort.entries + (size_t)b0 * ent_stride,
                                                    ent_stride, nullptr, 0, off_in, off_out, pout, pout_stride,
                                                    reinterpret_cast<Fq *>(ws.pb), ws.cap_b * 2, nb, g);
      else
        k_affine_level<false><<<grid, 128, 0, st>>>(nullptr, 0, 1, 0, nullptr, 0, pin, pin_stride, off_in, off_out, pout,
                                                     pout_stride, ws.park, ws.cap_b, nb, g);
      cap_in = cap_out;
    }
    const int last = ws.levels - 1;
    const Affine<Fq> *pts = (last & 1) ? ws.pb : ws.pa;
    const size_t pts_stride = (last & 1) ? ws.cap_b : ws.cap_a;
    k_accumulate_pts<128, 4><<<dim3(nb / 128, cnt), 128, 0, st>>>(pts, pts_stride, ws.lvl_off + (size_t)last * lstride,
                                                                 sort.order + (size_t)b0 * nb, nb,
                                                                 work.buckets + ((size_t)slot0 + b0) * nb);
  }
  return cudaGetLastError();
}

// executed work of the pair tree over the current sort (measurement aid, host arithmetic over the downloaded bucket
// offsets): out[0] = affine additions = sum over levels and buckets of floor(len_l / 2); out[1] = mixed additions of the
// XYZZ tail = sum over buckets of (len_levels - 1)+; out[2] = field inversions = threads of the level kernels
cudaError_t msm_affine_counts(const MsmSort &sort, const MsmAffineWs &ws, uint32_t nbatch, unsigned long long *out3,
                              cudaStream_t st) {
  const uint32_t nb = sort.cfg.buckets;
  std::vector<uint32_t> off((size_t)nbatch * (nb + 1));
  CK(cudaMemcpyAsync(off.data(), sort.offsets, off.size() * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  unsigned long long adds = 0, tail = 0, inv = 0;
  for (uint32_t b = 0; b < nbatch; b++) {
    const uint32_t *o = off.data() + (size_t)b * (nb + 1);
    unsigned long long total[MsmAffineWs::MAX_LEVELS] = {};
    for (uint32_t k = 0; k < nb; k++) {
      uint32_t len = o[k + 1] - o[k];
      for (int l = 0; l < ws.levels; l++) { adds += len >> 1; len = (len + 1) >> 1; total[l] += len; }
      if (len > 1) tail += len - 1;
    }
    for (int l = 0; l < ws.levels; l++) inv += (total[l] + ws.group[l] - 1) / ws.group[l];
  }
  out3[0] = adds; out3[1] = tail; out3[2] = inv;
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// bucket reduction  sum_i (i+1) B_i
// ---------------------------------------------------------------------------------------------
// COOP = true (small launches, latency shape): every "thread" below is a group of 4 lanes that shares each point
// addition / doubling (ec_coop.cuh), e.g. 5 instead of 14 product times for a G1 addition.
template <bool COOP, class F>
__device__ __forceinline__ void gadd(XYZZ<F> *a, const XYZZ<F> *b, const LaneGroup &g) {
  if constexpr (COOP) coop_add(a, b, g); else xyzz_add_ni(a, b);
}
template <bool COOP, class F>
__device__ __forceinline__ void gdbl(XYZZ<F> *a, const LaneGroup &g) {
  if constexpr (COOP) coop_dbl(a, g); else xyzz_dbl_ni(a);
}

// level 1: thread = chunk of 2^red_log buckets:  R = sum (u+1) B[chunk + u],  S = sum B[chunk + u]
template <class F, int THREADS, bool COOP>
__global__ void __launch_bounds__(THREADS) k_reduce1(const XYZZ<F> *buckets, XYZZ<F> *part_r, XYZZ<F> *part_s,
                                                     uint32_t nbuckets, int red_log) {
  const uint32_t t = (blockIdx.x * THREADS + threadIdx.x) >> (COOP ? 2 : 0);   // chunk
  const LaneGroup g = LaneGroup::of(threadIdx.x & 31);
  const uint32_t slot = blockIdx.y;
  const uint32_t parts = nbuckets >> red_log;
  const XYZZ<F> *B = buckets + (size_t)slot * nbuckets + ((size_t)t << red_log);
  XYZZ<F> run = XYZZ<F>::infinity(), acc = XYZZ<F>::infinity();
  for (int u = (1 << red_log) - 1; u >= 0; u--) {
    XYZZ<F> x = ldg_pod(B + u);
    gadd<COOP>(&run, &x, g);
    __syncwarp();
    gadd<COOP>(&acc, &run, g);
    __syncwarp();
  }
  if (!COOP || g.l == 0) {
    stg_pod(part_r + (size_t)slot * parts + t, acc);
    stg_pod(part_s + (size_t)slot * parts + t, run);
  }
}

// level 2: one CTA of min(parts, 256) threads per slot.  total = sum_t R_t + 2^red_log * sum_t t * S_t: a thread folds its
// span of parts, forms t * sigma_t by double-and-add over the bits of t, and both sums are folded by a shared-memory
// tree, so the depth is ~4 log2(threads) group operations instead of a serial pass over the parts.
template <bool COOP, class F>
__device__ void block_tree_sum(XYZZ<F> *sm, XYZZ<F> &v, uint32_t t, uint32_t parts, const LaneGroup &g) {
  const bool writer = !COOP || g.l == 0;
  if (writer) stg_pod(sm + t, v);
  __syncthreads();
  for (uint32_t stride = parts / 2; stride >= 1; stride >>= 1) {
    if (t < stride) {
      XYZZ<F> x = sm[t + stride];
      gadd<COOP>(&v, &x, g);
      if (writer) stg_pod(sm + t, v);
    }
    __syncthreads();
  }
}

template <class F, bool COOP>
__global__ void __launch_bounds__(COOP ? 512 : 256) k_reduce2(const XYZZ<F> *part_r, const XYZZ<F> *part_s, XYZZ<F> *out,
                                                              uint32_t parts, int red_log) {
  extern __shared__ uint4 sm_raw[];
  XYZZ<F> *sm = reinterpret_cast<XYZZ<F> *>(sm_raw);
  const LaneGroup g = LaneGroup::of(threadIdx.x & 31);
  const uint32_t slot = blockIdx.x, t = threadIdx.x >> (COOP ? 2 : 0), nthr = blockDim.x >> (COOP ? 2 : 0);
  const uint32_t span = parts / nthr;   // a power of two
  const XYZZ<F> *R = part_r + (size_t)slot * parts + t * span, *S = part_s + (size_t)slot * parts + t * span;
  // own span: r = sum R, sigma = sum S, rho = sum_u u * S_u
  XYZZ<F> r = XYZZ<F>::infinity(), sigma = XYZZ<F>::infinity(), rho = XYZZ<F>::infinity(), x;
  for (uint32_t u = 0; u < span; u++) {
    x = ldg_pod(R + u);
    gadd<COOP>(&r, &x, g);
  }
  for (uint32_t u = span - 1; u >= 1; u--) {
    x = ldg_pod(S + u);
    gadd<COOP>(&sigma, &x, g);
    gadd<COOP>(&rho, &sigma, g);
  }
  x = ldg_pod(S);
  gadd<COOP>(&sigma, &x, g);
  block_tree_sum<COOP, F>(sm, r, t, nthr, g);      // thread 0: sum_t R_t
  // sum over the span of (t*span + u) * S = span * (t * sigma) + rho
  XYZZ<F> w = XYZZ<F>::infinity();
  for (int bit = 31 - __clz((int)nthr) - 1; bit >= 0; bit--) {   // t < nthr
    gdbl<COOP>(&w, g);
    if ((t >> bit) & 1u) gadd<COOP>(&w, &sigma, g);
  }
  for (uint32_t s = span; s > 1; s >>= 1) gdbl<COOP>(&w, g);
  gadd<COOP>(&w, &rho, g);
  __syncthreads();
  block_tree_sum<COOP, F>(sm, w, t, nthr, g);      // thread 0: sum_t t * S_t
  if (t == 0) {
    for (int i = 0; i < red_log; i++) gdbl<COOP>(&w, g);   // * chunk size
    gadd<COOP>(&w, &r, g);
    if (!COOP || g.l == 0) stg_pod(out + slot, w);
  }
}

template <class F>
cudaError_t MsmWork<F>::alloc(uint32_t slots_, MsmCfg cfg_) {
  slots = slots_;
  cfg = cfg_;
  CK(cudaMalloc(&buckets, (size_t)slots * cfg.buckets * sizeof(XYZZ<F>)));
  CK(cudaMalloc(&part_r, (size_t)slots * cfg.parts() * sizeof(XYZZ<F>)));
  CK(cudaMalloc(&part_s, (size_t)slots * cfg.parts() * sizeof(XYZZ<F>)));
  return cudaSuccess;
}
template <class F>
void MsmWork<F>::free_all() {
  cudaFree(buckets); cudaFree(part_r); cudaFree(part_s);
  buckets = part_r = part_s = nullptr;
}
template struct MsmWork<Fq>;
template struct MsmWork<Fq2>;

template <class F> struct AccCfg;
template <> struct AccCfg<Fq> { static constexpr int THREADS = 128; static constexpr int RED_THREADS = 64; };
template <> struct AccCfg<Fq2> { static constexpr int THREADS = 64; static constexpr int RED_THREADS = 32; };

template <class F>
cudaError_t msm_accumulate(const MsmSort &sort, const MsmTable<F> *tables, int ntab, uint32_t nbatch, MsmWork<F> &work,
                           uint32_t slot0, cudaStream_t st, size_t tab_batch_stride, uint32_t tab_mod) {
  if (tab_mod == 0) tab_mod = nbatch;
  if (ntab < 1 || ntab > 4 || slot0 + nbatch * (uint32_t)ntab > work.slots) return cudaErrorInvalidValue;
  TablePtrs<F> tp;
  for (int i = 0; i < 4; i++) tp.tab[i] = i < ntab ? tables[i].tab : nullptr;
  for (int i = 0; i < ntab; i++)
    if (tables[i].n != sort.n || tables[i].cfg.c != sort.cfg.c) return cudaErrorInvalidValue;
  if (work.cfg.c != sort.cfg.c) return cudaErrorInvalidValue;
  constexpr int TH = AccCfg<F>::THREADS;
  const uint32_t nb = sort.cfg.buckets;
  dim3 grid(nb / TH, ntab, nbatch);
  XYZZ<F> *dst = work.buckets + (size_t)slot0 * nb;
  static const int variant = getenv("ZKB_ACC_VARIANT") ? atoi(getenv("ZKB_ACC_VARIANT")) : 0;
#define ZKB_ACC(MINB, INL) k_accumulate<F, TH, MINB, INL><<<grid, TH, 0, st>>>(tp, ntab, sort.n, sort.cfg.windows, nb, sort.offsets, sort.entries, sort.order, sort.n_long, dst, tab_batch_stride, tab_mod)
  if constexpr (sizeof(F) == 32) {
    switch (variant) {
      case 1: ZKB_ACC(4, false); break;
      case 2: ZKB_ACC(5, false); break;
      case 3: ZKB_ACC(3, true); break;
      case 4: ZKB_ACC(5, true); break;
      case 5: ZKB_ACC(4, 2); break;          // everything inlined, squares as plain products (the default until the squaring existed)
      case 6: ZKB_ACC(4, 3); break;          // inlined products, squares through the out-of-line call
      case 7: ZKB_ACC(4, 1); break;          // everything inlined, squares by Fp::sqr_dev
      default: ZKB_ACC(4, 3); break;         // measured best on B200: inlined products, the two squares through the
                                             // out-of-line 36-product squaring (128 registers, no stack)
    }
  } else {
    static const int variant2 = getenv("ZKB_ACC_VARIANT_G2") ? atoi(getenv("ZKB_ACC_VARIANT_G2")) : 0;
    dim3 gpair(nb * 2 / 128, ntab, nbatch);
#define ZKB_ACC2(MINB, INL, SMB) k_accumulate_g2pair<128, MINB, INL, SMB><<<gpair, 128, 0, st>>>(tp, ntab, sort.n, sort.cfg.windows, nb, sort.offsets, sort.entries, sort.order, sort.n_long, dst, tab_batch_stride, tab_mod)
    switch (variant2) {
      case 1: ZKB_ACC(4, false); break;      // one lane per bucket (round 1): 128 registers + 576 B of stack
      case 2: ZKB_ACC(8, false); break;
      case 3: ZKB_ACC2(4, false, true); break;
      case 4: ZKB_ACC2(3, true, true); break;
      case 5: ZKB_ACC2(4, true, true); break;
      case 6: ZKB_ACC2(2, false, false); break;
      case 7: ZKB_ACC2(3, false, false); break;
      default: ZKB_ACC2(3, false, true); break;    // lane pairs
    }
#undef ZKB_ACC2
  }
#undef ZKB_ACC
  dim3 glong(sort.max_long, ntab, nbatch);
  k_accumulate_long<F><<<glong, 32, 0, st>>>(tp, ntab, sort.n, sort.cfg.windows, nb, sort.offsets, sort.entries, sort.order,
                                             sort.n_long, dst, tab_batch_stride, tab_mod);
  return cudaGetLastError();
}

// out[i] = sum_b (b+1) * buckets[slot0 + i][b], i < nslots
template <class F>
cudaError_t msm_reduce(MsmWork<F> &work, uint32_t slot0, uint32_t nslots, XYZZ<F> *out, cudaStream_t st) {
  if (slot0 + nslots > work.slots) return cudaErrorInvalidValue;
  constexpr int RT = AccCfg<F>::RED_THREADS;
  const uint32_t nb = work.cfg.buckets;
  if ((uint64_t)nslots * nb < 262144) {
    // latency shape: lane groups of 4 share every point operation; chunks of 8 buckets (32 for the 2^15-bucket H sums)
    const int red_log = nb > 8192 ? RED_LOG : RED_LOG_LAT;
    const uint32_t parts = nb >> red_log, nthr = parts < 128 ? parts : 128;
    k_reduce1<F, 128, true><<<dim3(parts * 4 / 128, nslots), 128, 0, st>>>(
        work.buckets + (size_t)slot0 * nb, work.part_r + (size_t)slot0 * parts, work.part_s + (size_t)slot0 * parts, nb, red_log);
    k_reduce2<F, true><<<nslots, nthr * 4, nthr * sizeof(XYZZ<F>), st>>>(work.part_r + (size_t)slot0 * parts,
                                                                       work.part_s + (size_t)slot0 * parts, out, parts, red_log);
    return cudaGetLastError();
  }
  const int red_log = RED_LOG;
  const uint32_t parts = nb >> red_log;
  dim3 g1(parts / RT, nslots);
  k_reduce1<F, RT, false><<<g1, RT, 0, st>>>(work.buckets + (size_t)slot0 * nb, work.part_r + (size_t)slot0 * parts,
                                             work.part_s + (size_t)slot0 * parts, nb, red_log);
  const uint32_t nthr = parts < 256 ? parts : 256;
  static bool attr_set[64] = {};            // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(k_reduce2<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * (int)sizeof(XYZZ<F>));
    attr_set[dev & 63] = true;
  }
  k_reduce2<F, false><<<nslots, nthr, nthr * sizeof(XYZZ<F>), st>>>(work.part_r + (size_t)slot0 * parts,
                                                                     work.part_s + (size_t)slot0 * parts, out, parts, red_log);
  return cudaGetLastError();
}

// out[p * ntab + t] = sum over sub-ranges s of in[(p * subs + s) * ntab + t]   (one thread per output; subs is small)
template <class F>
__global__ void k_fold_subs(const XYZZ<F> *in, XYZZ<F> *out, uint32_t subs, uint32_t ntab, uint32_t total) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint32_t p = i / ntab, t = i % ntab;
  XYZZ<F> acc = XYZZ<F>::infinity();
  for (uint32_t s = 0; s < subs; s++) xyzz_add_ni(&acc, in + ((size_t)(p * subs + s) * ntab + t));
  stg_pod(out + i, acc);
}
template <class F>
cudaError_t msm_fold_subs(const XYZZ<F> *in, XYZZ<F> *out, uint32_t nproofs, uint32_t subs, uint32_t ntab, cudaStream_t st) {
  const uint32_t total = nproofs * ntab;
  k_fold_subs<F><<<(total + 31) / 32, 32, 0, st>>>(in, out, subs, ntab, total);
  return cudaGetLastError();
}
template cudaError_t msm_fold_subs<Fq>(const XYZZ<Fq> *, XYZZ<Fq> *, uint32_t, uint32_t, uint32_t, cudaStream_t);
template cudaError_t msm_fold_subs<Fq2>(const XYZZ<Fq2> *, XYZZ<Fq2> *, uint32_t, uint32_t, uint32_t, cudaStream_t);

template <class F>
cudaError_t msm_run(const MsmSort &sort, const MsmTable<F> *tables, int ntab, uint32_t nbatch, MsmWork<F> &work,
                    XYZZ<F> *out, cudaStream_t st) {
  cudaError_t e = msm_accumulate<F>(sort, tables, ntab, nbatch, work, 0, st, 0, 0);
  if (e != cudaSuccess) return e;
  return msm_reduce<F>(work, 0, nbatch * (uint32_t)ntab, out, st);
}

// executed mixed adds of one accumulate launch: per bucket (non-infinity entries - 1)+, summed
__global__ void k_count_madds(const uint8_t *inf_mask, uint32_t n, int windows, uint32_t nbuckets, const uint32_t *offsets,
                              const uint32_t *entries, unsigned long long *total) {
  uint32_t bucket = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t b = blockIdx.y;
  const uint32_t *off = offsets + (size_t)b * (nbuckets + 1);
  const uint32_t *ent = entries + (size_t)b * n * windows;
  uint32_t cnt = 0;
  for (uint32_t e = off[bucket]; e < off[bucket + 1]; e++) cnt += inf_mask[(ent[e] & 0x7fffffffu) % n] ? 0u : 1u;
  if (cnt > 1) atomicAdd(total, (unsigned long long)(cnt - 1));
}

template <class F>
__global__ void k_inf_mask(const Affine<F> *tab, uint32_t n, uint8_t *mask) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) mask[k] = ldg_pod(tab + k).is_inf() ? 1 : 0;
}

template <class F>
cudaError_t msm_count_madds(const MsmSort &sort, const MsmTable<F> &table, uint32_t nbatch, unsigned long long *host_total,
                            cudaStream_t st) {
  uint8_t *mask = nullptr;
  unsigned long long *dtot = nullptr;
  CK(cudaMalloc(&mask, table.n));
  CK(cudaMalloc(&dtot, 8));
  CK(cudaMemsetAsync(dtot, 0, 8, st));
  k_inf_mask<F><<<(table.n + 255) / 256, 256, 0, st>>>(table.tab, table.n, mask);
  dim3 grid(sort.cfg.buckets / 128, nbatch);
  k_count_madds<<<grid, 128, 0, st>>>(mask, sort.n, sort.cfg.windows, sort.cfg.buckets, sort.offsets, sort.entries, dtot);
  CK(cudaMemcpyAsync(host_total, dtot, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  cudaFree(mask);
  cudaFree(dtot);
  return cudaGetLastError();
}
template cudaError_t msm_count_madds<Fq>(const MsmSort &, const MsmTable<Fq> &, uint32_t, unsigned long long *, cudaStream_t);
template cudaError_t msm_count_madds<Fq2>(const MsmSort &, const MsmTable<Fq2> &, uint32_t, unsigned long long *, cudaStream_t);
template cudaError_t msm_accumulate<Fq>(const MsmSort &, const MsmTable<Fq> *, int, uint32_t, MsmWork<Fq> &, uint32_t, cudaStream_t, size_t, uint32_t);
template cudaError_t msm_accumulate<Fq2>(const MsmSort &, const MsmTable<Fq2> *, int, uint32_t, MsmWork<Fq2> &, uint32_t, cudaStream_t, size_t, uint32_t);
template cudaError_t msm_reduce<Fq>(MsmWork<Fq> &, uint32_t, uint32_t, XYZZ<Fq> *, cudaStream_t);
template cudaError_t msm_reduce<Fq2>(MsmWork<Fq2> &, uint32_t, uint32_t, XYZZ<Fq2> *, cudaStream_t);
template cudaError_t msm_run<Fq>(const MsmSort &, const MsmTable<Fq> *, int, uint32_t, MsmWork<Fq> &, XYZZ<Fq> *,
                                 cudaStream_t);
template cudaError_t msm_run<Fq2>(const MsmSort &, const MsmTable<Fq2> *, int, uint32_t, MsmWork<Fq2> &, XYZZ<Fq2> *,
                                  cudaStream_t);

}  // namespace zkb
